"""Developer tool (GPU box, needs tools/_exp/libscn_STALLS.so: `tools/build_exp.sh STALLS`): for one
gather-GEMM launch, the cycles CTA 0's roles spend inside each kind of mbarrier wait - which ring paces the step?"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["SCN_B200_LIB_PATH"] = os.path.join(ROOT, "tools", "_exp", "libscn_STALLS.so")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "automatic-as-built-reconstruction_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
import sparseconvnet as scn  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
s = int(sys.argv[2]) if len(sys.argv) > 2 else 1
xyz = bench.building(300000)
a = xyz * (50 / 2 ** s)
a -= a.min(0)
locs = torch.from_numpy(a).long()
locs = torch.cat([locs, torch.zeros(len(locs), 1, dtype=torch.long)], 1)
ss = [4096 >> s, 4096 >> s, 512 >> s]
dev = torch.device("cuda", 0)
feats = torch.randn(len(locs), C, device=dev)
conv = scn.SubmanifoldConvolution(3, C, C, 3, False).to(dev)
lib = ctypes.CDLL(os.environ["SCN_B200_LIB_PATH"])
roles = {0: "producer warp 0", 7: "producer warp 7", 8: "MMA issuer", 9: "weight loader", 10: "metadata loader",
         12: "epilogue warp 0", 16: "converter warp 0"}
kinds = ["metadata", "emptyA", "fullA", "fullB", "converted", "tmem", "mma-issue", "role total"]
# MMA issuer: "metadata" also counts its tcgen05 fences, "emptyA" is its tcgen05.commit time
for prec in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["fp32", "tf32", "bf16"]):
    scn.set_conv_precision(prec)
    x = scn.InputLayer(3, ss, 4)([locs, feats])
    buf = np.zeros((20, 8), dtype=np.int64)
    with torch.no_grad():
        for _ in range(3):
            conv(x)
        torch.cuda.synchronize()
        conv(x)
        lib.scn_debug_stalls_read(buf.ctypes.data_as(ctypes.c_void_p))
    cta = np.zeros((160, 4), dtype=np.int64)
    lib.scn_debug_cta_read(cta.ctypes.data_as(ctypes.c_void_p))
    cta = cta[cta[:, 3] > 0]
    t0 = cta[:, 2].min()
    end = (cta[:, 3] - t0) / 1e3
    dur = (cta[:, 3] - cta[:, 2]) / 1e3
    print("%s: %d CTAs, steps/CTA min %d mean %.0f max %d, items/CTA %d-%d; MMA role ends at %.1f .. %.1f us (mean %.1f), "
          "us per step: min %.3f mean %.3f max %.3f" % (prec, len(cta), cta[:, 0].min(), cta[:, 0].mean(), cta[:, 0].max(),
          cta[:, 1].min(), cta[:, 1].max(), end.min(), end.max(), end.mean(),
          (dur / cta[:, 0]).min(), (dur / cta[:, 0]).mean(), (dur / cta[:, 0]).max()))
    print("%s  C=%d scale %d  (us at 1.965 GHz, CTA 0)" % (prec, C, s))
    for w, name in roles.items():
        if buf[w, 7] == 0:
            continue
        print("  %-18s total %7.1f | " % (name, buf[w, 7] / 1965.0) +
              "  ".join("%s %6.1f" % (kinds[i], buf[w, i] / 1965.0) for i in range(7) if buf[w, i]))

# ---- weight-gradient kernel (work item 0) ---------------------------------------------------------------------
dw_roles = {0: "producer warp 0 (+epilogue)", 7: "producer warp 7", 8: "MMA issuer", 9: "pair-list loader", 10: "converter warp 0"}
dw_kinds = ["pair list", "stage empty", "stage full", "-", "converted", "done", "-", "role total"]
for prec in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["fp32", "tf32"]):
    scn.set_conv_precision(prec)
    x = scn.InputLayer(3, ss, 4)([locs, feats])
    x.features.requires_grad_(True)
    y = conv(x)
    g = torch.ones_like(y.features)
    for _ in range(2):
        y = conv(x)
        y.features.backward(g)
    buf = np.zeros((20, 8), dtype=np.int64)
    lib.scn_debug_stalls_dw_read(buf.ctypes.data_as(ctypes.c_void_p))
    print("dW %s  C=%d scale %d  (us at 1.965 GHz, work item 0)" % (prec, C, s))
    for w, name in dw_roles.items():
        if buf[w, 7] == 0:
            continue
        print("  %-28s total %7.1f | " % (name, buf[w, 7] / 1965.0) +
              "  ".join("%s %6.1f" % (dw_kinds[i], buf[w, i] / 1965.0) for i in range(6) if buf[w, i]))
