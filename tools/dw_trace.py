"""Developer tool (GPU box, needs tools/libscn_exp_TRDW.so): clock64 timeline of CTA 0's roles in one
weight-gradient launch."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["SCN_B200_LIB_PATH"] = os.path.join(ROOT, "tools", "libscn_exp_TRDW.so")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "automatic-as-built-reconstruction_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
import sparseconvnet as scn  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
s = int(sys.argv[2]) if len(sys.argv) > 2 else 1
prec = sys.argv[3] if len(sys.argv) > 3 else "tf32"
xyz = bench.building(300000)
a = xyz * (50 / 2 ** s)
a -= a.min(0)
locs = torch.from_numpy(a).long()
locs = torch.cat([locs, torch.zeros(len(locs), 1, dtype=torch.long)], 1)
ss = [4096 >> s, 4096 >> s, 512 >> s]
dev = torch.device("cuda", 0)
feats = torch.randn(len(locs), C, device=dev)
conv = scn.SubmanifoldConvolution(3, C, C, 3, False).to(dev)
scn.set_conv_precision(prec)
x = scn.InputLayer(3, ss, 4)([locs, feats])
x.features.requires_grad_(True)
lib = ctypes.CDLL(os.environ["SCN_B200_LIB_PATH"])
buf = np.zeros((5, 8192), dtype=np.uint64)
cnt = np.zeros(5, dtype=np.int32)
y = conv(x)
g = torch.ones_like(y.features)
for _ in range(3):
    y = conv(x)
    y.features.backward(g)
lib.scn_debug_trace_read(buf.ctypes.data_as(ctypes.c_void_p), cnt.ctypes.data_as(ctypes.c_void_p), 1)
y = conv(x)
y.features.backward(g)
lib.scn_debug_trace_read(buf.ctypes.data_as(ctypes.c_void_p), cnt.ctypes.data_as(ctypes.c_void_p), 1)
ev = {}
for r in range(3):
    v = buf[r, :cnt[r]]
    ev[r] = [(int(q >> np.uint64(56)), int(q & np.uint64(0xffffffffffffff))) for q in v]
t0 = min(e[0][1] for e in ev.values() if e)
print("counts", cnt)
for r, nm in enumerate(["loader", "producer0", "mma"]):
    print(nm, " ".join("%d:%.2f" % (tag, (t - t0) / 1965.0) for tag, t in ev[r][:72]))
m = [t for tag, t in ev[2] if tag == 1]
print("MMA step period us: mean %.3f over %d steps; first 10: %s" % (np.mean(np.diff(m)) / 1965.0, len(m),
      " ".join("%.2f" % (d / 1965.0) for d in np.diff(m)[:10])))
p1 = [t for tag, t in ev[1] if tag == 1]
p2 = [t for tag, t in ev[1] if tag == 2]
p3 = [t for tag, t in ev[1] if tag == 3]
n = min(len(p1), len(p2), len(p3))
print("producer per step us: wait-for-stage %.3f  issue %.3f  (wait pairs->next) %.3f" %
      (np.mean(np.array(p2[:n]) - np.array(p1[:n])) / 1965.0, np.mean(np.array(p3[:n]) - np.array(p2[:n])) / 1965.0,
       np.mean(np.array(p1[1:n]) - np.array(p3[:n - 1])) / 1965.0))
