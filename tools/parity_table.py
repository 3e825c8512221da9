"""per-tensor three-way parity table at BASELINE configs[1] size (developer tool; oracle/parity.py):
    python tools/parity_table.py [precision ...] > gpurun_out/parity_table.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("automatic-as-built-reconstruction_b200", "oracle", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch  # noqa: E402
import bench  # noqa: E402
import parity as P  # noqa: E402
import sparseconvnet as scn  # noqa: E402

locs, feats = bench.make_batch(300000, 1, 1, 0)
sd = bench.reference_state_dict()
torch.set_num_threads(os.cpu_count())
_, ref_maps, ref_grads = P.reference_step(sd, locs, feats, bench.REF_CFG)
truth_maps, truth_grads = P.truth_step(sd, locs, feats, bench.REF_CFG, device="cuda")
for prec in (sys.argv[1:] or ["fp32", "fp32_ffma"]):
    scn.set_conv_precision(prec)
    net = scn.FPN_Net(bench.FULL_SCALE, 3, ["xyz", "color", "normal"], 1, bench.PLANES, nPlaneM=128, residual_blocks=True,
                      fpn_scales_from_top=[4, 3, 2, 1], roi_scales_from_top=(4, 3),
                      downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8], rpn_map_sizes=bench.RPN_SIZES, voxel_scale=50,
                      rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95, track_running_stats=False)
    net.load_state_dict(sd, strict=False)   # (reference_state_dict omits the unused layers_out / linear)
    net = net.cuda().train()
    rpn, roi = net([locs, feats.cuda()])
    sum((m.features ** 2).sum() for m in list(rpn) + list(roi)).backward()
    maps = [(m.get_spatial_locations().numpy(), m.features.detach().cpu(), m.spatial_size.tolist())
            for m in list(rpn) + list(roi)]
    grads = {k: p.grad.detach().cpu() for k, p in net.named_parameters() if p.grad is not None}
    rep = P.three_way(maps, grads, ref_maps, ref_grads, truth_maps, truth_grads)
    print("==== %s: features %s gradients %s l2 %s" % (prec, rep["features"], rep["gradients"], rep["gradient_vector_l2"]))
    for i, e in enumerate(rep["maps"]):
        print("map%d rows %6d  gpu_vs_ref %.2e  gpu_vs_fp64 %.2e  ref_vs_fp64 %.2e" %
              (i, e["rows"], e["gpu_vs_ref"], e["gpu_vs_fp64"], e["ref_vs_fp64"]))
    for k, e in rep["grads"].items():
        if "missing" in e:
            print("%-28s missing %s" % (k, e["missing"]))
            continue
        print("%-28s gpu_vs_ref %.2e  gpu_vs_fp64 %.2e  ref_vs_fp64 %.2e  ratio %.2f  max|truth| %.3e %s" %
              (k, e["gpu_vs_ref"], e["gpu_vs_fp64"], e["ref_vs_fp64"], e["gpu_vs_fp64"] / max(e["ref_vs_fp64"], 1e-30),
               float(truth_grads[k].abs().max()), "" if e["gpu_vs_fp64"] <= e["ref_vs_fp64"] else "<<"))
