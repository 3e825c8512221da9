"""make_golden_nms.py - golden vectors for the rotated IoU of SURVEY.md section 8 row f4, produced by the REFERENCE'S
OWN numba kernels (second/core/non_max_suppression/nms_gpu.py:166-404, 552-704) executed by numba's CUDA simulator
on the CPU (NUMBA_ENABLE_CUDASIM=1 - no GPU needed).  TEST INFRASTRUCTURE, runs only in the build container.

The module imports `spconv.utils.non_max_suppression` (spconv 1.x, un-vendored) at load time for an unrelated
axis-aligned NMS; a stub module stands in for that import, nothing of it is called.

    NUMBA_ENABLE_CUDASIM=1 python oracle/make_golden_nms.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

assert os.environ.get("NUMBA_ENABLE_CUDASIM") == "1", "run with NUMBA_ENABLE_CUDASIM=1"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def reference_module():
    stub = types.ModuleType("spconv.utils")
    stub.non_max_suppression = None
    sys.modules.setdefault("spconv", types.ModuleType("spconv"))
    sys.modules["spconv.utils"] = stub
    spec = importlib.util.spec_from_file_location(
        "ref_nms_gpu", "/root/reference/second/core/non_max_suppression/nms_gpu.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def boxes(n, seed):
    """BEV boxes (x, y, size_x, size_y, yaw) like wall proposals: thin, long, clustered so that many overlap"""
    rng = np.random.RandomState(seed)
    c = rng.rand(n, 2) * 6
    d = np.stack([rng.rand(n) * 3 + 0.2, rng.rand(n) * 0.5 + 0.05], 1)
    swap = rng.rand(n) < 0.3
    d[swap] = d[swap][:, ::-1]
    yaw = rng.choice([0, np.pi / 2, -np.pi / 2, 0.3, -1.1, 2.8], n) + rng.randn(n) * 0.05
    b = np.concatenate([c, d, yaw[:, None]], 1).astype(np.float32)
    b[n // 2:n // 2 + 4] = b[:4]                 # exact duplicates
    b[-1, 2:4] = [1e-3, 1e-3]                    # a degenerate box
    return b


def main():
    from numba import cuda
    ref = reference_module()
    sys.modules["ref_nms_gpu"] = ref

    # The simulator runs every CUDA thread as a Python thread and its per-thread swap of the module globals races
    # when the reference's own 64-thread launches (rotate_iou_gpu_eval, nms_gpu.py:667-703) run on a module loaded
    # this way; the golden values are therefore produced by ONE simulated thread that walks the (box, query)
    # pairs exactly as rotate_iou_kernel_eval does (:660-664: dev_iou[n * K + k] = devRotateIoUEval(query[k],
    # box[n], criterion)) and calls the reference's device functions devRotateIoUEval / devRotateIoU themselves.
    @cuda.jit
    def walk(N, K, boxes, query, iou, criterion, plain):
        for n in range(N):
            for k in range(K):
                if plain:
                    iou[n * K + k] = ref.devRotateIoU(query[k * 5:k * 5 + 5], boxes[n * 5:n * 5 + 5])
                else:
                    iou[n * K + k] = ref.devRotateIoUEval(query[k * 5:k * 5 + 5], boxes[n * 5:n * 5 + 5], criterion)

    def run(a, q, criterion=-1, plain=False):
        out = np.zeros(len(a) * len(q), np.float32)
        walk[1, 1](len(a), len(q), a.reshape(-1), q.reshape(-1), out, criterion, plain)
        return out.reshape(len(a), len(q))

    a, q = boxes(48, 0), boxes(37, 1)
    out = {"boxes": a, "query": q}
    for crit in (-1, 0, 1, 2):
        out["iou_c%d" % crit] = run(a, q, crit)
    out["iou_self"] = run(a, a, -1)
    out["iou_plain"] = run(a, q, plain=True)
    np.savez_compressed(os.path.join(GOLD, "rotate_iou.npz"), **out)
    print("rotate_iou.npz", os.path.getsize(os.path.join(GOLD, "rotate_iou.npz")),
          {k: float(np.nanmax(v)) for k, v in out.items() if k.startswith("iou")})


if __name__ == "__main__":
    main()
