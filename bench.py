"""bench.py - the sparse3d backbone (FPN_Net) forward+backward on synthetic SUNCG-shaped buildings.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mode train|infer]
                    [--batch B] [--precision fp32|fp32_ffma|tf32|bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

--mode infer times the evaluation forward (no_grad, batch statistics as every shipped config has
track_running_stats=False) of one 2M-point 3-storey building per GPU with no collective (BASELINE configs[4]).

A step = one pass of the hot path over one batch: fresh Metadata (voxel hashing + every rulebook
rebuilt, as in real training), FPN_Net forward, loss = sum(features^2) over the 6 rpn + 2 roi maps,
backward (live graph only) and, for N > 1, the gradient all-reduce over NCCL.  Workload at N=1 is
BASELINE.json configs[1]: one 300k-point building, batch 1 (SURVEY.md section 8d generator).
Metric: active voxels/s (sum over samples of nActive at scale 0 / time), whole job over all GPUs.

  value     inputs resident in HBM when the timed region starts
  value_pruned  (extra, not the headline) the step with FPN_Net.prune_dead_branches = True
  value_inline  the same step through the reference's own call net([coords, feats]) - no prefetcher, voxel
            hashing and all rulebooks built inside the step
  e2e       same step through the public API from pinned HOST buffers (coords + features H2D and a
            D2H read of the loss inside the timed region)
  roofline  conv gather-GEMM class: algorithmic bytes (SURVEY.md 8d formula) / CUDA-event time of
            those launches, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference: the compiled reference CPU SparseConvNet (oracle/_ref) driving
            the same graph on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "automatic-as-built-reconstruction_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

FULL_SCALE = [4096, 4096, 512]
PLANES = [32, 64, 64, 128, 128, 128, 256, 256, 256]
RPN_SIZES = [[256, 256, 32], [128, 128, 16], [64, 64, 8], [32, 32, 4]]


# ---- synthetic input (benchmark-input spec of SURVEY.md Appendix D.3; not product code) ---------
PREFETCH_DEPTH = int(os.environ.get("SCN_BENCH_PREFETCH_DEPTH", "1"))


def building(n, L=(19.0, 16.3, 3.0), floors=1, seed=0):
    rng = np.random.RandomState(seed)
    L = np.array(L)
    pts = []
    per = n // (6 * floors)
    for f in range(floors):
        z0 = f * L[2]
        for z in (z0, z0 + L[2] - 1e-3):
            p = rng.rand(per, 3) * L
            p[:, 2] = z
            pts.append(p)
        k = got = 0
        while got < 4 * per:
            m = per // 2
            p = rng.rand(m, 3) * L
            p[:, 2] += z0
            if k % 2 == 0:
                p[:, 0] = (k // 2 % 5) * (L[0] - 1e-3) / 4
            else:
                p[:, 1] = (k // 2 % 5) * (L[1] - 1e-3) / 4
            pts.append(p)
            k += 1
            got += m
    p = np.concatenate(pts)
    if len(p) < n:
        p = np.concatenate([p, p[:n - len(p)]])
    return p[:n]


def to_input(xyz_list, scale=50, full=FULL_SCALE, seed=0):
    """the dataset + collate of the reference on raw float32 buildings (data3d/suncg_utils/suncg_dataset.py:126-188,
    data3d/data.py:25-37; numpy float64, benchmark-input spec - the product's GPU front end is scn.voxelize_batch).
    Returns (locs int64 [N,4], feats float32 [N,9], raw = list of float32 [n_i,9] buildings: xyz in metres + 6
    feature columns)."""
    torch.manual_seed(seed)
    locs, feats, raw = [], [], []
    for b, xyz in enumerate(xyz_list):
        pts = np.concatenate([xyz.astype(np.float32), torch.randn(len(xyz), 6).numpy()], 1)
        raw.append(torch.from_numpy(pts))
        a = np.matmul(pts[:, 0:3], np.eye(3) * scale)          # float32 @ float64 -> float64
        a = a + (-a.min(0))
        keep = (a.min(1) >= 0) * (a < np.array(full)[None]).all(1)
        f = pts.copy()
        f[:, 0:3] = a / scale
        a, f = a[keep], f[keep]
        l = torch.from_numpy(a).long()
        locs.append(torch.cat([l, torch.full((len(l), 1), b, dtype=torch.long)], 1))
        feats.append(torch.from_numpy(f))
    return torch.cat(locs), torch.cat(feats), raw


def make_batch(points, floors, batch, first_seed, with_raw=False):
    locs, feats, raw = to_input([building(points, floors=floors, seed=first_seed + i) for i in range(batch)])
    return (locs, feats, raw) if with_raw else (locs, feats)


def n_active0(locs):
    k = ((locs[:, 3] * 4096 + locs[:, 0]) * 4096 + locs[:, 1]) * 512 + locs[:, 2]
    return int(torch.unique(k).numel())


# ---- clocks ---------------------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region through NVML (in-process, every
    5 ms) - the same counters as the profiling recipe's nvidia-smi clocks line."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        self.idx, self.sm, self.mask, self.stop_flag, self.thread, self.h = gpu_index, [], 0, False, None, None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            self.h = None
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        while not self.stop_flag:
            try:
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self.stop_flag = True
        self.thread.join()
        busy = [v for v in self.sm if v > 0.5 * max(self.sm)] if self.sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": self.max_mhz,
                "reasons": [n for n, bit in self.REASONS if self.mask & bit], "samples": len(self.sm)}


# ---- reference (CPU) arm -----------------------------------------------------------------------------
def reference_state_dict():
    """random-init weights of the full-size architecture, built without touching the GPU library"""
    torch.manual_seed(0)
    sd = {}

    def conv(key, k, a, b):
        sd[key + ".weight"] = torch.empty(k, 1, a, b).normal_(0, (2.0 / a / k) ** 0.5)

    def bn(key, c):
        sd[key + ".weight"], sd[key + ".bias"] = torch.ones(c), torch.zeros(c)
        sd[key + ".running_mean"], sd[key + ".running_var"] = torch.zeros(c), torch.ones(c)

    def block(key, c):
        bn(key + ".1.0", c), conv(key + ".1.1", 27, c, c), bn(key + ".1.2", c), conv(key + ".1.3", 27, c, c)

    conv("layers_in.1", 27, 9, PLANES[0])
    for k, c in enumerate(PLANES):
        if k == 0:
            block("m_downs.0.0", c)
        else:
            bn("m_downs.%d.0.0" % k, PLANES[k - 1]), conv("m_downs.%d.0.1" % k, 8, PLANES[k - 1], c)
            block("m_downs.%d.1" % k, c)
        conv("m_shortcuts.%d" % k, 1, c, 128)
    for k in range(8):
        bn("m_ups.%d.0" % k, 128), conv("m_ups.%d.1" % k, 8, 128, 128), conv("m_mergeds.%d" % k, 27, 128, 128)
    for i, s in enumerate(RPN_SIZES):
        conv("convs_pro2d.%d" % i, s[2], 128, 128)
    return sd


REF_CFG = dict(full_scale=FULL_SCALE, n_planes=PLANES, rpn_map_sizes=RPN_SIZES)
# stated feature tolerances of the reduced-precision modes (tests/test_full_parity.py)
REDUCED_FEATURE_TOL = {"tf32": 1e-2, "bf16": 5e-2}


def _parity():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import parity
    return parity


def cpu_reference_step(sd, locs, feats, train=True):
    """one fwd+bwd (train) / eval forward of the compiled reference CPU path (fresh Metadata); returns seconds"""
    return _parity().reference_step(sd, locs, feats, REF_CFG, train=train)[0]


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count()
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank): the reference's own
    # OpenMP loops (per-sample rule building) read the variable when its extension is loaded, ATen follows torch
    os.environ["OMP_NUM_THREADS"] = str(cores)
    torch.set_num_threads(cores)
    # one reference step of the full 300k-point workload takes ~4.2 s on the box's host cores: up to 30 steps
    # (~2 minutes) run the workload itself, longer runs a bounded 60k-point sample of it
    train = args.mode == "train"
    points = args.points if (args.steps + args.warmup) <= 30 else min(args.points, 60000)
    if not train and (args.steps + args.warmup) * args.points > 12_000_000:      # 2M-point forward: ~10 s each
        points = min(args.points, 300000)
    locs, feats = make_batch(points, args.floors, args.batch, 0)
    na = n_active0(locs)
    sd = reference_state_dict()
    sample = "%d building(s) x %d points (nActive %d), %s, fresh Metadata each step" % (
        args.batch, points, na, "fwd+bwd" if train else "eval forward")
    for _ in range(args.warmup):
        cpu_reference_step(sd, locs, feats, train)
    ts = [cpu_reference_step(sd, locs, feats, train) for _ in range(args.steps)]
    sec = float(np.mean(ts))
    val = na / sec
    print(json.dumps({
        "impl": "reference", "metric": "active_voxels_per_sec", "value": val, "unit": "active voxels/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "buildings_per_sec": args.batch / sec,
        "config": workload_config(args, points),
        "cpu_baseline": {"value": val, "unit": "active voxels/s", "cores": cores, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": val, "unit": "active voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(args, points=None):
    pts = points or args.points
    train = args.mode == "train"
    if train:
        which = {(1, 300000, 1): "BASELINE configs[1]", (2, 300000, 1): "BASELINE configs[2] backbone, batch 2 per GPU",
                 (4, 300000, 1): "BASELINE configs[3] backbone, 4 buildings per GPU"}.get(
                     (args.batch, pts, args.floors), "BASELINE configs[1] geometry, non-default size")
    else:
        which = ("BASELINE configs[4]: large-scene inference, one building per GPU, no collective"
                 if (args.batch, pts, args.floors) == (1, 2000000, 3) else "inference, non-default size")
    return {"workload": "FPN_Net sparse3d backbone %s, %d x %dk-point synthetic building per GPU (%s)"
                        % ("fwd+bwd" if train else "eval forward (no_grad)", args.batch, pts // 1000, which),
            "mode": args.mode,
            "per_gpu_batch": args.batch, "points_per_building": points or args.points, "floors": args.floors,
            "full_scale": FULL_SCALE, "planes": PLANES, "precision": args.precision,
            "parallelism": "dp%d" % args.gpus,
            "prefetch": "Metadata (hash grids + rulebooks) of batch i+1 built on a side stream during batch i"
                        " (value, e2e); value_inline = plain net([coords, feats]), builds inside the step",
            "l2": "256 MiB flush buffer written between steps; per-step activations (GBs) exceed the 126 MB L2"}


# ---- B200 arm --------------------------------------------------------------------------------------
def run_b200(args, rank, local_rank, world):
    import torch.distributed as dist
    import sparseconvnet as scn
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    scn.set_conv_precision(args.precision)
    torch.manual_seed(0)
    net = scn.FPN_Net(FULL_SCALE, 3, ["xyz", "color", "normal"], 1, PLANES, nPlaneM=128, residual_blocks=True,
                      fpn_scales_from_top=[4, 3, 2, 1], roi_scales_from_top=(4, 3),
                      downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8], rpn_map_sizes=RPN_SIZES, voxel_scale=50,
                      rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95, track_running_stats=False)
    train = args.mode == "train"
    net = net.to(dev).train() if train else net.to(dev).eval()
    scn.broadcast_parameters(net)
    bucket = scn.GradBucket(net.parameters(), module=net) if train else None
    locs, feats, raw = make_batch(args.points, args.floors, args.batch, rank * args.batch, with_raw=True)   # weak scaling
    na_local = n_active0(locs)
    raw_pin = [r.pin_memory() for r in raw]                  # e2e input: the raw float32 buildings on the host
    locs_dev, feats_dev = locs.to(dev), feats.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def train_step(coords, f):
        bucket.zero()
        rpn, roi = net([coords, f])
        # loss = sum of squares over the 6 rpn + 2 roi maps, evaluated on their concatenation: 3 small torch
        # kernels per direction instead of ~20 (the maps are tiny, each of those kernels was pure launch latency)
        loss = torch.cat([m.features.reshape(-1) for m in list(rpn) + list(roi)]).square().sum()
        loss.backward()
        bucket.allreduce_mean()
        return loss

    def infer_step(coords, f):
        # inference (engine/inference_3d.py:17-30): eval mode, no_grad, whole buildings per rank, no collective;
        # the step's result is a checksum of the 8 output maps (what the detector heads would consume)
        with torch.no_grad():
            rpn, roi = net([coords, f])
            return torch.stack([m.features.sum() for m in list(rpn) + list(roi)]).sum()

    step = train_step if train else infer_step

    # The integer work of a batch (voxel hashing, 13 grids, 25 rulebooks) depends on its coordinates
    # only: like the reference's DataLoader workers it runs one batch ahead - here on a side stream
    # and a worker thread (scn.InputPrefetcher).  Every timed step still builds a fresh Metadata; the
    # first build of a timed region is inside the region and not overlapped.
    pf = scn.InputPrefetcher(net.prepare)

    loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_seen = []
    e2e_host_ms = []           # host time between consecutive batches of the e2e loop (diagnostic: stalls show here)

    cuprof_done = []

    def timed(n_steps, host_inputs):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = scn.SCN.launch_count()
        # SCN_BENCH_CUPROF=1: the device-resident timed region is a cudaProfilerStart / Stop range, so that
        # `ncu --profile-from-start off` lists exactly the launches of the timed steps (profiles/README.md)
        cuprof = os.environ.get("SCN_BENCH_CUPROF") == "1" and not host_inputs and not cuprof_done
        if cuprof:
            cuprof_done.append(True)                           # the headline region only, not value_pruned's
            torch.cuda.cudart().cudaProfilerStart()
        a.record()
        if host_inputs:
            # end to end from HOST buffers through the public API: scn.VoxelLoader uploads the raw float32 points
            # (36 B/point), voxelises them on the GPU (the dataset's float64 quantisation + collate) and builds the
            # batch's Metadata one batch ahead; the step's result (the loss) is read back every step
            # The loss reaches the host through a pinned 2-slot ring: copied D2H (non-blocking) every step and read
            # by the host one step later, so the read-back never drains the queue (a training loop that logs its
            # loss with one step of lag); the last values are read before the closing event.
            loader = scn.VoxelLoader((raw_pin for _ in range(n_steps)), net.prepare, 50, FULL_SCALE)
            pending = []
            t_prev = time.perf_counter()
            for i, (prepared, f) in enumerate(loader):
                t_now = time.perf_counter()
                e2e_host_ms.append((t_now - t_prev) * 1e3)
                t_prev = t_now
                flush.fill_(1)
                loss_host[i % 2].copy_(step(prepared, f).detach().reshape(1), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                pending.append((ev, i % 2))
                if len(pending) == 2:
                    e0, slot = pending.pop(0)
                    e0.synchronize()
                    loss_seen.append(float(loss_host[slot][0]))
            for e0, slot in pending:
                e0.synchronize()
                loss_seen.append(float(loss_host[slot][0]))
        else:
            depth = PREFETCH_DEPTH                             # batches in flight ahead of the step that runs
            for j in range(min(depth, n_steps)):
                pf.submit(locs_dev)
            for i in range(n_steps):
                flush.fill_(1)
                prepared = pf.get()
                if i + depth < n_steps:
                    pf.submit(locs_dev)                       # a later batch: overlaps this step
                step(prepared, feats_dev)
        b.record()
        torch.cuda.synchronize()
        if cuprof:
            torch.cuda.cudart().cudaProfilerStop()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), scn.SCN.launch_count() - l0

    for _ in range(args.warmup):
        step(locs_dev, feats_dev)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms, launches = timed(args.steps, False)
    clk = clocks.stop() if rank == 0 else None
    timed(4, True)                                           # warm the loader path (pinned staging, ring, voxeliser)
    del e2e_host_ms[:]
    ms_e2e, _ = timed(args.steps, True)

    def timed_inline(n_steps):
        """the reference's own call, net([coords, feats]) with nothing prepared: voxel hashing and every
        rulebook are built inside the step, on the step's stream (what tools/train_net_sparse3d.py does)"""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n_steps):
            flush.fill_(1)
            step(locs_dev, feats_dev)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    ms_inline = timed_inline(args.steps)

    # B200 extension, reported beside the headline and NOT part of it: FPN_Net.prune_dead_branches skips the layers no
    # returned map depends on (the reference computes them and drops the results, fpn_net.py:186-203; outputs and
    # gradients are bit-identical - tests/test_full_parity.py::test_pruned_dead_branches_change_nothing)
    net.prune_dead_branches = True
    net.invalidate_graph()
    if train:
        bucket = scn.GradBucket(net.parameters(), module=net)     # (its ranges follow the compiled graph's op order)
    for _ in range(args.warmup):
        step(locs_dev, feats_dev)
    ms_pruned, launches_pruned = timed(args.steps, False)
    dead_ops, live_ops = net._layer_graph().n_dead_ops, len(net._layer_graph().ops)
    net.prune_dead_branches = False
    net.invalidate_graph()
    if train:
        bucket = scn.GradBucket(net.parameters(), module=net)
    step(locs_dev, feats_dev)

    na_t = torch.tensor([na_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(na_t)
    na_total = float(na_t.item())

    # per-kernel-class roofline pass (CUDA events around each launch group, outside the headline timing)
    # (everything on one stream here - weight gradients included - so that the classes' event times are the
    # kernels' own durations and do not overlap each other; the rulebook build runs inline on the same stream)
    from sparseconvnet import _lib as _scn_lib
    _scn_lib.lib.scn_set_graph_overlap(0)
    scn.SCN.prof_enable(True)
    scn.SCN.prof_read()
    for _ in range(2):
        flush.fill_(1)
        step(locs_dev, feats_dev)
    prof = scn.SCN.prof_read()
    scn.SCN.prof_enable(False)
    _scn_lib.lib.scn_set_graph_overlap(2)

    if rank == 0:
        peaks = {}
        src = "fallback"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            src = "measured"
        except (OSError, ValueError):
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        g = prof["conv_gemm"]
        ach = g["bytes"] / (g["ms"] * 1e-3) / 1e9 if g["ms"] > 0 else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("conv_gemm")
        except (OSError, ValueError):
            pass
        classes = {k: {"launch_groups": v["regions"] // 2, "ms_per_step": v["ms"] / 2,
                       "algorithmic_gb_per_step": v["bytes"] / 2 / 1e9,
                       "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else None,
                       "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 and v["flops"] else None}
                   for k, v in prof.items()}
        sec = ms * 1e-3 / args.steps
        sec_e2e = ms_e2e * 1e-3 / args.steps
        out = {
            "metric": "active_voxels_per_sec", "value": na_total / sec, "unit": "active voxels/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32 (3xTF32 on tcgen05, fp32 accumulate)", "fp32_ffma": "f32", "tf32": "tf32",
                      "bf16": "bf16 (operands; fp32 accumulate, tf32 weight gradient)"}[args.precision],
            "data": "synthetic",
            "buildings_per_sec": args.batch * world / sec, "active_voxels_per_building": na_local / args.batch,
            "config": workload_config(args),
            "e2e": {"value": na_total / sec_e2e, "unit": "active voxels/s", "ms_per_step": sec_e2e * 1e3,
                    "h2d_bytes_per_step": int(sum(r.numel() for r in raw) * 4), "d2h_bytes_per_step": 4,
                    "host_ms_between_batches": {"median": float(np.median(e2e_host_ms)) if e2e_host_ms else None,
                                                "max": float(max(e2e_host_ms)) if e2e_host_ms else None,
                                                "argmax": int(np.argmax(e2e_host_ms)) if e2e_host_ms else None},
                    "input": "raw float32 points [n, 9] per building in pinned host memory -> scn.VoxelLoader "
                             "(upload, GPU voxelisation + collate, Metadata) -> step -> loss copied D2H into pinned memory every "
                             "step, read by the host one step later"},
            "value_inline": {"value": na_total / (ms_inline * 1e-3 / args.steps), "unit": "active voxels/s",
                             "ms_per_step": ms_inline / args.steps,
                             "what": "plain net([coords, feats]) - no prefetcher, rulebook builds inside the step"},
            "value_pruned": {"value": na_total / (ms_pruned * 1e-3 / args.steps), "unit": "active voxels/s",
                             "ms_per_step": ms_pruned / args.steps, "gpu_launches": int(launches_pruned),
                             "ops_skipped": int(dead_ops), "ops_run": int(live_ops),
                             "what": "NOT the headline: the same step with FPN_Net.prune_dead_branches = True (B200 extension, "
                                     "off by default) - the layers no returned map depends on are skipped; the reference "
                                     "computes them (fpn_net.py:186-203) and so does every other number of this line; "
                                     "outputs and gradients are bit-identical"},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "conv gather-GEMM (fwd + dX), all launches of a step",
                         "achieved": ach, "peak": hbm_peak, "peak_source": src + " (sustained copy)",
                         "unit": "GB/s", "frac": ach / hbm_peak if hbm_peak else None, "traffic": traffic,
                         "launches_per_step": g["regions"] // 2, "ms_per_step": g["ms"] / 2},
            "kernel_classes": classes,
            "kernel_classes_note": "separate pass (2 steps) with every kernel on ONE stream (rulebook build inline, weight "
                                   "gradients not on their companion stream): CUDA-event time per launch group = the "
                                   "kernels' own durations; in the timed steps the rulebook build and the weight "
                                   "gradients overlap the rest, so the classes' sum exceeds ms_per_step; dX and dW each count the dY / pair-list / "
                                   "weight bytes they read, so the classes' algorithmic bytes sum to more than the "
                                   "8.81 GB per-building contract of SURVEY 8d",
            "grad_allreduce_bytes": bucket.nbytes() if (world > 1 and train) else 0,
        }
        failed = None
        if world == 1 and not args.no_cpu_baseline and not train:
            P = _parity()
            cores = os.cpu_count()
            torch.set_num_threads(cores)
            sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
            with torch.no_grad():
                rpn_g, roi_g = net([locs_dev, feats_dev])
            gpu_maps = [(m.get_spatial_locations().numpy(), m.features.detach().cpu(), m.spatial_size.tolist())
                        for m in list(rpn_g) + list(roi_g)]
            t, ref_maps, _ = P.reference_step(sd, locs, feats, REF_CFG, train=False)
            out["cpu_baseline"] = {"value": na_local / t, "unit": "active voxels/s", "cores": cores,
                                   "kind": "reference", "seconds_per_step": t,
                                   "sample": "1 eval forward (same %d-point batch, nActive %d) on the compiled "
                                             "reference CPU SparseConvNet, fresh Metadata, no warm-up"
                                             % (len(locs), na_local)}
            worst, same = 0.0, True
            for g_, r_ in zip(gpu_maps, ref_maps):
                gl, gf = P._canon(*g_)
                rl, rf = P._canon(*r_)
                same = same and gl.shape == rl.shape and bool(np.array_equal(gl, rl))
                if same:
                    worst = max(worst, P._rel(gf, rf))
            tol = REDUCED_FEATURE_TOL.get(args.precision, 1e-4)
            out["parity"] = {"mode": "eval forward (batch statistics: track_running_stats=False)",
                             "active_site_sets_equal": same, "max_rel_feature_err_vs_reference_cpu": worst,
                             "maps_compared": len(ref_maps), "bound": tol,
                             "note": "the reference's own fp32 rounding at this size is ~1e-4 (train-mode "
                                     "three-way report); a float64 evaluation is not run at 2M points",
                             "ok": bool(same and worst <= 2 * tol)}
            if not out["parity"]["ok"]:
                failed = "parity outside the stated bound: %s" % json.dumps(out["parity"])
        if world == 1 and not args.no_cpu_baseline and train:
            # Parity at the benchmark's own size (oracle/parity.py): the 8 output maps and every live parameter
            # gradient of this very batch, library vs the compiled reference CPU run (which is also the
            # cpu_baseline sample) vs a float64 evaluation of the same graph.
            P = _parity()
            cores = os.cpu_count()
            torch.set_num_threads(cores)
            sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
            net.zero_grad(set_to_none=True)
            rpn_g, roi_g = net([locs_dev, feats_dev])
            sum((m.features ** 2).sum() for m in list(rpn_g) + list(roi_g)).backward()
            gpu_maps = [(m.get_spatial_locations().numpy(), m.features.detach().cpu(), m.spatial_size.tolist())
                        for m in list(rpn_g) + list(roi_g)]
            gpu_grads = {k: p.grad.detach().cpu() for k, p in net.named_parameters() if p.grad is not None}
            t, ref_maps, ref_grads = P.reference_step(sd, locs, feats, REF_CFG)
            out["cpu_baseline"] = {"value": na_local / t, "unit": "active voxels/s", "cores": cores,
                                   "kind": "reference", "seconds_per_step": t,
                                   "sample": "1 step (same %d-point batch, nActive %d) fwd+bwd on the compiled "
                                             "reference CPU SparseConvNet, fresh Metadata, no warm-up"
                                             % (len(locs), na_local)}
            if args.batch * args.points <= 700000:
                truth_maps, truth_grads = P.truth_step(sd, locs, feats, REF_CFG, device=dev)
                rep = P.summary(P.three_way(gpu_maps, gpu_grads, ref_maps, ref_grads, truth_maps, truth_grads))
                rep["max_rel_feature_err_vs_reference_cpu"] = rep["features"]["gpu_vs_ref"]
                rep["truth"] = "float64 evaluation of the same graph (oracle/scn_oracle.py OracleBackbone)"
                if args.precision in REDUCED_FEATURE_TOL:
                    tol = REDUCED_FEATURE_TOL[args.precision]
                    rep["stated_feature_tolerance"] = tol
                    rep["ok"] = bool(rep["active_site_sets_equal"] and rep["features"]["gpu_vs_ref"] <= tol)
                out["parity"] = rep
                if not rep["ok"]:
                    failed = "parity outside the stated bound: %s" % json.dumps(rep)
        print(json.dumps(out))
        if failed:
            sys.stdout.flush()
            sys.exit("bench.py: " + failed)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"],
                    help="train: fwd+bwd (+ gradient all-reduce for N > 1); infer: eval forward, no_grad, one "
                         "2M-point 3-storey building per GPU by default (BASELINE configs[4]), no collective")
    ap.add_argument("--batch", type=int, default=1, help="buildings per GPU")
    ap.add_argument("--points", type=int, default=None)
    ap.add_argument("--floors", type=int, default=None)
    ap.add_argument("--precision", default=os.environ.get("SCN_B200_PRECISION", "fp32"),
                    choices=["fp32", "fp32_ffma", "tf32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.points is None:
        args.points = 300000 if args.mode == "train" else 2000000
    if args.floors is None:
        args.floors = 1 if args.mode == "train" else 3
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
