"""Developer tool (GPU box): where the end-to-end (VoxelLoader) step time goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "automatic-as-built-reconstruction_b200"))
import torch
import bench
import sparseconvnet as scn

scn.set_conv_precision("fp32")
dev = torch.device("cuda", 0)
net = scn.FPN_Net(bench.FULL_SCALE, 3, ["xyz", "color", "normal"], 1, bench.PLANES, nPlaneM=128, residual_blocks=True,
                  fpn_scales_from_top=[4, 3, 2, 1], roi_scales_from_top=(4, 3),
                  downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8], rpn_map_sizes=bench.RPN_SIZES, voxel_scale=50,
                  rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95, track_running_stats=False).to(dev).train()
bucket = scn.GradBucket(net.parameters(), module=net)
locs, feats, raw = bench.make_batch(300000, 1, 1, 0, with_raw=True)
raw_pin = [r.pin_memory() for r in raw]
ld, fd = locs.to(dev), feats.to(dev)

def step(p, f):
    bucket.zero()
    rpn, roi = net([p, f])
    loss = sum((m.features ** 2).sum() for m in list(rpn) + list(roi))
    loss.backward()
    return loss

def wall(fn, n=10):
    torch.cuda.synchronize(); t = time.perf_counter(); fn(n); torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3

for _ in range(3): step(ld, fd)
def vox_only(n):
    for _ in range(n): scn.voxelize_batch(raw_pin, 50, bench.FULL_SCALE)
print("voxelize_batch alone      %.2f ms" % wall(vox_only))
def vox_prep(n):
    for _ in range(n):
        c, f = scn.voxelize_batch(raw_pin, 50, bench.FULL_SCALE); net.prepare(c)
print("voxelize + prepare alone  %.2f ms" % wall(vox_prep))
def prep_only(n):
    for _ in range(n): net.prepare(ld)
print("prepare(dev coords) alone %.2f ms" % wall(prep_only))
def loader_only(n):
    for p, f in scn.VoxelLoader((raw_pin for _ in range(n)), net.prepare, 50, bench.FULL_SCALE): pass
print("loader only               %.2f ms" % wall(loader_only))
def loader_steps(n, item=True):
    for p, f in scn.VoxelLoader((raw_pin for _ in range(n)), net.prepare, 50, bench.FULL_SCALE):
        l = step(p, f)
        if item: l.item()
print("loader + step + item      %.2f ms" % wall(loader_steps, 20))
print("loader + step (no item)   %.2f ms" % wall(lambda n: loader_steps(n, False), 20))
pf = scn.InputPrefetcher(net.prepare)
def pf_steps(n, item=True):
    pf.submit(ld)
    for i in range(n):
        p = pf.get()
        if i + 1 < n: pf.submit(ld)
        l = step(p, fd)
        if item: l.item()
print("prefetcher + step + item  %.2f ms" % wall(pf_steps, 20))
print("prefetcher + step         %.2f ms" % wall(lambda n: pf_steps(n, False), 20))
c, f = scn.voxelize_batch(raw_pin, 50, bench.FULL_SCALE)
print("voxeliser rows", c.shape, "dataset rows", ld.shape, "equal", torch.equal(c, ld), "feats equal", torch.equal(f, fd))
def pf_steps_vox(n):
    pf.submit(c)
    for i in range(n):
        p = pf.get()
        if i + 1 < n: pf.submit(c)
        step(p, f).item()
print("prefetcher(vox coords)+step+item %.2f ms" % wall(pf_steps_vox, 20))

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def bench_like(n, use_flush=True, zero=True):
    for p, f in scn.VoxelLoader((raw_pin for _ in range(n)), net.prepare, 50, bench.FULL_SCALE):
        if use_flush: flush.fill_(1)
        step(p, f).item()
print("bench-like loader loop (flush) x30  %.2f ms" % wall(bench_like, 30))
print("bench-like loader loop (no flush) x30 %.2f ms" % wall(lambda n: bench_like(n, False), 30))
import threading
print("threads", threading.active_count())
def ev_timed(n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    bench_like(n)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print("event-timed bench-like x30 %.2f ms" % ev_timed(30))
print("event-timed bench-like x30 %.2f ms" % ev_timed(30))

def per_step(n):
    ts = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    it = iter(scn.VoxelLoader((raw_pin for _ in range(n)), net.prepare, 50, bench.FULL_SCALE))
    while True:
        ta = time.perf_counter()
        try:
            p, f = next(it)
        except StopIteration:
            break
        tb = time.perf_counter()
        flush.fill_(1)
        l = step(p, f)
        tc = time.perf_counter()
        l.item()
        td = time.perf_counter()
        ts.append(((tb - ta) * 1e3, (tc - tb) * 1e3, (td - tc) * 1e3))
    return ts
for rep in range(3):
    ts = per_step(16)
    print("wait-loader / enqueue / wait-gpu ms:", " | ".join("%.1f %.1f %.1f" % t for t in ts))
