"""Developer / verification tool (>= 2 GPUs, torchrun): the overlapped per-range all-reduce of GradBucket(module=net)
must give exactly the mean of the ranks' gradients (compared with one blocking all-reduce of a copy)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "automatic-as-built-reconstruction_b200"))
import torch
import torch.distributed as dist
import bench
import sparseconvnet as scn

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
net = scn.FPN_Net(bench.FULL_SCALE, 3, ["xyz", "color", "normal"], 1, bench.PLANES, nPlaneM=128, residual_blocks=True,
                  fpn_scales_from_top=[4, 3, 2, 1], roi_scales_from_top=(4, 3),
                  downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8], rpn_map_sizes=bench.RPN_SIZES, voxel_scale=50,
                  rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95, track_running_stats=False).to(dev).train()
scn.broadcast_parameters(net)
bucket = scn.GradBucket(net.parameters(), module=net)
locs, feats = bench.make_batch(100000, 1, 1, rank)
ld, fd = locs.to(dev), feats.to(dev)
ok = True
def backward_only():
    bucket.zero()
    rpn, roi = net([ld, fd])
    sum((m.features ** 2).sum() for m in list(rpn) + list(roi)).backward()


for it in range(3):
    # reference: the same step, gradients copied only after the device has finished, one blocking all-reduce.  (A
    # copy taken on the main stream right before allreduce_mean() would race with the overlapped collectives - they
    # only wait for the progress events, not for later main-stream work.)
    backward_only()
    torch.cuda.synchronize()
    bucket._fired = False
    local = bucket.flat.clone()
    dist.all_reduce(local)
    local /= world
    torch.cuda.synchronize()
    backward_only()
    fired = bucket._fired
    bucket.allreduce_mean()
    torch.cuda.synchronize()
    err = float((bucket.flat - local).abs().max() / local.abs().max())
    ok = ok and fired and err <= 1e-5
    if rank == 0:
        print("iter %d: overlapped (marks fired: %s, %d ranges) vs blocking all-reduce: max rel diff %.2e" %
              (it, fired, len(bucket.chunks), err))
res = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print("OVERLAP_OK" if res.item() == 1.0 else "OVERLAP_MISMATCH")
dist.destroy_process_group()
