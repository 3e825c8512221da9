"""world_size-2 gloo test (CPU) of the data-parallel host logic: building sharding and the flat
gradient bucket's all-reduce-mean."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    from conftest import PKG  # noqa: F401  (sets sys.path)
    import sparseconvnet as scn
    torch.manual_seed(rank)                       # different initial weights per rank
    lin = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2, bias=False))
    dead = torch.nn.Linear(2, 2)                  # never used: gets no gradient (dead FPN branch)
    params = list(lin.parameters()) + list(dead.parameters())
    scn.broadcast_parameters(torch.nn.ModuleList([lin, dead]))
    bucket = scn.GradBucket(params)
    mine = scn.shard_indices(5, rank, world)
    x = torch.arange(20, dtype=torch.float32).view(5, 4)[mine]
    bucket.zero()
    lin(x).sum().backward()
    local = bucket.flat.clone()
    bucket.allreduce_mean()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    want = sum(gathered) / world
    ok = torch.allclose(bucket.flat, want) and all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in params)
    ok = ok and float(dead.weight.grad.abs().sum()) == 0.0
    w0 = [torch.zeros_like(lin[0].weight) for _ in range(world)]
    dist.all_gather(w0, lin[0].weight.data)
    ok = ok and torch.equal(w0[0], w0[1])
    out[rank] = (ok, mine)
    dist.destroy_process_group()


def test_grad_bucket_allreduce_world2():
    world = 2
    with mp.get_context("spawn").Manager() as mgr:       # (no fork() of the multi-threaded test process)
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert out[0][0] and out[1][0]
        assert out[0][1] == [0, 2, 4] and out[1][1] == [1, 3]


def test_gradient_ranges_end_at_geometric_fractions():
    """GradBucket(module=net): the ranges the reverse sweep finishes end at 1/2, 3/4, 7/8 ... of the graph's gradient
    bytes (the last, un-hidden collective is small), cover them contiguously and finish in bucket order (host logic
    only: no GPU needed)"""
    import torch
    import sparseconvnet as scn
    torch.manual_seed(0)
    net = scn.FPN_Net([512] * 3, 3, ["xyz", "color", "normal"], 1, [32, 64, 64, 128, 128, 128, 256, 256, 256],
                      nPlaneM=128, residual_blocks=True, fpn_scales_from_top=[4, 3, 2, 1], roi_scales_from_top=(4, 3),
                      downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8],
                      rpn_map_sizes=[[32] * 3, [16] * 3, [8] * 3, [4] * 3], voxel_scale=50,
                      rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95, track_running_stats=False)
    bucket = scn.GradBucket(net.parameters(), module=net)
    ch = bucket.chunks
    assert 2 <= len(ch) <= 6
    assert ch[0][0] == 0 and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
    ops = [c[2] for c in ch]
    assert ops == sorted(ops, reverse=True)
    total = ch[-1][1]
    assert ch[0][1] >= total // 2                                  # the first range: at least half of the bytes
    assert ch[-1][1] - ch[-1][0] <= total // 8                     # the last one: a small tail
    inside = set(id(p) for p in net._layer_graph().grad_params)
    assert total == sum((p.numel() + 3) // 4 * 4 for p in net.parameters() if id(p) in inside)
