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
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert out[0][0] and out[1][0]
        assert out[0][1] == [0, 2, 4] and out[1][1] == [1, 3]
