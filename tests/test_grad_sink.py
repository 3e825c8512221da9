"""The gradient bucket as the graph executor's sink (sparseconvnet/data_parallel.py GradBucket(module=net),
csrc/graph.cu scn_graph_backward_marked): parameter gradients written straight into the flat bucket must equal
the ones autograd accumulates (same kernels, same buffers' contents), the bucket is ordered by the reverse sweep,
every progress event fires, dead-branch ranges stay zero, and dropped .grad views are re-attached."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-300))


def _net_and_batch(scn):
    torch.manual_seed(3)
    net = scn.FPN_Net([512] * 3, 3, ["xyz", "color", "normal"], 1, [32, 64, 64, 128, 128, 128, 256, 256, 256],
                      nPlaneM=128, residual_blocks=True, fpn_scales_from_top=[4, 3, 2, 1], roi_scales_from_top=(4, 3),
                      downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8],
                      rpn_map_sizes=[[32] * 3, [16] * 3, [8] * 3, [4] * 3], voxel_scale=50,
                      rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95, track_running_stats=False).cuda().train()
    rng = np.random.RandomState(5)
    pts = rng.rand(20000, 3) * [400, 300, 1]
    pts[:, 2] = rng.randint(0, 3, 20000) * 40 + 10
    pts[8000:14000, 0] = rng.randint(0, 4, 6000) * 100 + 5
    pts[8000:14000, 2] = rng.rand(6000) * 120
    locs = torch.from_numpy(np.concatenate([pts, np.zeros((20000, 1))], 1)).long()
    return net, locs, torch.randn(20000, 9).cuda()


def _step(net, locs, feats):
    rpn, roi = net([locs, feats])
    sum((m.features ** 2).sum() + m.features.sum() for m in list(rpn) + list(roi)).backward()


def test_bucket_sink_equals_autograd_accumulation():
    import sparseconvnet as scn
    net, locs, feats = _net_and_batch(scn)
    net.zero_grad(set_to_none=True)
    _step(net, locs, feats)
    want = {k: (None if p.grad is None else p.grad.clone()) for k, p in net.named_parameters()}
    bucket = scn.GradBucket(net.parameters(), module=net, n_chunks=4)
    graph = net._layer_graph()
    assert graph.grad_sink is bucket and len(bucket.chunks) == 4
    ops = [c[2] for c in bucket.chunks]
    assert ops == sorted(ops, reverse=True)                         # ranges finish in bucket order
    assert bucket.chunks[0][0] == 0 and all(a[1] == b[0] for a, b in zip(bucket.chunks, bucket.chunks[1:]))
    launches = []
    for it in range(2):                                             # twice: direct writes overwrite, never accumulate
        bucket.zero()
        k0 = scn.SCN.launch_count()
        _step(net, locs, feats)
        launches.append(scn.SCN.launch_count() - k0)
        torch.cuda.synchronize()
        live = 0
        for k, p in net.named_parameters():
            assert p.grad.data_ptr() == bucket.view_of(p).data_ptr(), k
            if want[k] is None or float(want[k].abs().max()) == 0.0:
                assert float(p.grad.abs().max()) == 0.0, k          # dead branch: stays zero
            else:
                assert rel(p.grad, want[k]) <= 1e-6, (k, it)
                live += 1
        assert live > 60
    # every progress event was recorded: a stream that waits for them completes
    side = torch.cuda.Stream()
    from sparseconvnet import _lib
    for ev in bucket._events:
        _lib.check(_lib.lib.scn_stream_wait_event(side.cuda_stream, ev))
    side.synchronize()
    # views dropped by zero_grad(set_to_none=True) come back before the next sweep writes
    net.zero_grad(set_to_none=True)
    bucket.zero()
    _step(net, locs, feats)
    inside = set(id(p) for p in graph.grad_params)
    assert all(p.grad is not None and p.grad.data_ptr() == bucket._view[id(p)].data_ptr()
               for p in net.parameters() if id(p) in inside)
    bucket.check_views()                                            # parameters outside the graph: re-attached here
    assert all(p.grad is not None and p.grad.data_ptr() == bucket._view[id(p)].data_ptr() for p in net.parameters())
    k = "layers_in.1.weight"
    assert rel(dict(net.named_parameters())[k].grad, want[k]) <= 1e-6
    assert bucket.allreduce_mean() is None                          # world 1: nothing to do


def test_shared_parameter_falls_back_to_per_layer_path():
    import sparseconvnet as scn
    from sparseconvnet import graph as G
    conv = scn.SubmanifoldConvolution(3, 16, 16, 3, False)
    seq = scn.Sequential().add(conv).add(scn.BatchNormReLU(16)).add(conv)
    g = G.LayerGraph(16, [64, 64, 64])
    v = g.emit(seq, 0)
    with pytest.raises(G.Unsupported):
        g.finalize([v])


def test_layer_graph_guards_its_backward():
    """the one-call backward re-reads live Parameters and releases the step's arena: an in-place parameter edit between
    forward and backward, and a second backward, are errors (not silently wrong gradients)"""
    import sparseconvnet as scn
    net, locs, feats = _net_and_batch(scn)
    rpn, roi = net([locs, feats])
    loss = sum((m.features ** 2).sum() for m in list(rpn) + list(roi))
    with torch.no_grad():
        net.layers_in[1].weight.mul_(1.0)                              # in place: bumps the version counter
    with pytest.raises(RuntimeError, match="modified in place"):
        loss.backward()
    rpn, roi = net([locs, feats])
    loss = sum((m.features ** 2).sum() for m in list(rpn) + list(roi))
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second time"):
        loss.backward()
