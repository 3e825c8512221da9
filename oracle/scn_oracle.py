"""scn_oracle.py - CPU restatement of the reference's sparse3d-backbone algorithms.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this file; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg do, and only as the checker.

Plain numpy (integer work) + torch CPU fp32 (the reference's CPU path is ATen fp32 matmul).
Every function cites the reference file:line it restates (paths relative to
/root/reference/SparseConvNet/sparseconvnet/SCN/).  Parity pinning: the reference ships no tests
or golden vectors for this path, so this restatement is pinned against
outputs of the reference itself compiled in this container (oracle/_ref, built by
oracle/build_ref.py) - tests/test_oracle_vs_ref.py does that whenever oracle/_ref is present -
and against the committed fixtures in tests/golden/ that oracle/make_golden.py generated from
the compiled reference.

Row numbering: the input layer numbers sites by first occurrence (deterministic).  Sites created
by strided convolutions are numbered by the reference in sparsehash iteration order, which is
not reproducible (sparsehash is an un-vendored dependency); here they are numbered per sample by
first touch in input-row order.  All comparisons are therefore made on canonical
(coordinate-level) form - see `canonical_pairs`.
"""
import numpy as np
import torch


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def _keys(coords4, ss):
    """(x,y,z,b) int64 [n,4] -> unique int64 key, batch-major (so sorting keys = canonical order)"""
    c = np.asarray(coords4, dtype=np.int64)
    X, Y, Z = (int(ss[0]) + 2, int(ss[1]) + 2, int(ss[2]) + 2)
    return ((c[:, 3] * X + (c[:, 0] + 1)) * Y + (c[:, 1] + 1)) * Z + (c[:, 2] + 1)


class Lookup(object):
    """coordinate -> row lookup over one scale (stands in for SparseGridMap, Metadata.h:24-27)"""

    def __init__(self, coords4, ss):
        self.ss = [int(s) for s in ss]
        k = _keys(coords4, ss)
        self.order = np.argsort(k, kind="stable")
        self.sorted = k[self.order]

    def find(self, coords4):
        """rows (or -1) of the given coordinates; out-of-range coordinates are never present"""
        c = np.asarray(coords4, dtype=np.int64)
        ok = np.ones(len(c), bool)
        for d in range(3):
            ok &= (c[:, d] >= -1) & (c[:, d] <= self.ss[d])
        k = _keys(np.where(ok[:, None], c, 0), self.ss)
        pos = np.searchsorted(self.sorted, k)
        pos = np.minimum(pos, max(len(self.sorted) - 1, 0))
        hit = ok & (len(self.sorted) > 0)
        if len(self.sorted):
            hit &= self.sorted[pos] == k
        out = np.full(len(c), -1, np.int64)
        out[hit] = self.order[pos[hit]]
        return out


def canonical_rank(coords4, ss):
    """rank of every row in (batch,x,y,z) lexicographic order: the canonical relabelling of
    SURVEY.md section 8c"""
    k = _keys(coords4, ss)
    rank = np.empty(len(k), np.int64)
    rank[np.argsort(k, kind="stable")] = np.arange(len(k))
    return rank


def canonical_pairs(pairs, in_rank, out_rank):
    """(in,out) row pairs -> sorted array of canonical (in,out) pairs ("set of pairs")"""
    p = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    if len(p) == 0:
        return p
    q = np.stack([in_rank[p[:, 0]], out_rank[p[:, 1]]], 1)
    return q[np.lexsort((q[:, 1], q[:, 0]))]


# --------------------------------------------------------------------------------------------
# voxel quantisation (data3d/suncg_utils/suncg_dataset.py:126-188, data3d/data.py:25-37)
# --------------------------------------------------------------------------------------------
def quantize_points(xyz, scale, full_scale, batch_idx=0):
    """a = xyz*scale in fp64; a -= a.min(0); keep 0 <= a < full_scale; truncate to int64"""
    a = np.asarray(xyz, dtype=np.float64) * float(scale)
    if len(a):
        a = a - a.min(0)
    keep = ((a >= 0) & (a < np.asarray(full_scale, dtype=np.float64)[None])).all(1)
    c = a[keep].astype(np.int64)
    return np.concatenate([c, np.full((len(c), 1), batch_idx, np.int64)], 1), keep


def voxelize_batch(buildings, scale, full_scale, matrix=None, xyz_feature=True):
    """dataset + collate for a batch of raw buildings (float32 [n_i, C], columns 0-2 = xyz in metres):
    suncg_dataset.py:126-137 a = np.matmul(a, m) in float64 with m = eye(3) * scale (times the augmentations);
    :140-147 a += -a.min(0); :160-162 b[:, 0:3] = a / scale; :171-188 keep 0 <= a < full_scale, a.long();
    data3d/data.py:25-37 concatenate, append the sample index.  Returns (locs int64 [N,4], feats float32 [N,C])."""
    m = np.eye(3) * float(scale) if matrix is None else np.asarray(matrix, dtype=np.float64).reshape(3, 3)
    locs, feats = [], []
    for i, pts in enumerate(buildings):
        pts = np.asarray(pts)
        assert pts.dtype == np.float32
        a = np.matmul(pts[:, 0:3], m)                      # float32 @ float64 -> float64, as in the reference
        if len(a):
            a = a + (-a.min(0))
        b = pts.copy()
        if xyz_feature:
            b[:, 0:3] = a / scale
        keep = (a.min(1) >= 0) * np.all(a < np.asarray(full_scale)[np.newaxis, :], 1) if len(a) else np.zeros(0, bool)
        a, b = a[keep], b[keep]
        locs.append(np.concatenate([a.astype(np.int64), np.full((len(a), 1), i, np.int64)], 1))
        feats.append(b)
    return np.concatenate(locs), np.concatenate(feats)


# --------------------------------------------------------------------------------------------
# input layer (Metadata/IOLayersRules.h:19-125, CPU/IOLayers.cpp:12-46)
# --------------------------------------------------------------------------------------------
def input_layer_rules(coords, mode=4):
    """coords int64 [N,3|4].  Returns (site_coords [nActive,4] in first-occurrence order,
    point_row [N], header [mode,maxActive,nIn,nOut], table [nActive,1+maxActive] int32) -
    the reference rulebook format of IOLayersRules.h:112-124 (mode 3/4) / :100-111 (mode 1/2)."""
    c = np.asarray(coords, dtype=np.int64)
    if c.shape[1] == 3:
        c = np.concatenate([c, np.zeros((len(c), 1), np.int64)], 1)
    n = len(c)
    if mode == 0:  # IOLayersRules.h:29-58: points are guaranteed unique
        return c.copy(), np.arange(n), [0, 1, n, n], np.zeros((n, 2), np.int32)
    ss = c[:, :3].max(0) + 1 if n else np.array([1, 1, 1])
    k = _keys(c, ss)
    _, first, inv = np.unique(k, return_index=True, return_inverse=True)
    # site id = rank of the first occurrence among first occurrences (IOLayersRules.h:78-92)
    site_of_unique = np.empty(len(first), np.int64)
    site_of_unique[np.argsort(first, kind="stable")] = np.arange(len(first))
    point_row = site_of_unique[inv]
    n_active = len(first)
    site_coords = np.zeros((n_active, 4), np.int64)
    site_coords[point_row[::-1]] = c[::-1]          # first occurrence wins
    counts = np.bincount(point_row, minlength=n_active)
    if mode in (1, 2):
        max_active = 1 if n else 0
        table = np.zeros((n_active, 2), np.int32)
        table[:, 0] = 1
        if mode == 1:    # front()
            table[point_row[::-1], 1] = np.arange(n)[::-1]
        else:            # back()
            table[point_row, 1] = np.arange(n)
    else:
        max_active = int(counts.max()) if n else 0
        table = np.zeros((n_active, 1 + max_active), np.int32)
        table[:, 0] = counts
        order = np.argsort(point_row, kind="stable")          # members ascending point id
        start = np.concatenate([[0], np.cumsum(counts)])[:-1]
        slot = np.arange(n) - start[point_row[order]]
        table[point_row[order], 1 + slot] = order
    return site_coords, point_row, [mode, max_active, n, n_active], table


def input_layer_forward(feats, header, table):
    """CPU/IOLayers.cpp:12-29: out[row] = sum_i mult*in[idx_i], mult = 1/count in mode 4"""
    mode, max_active, n_in, n_out = header
    f = torch.as_tensor(feats, dtype=torch.float32)
    if mode == 0:
        return f.clone()
    out = torch.zeros(n_out, f.shape[1])
    t = torch.as_tensor(np.asarray(table), dtype=torch.int64)
    for j in range(max_active):
        sel = t[:, 0] > j
        mult = (1.0 / t[sel, 0].float())[:, None] if mode == 4 else 1.0
        out[sel] += mult * f[t[sel, 1 + j]]
    return out


# --------------------------------------------------------------------------------------------
# rulebooks
# --------------------------------------------------------------------------------------------
def _offsets(filter_size):
    """row-major enumeration of the filter box, LAST dim fastest (RectangularRegions.h:31-38)"""
    fx, fy, fz = [int(v) for v in filter_size]
    return [(ix, iy, iz) for ix in range(fx) for iy in range(fy) for iz in range(fz)]


def submanifold_rules(site_coords, ss, filter_size):
    """Metadata/SubmanifoldConvolutionRules.h:13-45: for each active output site o and offset k
    (in = o + (i - floor(K/2))): emit (in,out) into rules[k] when `in` is active."""
    c = np.asarray(site_coords, dtype=np.int64)
    lk = Lookup(c, ss)
    f = [int(v) for v in filter_size]
    rules = []
    for (ix, iy, iz) in _offsets(f):
        q = c.copy()
        q[:, 0] += ix - f[0] // 2
        q[:, 1] += iy - f[1] // 2
        q[:, 2] += iz - f[2] // 2
        rows = lk.find(q)
        out = np.nonzero(rows >= 0)[0]
        rules.append(np.stack([rows[out], out], 1).astype(np.int32))
    return rules


def conv_rules(in_coords, in_ss, filter_size, stride, out_ss):
    """Metadata/ConvolutionRules.h:12-34 + RectangularRegions.h:97-119: every input site i feeds
    the output cells j with j*stride <= i <= j*stride+size-1 (clipped to the output extent) at
    offset k = offset of i inside j's window; output sites are created on first touch.
    Returns (out_coords [nOut,4], rules list of [n_k,2] (in,out) int32).  Output rows are
    batch-contiguous ascending (ConvolutionRules.h:80-88), first-touch order inside a sample."""
    c = np.asarray(in_coords, dtype=np.int64)
    f = [int(v) for v in filter_size]
    s = [int(v) for v in stride]
    o = [int(v) for v in out_ss]
    n = len(c)
    cand_in, cand_out, cand_k = [], [], []
    lb = np.zeros((n, 3), np.int64)
    cnt = np.zeros((n, 3), np.int64)
    for d in range(3):
        # RectangularRegions.h:111-119, C++ truncating division (operands may be negative)
        num = c[:, d] - f[d] + s[d]
        l = np.where(num >= 0, num // s[d], -((-num) // s[d]))
        l = np.maximum(l, 0)
        u = np.minimum(c[:, d] // s[d], o[d] - 1)
        lb[:, d] = l
        cnt[:, d] = u - l + 1
    R = [(f[d] + s[d] - 1) // s[d] for d in range(3)]
    for jx in range(R[0]):
        for jy in range(R[1]):
            for jz in range(R[2]):
                ok = (jx < cnt[:, 0]) & (jy < cnt[:, 1]) & (jz < cnt[:, 2])
                idx = np.nonzero(ok)[0]
                j = lb[idx] + np.array([jx, jy, jz])
                off = c[idx, :3] - j * np.array(s)
                k = (off[:, 0] * f[1] + off[:, 1]) * f[2] + off[:, 2]
                cand_in.append(idx)
                cand_out.append(np.concatenate([j, c[idx, 3:4]], 1))
                cand_k.append(k)
    K = f[0] * f[1] * f[2]
    if n == 0:
        return np.zeros((0, 4), np.int64), [np.zeros((0, 2), np.int32) for _ in range(K)]
    cin = np.concatenate(cand_in)
    cout = np.concatenate(cand_out)
    ck = np.concatenate(cand_k)
    # first touch in input-row order (then candidate order): sort candidates by input row
    order = np.argsort(cin, kind="stable")
    cin, cout, ck = cin[order], cout[order], ck[order]
    keys = _keys(cout, o)
    _, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    batch_of_unique = cout[first, 3]
    rank_key = batch_of_unique * (len(cin) + 1) + first     # batch-major, then first touch
    row_of_unique = np.empty(len(first), np.int64)
    row_of_unique[np.argsort(rank_key, kind="stable")] = np.arange(len(first))
    out_row = row_of_unique[inv]
    out_coords = np.zeros((len(first), 4), np.int64)
    out_coords[row_of_unique] = cout[first]
    rules = []
    for k in range(K):
        sel = ck == k
        rules.append(np.stack([cin[sel], out_row[sel]], 1).astype(np.int32))
    return out_coords, rules


def sparse_to_dense_rules(site_coords, ss):
    """Metadata/ConvolutionRules.h:110-128: (row, linear offset (x*Y+y)*Z+z) per sample"""
    c = np.asarray(site_coords, dtype=np.int64)
    off = (c[:, 0] * int(ss[1]) + c[:, 1]) * int(ss[2]) + c[:, 2]
    return np.stack([np.arange(len(c)), off], 1).astype(np.int64), c[:, 3].copy()


# --------------------------------------------------------------------------------------------
# compute (CPU/Convolution.cpp:9-185, CPU/Deconvolution.cpp:8-77)
# --------------------------------------------------------------------------------------------
def conv_forward(x, weight, rules, n_out, bias=None, swap=False):
    """Y = bias; for k: Y[out_k] += X[in_k] @ W[k]   (swap=True: deconvolution, roles swapped,
    CPU/Deconvolution.cpp:34-37).  weight [K,1,Cin,Cout]."""
    x = torch.as_tensor(x, dtype=torch.float32)
    w = torch.as_tensor(weight, dtype=torch.float32)
    y = torch.zeros(n_out, w.shape[3])
    if bias is not None and bias.numel():
        y += bias
    for k, r in enumerate(rules):
        if len(r) == 0:
            continue
        r = torch.as_tensor(np.asarray(r), dtype=torch.int64)
        i, o = (r[:, 1], r[:, 0]) if swap else (r[:, 0], r[:, 1])
        y.index_add_(0, o, x[i] @ w[k, 0])
    return y


def conv_backward(x, d_out, weight, rules, swap=False):
    """dX = 0; for k: dW[k] = X[in_k]^T @ dY[out_k]; dX[in_k] += dY[out_k] @ W[k]^T
    (CPU/Convolution.cpp:82-115).  Returns (dX, dW, dBias)."""
    x = torch.as_tensor(x, dtype=torch.float32)
    dy = torch.as_tensor(d_out, dtype=torch.float32)
    w = torch.as_tensor(weight, dtype=torch.float32)
    dx = torch.zeros_like(x)
    dw = torch.zeros_like(w)
    for k, r in enumerate(rules):
        if len(r) == 0:
            continue
        r = torch.as_tensor(np.asarray(r), dtype=torch.int64)
        i, o = (r[:, 1], r[:, 0]) if swap else (r[:, 0], r[:, 1])
        dw[k, 0] = x[i].t() @ dy[o]
        dx.index_add_(0, i, dy[o] @ w[k, 0].t())
    return dx, dw, dy.sum(0)


# --------------------------------------------------------------------------------------------
# BatchNormalization + (Leaky)ReLU (CPU/BatchNormalization.cpp:13-107)
# --------------------------------------------------------------------------------------------
def bn_forward(x, weight, bias, running_mean, running_var, eps, momentum, train, leakiness):
    """Returns (y, save_mean, save_invstd); running stats updated in place when train."""
    x = torch.as_tensor(x, dtype=torch.float32)
    n = x.shape[0]
    if train:
        mean = x.sum(0) / n
        s = (x * x).sum(0) - mean * mean * n                  # :30-33
        running_mean.mul_(momentum).add_((1 - momentum) * mean)
        running_var.mul_(momentum).add_((1 - momentum) * s / (n - 1))
        invstd = torch.pow(s / n + eps, -0.5)                 # :39 (biased)
    else:
        mean = running_mean.clone()
        invstd = torch.pow(running_var + eps, -0.5)           # :43-44
    w = invstd * (weight if weight is not None else 1.0)
    b = -mean * w + (bias if bias is not None else 0.0)
    y = x * w + b
    y = torch.where(y > 0, y, y * leakiness)
    return y, mean, invstd


def bn_backward(x, y, d_out, save_mean, save_invstd, weight, leakiness):
    """CPU/BatchNormalization.cpp:64-107.  Returns (dX, dWeight, dBias)."""
    x = torch.as_tensor(x, dtype=torch.float32)
    n = x.shape[0]
    d = d_out * torch.where(y > 0, torch.ones_like(y), torch.full_like(y, leakiness))
    gsum = d.sum(0)
    dotp = ((x - save_mean) * d).sum(0)
    g = weight if weight is not None else torch.ones_like(save_mean)
    k = dotp * save_invstd * save_invstd / n
    dx = (d - gsum / n - (x - save_mean) * k) * save_invstd * g
    return dx, dotp * save_invstd, gsum


# --------------------------------------------------------------------------------------------
# SparseToDense (CPU/SparseToDense.cpp:8-64)
# --------------------------------------------------------------------------------------------
def sparse_to_dense(x, site_coords, ss, batch_size):
    x = torch.as_tensor(x, dtype=torch.float32)
    c = torch.as_tensor(np.asarray(site_coords), dtype=torch.int64)
    out = torch.zeros(batch_size, x.shape[1], int(ss[0]), int(ss[1]), int(ss[2]))
    out[c[:, 3], :, c[:, 0], c[:, 1], c[:, 2]] = x
    return out


# --------------------------------------------------------------------------------------------
# synthetic SUNCG-shaped building (benchmark-input spec)
# --------------------------------------------------------------------------------------------
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
        k = 0
        got = 0
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


def to_input(xyz_list, scale=50, full=(4096, 4096, 512), seed=0):
    """mirrors suncg_dataset.py:126-188 + data.py:25-37: (locs int64 [N,4], feats f32 [N,9])"""
    torch.manual_seed(seed)
    locs, feats = [], []
    for b, xyz in enumerate(xyz_list):
        a = xyz * scale
        a -= a.min(0)
        keep = (a < np.array(full)[None]).all(1)
        a = a[keep]
        l = torch.from_numpy(a).long()
        locs.append(torch.cat([l, torch.full((len(l), 1), b, dtype=torch.long)], 1))
        f = torch.randn(len(l), 9)
        f[:, 0:3] = torch.from_numpy(a / scale).float()
        feats.append(f)
    return torch.cat(locs), torch.cat(feats)


def seeded_state_dict(shapes, seed=0):
    """deterministic FPN_Net parameters from (key -> shape), independent of any module's init order: fixtures
    that would mostly consist of random weights store the seed instead (tests/golden/wide_net.npz).  Keys
    are visited in sorted order, one CPU generator per key; convolution weights N(0, 2/(K Cin)), BN scales
    1 + 0.1 N(0,1), BN shifts 0.1 N(0,1), running statistics 0 / 1."""
    sd = {}
    for i, k in enumerate(sorted(shapes)):
        shp = tuple(int(v) for v in shapes[k])
        g = torch.Generator().manual_seed(1000003 * (seed + 1) + i)
        if k.endswith("running_mean"):
            sd[k] = torch.zeros(shp)
        elif k.endswith("running_var"):
            sd[k] = torch.ones(shp)
        elif len(shp) == 4:
            sd[k] = torch.randn(shp, generator=g) * (2.0 / (shp[0] * shp[2])) ** 0.5
        elif k.endswith("weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            sd[k] = 0.1 * torch.randn(shp, generator=g)
    return sd


def subsample(a, limit=4096, stride=5):
    """fixture view of a large tensor: flat, every `stride`-th element when it has more than `limit`"""
    a = np.asarray(a).reshape(-1)
    return a[::stride] if a.size > limit else a


# --------------------------------------------------------------------------------------------
# the whole backbone on the restated ops, any dtype (float64 = ground truth for tolerance studies)
# --------------------------------------------------------------------------------------------
class OracleBackbone(object):
    """FPN_Net graph (fpn_net.py:95-265) evaluated with the restated ops above through torch autograd
    in `dtype`.  With dtype=float64 it serves as ground truth to separate rounding noise from real
    discrepancies: tests compare |impl - f64| with |reference_fp32 - f64|."""

    def __init__(self, state_dict, full_scale, n_planes, rpn_map_sizes, dtype=torch.float64, leakiness=0.0,
                 eps=1e-4, fpn_scales_from_top=(4, 3, 2, 1), roi_scales_from_top=(4, 3),
                 rpn_3d_2d_selector=(1, 2, 3, 4, 5, 6), device="cpu"):
        # device: where the float arithmetic of this CHECKER runs (the integer rules stay numpy on the host);
        # "cuda" makes the float64 truth of a full-size building a matter of seconds
        self.dt, self.dev = dtype, torch.device(device)
        self.p = {k: v.detach().clone().to(device=self.dev, dtype=dtype) for k, v in state_dict.items()}
        for k, v in self.p.items():
            if "running_" not in k:
                v.requires_grad_(True)
        self.full_scale = [int(v) for v in full_scale]
        self.n_scales = len(n_planes)
        self.rpn_map_sizes = [list(s) for s in rpn_map_sizes]
        self.leak, self.eps = leakiness, eps
        self.fpn, self.roi, self.sel = list(fpn_scales_from_top), list(roi_scales_from_top), list(rpn_3d_2d_selector)
        self.rules = {}

    def _conv(self, x, w, rules, n_out, swap=False):
        y = torch.zeros(n_out, w.shape[3], dtype=self.dt, device=self.dev)
        for k, r in enumerate(rules):
            if len(r):
                r = torch.as_tensor(np.asarray(r), dtype=torch.int64).to(self.dev)
                i, o = (r[:, 1], r[:, 0]) if swap else (r[:, 0], r[:, 1])
                y = y.index_add(0, o, x[i] @ w[k, 0])
        return y

    def _bn(self, key, x):
        n = x.shape[0]
        mean = x.sum(0) / n
        s = (x * x).sum(0) - mean * mean * n
        invstd = torch.pow(s / n + self.eps, -0.5)
        w = invstd * self.p[key + ".weight"]
        y = x * w + (self.p[key + ".bias"] - mean * w)
        return torch.where(y > 0, y, y * self.leak)

    def _subm(self, key, x, loc, ss, fs):
        rk = ("s", tuple(ss), fs)
        if rk not in self.rules:
            self.rules[rk] = submanifold_rules(loc, ss, [fs] * 3)
        return self._conv(x, self.p[key + ".weight"], self.rules[rk], len(loc))

    def _down_rules(self, loc, ss, fs, st):
        out_ss = [(a - f) // s + 1 for a, f, s in zip(ss, fs, st)]
        rk = ("c", tuple(ss), tuple(fs), tuple(st))
        if rk not in self.rules:
            self.rules[rk] = conv_rules(loc, ss, fs, st, out_ss)
        return self.rules[rk], out_ss

    def _block(self, key, x, loc, ss):
        h = self._subm(key + ".1.1", self._bn(key + ".1.0", x), loc, ss, 3)
        h = self._subm(key + ".1.3", self._bn(key + ".1.2", h), loc, ss, 3)
        return x + h

    def forward(self, coords, feats):
        loc, prow, header, table = input_layer_rules(np.asarray(coords), 4)
        t = torch.as_tensor(np.asarray(table), dtype=torch.int64).to(self.dev)
        f = torch.as_tensor(feats).to(device=self.dev, dtype=self.dt)
        x = torch.zeros(header[3], f.shape[1], dtype=self.dt, device=self.dev)
        for j in range(header[1]):
            sel = t[:, 0] > j
            x[sel] += (1.0 / t[sel, 0].to(self.dt))[:, None] * f[t[sel, 1 + j]]
        ss = self.full_scale
        x = self._subm("layers_in.1", x, loc, ss, 3)
        downs = []
        for k in range(self.n_scales):
            if k == 0:
                x = self._block("m_downs.0.0", x, loc, ss)
            else:
                (nloc, rules), nss = self._down_rules(loc, ss, [2, 2, 2], [2, 2, 2])
                x = self._conv(self._bn("m_downs.%d.0.0" % k, x), self.p["m_downs.%d.0.1.weight" % k], rules, len(nloc))
                loc, ss = nloc, nss
                x = self._block("m_downs.%d.1" % k, x, loc, ss)
            downs.append((x, loc, ss))
        x = self._subm("m_shortcuts.%d" % (self.n_scales - 1), x, loc, ss, 1)
        ups = [(x, loc, ss)]
        for k in range(self.n_scales - 1):
            j = self.n_scales - 2 - k
            dx, dloc, dss = downs[j]
            (_, rules), _ = self._down_rules(dloc, dss, [2, 2, 2], [2, 2, 2])
            x = self._conv(self._bn("m_ups.%d.0" % k, x), self.p["m_ups.%d.1.weight" % k], rules, len(dloc), swap=True)
            x = x + self._subm("m_shortcuts.%d" % j, dx, dloc, dss, 1)
            ups.append((self._subm("m_mergeds.%d" % k, x, dloc, dss, 3), dloc, dss))
        maps3d = [ups[i] for i in self.fpn]
        maps2d = []
        for i, (mx, mloc, mss) in enumerate(maps3d):
            (zloc, rules), zss = self._down_rules(mloc, mss, [1, 1, self.rpn_map_sizes[i][2]], [1, 1, 1])
            maps2d.append((self._conv(mx, self.p["convs_pro2d.%d.weight" % i], rules, len(zloc)), zloc, zss))
        both = maps3d + maps2d
        return [both[i] for i in self.sel], [ups[i] for i in self.roi]

    def grads(self):
        return {k: v.grad for k, v in self.p.items() if v.requires_grad and v.grad is not None}
