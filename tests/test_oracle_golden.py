"""The CPU restatement (oracle/scn_oracle.py) against the golden vectors generated from the
compiled reference (tests/golden/*.npz, oracle/make_golden.py) and against the hand-checked
known answers of SURVEY.md Appendix C."""
import numpy as np
import pytest

import scn_oracle as O


def _check_rulebooks(g):
    coords, ss = g["coords"], g["ss"].tolist()
    loc0, prow, header, table = O.input_layer_rules(coords, 4)
    assert header == g["in_header"].tolist()
    assert np.array_equal(table, g["in_table"])
    assert np.array_equal(loc0, g["loc0"])                       # first-occurrence order is canonical
    r0 = O.canonical_rank(loc0, ss)
    for k, r in enumerate(O.submanifold_rules(loc0, ss, [3, 3, 3])):
        assert np.array_equal(O.canonical_pairs(r, r0, r0), g["sub3_%d" % k]), k
    ss1 = [s // 2 for s in ss]
    loc1, rules = O.conv_rules(loc0, ss, [2, 2, 2], [2, 2, 2], ss1)
    r1 = O.canonical_rank(loc1, ss1)
    assert np.array_equal(loc1[np.argsort(r1)], g["loc1_sorted"])
    assert (np.diff(loc1[:, 3]) >= 0).all()                      # batch-contiguous ascending
    assert (np.diff(g["loc1_batch_col"]) >= 0).all()
    for k, r in enumerate(rules):
        assert np.array_equal(O.canonical_pairs(r, r0, r1), g["conv2_%d" % k]), k
    # overlapping 3/2
    keep = g["odd_keep"]
    so, soo = g["odd_ss"].tolist(), g["odd_out_ss"].tolist()
    lo, _, _, _ = O.input_layer_rules(coords[keep], 4)
    ro = O.canonical_rank(lo, so)
    lout, rules = O.conv_rules(lo, so, [3, 3, 3], [2, 2, 2], soo)
    rout = O.canonical_rank(lout, soo)
    assert np.array_equal(lout[np.argsort(rout)], g["odd_loc_out_sorted"])
    for k, r in enumerate(rules):
        assert np.array_equal(O.canonical_pairs(r, ro, rout), g["conv3s2_%d" % k]), k
    # z-collapse
    ssz = [ss1[0], ss1[1], 1]
    lz, rules = O.conv_rules(loc1, ss1, [1, 1, ss1[2]], [1, 1, 1], ssz)
    rz = O.canonical_rank(lz, ssz)
    assert np.array_equal(lz[np.argsort(rz)], g["locz_sorted"])
    for k, r in enumerate(rules):
        assert np.array_equal(O.canonical_pairs(r, r1, rz), g["zc_%d" % k]), k
    rows, sample = O.sparse_to_dense_rules(loc1, ss1)
    got = np.stack([r1[rows[:, 0]], rows[:, 1], sample], 1)
    assert np.array_equal(got[np.argsort(got[:, 0])], g["s2d"])


def test_appendix_c_golden(gold):
    _check_rulebooks(gold("appendix_c"))


def test_cloud_golden(gold):
    _check_rulebooks(gold("cloud_rulebooks"))


def test_appendix_c_hand_checked():
    """literal values of SURVEY.md Appendix C (independent of any .npz)"""
    coords = np.array([[0, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 0], [3, 3, 3, 0], [2, 2, 2, 0], [1, 0, 0, 1],
                       [0, 0, 0, 1]])
    loc, prow, header, table = O.input_layer_rules(coords, 4)
    assert header == [4, 2, 7, 6]
    assert table.reshape(-1).tolist() == [2, 0, 2, 1, 1, 0, 1, 3, 0, 1, 4, 0, 1, 5, 0, 1, 6, 0]
    assert loc.tolist() == [[0, 0, 0, 0], [0, 0, 1, 0], [3, 3, 3, 0], [2, 2, 2, 0], [1, 0, 0, 1], [0, 0, 0, 1]]
    sub = O.submanifold_rules(loc, [8, 8, 8], [3, 3, 3])
    want = {0: [(3, 2)], 4: [(5, 4)], 12: [(0, 1)], 13: [(i, i) for i in range(6)], 14: [(1, 0)], 22: [(4, 5)],
            26: [(2, 3)]}
    for k in range(27):
        assert sorted(map(tuple, sub[k].tolist())) == sorted(want.get(k, [])), k
    loc1, rules = O.conv_rules(loc, [8, 8, 8], [2, 2, 2], [2, 2, 2], [4, 4, 4])
    assert sorted(map(tuple, loc1.tolist())) == [(0, 0, 0, 0), (0, 0, 0, 1), (1, 1, 1, 0)]
    coord_pairs = {k: sorted((tuple(loc[i]), tuple(loc1[o])) for i, o in r.tolist()) for k, r in enumerate(rules)}
    assert coord_pairs[0] == sorted([((0, 0, 0, 0), (0, 0, 0, 0)), ((2, 2, 2, 0), (1, 1, 1, 0)),
                                     ((0, 0, 0, 1), (0, 0, 0, 1))])
    assert coord_pairs[1] == [((0, 0, 1, 0), (0, 0, 0, 0))]
    assert coord_pairs[4] == [((1, 0, 0, 1), (0, 0, 0, 1))]
    assert coord_pairs[7] == [((3, 3, 3, 0), (1, 1, 1, 0))]
    assert all(len(coord_pairs[k]) == 0 for k in (2, 3, 5, 6))
    rows, sample = O.sparse_to_dense_rules(loc1, [4, 4, 4])
    got = sorted((int(s), tuple(loc1[r][:3]), int(off)) for (r, off), s in zip(rows.tolist(), sample.tolist()))
    assert got == [(0, (0, 0, 0), 0), (0, (1, 1, 1), 21), (1, (0, 0, 0), 0)]
    lz, rz = O.conv_rules(loc1, [4, 4, 4], [1, 1, 4], [1, 1, 1], [4, 4, 1])
    zc = sorted((tuple(loc1[i]), tuple(lz[o]), k) for k, r in enumerate(rz) for i, o in r.tolist())
    assert zc == sorted([((0, 0, 0, 0), (0, 0, 0, 0), 0), ((1, 1, 1, 0), (1, 1, 0, 0), 1),
                         ((0, 0, 0, 1), (0, 0, 0, 1), 0)])


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_input_modes_shapes(mode):
    rng = np.random.RandomState(3)
    c = rng.randint(0, 5, (60, 4))
    c[:, 3] = np.sort(rng.randint(0, 2, 60))
    if mode == 0:
        c = np.unique(c, axis=0)
    loc, prow, header, table = O.input_layer_rules(c, mode)
    assert header[0] == mode and header[2] == len(c) and header[3] == len(loc)
    assert np.array_equal(loc[prow], c)


def test_quantize_matches_numpy_spec():
    xyz = O.building(5000, seed=2)
    coords, keep = O.quantize_points(xyz, 50, [4096, 4096, 512], 1)
    locs, _ = O.to_input([xyz])
    assert np.array_equal(coords[:, :3], locs[:, :3].numpy()) and keep.all() and (coords[:, 3] == 1).all()


def test_empty_inputs():
    loc, prow, header, table = O.input_layer_rules(np.zeros((0, 4), np.int64), 4)
    assert header == [4, 0, 0, 0] and len(loc) == 0
    assert all(len(r) == 0 for r in O.submanifold_rules(loc, [8, 8, 8], [3, 3, 3]))
    l1, rules = O.conv_rules(loc, [8, 8, 8], [2, 2, 2], [2, 2, 2], [4, 4, 4])
    assert len(l1) == 0 and len(rules) == 8


def test_zcollapse64_rules_match_reference_golden(gold):
    """filter volume 64 ([1,1,64] z-collapse): the restated rule generator against the reference Metadata"""
    g = gold("zcollapse64")
    ss, zs = g["ss"].tolist(), [int(g["ss"][0]), int(g["ss"][1]), 1]
    loc0, _, _, _ = O.input_layer_rules(g["coords"], 4)
    locz, rules = O.conv_rules(loc0, ss, [1, 1, 64], [1, 1, 1], zs)
    r0, rz = O.canonical_rank(loc0, ss), O.canonical_rank(locz, zs)
    assert np.array_equal(locz[np.argsort(rz)], g["locz_sorted"])
    assert len(rules) == 64
    for k, r in enumerate(rules):
        assert np.array_equal(O.canonical_pairs(r, r0, rz), g["zc_%d" % k]), k


def test_voxelize_batch_is_the_per_building_quantiser_plus_collate():
    raw = []
    for i, n in enumerate((4000, 0, 2500)):
        xyz = O.building(n, seed=i).astype(np.float32) if n else np.zeros((0, 3), np.float32)
        raw.append(np.concatenate([xyz, np.random.RandomState(i).randn(n, 6).astype(np.float32)], 1))
    locs, feats = O.voxelize_batch(raw, 50, [4096, 4096, 512])
    want = [O.quantize_points(b[:, :3].astype(np.float64), 50, [4096, 4096, 512], i)[0] for i, b in enumerate(raw)]
    assert np.array_equal(locs, np.concatenate(want))
    assert feats.dtype == np.float32 and feats.shape == (len(locs), 9)
    assert np.array_equal(feats[:, 3:], np.concatenate([b[:, 3:] for b in raw]))
    assert np.abs(feats[:, :3] * 50 - locs[:, :3]).max() < 1.0 + 1e-3     # feature xyz = voxel-space position / scale
