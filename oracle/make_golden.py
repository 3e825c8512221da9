"""make_golden.py - generate tests/golden/*.npz from the REFERENCE ITSELF (TEST INFRASTRUCTURE).

Runs only in the build container (needs /root/reference and oracle/_ref/*.so):
  * assembles a throw-away package /tmp/scn_refpkg/sparseconvnet from the reference's own Python
    layer files (copied to /tmp, never into this repo), with `SCN` bound to oracle/_ref/SCN_ref.so
    and the one py3.12/torch-2.11 fix SURVEY.md section 8c documents
    (convolution.py:36 `/` -> `//`);
  * rulebooks: the reference Metadata<3> through oracle/_ref/SCN_refdump.so;
  * features / gradients: the reference's real fpn_net.py FPN_Net on a reduced-width config.
Fixtures are small (a few hundred kB) and are committed together with this script.

    python oracle/make_golden.py
"""
import os
import shutil
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF_PY = "/root/reference/SparseConvNet/sparseconvnet"
PKG = "/tmp/scn_refpkg"

sys.path.insert(0, HERE)
import ref_backbone  # noqa: E402
import scn_oracle  # noqa: E402

def small_net_cfg():
    """reduced-width FPN_Net on a 512^3 index space (8 stride-2 downsamplings -> [2,2,2] at the top);
    rpn maps are ups[4..1] = scales 4..7 -> sizes 32,16,8,4 cubed"""
    return dict(full_scale=[512, 512, 512], n_planes=[8, 16, 16, 16, 16, 16, 16, 16, 16], n_plane_m=16,
                rpn_map_sizes=[[32, 32, 32], [16, 16, 16], [8, 8, 8], [4, 4, 4]])


def wide_net_cfg():
    """tensor-core-width FPN_Net (every convolution but the 9-channel stem has Cin, Cout in {32, 64}: the
    widths the tcgen05 gather-GEMM / weight-gradient kernels take), same index space and cloud as small_net"""
    return dict(full_scale=[512, 512, 512], n_planes=[32, 64, 32, 32, 32, 32, 32, 32, 32], n_plane_m=32,
                rpn_map_sizes=[[32, 32, 32], [16, 16, 16], [8, 8, 8], [4, 4, 4]])


def reference_package():
    """import the reference's own sparseconvnet python on top of SCN_ref.so"""
    dst = os.path.join(PKG, "sparseconvnet")
    if os.path.isdir(PKG):
        shutil.rmtree(PKG)
    os.makedirs(dst)
    for f in os.listdir(REF_PY):
        if f.endswith(".py"):
            shutil.copy(os.path.join(REF_PY, f), os.path.join(dst, f))
    src = open(os.path.join(dst, "convolution.py")).read()
    assert "/ self.filter_stride" in src
    open(os.path.join(dst, "convolution.py"), "w").write(
        src.replace("/ self.filter_stride", "// self.filter_stride"))
    with open(os.path.join(dst, "SCN.py"), "w") as f:
        f.write("import importlib.util\n"
                "_s = importlib.util.spec_from_file_location('SCN_ref', %r)\n"
                "_m = importlib.util.module_from_spec(_s); _s.loader.exec_module(_m)\n"
                "globals().update({k: getattr(_m, k) for k in dir(_m) if not k.startswith('__')})\n"
                % os.path.join(HERE, "_ref", "SCN_ref.so"))
    sys.path.insert(0, PKG)
    import sparseconvnet as ref_scn
    return ref_scn


def random_cloud(n, extent, batch, seed):
    rng = np.random.RandomState(seed)
    pts = []
    for b in range(batch):
        # clustered: points on a few planes so that neighbourhoods are populated
        c = (rng.rand(n, 3) * np.array(extent)).astype(np.int64)
        c[: n // 3, 2] = extent[2] // 2
        c[n // 3: 2 * n // 3, 0] = extent[0] // 3
        pts.append(np.concatenate([c, np.full((n, 1), b, np.int64)], 1))
    c = np.concatenate(pts)
    # interleave duplicates to exercise mode-4 averaging
    c = np.concatenate([c, c[rng.randint(0, len(c), n // 4)]])
    order = np.argsort(c[:, 3], kind="stable")
    return c[order]


def dump_rulebooks(coords, ss, name):
    D = ref_backbone.scn_refdump()
    L = ref_backbone.L
    m = D.RefMetadata3()
    coords_t = torch.from_numpy(coords)
    m.inputLayer(L(ss), coords_t, 0, 4)
    hdr_tab = m.inputLayerRuleBook()
    out = {"coords": coords, "ss": np.array(ss)}
    out["in_header"] = hdr_tab[0].numpy().reshape(-1)
    out["in_table"] = hdr_tab[1].numpy()
    loc0 = m.getSpatialLocations(L(ss)).numpy()
    out["loc0"] = loc0
    r0 = scn_oracle.canonical_rank(loc0, ss)
    sub = m.getSubmanifoldRuleBook(L(ss), L([3, 3, 3]))
    for k, r in enumerate(sub):
        out["sub3_%d" % k] = scn_oracle.canonical_pairs(r.numpy(), r0, r0).astype(np.int32)
    ss1 = [s // 2 for s in ss]
    conv = m.getRuleBook(L(ss), L(ss1), L([2, 2, 2]), L([2, 2, 2]))
    loc1 = m.getSpatialLocations(L(ss1)).numpy()
    r1 = scn_oracle.canonical_rank(loc1, ss1)
    out["loc1_sorted"] = loc1[np.argsort(r1)]
    out["loc1_batch_col"] = loc1[:, 3]
    for k, r in enumerate(conv):
        out["conv2_%d" % k] = scn_oracle.canonical_pairs(r.numpy(), r0, r1).astype(np.int32)
    # overlapping strided conv 3/2 (general region arithmetic, RectangularRegions.h:111-119)
    ss_o = [(s - 3) // 2 + 1 for s in [s1 * 2 + 1 for s1 in ss1]]
    m2 = D.RefMetadata3()
    ss_odd = [s1 * 2 + 1 for s1 in ss1]
    keep = (coords[:, :3] < np.array(ss_odd)).all(1)
    m2.inputLayer(L(ss_odd), torch.from_numpy(coords[keep]), 0, 4)
    loc_odd = m2.getSpatialLocations(L(ss_odd)).numpy()
    ro = scn_oracle.canonical_rank(loc_odd, ss_odd)
    conv3 = m2.getRuleBook(L(ss_odd), L(ss_o), L([3, 3, 3]), L([2, 2, 2]))
    loc_o = m2.getSpatialLocations(L(ss_o)).numpy()
    r_o = scn_oracle.canonical_rank(loc_o, ss_o)
    out["odd_keep"] = keep
    out["odd_ss"] = np.array(ss_odd)
    out["odd_out_ss"] = np.array(ss_o)
    out["odd_loc_out_sorted"] = loc_o[np.argsort(r_o)]
    for k, r in enumerate(conv3):
        out["conv3s2_%d" % k] = scn_oracle.canonical_pairs(r.numpy(), ro, r_o).astype(np.int32)
    # z-collapse [1,1,Z]/1 from the coarse scale
    zc = m.getRuleBook(L(ss1), L([ss1[0], ss1[1], 1]), L([1, 1, ss1[2]]), L([1, 1, 1]))
    locz = m.getSpatialLocations(L([ss1[0], ss1[1], 1])).numpy()
    rz = scn_oracle.canonical_rank(locz, [ss1[0], ss1[1], 1])
    out["locz_sorted"] = locz[np.argsort(rz)]
    for k, r in enumerate(zc):
        out["zc_%d" % k] = scn_oracle.canonical_pairs(r.numpy(), r1, rz).astype(np.int32)
    s2d = m.getSparseToDenseRuleBook(L(ss1))
    # per-sample rules (row, linear offset); canonicalise rows
    rows = []
    for b, r in enumerate(s2d):
        r = r.numpy().astype(np.int64)
        rows.append(np.stack([r1[r[:, 0]], r[:, 1], np.full(len(r), b)], 1))
    rows = np.concatenate(rows) if rows else np.zeros((0, 3), np.int64)
    out["s2d"] = rows[np.argsort(rows[:, 0])]
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    return out


def zcollapse64():
    """[1,1,64] z-collapse (FPN_Net with RPN_SCALES_FROM_TOP [5,4,3] on a z = 512 index space,
    configs/Stanford_walls/*.yaml): a filter volume of 64 through the reference Metadata"""
    D = ref_backbone.scn_refdump()
    L = ref_backbone.L
    ss, zs = [24, 20, 64], [24, 20, 1]
    rng = np.random.RandomState(4)
    c = np.concatenate([np.concatenate([(rng.rand(1200, 3) * np.array(ss)).astype(np.int64),
                                        np.full((1200, 1), b, np.int64)], 1) for b in range(2)])
    m = D.RefMetadata3()
    m.inputLayer(L(ss), torch.from_numpy(c), 0, 4)
    loc0 = m.getSpatialLocations(L(ss)).numpy()
    r0 = scn_oracle.canonical_rank(loc0, ss)
    rules = m.getRuleBook(L(ss), L(zs), L([1, 1, 64]), L([1, 1, 1]))
    locz = m.getSpatialLocations(L(zs)).numpy()
    rz = scn_oracle.canonical_rank(locz, zs)
    out = {"coords": c, "ss": np.array(ss), "locz_sorted": locz[np.argsort(rz)].astype(np.int16)}
    for k, r in enumerate(rules):
        out["zc_%d" % k] = scn_oracle.canonical_pairs(r.numpy(), r0, rz).astype(np.int32)
    np.savez_compressed(os.path.join(GOLD, "zcollapse64.npz"), **out)
    return out


def appendix_c():
    coords = np.array([[0, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 0], [3, 3, 3, 0], [2, 2, 2, 0], [1, 0, 0, 1],
                       [0, 0, 0, 1]], np.int64)
    return dump_rulebooks(coords, [8, 8, 8], "appendix_c")


def small_net_golden(ref_scn):
    cfg = small_net_cfg()
    torch.manual_seed(0)
    net = ref_scn.FPN_Net(cfg["full_scale"], 3, ["xyz", "color", "normal"], 1, cfg["n_planes"],
                          nPlaneM=cfg["n_plane_m"], residual_blocks=True, fpn_scales_from_top=[4, 3, 2, 1],
                          roi_scales_from_top=(4, 3), downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8],
                          rpn_map_sizes=cfg["rpn_map_sizes"], voxel_scale=50,
                          rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95,
                          track_running_stats=False)
    net.train()
    xyz = [scn_oracle.building(6000, L=(9.0, 8.0, 3.0), floors=1, seed=s) for s in range(2)]
    locs, feats = scn_oracle.to_input(xyz, scale=50, full=cfg["full_scale"], seed=0)
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    rpn, roi = net([locs, feats])
    loss = sum((m.features ** 2).sum() for m in list(rpn) + list(roi))
    loss.backward()
    out = {"locs": locs.numpy(), "feats": feats.numpy(), "loss": np.array(loss.item())}
    for k, v in sd0.items():
        out["sd/" + k] = v.numpy()
    for i, m in enumerate(list(rpn) + list(roi)):
        loc = m.get_spatial_locations().numpy()
        order = np.argsort(scn_oracle.canonical_rank(loc, m.spatial_size.tolist()))
        out["out%d_loc" % i] = loc[order]
        out["out%d_feat" % i] = m.features.detach().numpy()[order]
        out["out%d_batchcol" % i] = loc[:, 3]
    for k, p in net.named_parameters():
        if p.grad is not None:
            out["grad/" + k] = p.grad.numpy()
    for k, v in net.state_dict().items():
        if "running_" in k:
            out["after/" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, "small_net.npz"), **out)
    # eval-mode forward (track_running_stats=False => batch statistics, unbiased variance)
    net.eval()
    with torch.no_grad():
        rpn, roi = net([locs, feats])
    ev = {}
    for i, m in enumerate(list(rpn) + list(roi)):
        loc = m.get_spatial_locations().numpy()
        order = np.argsort(scn_oracle.canonical_rank(loc, m.spatial_size.tolist()))
        ev["out%d_feat" % i] = m.features.numpy()[order]
    np.savez_compressed(os.path.join(GOLD, "small_net_eval.npz"), **ev)
    return out


def wide_net_golden(ref_scn):
    """the reference's own fpn_net.py at tensor-core widths.  Parameters come from scn_oracle.seeded_state_dict
    (the fixture stores the seed, not 4 MB of random weights); parameter gradients above 4096 elements are
    stored subsampled (scn_oracle.subsample)."""
    cfg = wide_net_cfg()
    net = ref_scn.FPN_Net(cfg["full_scale"], 3, ["xyz", "color", "normal"], 1, cfg["n_planes"],
                          nPlaneM=cfg["n_plane_m"], residual_blocks=True, fpn_scales_from_top=[4, 3, 2, 1],
                          roi_scales_from_top=(4, 3), downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8],
                          rpn_map_sizes=cfg["rpn_map_sizes"], voxel_scale=50,
                          rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95,
                          track_running_stats=False)
    seed = 7
    net.load_state_dict(scn_oracle.seeded_state_dict({k: v.shape for k, v in net.state_dict().items()}, seed))
    net.train()
    xyz = [scn_oracle.building(6000, L=(9.0, 8.0, 3.0), floors=1, seed=10 + s) for s in range(2)]
    locs, feats = scn_oracle.to_input(xyz, scale=50, full=cfg["full_scale"], seed=1)
    rpn, roi = net([locs, feats])
    loss = sum((m.features ** 2).sum() for m in list(rpn) + list(roi))
    loss.backward()
    out = {"locs": locs.numpy().astype(np.int16), "feats": feats.numpy(), "loss": np.array(loss.item()),
           "seed": np.array(seed)}
    for k, v in net.state_dict().items():
        out["shape/" + k] = np.array(v.shape)
    for i, m in enumerate(list(rpn) + list(roi)):
        loc = m.get_spatial_locations().numpy()
        order = np.argsort(scn_oracle.canonical_rank(loc, m.spatial_size.tolist()))
        out["out%d_loc" % i] = loc[order].astype(np.int16)
        out["out%d_feat" % i] = m.features.detach().numpy()[order]
    for k, p in net.named_parameters():
        if p.grad is not None:
            out["grad/" + k] = scn_oracle.subsample(p.grad.numpy())
    net.eval()
    with torch.no_grad():
        rpn, roi = net([locs, feats])
    for i, m in enumerate(list(rpn) + list(roi)):
        loc = m.get_spatial_locations().numpy()
        order = np.argsort(scn_oracle.canonical_rank(loc, m.spatial_size.tolist()))
        out["eval%d_feat" % i] = scn_oracle.subsample(m.features.numpy()[order], stride=3)
    np.savez_compressed(os.path.join(GOLD, "wide_net.npz"), **out)
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "wide":      # only the fixtures added in round 2
        wide_net_golden(reference_package())
        print("wide_net.npz", os.path.getsize(os.path.join(GOLD, "wide_net.npz")))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "zc64":
        zcollapse64()
        print("zcollapse64.npz", os.path.getsize(os.path.join(GOLD, "zcollapse64.npz")))
        return
    appendix_c()
    dump_rulebooks(random_cloud(1500, [40, 36, 24], 2, seed=1), [48, 48, 32], "cloud_rulebooks")
    ref_scn = reference_package()
    small_net_golden(ref_scn)
    wide_net_golden(ref_scn)
    zcollapse64()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
