"""INTEGRATION.md Option B, executed as far as a machine without a GPU can take it: the reference's OWN, unmodified
Python package (every sparseconvnet/*.py, linked - not copied - from /root/reference into a throw-away directory) with
its compiled extension replaced by this repo's two files SCN.py + _lib.py, in a fresh interpreter.  The reference's
`import sparseconvnet` succeeds on it, its FPN_Net builds, `scn.Metadata(3)` is this library's Metadata, and a forward
call travels through the reference's own InputLayer / autograd Function code into libscn_b200's entry point (which
refuses the CPU tensor loudly - there is no GPU here; the arithmetic behind the entry points is what the -m gpu tests
check).  Only runs where the reference tree exists (the build container)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import PKG, ROOT

REF_PKG = "/root/reference/SparseConvNet/sparseconvnet"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_PKG), reason="reference tree not present (GPU box)")

SCRIPT = r'''
import json, os, sys
import torch
import sparseconvnet as scn
out = {"package_file": os.path.realpath(scn.__file__), "scn_file": os.path.realpath(scn.SCN.__file__)}
m = scn.Metadata(3)                                   # metadata.py:16-17 -> getattr(sparseconvnet.SCN, "Metadata_3")()
out["metadata_class"] = type(m).__module__ + "." + type(m).__name__
out["rulebook_bits"] = scn.SCN.n_rulebook_bits()
net = scn.FPN_Net([512] * 3, 3, ["xyz", "color", "normal"], 1, [32, 64, 32, 32, 32, 32, 32, 32, 32], nPlaneM=32,
                  residual_blocks=True, fpn_scales_from_top=[4, 3, 2, 1], roi_scales_from_top=(4, 3),
                  downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8], rpn_map_sizes=[[32] * 3, [16] * 3, [8] * 3, [4] * 3],
                  voxel_scale=50, rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95, track_running_stats=False)
out["fpn_net_module"] = type(net).__module__
out["state_dict"] = {k: list(v.shape) for k, v in net.state_dict().items()}
coords = torch.randint(0, 64, (50, 3))
coords = torch.cat([coords, torch.zeros(50, 1, dtype=torch.long)], 1)
try:                                                   # the reference's forward: fpn_net.py:140-152 -> ioLayers.py:51-64
    net([coords, torch.randn(50, 9)])
    out["forward"] = "returned"
except RuntimeError as e:
    out["forward"] = "RuntimeError: " + str(e)
try:
    scn.SCN.MaxPooling_updateOutput(None, None, None, None, None, None, None, None, 0)
    out["off_path"] = "returned"
except NotImplementedError as e:
    out["off_path"] = "NotImplementedError"
print("RESULT " + json.dumps(out))
'''


def test_reference_package_runs_on_this_extension(tmp_path):
    pkg = tmp_path / "sparseconvnet"
    pkg.mkdir()
    for f in sorted(os.listdir(REF_PKG)):
        if f.endswith(".py"):
            os.symlink(os.path.join(REF_PKG, f), pkg / f)            # the reference's files, where they lie
    assert not (pkg / "SCN.py").exists()                              # (its SCN is a compiled extension, absent here)
    for f in ("SCN.py", "_lib.py"):                                   # the whole of Option B: two files
        os.symlink(os.path.join(PKG, "sparseconvnet", f), pkg / f)
    env = dict(os.environ, PYTHONPATH=str(tmp_path), SCN_B200_LIB_PATH=os.path.join(PKG, "libscn_b200.so"))
    env.pop("SCN_B200_PRECISION", None)
    r = subprocess.run([sys.executable, "-c", SCRIPT], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, r.stdout[-2000:]
    out = json.loads(line[0][7:])
    # the package that ran is the reference's, the extension under it this repo's
    assert out["package_file"] == os.path.join(REF_PKG, "__init__.py")
    assert out["scn_file"] == os.path.join(PKG, "sparseconvnet", "SCN.py")
    assert out["metadata_class"] == "sparseconvnet.SCN.Metadata_3" and out["rulebook_bits"] == 32
    assert out["fpn_net_module"] == "sparseconvnet.fpn_net"
    # the network the reference built: the same parameters as this repo's FPN_Net and as the golden the reference wrote
    import numpy as np
    gold = np.load(os.path.join(ROOT, "tests", "golden", "wide_net.npz"))
    want = {k[6:]: [int(v) for v in gold[k]] for k in gold.files if k.startswith("shape/")}
    assert out["state_dict"] == want
    # its forward reached this library's entry point through the reference's own layer code
    # (ioLayers.py:177 -> SCN.InputLayer_updateOutput -> _lib.require_cuda_f32)
    assert out["forward"].startswith("RuntimeError: sparseconvnet (B200): InputLayer features must be a CUDA tensor"), out["forward"]
    assert out["off_path"] == "NotImplementedError"
