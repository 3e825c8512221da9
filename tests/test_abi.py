"""The C-ABI library loads (no GPU needed) and exports every symbol include/scn_b200.h declares;
the ctypes binding table covers exactly the same set."""
import ctypes
import os
import re

from conftest import PKG, ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "scn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(PKG, "libscn_b200.so"))
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "libscn_b200.so does not export " + n


def test_binding_table_matches_header():
    import sparseconvnet._lib as L
    assert sorted(L.EXPORTS) == _declared()
    assert L.lib.scn_n_rulebook_bits() == 32
    assert L.lib.scn_version() >= 100


def test_product_does_not_import_oracle():
    """the product package must never reach into oracle/ (no CPU path behind the CUDA library)"""
    for d, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                s = open(os.path.join(d, f)).read()
                assert "scn_oracle" not in s and "ref_backbone" not in s and "oracle/" not in s, f


def test_cpu_tensor_is_rejected_loudly():
    import pytest
    import torch
    import sparseconvnet as scn
    conv = scn.SubmanifoldConvolution(3, 4, 4, 3, False)
    t = scn.SparseConvNetTensor(torch.zeros(3, 4), scn.Metadata(3), torch.tensor([8, 8, 8]))
    with pytest.raises(RuntimeError, match="CUDA"):
        conv(t)


def test_reference_import_paths_resolve():
    """the import lines of the reference files around the path work unchanged: `import sparseconvnet as scn`
    (fpn_net.py:6) and `from SparseConvNet.sparseconvnet.tools_3d_2d import sparse_3d_to_dense_2d`
    (maskrcnn_benchmark/layers/roi_align_rotated_3d.py:7)"""
    import sparseconvnet as scn
    from SparseConvNet.sparseconvnet.tools_3d_2d import sparse_3d_to_dense_2d
    import SparseConvNet.sparseconvnet as scn2
    assert scn2 is scn and callable(sparse_3d_to_dense_2d)
    assert scn2.FPN_Net is scn.FPN_Net
