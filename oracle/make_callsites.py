"""TEST INFRASTRUCTURE - not part of the product.

Generates tests/golden/scn_callsites.json: every call the reference's own Python layer files make into the
extension module `sparseconvnet.SCN` (the pybind11 boundary of SURVEY 8b), as found by parsing those files where
they lie under /root/reference (nothing is copied: only the call's name, its argument count, its keyword names and
its file:line are recorded).  tests/test_abi.py::test_reference_call_sites_bind checks this repo's
`sparseconvnet.SCN` against the list, so "the reference's layer files bind to this module unchanged" is checked
against the reference's real call sites and not against a hand-written list.

Run here (the build container; /root/reference is absent on the GPU box):  python oracle/make_callsites.py
"""
import ast
import json
import os
import sys

REF = "/root/reference"
PKG = os.path.join(REF, "SparseConvNet", "sparseconvnet")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "scn_callsites.json")

# layer files FPN_Net and its callers instantiate; every other file of the package is off the
# path and its SCN entry points only have to exist as loud NotImplementedError stubs
ON_PATH = {
    "batchNormalization.py", "convolution.py", "deconvolution.py", "ioLayers.py", "networkInNetwork.py",
    "sparseToDense.py", "submanifoldConvolution.py", "metadata.py", "sparseConvNetTensor.py", "tools_3d_2d.py",
    "fpn_net.py", "sequential.py", "tables.py", "identity.py", "utils.py",
}
# ioLayers.py also holds the BL (batch x length) layers, which FPN_Net never builds
OFF_PATH_PREFIX = ("BLInputLayer_", "BLOutputLayer_")

# Metadata methods the on-path Python calls on the object returned by scn.Metadata(3)
METADATA_CALLERS = [
    ("SparseConvNet/sparseconvnet/sparseConvNetTensor.py", "getSpatialLocations"),
    ("maskrcnn_benchmark/modeling/rpn/anchor_generator_sparse3d.py", "get_spatial_locations"),
]


def _dotted(node):
    parts = []
    while isinstance(node, ast.Attribute):
        parts.append(node.attr)
        node = node.value
    if isinstance(node, ast.Name):
        parts.append(node.id)
        return ".".join(reversed(parts))
    return None


def scn_calls(path):
    tree = ast.parse(open(path).read(), filename=path)
    out = []
    for node in ast.walk(tree):
        if not isinstance(node, ast.Call):
            continue
        name = _dotted(node.func)
        if name and name.startswith("sparseconvnet.SCN."):
            out.append({
                "name": name.split(".", 2)[2],
                "nargs": len(node.args),
                "kwargs": sorted(k.arg for k in node.keywords if k.arg),
                "line": node.lineno,
            })
    return out


def method_calls(path, attr):
    tree = ast.parse(open(path).read(), filename=path)
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == attr:
            out.append({"name": attr, "nargs": len(node.args), "line": node.lineno})
    return out


# on-path module classes and the constructor calls fpn_net.py / tools_3d_2d.py / the RPN make
SURFACE_CASES = [
    ("InputLayer", (3, [64, 64, 32]), {"mode": 4}),
    ("OutputLayer", (3,), {}),
    ("SubmanifoldConvolution", (3, 9, 32, 3, False), {}),
    ("SubmanifoldConvolution", (3, 16, 16, [1, 1, 4], True), {}),
    ("Convolution", (3, 32, 64, 2, 2, False), {}),
    ("Convolution", (3, 16, 16, [1, 1, 8], [1, 1, 1], False), {}),
    ("Deconvolution", (3, 64, 32, 2, 2, False), {}),
    ("BatchNormalization", (16,), {}),
    ("BatchNormReLU", (16,), {}),
    ("BatchNormLeakyReLU", (16,), {"leakiness": 0}),
    ("BatchNormLeakyReLU", (32,), {"leakiness": 0.333, "momentum": 0.95, "track_running_stats": False}),
    ("NetworkInNetwork", (16, 32, False), {}),
    ("NetworkInNetwork", (16, 32, True), {}),
    ("SparseToDense", (3, 16), {}),
    ("Sequential", (), {}),
    ("ConcatTable", (), {}),
    ("AddTable", (), {}),
    ("JoinTable", (), {}),
    ("Identity", (), {}),
]


def module_surface():
    """constructor signatures, reprs, parameter / buffer names and shapes of the on-path module classes, taken
    from the reference's own Python package imported on oracle/_ref/SCN_ref.so (see make_golden.py)"""
    import inspect
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import make_golden
    ref = make_golden.reference_package()
    out = []
    for cls, args, kw in SURFACE_CASES:
        C = getattr(ref, cls)
        sig = inspect.signature(C.__init__)
        params = [[p.name, None if p.default is inspect.Parameter.empty else repr(p.default)]
                  for p in list(sig.parameters.values())[1:]]
        m = C(*args, **kw)
        out.append({
            "class": cls, "args": list(args), "kwargs": kw, "init_params": params, "repr": repr(m),
            "parameters": {k: list(v.shape) for k, v in m.named_parameters()},
            "buffers": {k: list(v.shape) for k, v in m.named_buffers()},
        })
    return out


def main():
    if not os.path.isdir(PKG):
        sys.exit("reference tree not found at " + REF)
    sites = []
    for f in sorted(os.listdir(PKG)):
        if not f.endswith(".py"):
            continue
        for c in scn_calls(os.path.join(PKG, f)):
            c["file"] = "SparseConvNet/sparseconvnet/" + f
            c["on_path"] = f in ON_PATH and not c["name"].startswith(OFF_PATH_PREFIX)
            sites.append(c)
    methods = []
    for rel, attr in METADATA_CALLERS:
        for c in method_calls(os.path.join(REF, rel), attr):
            c["file"] = rel
            methods.append(c)
    # `Metadata(dim)` resolves `Metadata_%d` by getattr (metadata.py:16-17)
    doc = {"generated_by": "oracle/make_callsites.py", "scn_calls": sites, "method_calls": methods,
           "metadata_factory": "Metadata_3", "module_surface": module_surface()}
    with open(OUT, "w") as fh:
        json.dump(doc, fh, indent=1, sort_keys=True)
        fh.write("\n")
    print("wrote %s: %d SCN call sites (%d on the path), %d method call sites"
          % (os.path.normpath(OUT), len(sites), sum(c["on_path"] for c in sites), len(methods)))


if __name__ == "__main__":
    main()
