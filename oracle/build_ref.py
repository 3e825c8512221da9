"""Build the reference's own CPU SparseConvNet code into oracle/_ref/ (TEST INFRASTRUCTURE).

Two extension modules, both compiled from the reference sources WHERE THEY LIE under
/root/reference (nothing is copied into this repo), with /usr/bin/g++ and the sparsehash shim
in oracle/shim/ (SURVEY.md Appendix D.1/D.2):

  oracle/_ref/SCN_ref.so      pybind.cpp + sparseconvnet_cpu.cpp, unmodified: the reference's
                              `sparseconvnet.SCN` extension (CPU ops + Metadata_3).
  oracle/_ref/SCN_refdump.so  oracle/ref_dump.cpp, which #includes the reference Metadata.cpp and
                              exports rulebooks as tensors.

oracle/_ref/ is git-ignored but travels to the GPU box with gpurun.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the results.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SCN = "/root/reference/SparseConvNet/sparseconvnet/SCN/"
OUT = os.path.join(HERE, "_ref")


def _load(name, sources, verbose):
    target = os.path.join(OUT, name + ".so")
    if os.path.exists(target):
        return target
    os.environ["CXX"] = "/usr/bin/g++"
    os.environ["CC"] = "/usr/bin/gcc"
    from torch.utils.cpp_extension import load
    bdir = os.path.join(OUT, "build_" + name)
    os.makedirs(bdir, exist_ok=True)
    load(name=name, sources=sources,
         extra_include_paths=[os.path.join(HERE, "shim"), REF_SCN],
         extra_cflags=["-std=c++17", "-fopenmp", "-O2", "-w"],
         extra_ldflags=["-fopenmp"], build_directory=bdir, verbose=verbose)
    shutil.copy(os.path.join(bdir, name + ".so"), target)
    shutil.rmtree(bdir, ignore_errors=True)
    return target


def build(verbose=False):
    """Returns the list of built .so paths, or None when /root/reference is absent (GPU box:
    the prebuilt files that travelled with the snapshot are used)."""
    if not os.path.isdir(REF_SCN):
        return None
    return [
        _load("SCN_ref", [REF_SCN + "pybind.cpp", REF_SCN + "sparseconvnet_cpu.cpp"], verbose),
        _load("SCN_refdump", [os.path.join(HERE, "ref_dump.cpp")], verbose),
    ]


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
