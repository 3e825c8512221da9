"""bench.py's reference arm prints the contract's JSON line (CPU only: a 20k-point building through the compiled
reference); under a 2-rank launch only rank 0 prints.  The GPU arm's line has the same base keys (checked on the box
by the driver); here the static parts of it - argument defaults, the workload description - are checked too."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

import ref_backbone as RB

pytestmark = pytest.mark.skipif(not RB.available(), reason="oracle/_ref not built")

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


def _lines(cmd):
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]


def _check(d, n_gpus):
    assert BASE_KEYS <= set(d), sorted(BASE_KEYS - set(d))
    assert d["impl"] == "reference" and d["metric"] == "active_voxels_per_sec" and d["unit"] == "active voxels/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["n_gpus"] == n_gpus and d["steps"] == 1 and d["warmup"] == 0 and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] == os.cpu_count() and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_prints_the_contract_line():
    lines = _lines([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--points", "20000"])
    assert len(lines) == 1
    _check(lines[0], 1)


def test_reference_arm_two_ranks_print_one_line():
    lines = _lines([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                    "--master-addr", "127.0.0.1", "--master-port", "29541", "bench.py", "--impl", "reference",
                    "--gpus", "2", "--steps", "1", "--warmup", "0", "--points", "20000"])
    assert len(lines) == 1
    _check(lines[0], 2)


def test_defaults_are_the_baseline_workload():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.FULL_SCALE == [4096, 4096, 512] and bench.PLANES == [32, 64, 64, 128, 128, 128, 256, 256, 256]
    # the synthetic building is the generator of SURVEY appendix D.3: sizes the survey measured
    locs, feats = bench.make_batch(300000, 1, 1, 0)
    assert tuple(locs.shape) == (300000, 4) and tuple(feats.shape) == (300000, 9)
    assert bench.n_active0(locs) == 278639
