"""bench.py contract pieces that can be checked without a GPU: the reference arm prints exactly
one JSON line on stdout with the agreed keys; our arm refuses to run without a device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
        "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


def test_reference_arm_prints_one_json_line(ref):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "2", "--warmup", "1", "--ref-nx", "24"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert KEYS <= set(d) and d["impl"] == "reference" and d["unit"] == "elements/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "elements/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert "flat plate 1000x1000" in d["config"]["workload"] and "24x24" in d["reference_run"]["sample"]
    # the reference arm is timed on OUR arm's configuration: the same `config` object
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    ours = bench.arm_config(argparse.Namespace(workload="plate", scatter="atomic"), 1, d["config"]["workload"],
                            1000 * 1000)
    assert d["config"] == ours
    assert d["steps"] == 2 and d["warmup"] == 1


def test_reference_arm_other_ranks_stay_silent(ref):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_our_arm_needs_a_device():
    from conftest import has_gpu
    if has_gpu():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and out.stdout.strip() == ""
    assert "no CUDA device" in out.stderr
