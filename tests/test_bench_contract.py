"""bench.py contract checks that need no GPU: the reference (CPU) arm prints exactly one JSON line with the agreed keys, and the
GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    # the arm proves real 2^20-row segments (~1 min each on 8 cores); the contract check runs it on 2^14 rows through the test hook,
    # which the line itself flags as not valid for the headline
    env = dict(os.environ, B200_BENCH_REF_PO2="14")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "segments_per_sec" and d["unit"] == "segments/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("synthetic 1M-cycle segment")
    assert d["config"]["sample_po2"] == 14 and "invalid_for_headline" in d and d["config"]["verifies"] is True


def test_gpu_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
