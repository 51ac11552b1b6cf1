"""CPU: the reference arm of bench.py honours the one-line JSON contract (the CUDA arm needs a
GPU; its line carries the same keys plus clocks / gpu_launches / roofline)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_contract_line():
    from tests.checker import have
    if not (have("ref_fast") or have("ref") or have("orc")):
        pytest.skip("no CPU checker library built")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-1500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                       # exactly ONE line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mpixels/sec SIFT+match+blend" and d["unit"] == "Mpx/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["config"]["workload"] == "ordered_13x1500x1112" and d["higher_is_better"] is True and d["steps"] == 1
    assert d["value"] > 0 and abs(d["value"] - 21.684 / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"] + 1e-3
    cb, e2e = d["cpu_baseline"], d["e2e"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_other_ranks_of_the_reference_arm_stay_silent():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
