"""CPU: the reference arm of bench.py (the one leg allowed to execute oracle/) prints the contract's JSON line."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="8", DEMFI_REF_BUDGET_S="1")  # one whole-frame forward (a minute on 8 cores), no more
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "12", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "interpolated_frames_per_sec" and d["unit"] == "frames/s"
    # `steps` is what was actually run (whole frames, bounded by the CPU-time budget), never an extrapolation
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["steps_requested"] == 12 and d["value"] > 0
    assert abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-6
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "whole 736x1280 padded frame" in d["cpu_baseline"]["sample"] and "crop" not in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
