"""bench.py's CPU-side contract: the reference arm's JSON line (what it measured vs what it extrapolates) and the argument surface.
The GPU arms are exercised by the driver; here only what runs without a device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_reports_measured_and_extrapolated_separately():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--skip-a0",
                        "--cpu-rows", "128"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "volumes/s" and line["higher_is_better"] is True
    assert line["extrapolated"] is True and 0 < line["measured_fraction"] < 1
    # ms_per_step is what one timed step of this run took (the bounded sample), NOT the per-volume estimate
    assert line["ms_per_step"] < line["estimated_ms_per_volume"]
    assert abs(line["value"] - 1e3 / line["estimated_ms_per_volume"]) < 1e-9
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["extrapolated"] is True and "train mode" in cb["sample"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    assert line["dropout_off"]["extrapolated"] is True and line["dropout_off"]["value"] > line["value"]      # bernoulli_ dominates the train-mode step


def test_b200_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
