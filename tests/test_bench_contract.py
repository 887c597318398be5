"""The bench.py contract the driver depends on: exactly one JSON line on stdout with the required keys, for the
reference arm (CPU, runs here) and for the GPU arm (gpu-marked)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def run(*flags, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                      # ONE JSON line, everything else goes to stderr
    return json.loads(lines[0])


@pytest.mark.timeout(900)
def test_reference_arm_prints_one_contract_line():
    d = run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["metric"] == "unet_train_tiles_per_sec_256px" and d["unit"] == "tiles/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_non_zero_rank_prints_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_gpu_arm_prints_one_contract_line():
    d = run("--steps", "3", "--warmup", "3", "--no-cpu-baseline")
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d)
    assert d["metric"] == "unet_train_tiles_per_sec_256px" and d["n_gpus"] == 1 and d["steps"] == 3
    assert d["dtype"] == "bf16" and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["gpu_launches"] > 100 * d["steps"]          # our kernels, counted per step (graph replay included)
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0 < r["frac"] < 1.2 and r["achieved"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["unit"] == "tiles/s" and e["h2d_bytes_per_step"] == 32 * 256 * 256 * (8 * 2 + 1) and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] <= d["value"] * 1.05
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
