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
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.headline_config(1)      # the reference arm times the GPU arm's own config
    assert d["steps"] == 1 and d["warmup"] == 0         # K and W as passed


def test_timer_accounts_every_algorithmic_flop_of_a_step():
    """bench.OpTimer over the CPU operator oracle: the FLOPs it attributes to the tensor-core launches of one
    training step equal train_flops_per_tile minus the 1x1 head (x3 passes), with the first layer counted at its 8
    logical input channels -- and only the first layer (every other 64-channel input is real)."""
    import torch

    sys.path.insert(0, ROOT)
    import bench
    from kcl_ltss_bioatm_b200.spec import UNetSpec, fwd_flops_per_tile, train_flops_per_tile
    from kcl_ltss_bioatm_b200.unet import UNetB200
    from oracle.ops_ref import RefOps

    spec = UNetSpec(base_filters=64, depth=2)
    net = UNetB200(spec, ops=RefOps(), device="cpu", seed=0)
    timer = bench.OpTimer(net, events=False)
    n, h, w = 2, 16, 16
    x = torch.randn(n, h, w, spec.in_channels).to(torch.bfloat16)
    t = (torch.rand(n, h, w) > 0.8).to(torch.uint8)
    timer.enabled = True
    net.train_step(x, t)
    timer.enabled = False
    expect = (train_flops_per_tile(spec, h, w) - 3 * fwd_flops_per_tile(spec, h, w)["head"]) * n
    assert abs(timer.total_flops() - expect) <= 1e-9 * expect
    rows, summary = timer.per_layer(1, 1369.9, 1624.1)
    assert len(rows) == 3 * (len(net.convs) + len(net.ups)) - 1          # no dgrad for the first layer
    assert {r["layer"] for r in rows} == set(net.convs) | set(net.ups)
    first = [r for r in rows if r["layer"] == "enc0.conv1" and r["pass"] == "fwd"][0]
    second = [r for r in rows if r["layer"] == "enc0.conv2" and r["pass"] == "fwd"][0]
    assert abs(second["gflop"] / first["gflop"] - 64 / 8) < 1e-9
    hbm = timer.hbm_by_kernel(1, 6541.1)
    assert {"scale_shift_act", "scale_shift_act_pool", "maxpool_bwd", "bn_bwd_reduce", "bn_bwd_apply", "head_fwd",
            "head_bwd", "adam", "pad_channels", "pack_batch", "channel_sum"} <= set(hbm)
    assert abs(hbm["adam"]["algorithmic_mb_per_step"] * 1e6 - 28.0 * net.layout.total) < 1


def test_reference_arm_non_zero_rank_prints_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_gpu_arm_prints_one_contract_line():
    d = run("--steps", "3", "--warmup", "3", "--no-cpu-baseline")
    assert BASE_KEYS | {"roofline", "clocks", "per_layer", "per_layer_summary", "other_configs"} <= set(d)
    assert abs(d["per_layer_summary"]["flops_accounted"] - 1.0) < 1e-6
    assert abs(d["roofline"]["algorithmic_tflop_per_step"] + d["roofline_wgrad"]["algorithmic_tflop_per_step"]
               - d["per_layer_summary"]["expected_gemm_flops_per_step"] / 1e12) < 1e-6
    assert len(d["per_layer"]) == 65 and {"adam", "head_fwd", "maxpool_bwd"} <= set(d["roofline_hbm"]["by_kernel"])
    for key in ("configs[2]", "configs[3]", "configs[4]"):
        assert "error" not in d["other_configs"][key], d["other_configs"][key]
        assert d["other_configs"][key]["tiles_per_s"] > 0
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
