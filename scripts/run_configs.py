"""Functional + timing pass over the other BASELINE.json configs on one GPU (per-GPU shares of the
multi-GPU configs): configs[2] 512^2 tiles, configs[3] tiled 4096^2 inference, configs[4] wide UNet 1024^2.
Prints one line per config.  Usage: python scripts/run_configs.py [c3] [c4] [c5]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kcl_ltss_bioatm_b200.data import synthetic_batch, synthetic_scene  # noqa: E402
from kcl_ltss_bioatm_b200.predict import ScenePredictor  # noqa: E402
from kcl_ltss_bioatm_b200.spec import UNetSpec, train_flops_per_tile, fwd_flops_per_tile  # noqa: E402
from kcl_ltss_bioatm_b200.trainer import Trainer  # noqa: E402
from kcl_ltss_bioatm_b200.unet import UNetB200  # noqa: E402

which = sys.argv[1:] or ["c3", "c4", "c5"]
dev = "cuda:0"


def timed_steps(tr, x, t, warm, steps):
    for _ in range(warm):
        tr.step(x, t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tr.step(x, t)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


if "c3" in which:
    spec = UNetSpec()
    tr = Trainer(spec, device=dev)
    # base tiles generated at 128^2 and tiled up: generation on CPU is slow for 32 x 512^2
    xs, ts = synthetic_batch(8, 512, 512, spec.in_channels, seed=3)
    x, t = xs.repeat(4, 1, 1, 1).to(dev), ts.repeat(4, 1, 1).to(dev)
    ms = timed_steps(tr, x, t, 2, 5)
    tf = train_flops_per_tile(spec, 512, 512) * 32 / (ms * 1e-3) / 1e12
    print(f"configs[2] per-GPU micro-step 32 x 512^2: {ms:.2f} ms/step, {32 / ms * 1e3:.0f} tiles/s/GPU, {tf:.0f} TFLOP/s, "
          f"loss {tr.model.loss_out[0].item():.4f}, workspace {tr.model.activation_bytes() / 1e9:.1f} GB", flush=True)
    del tr, x, t
    torch.cuda.empty_cache()

if "c4" in which:
    spec = UNetSpec()
    model = UNetB200(spec, device=dev, seed=0)
    pred = ScenePredictor(model, tile=256, margin=16, batch_tiles=64)
    scene = synthetic_scene(4096, 4096, spec.in_channels, seed=1).to(dev)
    pred.predict_scene(scene)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        mask = pred.predict_scene(scene)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    tiles = pred.num_tiles(4096, 4096)
    tf = fwd_flops_per_tile(spec, 256, 256)["total"] * tiles / dt / 1e12
    print(f"configs[3] tiled inference 4096^2: {tiles} tiles, {dt * 1e3:.1f} ms/scene, {tiles / dt:.0f} tiles/s/GPU, "
          f"{tf:.0f} TFLOP/s, plume fraction {mask.float().mean().item():.3f}", flush=True)
    del model, pred, scene
    torch.cuda.empty_cache()

if "c5" in which:
    spec = UNetSpec.wide()
    tr = Trainer(spec, device=dev)
    xs, ts = synthetic_batch(1, 1024, 1024, spec.in_channels, seed=5)
    x, t = xs.to(dev), ts.to(dev)
    ms = timed_steps(tr, x, t, 2, 5)
    tf = train_flops_per_tile(spec, 1024, 1024) / (ms * 1e-3) / 1e12
    print(f"configs[4] wide UNet (F=128, depth 5) 1 x 1024^2: {ms:.2f} ms/step, {1e3 / ms:.1f} tiles/s/GPU, {tf:.0f} TFLOP/s, "
          f"params {tr.model.num_parameters() / 1e6:.1f} M, loss {tr.model.loss_out[0].item():.4f}, "
          f"workspace {tr.model.activation_bytes() / 1e9:.1f} GB", flush=True)
