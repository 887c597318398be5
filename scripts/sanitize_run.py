"""One small launch of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):

    compute-sanitizer --tool memcheck  python scripts/sanitize_run.py
    compute-sanitizer --tool racecheck python scripts/sanitize_run.py

Shapes are tiny (the tools slow kernels down 10-100x) but chosen so that every kernel template that the training
step of the default network uses is instantiated: igemm_conv3_kernel <64,0> <64,1> <128,1> <256,2>, the generic
igemm_fwd_kernel (transposed conv, 8x8 images), igemm_wgrad3_kernel <64> <128>, the generic igemm_wgrad_kernel, all
bandwidth kernels, tile cut / stitch, and the label-generation kernels.  Prints one line per stage so that a
sanitizer report can be attributed."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kcl_ltss_bioatm_b200.data import synthetic_batch, synthetic_scene  # noqa: E402
from kcl_ltss_bioatm_b200.predict import ScenePredictor  # noqa: E402
from kcl_ltss_bioatm_b200.spec import UNetSpec  # noqa: E402
from kcl_ltss_bioatm_b200.trainer import Trainer  # noqa: E402

DEV = "cuda:0"


def stage(name):
    torch.cuda.synchronize()
    print(f"[sanitize] {name}: ok", flush=True)


def main():
    which = sys.argv[1:] or ["train", "predict", "labels"]
    spec = UNetSpec(base_filters=64, depth=2)
    if "train" in which:
        tr = Trainer(spec, device=DEV, seed=0)
        x, t = synthetic_batch(2, 32, 32, spec.in_channels, seed=1)
        for _ in range(2):
            tr.step(x.to(DEV), t.to(DEV))
        stage("two eager training steps, depth 2, 2 x 32 x 32 (conv3 <64,0> <64,1> <128,1> <256,2>, generic fwd, "
              "wgrad3 <64> <128>, generic wgrad, BatchNorm / pool / head / Adam / pack kernels)")
        tr2 = Trainer(spec, device=DEV, seed=0, micro_batches=2)
        tr2.step(x.to(DEV), t.to(DEV))
        stage("micro-batched step (gradient accumulation)")
        x2, t2 = synthetic_batch(1, 24, 40, spec.in_channels, seed=2)
        tr.step(x2.to(DEV), t2.to(DEV))
        stage("ragged tile 24 x 40 (partial GEMM tiles, TMA out-of-bounds fill)")
        tr.step_graphed(x.to(DEV), t.to(DEV))
        tr.step_graphed(x.to(DEV), t.to(DEV))
        stage("CUDA-graph step: capture + two replays")
        tr.release_graphs()
    if "predict" in which:
        from kcl_ltss_bioatm_b200.unet import UNetB200

        net = UNetB200(spec, device=DEV, seed=0)
        pred = ScenePredictor(net, tile=32, margin=4, batch_tiles=5)
        mask, prob = pred.predict_scene(synthetic_scene(70, 90, spec.in_channels, seed=3).to(DEV), want_prob=True)
        assert mask.shape == (70, 90)
        stage("tiled scene inference 70 x 90 (extract_tiles, eval forward with folded BatchNorm, stitch_threshold)")
    if "labels" in which:
        from kcl_ltss_bioatm_b200.fires import FireLocator
        from kcl_ltss_bioatm_b200.labels import LabelRasterizer
        from kcl_ltss_bioatm_b200.ops import CudaOps
        from kcl_ltss_bioatm_b200.sweep import ThresholdSweep

        ops = CudaOps()
        hulls = [(np.array([5.0, 40, 52, 30, 9]), np.array([6.0, 3, 30, 55, 41])),
                 (np.array([60.0, 95, 80]), np.array([70.0, 75, 99]))]
        m = LabelRasterizer(DEV, ops=ops).scene_mask(hulls, 100, 130)
        assert int(m.sum()) > 0
        stage("hull rasteriser")
        rng = np.random.default_rng(0)
        lat, lon = np.meshgrid(np.linspace(10, 11, 60), np.linspace(20, 21, 70), indexing="ij")
        FireLocator(lat, lon, device=DEV, ops=ops).nearest_pixels(10 + rng.random(40), 20 + rng.random(40))
        stage("fire locator")
        aod = rng.random((90, 110)).astype("float32")
        ThresholdSweep(DEV, ops=ops).extents(aod, np.arange(0.02, 0.5, 0.02), [20, 50, 70], [20, 60, 90])
        stage("threshold sweep (masks, connected components, fire extents)")
    print("[sanitize] done", flush=True)


if __name__ == "__main__":
    main()
