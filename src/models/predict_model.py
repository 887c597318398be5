"""Predict plume masks for large scenes with the B200 UNet (tiled, overlap-stitched).

Named in the reference's layout (README.md:44-47), never shipped there.  Scenes are
``<path_to_viirs_ml_reprojected_h5>/*.pt`` tensors [H, W, C]; with several GPUs they are dealt round-robin to
the ranks (no collective).  Masks (uint8, 1 = plume) are written next to the reference's mask folder
(``filepaths.path_to_viirs_ml_plume_masks``, filepaths.py:26) as ``<scene>.mask.pt``.
"""
import glob
import logging
import os
import time

import click
import torch

from kcl_ltss_bioatm_b200.data import synthetic_scene
from kcl_ltss_bioatm_b200.predict import ScenePredictor, shard_round_robin
from kcl_ltss_bioatm_b200.spec import UNetSpec
from kcl_ltss_bioatm_b200.trainer import LOG_FMT, init_distributed
from kcl_ltss_bioatm_b200.unet import UNetB200
from src.config import filepaths as fp

logging.basicConfig(level=logging.INFO, format=LOG_FMT)
logger = logging.getLogger(__name__)


@click.command()
@click.option("--name", default="unet_plume", show_default=True)
@click.option("--tile", default=256, show_default=True)
@click.option("--margin", default=16, show_default=True, help="overlap margin; stride = tile - 2*margin")
@click.option("--batch-tiles", default=64, show_default=True)
@click.option("--synthetic-scenes", default=0, show_default=True, help="use N synthetic 4096^2 scenes")
@click.option("--scene-size", default=4096, show_default=True)
def main(name, tile, margin, batch_tiles, synthetic_scenes, scene_size):
    rank, world, local, _ = init_distributed("cuda")
    device = torch.device("cuda", local)
    opt_path = os.path.join(fp.path_to_model_folder, f"{name}.opt.pt")
    spec = UNetSpec.from_dict(torch.load(opt_path, map_location="cpu")["spec"]) if os.path.exists(opt_path) \
        else UNetSpec()
    model = UNetB200(spec, device=device, seed=0)
    wpath = os.path.join(fp.path_to_model_folder, f"{name}.pt")
    if os.path.exists(wpath):
        model.load_state_dict(torch.load(wpath, map_location="cpu"))
        logger.info("loaded %s", wpath)
    else:
        logger.warning("no weights at %s: predicting with the initial weights", wpath)
    pred = ScenePredictor(model, tile=tile, margin=margin, batch_tiles=batch_tiles)
    files = sorted(glob.glob(os.path.join(fp.path_to_viirs_ml_reprojected_h5, "*.pt")))
    n_scenes = synthetic_scenes if synthetic_scenes else len(files)
    mine = shard_round_robin(n_scenes, rank, world)
    os.makedirs(fp.path_to_viirs_ml_plume_masks, exist_ok=True)
    t0, tiles = time.time(), 0
    for idx in mine:
        if synthetic_scenes:
            scene, stem = synthetic_scene(scene_size, scene_size, spec.in_channels, seed=idx), f"synthetic_{idx:04d}"
        else:
            scene, stem = torch.load(files[idx], map_location="cpu"), os.path.splitext(os.path.basename(files[idx]))[0]
        mask = pred.predict_scene(scene.to(torch.bfloat16).to(device))
        tiles += pred.num_tiles(scene.shape[0], scene.shape[1])
        torch.save(mask.cpu(), os.path.join(fp.path_to_viirs_ml_plume_masks, f"{stem}.mask.pt"))
        logger.info("rank %d scene %s: %.2f %% plume", rank, stem, 100.0 * mask.float().mean().item())
    torch.cuda.synchronize()
    logger.info("rank %d: %d scenes, %.0f tiles/s", rank, len(mine), tiles / max(time.time() - t0, 1e-9))


if __name__ == "__main__":
    main()
