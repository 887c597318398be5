"""Train the smoke-plume UNet on the B200 path.

The reference names this script in its layout (README.md:44-47) but never shipped it; conventions kept
from the reference's scripts: ``main()`` guarded by ``__main__`` with module-level stdlib logging
(src/features/plume_identifier_rg.py:23-25, 514, 602-603), data / model locations from
``src.config.filepaths`` (filepaths.py:32-33), click for the CLI (requirements.txt:5).

    python -m src.models.train_model --steps 200 --batch 32 --tile 256
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m src.models.train_model ...

Inputs: ``<path_to_model_data_folder>/*.pt`` files holding {"x": [n,h,w,c] bf16/float, "mask": [n,h,w] uint8};
when the folder has none (the reference ships no data) synthetic MODIS-like tiles are generated.
Output: ``<path_to_model_folder>/<name>.pt`` (state_dict, oracle key layout) + ``<name>.opt.pt``.
"""
import glob
import logging
import os
import time

import click
import torch

from kcl_ltss_bioatm_b200.data import synthetic_batch
from kcl_ltss_bioatm_b200.spec import UNetSpec
from kcl_ltss_bioatm_b200.trainer import LOG_FMT, Trainer, init_distributed
from src.config import filepaths as fp

logging.basicConfig(level=logging.INFO, format=LOG_FMT)
logger = logging.getLogger(__name__)


def _file_batches(folder, per_rank, rank, world, device, spec):
    files = sorted(glob.glob(os.path.join(folder, "*.pt")))
    if not files:
        return None
    logger.info("training on %d tile files under %s", len(files), folder)

    def batch_fn(it):
        blob = torch.load(files[(it * world + rank) % len(files)], map_location="cpu")
        x, m = blob["x"][:per_rank], blob["mask"][:per_rank]
        if x.shape[-1] != spec.in_channels:
            raise ValueError(f"{files[0]}: expected {spec.in_channels} bands, got {x.shape[-1]}")
        return x.to(torch.bfloat16).to(device, non_blocking=True), m.to(torch.uint8).to(device, non_blocking=True)

    return batch_fn


@click.command()
@click.option("--steps", default=200, show_default=True)
@click.option("--batch", default=32, show_default=True, help="tiles per GPU per step")
@click.option("--tile", default=256, show_default=True)
@click.option("--micro-batches", default=1, show_default=True)
@click.option("--base-filters", default=64, show_default=True)
@click.option("--depth", default=4, show_default=True)
@click.option("--in-channels", default=8, show_default=True)
@click.option("--norm", default="batch", type=click.Choice(["batch", "none"]), show_default=True)
@click.option("--lr", default=1e-3, show_default=True)
@click.option("--name", default="unet_plume", show_default=True)
@click.option("--resume/--no-resume", default=False)
@click.option("--seed", default=0, show_default=True)
@click.option("--log-every", default=10, show_default=True)
def main(steps, batch, tile, micro_batches, base_filters, depth, in_channels, norm, lr, name, resume, seed,
         log_every):
    rank, world, local, pg = init_distributed("cuda")
    device = torch.device("cuda", local)
    spec = UNetSpec(in_channels=in_channels, base_filters=base_filters, depth=depth, norm=norm, lr=lr)
    trainer = Trainer(spec, device=device, process_group=pg, seed=seed, micro_batches=micro_batches)
    if resume:
        trainer.load_checkpoint(fp.path_to_model_folder, name)
        logger.info("resumed %s at step %d", name, trainer.model.step_count)
    batch_fn = _file_batches(fp.path_to_model_data_folder, batch, rank, world, device, spec)
    if batch_fn is None:
        logger.info("no tile files under %s: using synthetic tiles", fp.path_to_model_data_folder)
        pool = [synthetic_batch(batch, tile, tile, in_channels, seed=1234 + 97 * rank + i, device=device)
                for i in range(4)]
        batch_fn = lambda it: pool[it % len(pool)]  # noqa: E731
    t0 = time.time()
    trainer.fit(steps, batch_fn, log_every=log_every if rank == 0 else 0)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        logger.info("%d steps, %.1f tiles/s over %d GPU(s)", steps, steps * batch * world / dt, world)
        path = trainer.save_checkpoint(fp.path_to_model_folder, name)
        logger.info("saved %s", path)


if __name__ == "__main__":
    main()
