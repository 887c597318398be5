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


class TileFiles:
    """Index of every tile in ``<folder>/*.pt`` ({"x": [n,h,w,c], "mask": [n,h,w]} as written by
    ``labels.write_training_tiles``).  Tiles are numbered globally across files; batch `it` of rank `r` is the
    `per_rank` consecutive tiles starting at ``(it * world + r) * per_rank`` (wrapping around), so every tile of
    every file is visited and every rank gets the same count.  Shapes are validated once, up front."""

    def __init__(self, files, spec):
        self.files, self.counts = files, []
        self.hw = None
        div = spec.divisor()
        for f in files:
            blob = torch.load(f, map_location="cpu")
            x, m = blob["x"], blob["mask"]
            if x.dim() != 4 or m.dim() != 3 or x.shape[0] != m.shape[0] or tuple(x.shape[1:3]) != tuple(m.shape[1:]):
                raise ValueError(f"{f}: expected x [n,h,w,c] and mask [n,h,w], got {tuple(x.shape)} / {tuple(m.shape)}")
            if x.shape[-1] != spec.in_channels:
                raise ValueError(f"{f}: expected {spec.in_channels} bands, got {x.shape[-1]}")
            if x.shape[1] % div or x.shape[2] % div:
                raise ValueError(f"{f}: tile size {x.shape[1]}x{x.shape[2]} is not a multiple of {div}")
            if self.hw is None:
                self.hw = tuple(x.shape[1:3])
            elif tuple(x.shape[1:3]) != self.hw:
                raise ValueError(f"{f}: tile size {tuple(x.shape[1:3])} differs from {self.hw} in {files[0]}")
            self.counts.append(int(x.shape[0]))
        self.total = sum(self.counts)
        if self.total == 0:
            raise ValueError(f"no tiles in {len(files)} files")
        self._cache = (None, None)

    def locate(self, g):
        """Global tile index -> (file index, index inside the file)."""
        g %= self.total
        for fi, c in enumerate(self.counts):
            if g < c:
                return fi, g
            g -= c
        raise AssertionError

    def _blob(self, fi):
        if self._cache[0] != fi:
            self._cache = (fi, torch.load(self.files[fi], map_location="cpu"))
        return self._cache[1]

    def batch(self, it, per_rank, rank, world):
        start = (it * world + rank) * per_rank
        xs, ms = [], []
        for g in range(start, start + per_rank):
            fi, k = self.locate(g)
            blob = self._blob(fi)
            xs.append(blob["x"][k])
            ms.append(blob["mask"][k])
        return torch.stack(xs).to(torch.bfloat16), torch.stack(ms).to(torch.uint8)


def _file_batches(folder, per_rank, rank, world, device, spec):
    files = sorted(glob.glob(os.path.join(folder, "*.pt")))
    if not files:
        return None
    index = TileFiles(files, spec)
    logger.info("training on %d tiles in %d files under %s", index.total, len(files), folder)

    def batch_fn(it):
        x, m = index.batch(it, per_rank, rank, world)
        return x.to(device, non_blocking=True), m.to(device, non_blocking=True)

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
@click.option("--precision", default="bf16", type=click.Choice(["bf16", "bf16x3", "tf32"]), show_default=True,
              help="bf16x3 (alias tf32): hi+lo bf16 storage, three tensor-core passes, logits within 1e-3 of fp32")
@click.option("--graph/--no-graph", default=True, show_default=True, help="replay the step from a CUDA graph")
@click.option("--deterministic/--no-deterministic", default=False, show_default=True,
              help="ordered reductions instead of floating-point atomics: bit-identical runs")
def main(steps, batch, tile, micro_batches, base_filters, depth, in_channels, norm, lr, name, resume, seed,
         log_every, precision, graph, deterministic):
    rank, world, local, pg = init_distributed("cuda")
    device = torch.device("cuda", local)
    spec = UNetSpec(in_channels=in_channels, base_filters=base_filters, depth=depth, norm=norm, lr=lr,
                    precision=precision)
    trainer = Trainer(spec, device=device, process_group=pg, seed=seed, micro_batches=micro_batches)
    trainer.model.ops.set_deterministic(deterministic)
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
    trainer.fit(steps, batch_fn, log_every=log_every if rank == 0 else 0, graphed=graph)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        logger.info("%d steps, %.1f tiles/s over %d GPU(s)", steps, steps * batch * world / dt, world)
        path = trainer.save_checkpoint(fp.path_to_model_folder, name)
        logger.info("saved %s", path)


if __name__ == "__main__":
    main()
