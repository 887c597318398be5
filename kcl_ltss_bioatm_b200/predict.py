"""Tiled large-scene inference: cut overlapping tiles, forward, stitch by centre crop, threshold.

A scene [Hs, Ws, C] is covered by T x T tiles at stride T - 2*margin.  Each tile owns its interior (the
pixels at least `margin` from its border) plus any border strip that touches the scene edge, so every
scene pixel is owned by exactly one tile and the stitched mask does not depend on tile order.  Scenes are
independent: with several GPUs they are dealt round-robin to the ranks and no collective is needed
(BASELINE.json config 4).

The reference has no predict path (README.md:44-47 names ``predict_model.py``; it was never committed);
output format defined here: uint8 mask [Hs, Ws], 1 = plume, i.e. sigmoid(logit) >= spec.mask_threshold.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch

from .unet import UNetB200


def tile_origins(size: int, tile: int, margin: int) -> List[int]:
    """Origins k*(tile - 2*margin), enough of them that the last tile reaches the scene edge."""
    stride = tile - 2 * margin
    if stride <= 0:
        raise ValueError("margin too large for the tile size")
    n = max(1, math.ceil((size - 2 * margin) / stride))
    return [k * stride for k in range(n)]


def tile_grid(hs: int, ws: int, tile: int, margin: int) -> Tuple[List[int], List[int]]:
    ys, xs = tile_origins(hs, tile, margin), tile_origins(ws, tile, margin)
    return [y for y in ys for _ in xs], [x for _ in ys for x in xs]


def shard_round_robin(n_items: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_items, world))


class ScenePredictor:
    def __init__(self, model: UNetB200, tile: int = 256, margin: int = 16, batch_tiles: int = 64):
        if tile % model.spec.divisor():
            raise ValueError(f"tile must be a multiple of {model.spec.divisor()}")
        self.model, self.tile, self.margin, self.batch_tiles = model, tile, margin, batch_tiles
        thr = model.spec.mask_threshold
        self.logit_threshold = math.log(thr / (1.0 - thr))
        self._tiles: Optional[torch.Tensor] = None

    def num_tiles(self, hs: int, ws: int) -> int:
        return len(tile_origins(hs, self.tile, self.margin)) * len(tile_origins(ws, self.tile, self.margin))

    def predict_scene(self, scene: torch.Tensor, want_prob: bool = False):
        """scene: [Hs, Ws, C_in] in the model's activation dtype on the model's device.
        Returns the uint8 mask [Hs, Ws] (and the fp32 probability map if want_prob)."""
        m, ops, T = self.model, self.model.ops, self.tile
        dev = m.device
        hs, ws, cs = scene.shape
        if cs != m.spec.in_channels:
            raise ValueError(f"scene has {cs} channels, model expects {m.spec.in_channels}")
        ys_l, xs_l = tile_grid(hs, ws, T, self.margin)
        ys = torch.tensor(ys_l, dtype=torch.int32, device=dev)
        xs = torch.tensor(xs_l, dtype=torch.int32, device=dev)
        mask = torch.empty(hs, ws, dtype=torch.uint8, device=dev)
        prob = torch.empty(hs, ws, dtype=torch.float32, device=dev) if want_prob else None
        bt = min(self.batch_tiles, len(ys_l))
        cd = m.spec.cin_padded  # tiles are cut already zero padded to the first layer's K
        if self._tiles is None or self._tiles.shape[0] != bt or self._tiles.shape[1] != T:
            shape = (bt, T, T, cd) if getattr(m, "planes", 1) == 1 else (bt, T, T, 2, cd)
            self._tiles = torch.empty(*shape, dtype=m.act_dtype, device=dev)
        was = m.training
        m.eval()
        for b0 in range(0, len(ys_l), bt):
            b1 = min(b0 + bt, len(ys_l))
            tiles = self._tiles[: b1 - b0]
            ops.extract_tiles(scene, ys[b0:b1], xs[b0:b1], T, tiles)
            logits = m.forward(tiles)
            ops.stitch_threshold(logits, ys[b0:b1], xs[b0:b1], T, self.margin, self.logit_threshold, mask, prob)
        m.train(was)
        return (mask, prob) if want_prob else mask
