"""Label geometry on the GPU: the reference's plume-hull tables -> masks -> training tiles.

The reference writes one CSV per MAIAC file with the convex hull of every accepted plume
(``id, hull_lats, hull_lons, hull_x, hull_y, datetime``; plume_identifier_gaussian_profile.py:283-301, 639-644) and
its curation script turns a hull back into pixels with a Delaunay point-in-hull test over the whole pixel grid
(plume_selector.py:88-116).  This module keeps those function names and argument meanings

    remove_duplicated_plumes(plume_df)            plume_selector.py:26-49
    subset_plume(aod, plume_df)                   plume_selector.py:53-85
    in_hull(p, hull)                              plume_selector.py:88-98
    find_plume_aod(plume_image, hull_x, hull_y)   plume_selector.py:101-116

and adds what the UNet needs from them: ``LabelRasterizer`` (all hulls of a scene -> one uint8 mask, or directly
-> the masks of a set of tiles) and ``build_training_tiles`` (scene + hulls -> {"x", "mask"} tiles in the format
``src/models/train_model.py`` reads).  The per-pixel decision runs in ``plume_rasterize_hulls`` (exact int64 edge
functions); the host only orders each hull's vertices.  There is no CPU path: without the CUDA library these
functions raise.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .ops import CudaOps

Hull = Tuple[np.ndarray, np.ndarray]  # (hull_x, hull_y)


# ---------------------------------------------------------------------------------------------- host geometry
def convex_polygon(hull_x, hull_y) -> np.ndarray:
    """int32 [m, 2] (x, y): the convex hull of the given integer vertices in counter-clockwise order (Andrew's
    monotone chain; collinear points dropped).  The reference hands the bare point set to Delaunay
    (plume_selector.py:96-97), so the order in the CSV carries no meaning."""
    x = np.asarray(hull_x, dtype=np.float64)
    y = np.asarray(hull_y, dtype=np.float64)
    if np.isnan(x).any() or np.isnan(y).any():
        raise ValueError("hull holds NaN coordinates")
    if (x != np.round(x)).any() or (y != np.round(y)).any():
        raise ValueError("hull vertices must be integer pixel coordinates")
    pts = np.unique(np.stack([x, y], 1).astype(np.int64), axis=0)
    if len(pts) < 3:
        raise ValueError("degenerate hull: fewer than three distinct vertices")
    chain: List[List[Tuple[int, int]]] = []
    for seq in (pts, pts[::-1]):
        out: List[Tuple[int, int]] = []
        for px, py in seq.tolist():
            while len(out) >= 2 and ((out[-1][0] - out[-2][0]) * (py - out[-2][1])
                                     - (out[-1][1] - out[-2][1]) * (px - out[-2][0])) <= 0:
                out.pop()
            out.append((px, py))
        chain.append(out[:-1])
    poly = np.array(chain[0] + chain[1], dtype=np.int64)
    if len(poly) < 3:
        raise ValueError("degenerate hull: all vertices are collinear")
    if np.abs(poly).max() >= 2 ** 30:
        raise ValueError("hull coordinates exceed the int32 pixel range")
    return poly.astype(np.int32)


def pack_polygons(hulls: Iterable[Hull], device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(verts int32 [nv, 2], offsets int32 [n+1], bbox int32 [n, 4] = xmin, ymin, xmax, ymax) on `device`."""
    polys = [convex_polygon(hx, hy) for hx, hy in hulls]
    offs = np.zeros(len(polys) + 1, dtype=np.int32)
    if polys:
        offs[1:] = np.cumsum([len(p) for p in polys])
        verts = np.concatenate(polys, 0)
        bbox = np.array([[p[:, 0].min(), p[:, 1].min(), p[:, 0].max(), p[:, 1].max()] for p in polys], dtype=np.int32)
    else:
        verts, bbox = np.zeros((0, 2), dtype=np.int32), np.zeros((0, 4), dtype=np.int32)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
    return to(verts), to(offs), to(bbox)


class LabelRasterizer:
    """Hull lists -> uint8 masks through ``plume_rasterize_hulls``."""

    def __init__(self, device="cuda", ops: Optional[CudaOps] = None):
        self.device = torch.device(device)
        self.ops = ops if ops is not None else CudaOps()

    def scene_mask(self, hulls: Sequence[Hull], height: int, width: int) -> torch.Tensor:
        """[height, width] uint8 on the device: 1 inside (or on the boundary of) any hull."""
        verts, offs, bbox = pack_polygons(hulls, self.device)
        mask = torch.empty(height, width, dtype=torch.uint8, device=self.device)
        self.ops.rasterize_hulls(verts, offs, bbox, mask)
        return mask

    def tile_masks(self, hulls: Sequence[Hull], ys, xs, tile: int) -> torch.Tensor:
        """[n, tile, tile] uint8: the masks of the tiles whose top-left corners are (ys[k], xs[k]); pixels past the
        scene edge are simply outside every hull."""
        verts, offs, bbox = pack_polygons(hulls, self.device)
        ys_t = torch.as_tensor(ys, dtype=torch.int32, device=self.device)
        xs_t = torch.as_tensor(xs, dtype=torch.int32, device=self.device)
        out = torch.empty(ys_t.numel(), tile, tile, dtype=torch.uint8, device=self.device)
        self.ops.rasterize_hulls(verts, offs, bbox, out, ys_t, xs_t)
        return out


_default: Optional[LabelRasterizer] = None


def _rasterizer() -> LabelRasterizer:
    global _default
    if _default is None:
        _default = LabelRasterizer()
    return _default


# ---------------------------------------------------------------------------------------------- reference-named API
def in_hull(p, hull) -> np.ndarray:
    """Test if the integer points `p` ([N, 2] as (x, y)) are inside the convex hull of `hull` ([M, 2]); boundary
    points count as inside, as with the reference's ``Delaunay(hull).find_simplex(p) >= 0``.  Evaluated on the GPU
    by rasterising the points' bounding window and gathering."""
    p = np.asarray(p)
    if p.ndim != 2 or p.shape[1] != 2:
        raise ValueError("p must be [N, 2]")
    if (p != np.round(p)).any():
        raise ValueError("in_hull evaluates pixel (integer) coordinates")
    hull = np.asarray(hull)
    pi = p.astype(np.int64)
    x0, y0 = int(pi[:, 0].min()), int(pi[:, 1].min())
    w, h = int(pi[:, 0].max()) - x0 + 1, int(pi[:, 1].max()) - y0 + 1
    r = _rasterizer()
    m = r.tile_masks([(hull[:, 0], hull[:, 1])], [y0], [x0], max(h, w))[0].cpu().numpy()
    return m[pi[:, 1] - y0, pi[:, 0] - x0].astype(bool)


def find_plume_aod(plume_image, hull_x, hull_y):
    """The values of `plume_image` ([H, W], numpy or torch) at the pixels inside the hull, row-major order."""
    img = torch.as_tensor(plume_image)
    h, w = img.shape
    m = _rasterizer().scene_mask([(hull_x, hull_y)], h, w).bool()
    vals = img.to(m.device)[m]
    return vals.cpu().numpy() if isinstance(plume_image, np.ndarray) else vals


def subset_plume(aod, plume_df, buffer: int = 40):
    """Crop `aod` to the hull's bounding box grown by `buffer` pixels (clipped at the image edges) and express
    the hull in the crop's coordinates.  Returns (crop, hull_x, hull_y) or (None, None, None) for NaN hulls."""
    hull_x = np.asarray(plume_df["hull_x"], dtype=np.float64)
    hull_y = np.asarray(plume_df["hull_y"], dtype=np.float64)
    min_x, max_x, min_y, max_y = hull_x.min(), hull_x.max(), hull_y.min(), hull_y.max()
    if np.isnan([min_x, max_x, min_y, max_y]).any():
        return None, None, None
    if min_x - buffer < 0:
        min_x = 0
    else:
        hull_x, min_x = hull_x - min_x + buffer, min_x - buffer
    if min_y - buffer < 0:
        min_y = 0
    else:
        hull_y, min_y = hull_y - min_y + buffer, min_y - buffer
    max_x = aod.shape[1] if max_x + buffer > aod.shape[1] else max_x + buffer
    max_y = aod.shape[0] if max_y + buffer > aod.shape[0] else max_y + buffer
    return aod[int(min_y):int(max_y), int(min_x):int(max_x)], hull_x, hull_y


def remove_duplicated_plumes(plume_df):
    """Drop every plume (all rows of an (id, datetime) pair) whose centroid -- mean hull latitude / longitude
    rounded to 3 decimals -- repeats that of an earlier plume of the same datetime, earlier in (id, datetime)
    order.  Takes and returns a pandas DataFrame with the reference's columns."""
    ids = plume_df["id"].to_numpy()
    dts = plume_df["datetime"].to_numpy()
    lats = plume_df["hull_lats"].to_numpy(dtype=np.float64)
    lons = plume_df["hull_lons"].to_numpy(dtype=np.float64)
    order = {d: i for i, d in enumerate(dict.fromkeys(dts.tolist()))}
    dti = np.array([order[d] for d in dts.tolist()])
    seen, keep = set(), set()
    for pid, di in sorted(set(zip(ids.tolist(), dti.tolist()))):
        sel = (ids == pid) & (dti == di)
        key = (di, float(np.round(lats[sel].mean(), 3)), float(np.round(lons[sel].mean(), 3)))
        if key not in seen:
            seen.add(key)
            keep.add((pid, di))
    rows = np.array([(i, d) in keep for i, d in zip(ids.tolist(), dti.tolist())], dtype=bool)
    return plume_df[rows].reset_index(drop=True)


# ---------------------------------------------------------------------------------------------- dataset builder
def hulls_of(plume_df, datetime=None) -> List[Hull]:
    """The hulls of a hull table (optionally of one datetime), one (hull_x, hull_y) pair per plume id."""
    df = plume_df if datetime is None else plume_df[plume_df["datetime"] == datetime]
    out = []
    for _, g in df.groupby(["id", "datetime"] if "datetime" in df.columns else ["id"], sort=True):
        out.append((g["hull_x"].to_numpy(), g["hull_y"].to_numpy()))
    return out


def build_training_tiles(scene: torch.Tensor, hulls: Sequence[Hull], tile: int = 256, stride: Optional[int] = None,
                         min_plume_pixels: int = 1, rasterizer: Optional[LabelRasterizer] = None):
    """scene: [H, W, C] bf16 on the device (the AOD / band stack); hulls: the scene's plumes.  Cuts the scene into
    `tile` x `tile` windows every `stride` pixels (default: non-overlapping), rasterises the hull masks straight
    into those windows and keeps the windows holding at least `min_plume_pixels` plume pixels.
    Returns (x [n, tile, tile, C] bf16, mask [n, tile, tile] uint8, ys, xs)."""
    r = rasterizer if rasterizer is not None else _rasterizer()
    h, w, c = scene.shape
    stride = stride or tile
    ys = [y for y in range(0, max(h - tile, 0) + 1, stride) for _ in range(0, max(w - tile, 0) + 1, stride)]
    xs = [x for _ in range(0, max(h - tile, 0) + 1, stride) for x in range(0, max(w - tile, 0) + 1, stride)]
    masks = r.tile_masks(hulls, ys, xs, tile)
    keep = (masks.flatten(1).sum(1, dtype=torch.int32) >= min_plume_pixels).nonzero().flatten()
    ys_t = torch.tensor(ys, dtype=torch.int32, device=scene.device)[keep].contiguous()
    xs_t = torch.tensor(xs, dtype=torch.int32, device=scene.device)[keep].contiguous()
    x = torch.empty(len(keep), tile, tile, c, dtype=scene.dtype, device=scene.device)
    if len(keep):
        r.ops.extract_tiles(scene, ys_t, xs_t, tile, x)
    return x, masks[keep].contiguous(), ys_t, xs_t


def write_training_tiles(folder: str, name: str, x: torch.Tensor, mask: torch.Tensor) -> str:
    """``<folder>/<name>.pt`` = {"x": [n,h,w,c], "mask": [n,h,w] uint8} -- the file format train_model reads from
    ``path_to_model_data_folder``."""
    os.makedirs(folder, exist_ok=True)
    path = os.path.join(folder, f"{name}.pt")
    torch.save({"x": x.cpu(), "mask": mask.cpu()}, path)
    return path
