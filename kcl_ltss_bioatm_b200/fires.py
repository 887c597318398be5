"""Fire -> pixel geolocation on the GPU.

The reference finds, for every active-fire detection, the image pixel nearest to it with a Python loop that
builds four full-image boolean masks per fire (plume_identifier_gaussian_profile.py:85-123; the older variants
:135-161 of plume_identifier_basic.py do the same with a global argmin).  This module keeps the reference's
function names and argument meanings

    subset_fires_to_image(lat, lon, fire_df, date_to_find)       gaussian_profile.py:46-54
    grid_indexes(lat)                                            gaussian_profile.py:57-62
    haversine(lon1, lat1, lon2, lat2)                            gaussian_profile.py:65-82
    locate_fire_in_image(fire_coords, lats, lons, rows, cols)    gaussian_profile.py:85-123

and runs the O(fires x pixels) search in ``plume_locate_fires`` (all fires of a timestamp in one call, float64
haversine in the reference's operation order).  ``FireLocator`` keeps the lat/lon grids of a tile on the device
across timestamps.  There is no CPU path: without the CUDA library ``locate_fire_in_image`` raises.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from .ops import CudaOps

P_ID_WIN_SIZE = 15      # gaussian_profile.py:37: fires closer than this (+1) to the image edge are dropped
HALF_BOX_DEG = 0.05     # gaussian_profile.py:97-98


def subset_fires_to_image(lat, lon, fire_df, date_to_find):
    """Fires (pandas DataFrame with latitude, longitude, date_time) of one timestamp strictly inside the image's
    latitude / longitude extent."""
    m = (fire_df.date_time == date_to_find) & (fire_df.latitude > np.min(lat)) & (fire_df.latitude < np.max(lat)) \
        & (fire_df.longitude > np.min(lon)) & (fire_df.longitude < np.max(lon))
    return fire_df[m]


def grid_indexes(lat):
    rows = np.arange(lat.shape[0])
    cols = np.arange(lat.shape[1])
    cols, rows = np.meshgrid(cols, rows)
    return rows, cols


def haversine(lon1, lat1, lon2, lat2):
    """Great-circle distance in km between points given in decimal degrees (earth radius 6367 km)."""
    lon1, lat1, lon2, lat2 = map(np.radians, [lon1, lat1, lon2, lat2])
    a = np.sin((lat2 - lat1) / 2.0) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin((lon2 - lon1) / 2.0) ** 2
    return 6367 * (2 * np.arcsin(np.sqrt(a)))


class FireLocator:
    """Nearest-pixel search for batches of fires against one lat/lon grid kept on the device."""

    def __init__(self, lats, lons, device="cuda", ops: Optional[CudaOps] = None):
        self.device = torch.device(device)
        self.ops = ops if ops is not None else CudaOps()
        self.lats = torch.tensor(np.asarray(lats, dtype=np.float64)).to(self.device).contiguous()
        self.lons = torch.tensor(np.asarray(lons, dtype=np.float64)).to(self.device).contiguous()
        if self.lats.dim() != 2 or self.lats.shape != self.lons.shape:
            raise ValueError("lats / lons must be [H, W] grids of the same shape")

    def nearest_pixels(self, fire_lat, fire_lon) -> np.ndarray:
        """int32 [n, 2] (row, col) per fire, (-1, -1) where no pixel lies inside the fire's +-0.05 degree box."""
        fl = torch.tensor(np.asarray(fire_lat, dtype=np.float64)).to(self.device)
        fo = torch.tensor(np.asarray(fire_lon, dtype=np.float64)).to(self.device)
        out = torch.empty(fl.numel(), 2, dtype=torch.int32, device=self.device)
        if fl.numel():
            self.ops.locate_fires(self.lats, self.lons, fl, fo, HALF_BOX_DEG, out)
        return out.cpu().numpy()

    def locate(self, fire_lat, fire_lon, win: int = P_ID_WIN_SIZE) -> Tuple[List[int], List[int]]:
        """The reference's result: rows / cols of the fires that were found and are not within win+1 pixels of
        the image edge (its exact comparisons, gaussian_profile.py:108-114), in fire order."""
        rc = self.nearest_pixels(fire_lat, fire_lon)
        h, w = self.lats.shape
        r, c = rc[:, 0], rc[:, 1]
        ok = (r >= 0) & ~((r < win + 1) | (r > h - win - 1)) & ~((c < win + 1) | (c > w - win - 1))
        return r[ok].tolist(), c[ok].tolist()


def locate_fire_in_image(fire_coords, lats, lons, rows=None, cols=None):
    """fire_coords: DataFrame with latitude / longitude columns.  `rows` / `cols` (the index grids of
    ``grid_indexes``) are accepted for signature compatibility; the pixel indices are implied by the grid."""
    loc = FireLocator(lats, lons)
    return loc.locate(fire_coords.latitude.values, fire_coords.longitude.values)
