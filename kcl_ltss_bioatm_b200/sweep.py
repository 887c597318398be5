"""Threshold sweep on the GPU: for every threshold of a sweep, the AOD mask, its connected
components and, per fire, the size of the nearest component -- the reference's real CPU hot loop (about 75
``skimage.measure.label`` calls on a 1200 x 1200 grid per timestamp, plume_identifier_gaussian_profile.py:489-503).

Reference-named functions (same arguments and results):

    generate_mask_dict(aod, threshold_range)                     gaussian_profile.py:142-154
    find_plume_extents(masks_dict, fire_rows, fire_cols)         gaussian_profile.py:157-179
    find_threshold_index(plume_extents_across_all_fires)         gaussian_profile.py:204-240  (host, numpy)

``ThresholdSweep.extents`` is the fused form that keeps masks, labels and sizes on the device.  The masks use the
cross-shaped footprint with scikit-image's border rules (erosion: set beyond the border; dilation: unset) and the
labelling is 8-connected, scikit-image's default for 2-D.  There is no CPU path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .ops import CudaOps

P_ID_WIN_SIZE = 15  # gaussian_profile.py:37


class ThresholdSweep:
    def __init__(self, device="cuda", ops: Optional[CudaOps] = None):
        self.device = torch.device(device)
        self.ops = ops if ops is not None else CudaOps()

    def masks(self, aod, thresholds) -> torch.Tensor:
        """uint8 [T, H, W] on the device."""
        a = torch.as_tensor(np.asarray(aod, dtype=np.float32) if not torch.is_tensor(aod) else aod,
                            dtype=torch.float32).to(self.device).contiguous()
        thr = torch.tensor(np.asarray(thresholds, dtype=np.float64)).to(self.device)
        out = torch.empty(thr.numel(), *a.shape, dtype=torch.uint8, device=self.device)
        self.ops.threshold_masks(a, thr, out)
        return out

    def label(self, masks: torch.Tensor):
        """(labels int32 [T, H, W]: -1 background, else the component's smallest row-major pixel index;
        sizes int32 [T, H, W]: the component's pixel count at that index)."""
        labels = torch.empty(masks.shape, dtype=torch.int32, device=masks.device)
        sizes = torch.empty(masks.shape, dtype=torch.int32, device=masks.device)
        self.ops.label_components(masks, labels, sizes)
        return labels, sizes

    def extents_of_masks(self, masks: torch.Tensor, fire_rows, fire_cols, win: int = P_ID_WIN_SIZE) -> np.ndarray:
        h, w = masks.shape[1:]
        rc = np.stack([np.asarray(fire_rows, dtype=np.int64), np.asarray(fire_cols, dtype=np.int64)], 1)
        if len(rc) and (rc[:, 0].min() < win or rc[:, 0].max() > h - win - 1 or rc[:, 1].min() < win
                        or rc[:, 1].max() > w - win - 1):
            raise ValueError("a fire is closer than the window to the image edge (locate_fire_in_image filters these)")
        labels, sizes = self.label(masks)
        out = torch.zeros(masks.shape[0], len(rc), dtype=torch.int32, device=masks.device)
        if len(rc):
            self.ops.fire_extents(labels, sizes, torch.tensor(rc, dtype=torch.int32).to(masks.device), win, out)
        return out.cpu().numpy().astype(np.float64)

    def extents(self, aod, thresholds, fire_rows, fire_cols, win: int = P_ID_WIN_SIZE) -> np.ndarray:
        """float64 [T, n_fires]: generate_mask_dict + find_plume_extents in one pass on the device."""
        return self.extents_of_masks(self.masks(aod, thresholds), fire_rows, fire_cols, win)


_default: Optional[ThresholdSweep] = None


def _sweep() -> ThresholdSweep:
    global _default
    if _default is None:
        _default = ThresholdSweep()
    return _default


def generate_mask_dict(aod, threshold_range) -> Dict[float, np.ndarray]:
    """{threshold: bool mask [H, W]} -- aod > t with singleton pixels removed (erosion then dilation)."""
    m = _sweep().masks(aod, threshold_range).cpu().numpy().astype(bool)
    return {t: m[i] for i, t in enumerate(threshold_range)}


def find_plume_extents(masks_dict, fire_rows, fire_cols) -> np.ndarray:
    """[len(masks_dict), len(fires)]: per threshold (dict order) and fire the pixel count of the labelled region
    nearest to the fire within its 31 x 31 window, 0 where there is none."""
    stack = np.stack([np.asarray(masks_dict[k]) for k in masks_dict]).astype(np.uint8)
    s = _sweep()
    return s.extents_of_masks(torch.from_numpy(stack).to(s.device), fire_rows, fire_cols)


def find_threshold_index(plume_extents_across_all_fires) -> List[Optional[int]]:
    """Per fire the threshold index with the largest jump in plume size (ratio of consecutive extents), None when
    no plume can be told apart: all ratios undefined, or the maximum directly follows an undefined ratio."""
    best: List[Optional[int]] = []
    for extents in np.asarray(plume_extents_across_all_fires, dtype=np.float64).T:
        null = extents[:-1] == 0
        with np.errstate(divide="ignore", invalid="ignore"):
            ratios = extents[1:] / extents[:-1]
        ratios[null] = np.nan
        if np.all(np.isnan(ratios)):
            best.append(None)
            continue
        k = int(np.nanargmax(ratios))
        if np.any(np.isnan(ratios)) and k == np.where(np.isnan(ratios))[0][-1] + 1:
            best.append(None)
            continue
        best.append(None if k == ratios.size else k)
    return best
