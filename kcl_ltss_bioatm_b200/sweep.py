"""Threshold sweep on the GPU: for every threshold of a sweep, the AOD mask, its connected
components and, per fire, the size of the nearest component -- the reference's real CPU hot loop (about 75
``skimage.measure.label`` calls on a 1200 x 1200 grid per timestamp, plume_identifier_gaussian_profile.py:489-503).

Reference-named functions (same arguments and results):

    generate_mask_dict(aod, threshold_range)                     gaussian_profile.py:142-154
    find_plume_extents(masks_dict, fire_rows, fire_cols)         gaussian_profile.py:157-179
    find_threshold_index(plume_extents_across_all_fires)         gaussian_profile.py:204-240  (host, numpy)
    plume_masks(masks_dict, threshold_index_for_fires, rows, cols)   the label / extract_label / == part of find_plume_mask, :306-331
    cluster_fires(aod, fire_rows, fire_cols)                     gaussian_profile.py:126-139
    fire_cluster_centroids(fire_labels)                          gaussian_profile.py:474-477  (host, numpy)
    interpolate_aod_nearest(aod)                                 gaussian_profile.py:451-461

``ThresholdSweep.extents`` is the fused form: masks as bit planes (32 pixels per word), components over runs of set
bits, only the [T, n_fires] extents leave the device.  ``masks`` / ``label`` give dense byte / int32 planes.  The masks use the
cross-shaped footprint with scikit-image's border rules (erosion: set beyond the border; dilation: unset) and the
labelling is 8-connected, scikit-image's default for 2-D.  There is no CPU path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .ops import CudaOps

P_ID_WIN_SIZE = 15  # gaussian_profile.py:37
NULL_VALUE = -999   # gaussian_profile.py:41
THRESHOLD_STEP_SIZES = [0.02, 0.03, 0.04]   # gaussian_profile.py:34
THRESHOLD_MAX = [0.5, 0.75, 1]               # gaussian_profile.py:35


class ThresholdSweep:
    def __init__(self, device="cuda", ops: Optional[CudaOps] = None):
        self.device = torch.device(device)
        self.ops = ops if ops is not None else CudaOps()

    def _image(self, aod) -> torch.Tensor:
        """The image on the device in its own precision: float64 stays float64 (the reference's AOD is int16 * 0.001 in
        float64 and is compared in float64, tools.py:88), everything else becomes float32."""
        if torch.is_tensor(aod):
            dt = torch.float64 if aod.dtype == torch.float64 else torch.float32
            return aod.to(device=self.device, dtype=dt).contiguous()
        a = np.asarray(aod)
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64 if a.dtype == np.float64 else np.float32)).to(self.device)

    def masks(self, aod, thresholds) -> torch.Tensor:
        """uint8 [T, H, W] on the device."""
        a = self._image(aod)
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

    # ---- bit-plane path (32 pixels per word; csrc/sweep_bits.cuh): what extents / find_plume_extents run on
    def _workspace(self, h, w, t) -> torch.Tensor:
        need = self.ops.sweep_workspace_bytes(h, w, t)
        ws = getattr(self, "_ws", None)
        if ws is None or ws.numel() < need:
            self._ws = ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return ws

    def _fires(self, fire_rows, fire_cols, h, w, win) -> torch.Tensor:
        rc = np.stack([np.asarray(fire_rows, dtype=np.int64).reshape(-1), np.asarray(fire_cols, dtype=np.int64).reshape(-1)], 1)
        if len(rc) and (rc[:, 0].min() < win or rc[:, 0].max() > h - win - 1 or rc[:, 1].min() < win
                        or rc[:, 1].max() > w - win - 1):
            raise ValueError("a fire is closer than the window to the image edge (locate_fire_in_image filters these)")
        return torch.tensor(rc, dtype=torch.int32).to(self.device)

    def mask_bits(self, aod, thresholds) -> torch.Tensor:
        """int32 [T, H, ceil(W / 32)] on the device: bit i of word s = masks[t, y, 32 s + i]."""
        a = self._image(aod)
        thr = torch.tensor(np.asarray(thresholds, dtype=np.float64).reshape(-1)).to(self.device)
        bits = torch.empty(thr.numel(), a.shape[0], (a.shape[1] + 31) // 32, dtype=torch.int32, device=self.device)
        if bits.numel():
            self.ops.threshold_mask_bits(a, thr, bits)
        return bits

    @staticmethod
    def unpack_bits(bits: torch.Tensor, w: int) -> np.ndarray:
        """bool [T, H, W] on the host from device bit planes (the device -> host copy is 8 x smaller than byte masks)."""
        b = bits.cpu().numpy()
        t, h, segs = b.shape
        return np.unpackbits(b.view(np.uint8).reshape(t, h, segs * 4), axis=2, bitorder="little")[:, :, :w].astype(bool)

    def pack_bits_host(self, masks: np.ndarray) -> torch.Tensor:
        """bool [T, H, W] on the host -> device bit planes (packed on the host: the copy is 8 x smaller)."""
        t, h, w = masks.shape
        segs = (w + 31) // 32
        padded = np.zeros((t, h, segs * 32), dtype=bool)
        padded[:, :, :w] = masks
        words = np.packbits(padded, axis=2, bitorder="little").view("<u4").view(np.int32)
        return torch.from_numpy(np.ascontiguousarray(words)).to(self.device)

    def extents_of_bits(self, bits: torch.Tensor, w: int, fire_rows, fire_cols, win: int = P_ID_WIN_SIZE) -> np.ndarray:
        """device bit planes [T, H, ceil(W / 32)] -> float64 [T, n_fires] (find_plume_extents)."""
        t, h, _ = bits.shape
        rc = self._fires(fire_rows, fire_cols, h, w, win)
        out = torch.zeros(t, len(rc), dtype=torch.int32, device=bits.device)
        if len(rc) and t:
            self.ops.bits_extents(bits, w, rc, win, self._workspace(h, w, t), out)
        return out.cpu().numpy().astype(np.float64)

    def fire_components(self, bits: torch.Tensor, w: int, plane_of_fire, fire_rows, fire_cols, win: int = P_ID_WIN_SIZE):
        """Per fire the connected component nearest to it in the plane chosen for it (``find_plume_mask``: label,
        extract_label, ``labelled == label`` -- without a label plane).  bits: device bit planes [T, H, ceil(W / 32)];
        plane_of_fire: plane index per fire, negative / None = no plane.  Returns (crops, stats): stats int64 [n, 6] on
        the host = area, min_row, min_col, max_row + 1, max_col + 1, root (area 0 where the fire has no plane or its
        window holds no component); crops[f] = bool array of the component inside its bounding box
        [min_row:max_row + 1, min_col:max_col + 1], None where area is 0.  Only fires with a plane get a component
        plane on the device, and only bounding boxes are unpacked on the host."""
        t, h, segs = bits.shape
        rc = self._fires(fire_rows, fire_cols, h, w, win)
        planes = np.array([-1 if p is None else int(p) for p in plane_of_fire], dtype=np.int64)
        if len(planes) != len(rc) or (planes >= t).any():
            raise ValueError("plane_of_fire needs one valid plane index (or None) per fire")
        stats = np.zeros((len(rc), 6), dtype=np.int64)
        stats[:, 5] = -1
        crops = [None] * len(rc)
        sel = np.nonzero(planes >= 0)[0]
        if len(sel) == 0 or t == 0:
            return crops, stats
        sel_t = torch.from_numpy(sel).to(bits.device)
        comp = torch.empty(len(sel), h, segs, dtype=torch.int32, device=bits.device)
        st = torch.empty(len(sel), 8, dtype=torch.int32, device=bits.device)
        ws = self._workspace(h, w, t)
        scratch = torch.empty(t, len(sel), dtype=torch.int32, device=bits.device)
        rc_sel = rc[sel_t].contiguous()
        self.ops.bits_extents(bits, w, rc_sel, win, ws, scratch)                    # labels the planes into ws
        self.ops.fire_components(bits, w, rc_sel, torch.from_numpy(planes[sel].astype(np.int32)).to(bits.device), win, ws,
                                 comp, st)
        stats[sel] = st[:, :6].cpu().numpy()
        words = comp.cpu().numpy().view(np.uint32)
        for j, f in enumerate(sel):
            area, y0, x0, y1, x1 = stats[f, :5]
            if area > 0:
                s0, s1 = x0 >> 5, (x1 + 31) >> 5
                sub = np.ascontiguousarray(words[j, y0:y1, s0:s1])
                px = np.unpackbits(sub.view(np.uint8).reshape(y1 - y0, (s1 - s0) * 4), axis=1, bitorder="little")
                crops[f] = px[:, x0 - 32 * s0:x1 - 32 * s0].astype(bool)
        return crops, stats

    @staticmethod
    def full_mask(crop, stats_row, shape) -> np.ndarray:
        """The [H, W] mask of a component from its bounding-box crop."""
        out = np.zeros(shape, dtype=bool)
        out[stats_row[1]:stats_row[3], stats_row[2]:stats_row[4]] = crop
        return out

    def extents_of_masks(self, masks: torch.Tensor, fire_rows, fire_cols, win: int = P_ID_WIN_SIZE) -> np.ndarray:
        """uint8 masks [T, H, W] on the device -> float64 [T, n_fires] (find_plume_extents)."""
        t, h, w = masks.shape
        rc = self._fires(fire_rows, fire_cols, h, w, win)
        out = torch.zeros(t, len(rc), dtype=torch.int32, device=masks.device)
        if len(rc) and t:
            bits = torch.empty(t, h, (w + 31) // 32, dtype=torch.int32, device=masks.device)
            self.ops.pack_mask_bits(masks.contiguous(), bits)
            self.ops.bits_extents(bits, w, rc, win, self._workspace(h, w, t), out)
        return out.cpu().numpy().astype(np.float64)

    def extents_dense(self, masks: torch.Tensor, fire_rows, fire_cols, win: int = P_ID_WIN_SIZE) -> np.ndarray:
        """The same through the dense int32 label planes (label + fire_extents)."""
        h, w = masks.shape[1:]
        rc = self._fires(fire_rows, fire_cols, h, w, win)
        labels, sizes = self.label(masks)
        out = torch.zeros(masks.shape[0], len(rc), dtype=torch.int32, device=masks.device)
        if len(rc):
            self.ops.fire_extents(labels, sizes, rc, win, out)
        return out.cpu().numpy().astype(np.float64)

    def extents(self, aod, thresholds, fire_rows, fire_cols, win: int = P_ID_WIN_SIZE) -> np.ndarray:
        """float64 [T, n_fires]: generate_mask_dict + find_plume_extents in one call on the device (any T: the
        reference's three sweeps of a timestamp can be passed as one concatenated threshold list)."""
        a = self._image(aod)
        thr = torch.tensor(np.asarray(thresholds, dtype=np.float64).reshape(-1)).to(self.device)
        h, w = a.shape
        rc = self._fires(fire_rows, fire_cols, h, w, win)
        out = torch.zeros(thr.numel(), len(rc), dtype=torch.int32, device=self.device)
        if len(rc) and thr.numel():
            self.ops.sweep_extents(a, thr, rc, win, self._workspace(h, w, thr.numel()), out)
        return out.cpu().numpy().astype(np.float64)


    def fill_nearest(self, aod, null_value=NULL_VALUE) -> torch.Tensor:
        """The image (device tensor, float32 or float64 like the input) with every pixel == null_value replaced by the
        value of the nearest pixel != null_value (Euclidean pixel distance; among equidistant ones the first in
        row-major order)."""
        a = self._image(aod)
        if not bool((a != null_value).any()):
            raise ValueError("no valid pixel to interpolate from (NearestNDInterpolator needs at least one point)")
        ws = torch.empty(self.ops.fill_nearest_workspace_bytes(*a.shape), dtype=torch.uint8, device=self.device)
        out = torch.empty_like(a)
        self.ops.fill_nearest(a, null_value, ws, out)
        return out

    def cluster_fires(self, shape, fire_rows, fire_cols, min_size: int = 3) -> np.ndarray:
        """int64 [H, W]: the fire pixels labelled by 8-connected cluster, numbered 1..n in raster order of each
        cluster's first pixel, clusters smaller than ``min_size`` removed (their numbers stay unused).  The labelling
        runs on the device (label_components on the fire grid); only the labels and sizes AT the fire pixels come
        back, and the numbering -- a rank over a handful of roots -- is done on the host."""
        h, w = int(shape[0]), int(shape[1])
        rows = torch.as_tensor(np.asarray(fire_rows, dtype=np.int64).reshape(-1)).to(self.device)
        cols = torch.as_tensor(np.asarray(fire_cols, dtype=np.int64).reshape(-1)).to(self.device)
        out = np.zeros((h, w), dtype=np.int64)
        if rows.numel() == 0:
            return out
        if int(rows.min()) < 0 or int(rows.max()) >= h or int(cols.min()) < 0 or int(cols.max()) >= w:
            raise IndexError("a fire lies outside the image")
        grid = torch.zeros(1, h, w, dtype=torch.uint8, device=self.device)
        grid[0, rows, cols] = 1
        labels, sizes = self.label(grid)
        root = labels[0, rows, cols].to(torch.int64)                     # canonical label = first pixel of the cluster
        size = sizes.view(-1)[root].cpu().numpy()
        root = root.cpu().numpy()
        ranks = {r: k + 1 for k, r in enumerate(np.unique(root))}         # raster order of first pixels = label numbers
        keep = size >= min_size
        r, c = rows.cpu().numpy()[keep], cols.cpu().numpy()[keep]
        out[r, c] = [ranks[v] for v in root[keep]]
        return out


    def timestamp(self, aod, fire_rows, fire_cols, fill: bool = True, win: int = P_ID_WIN_SIZE):
        """The data-parallel front half of the reference's per-timestamp work in one go -- ``main`` :611-613 and
        ``identify`` :478-499: nearest-valid fill of the AOD grid, fire clustering and centroids, then for each of the
        three sweeps (THRESHOLD_STEP_SIZES / THRESHOLD_MAX) the masks, the plume extents of every fire cluster, the
        threshold index per cluster (host) and, for the clusters that have one, the plume mask that
        ``find_plume_mask`` would hand to ``assess_plume``.  The image crosses PCIe once; the 75 thresholds are
        labelled in one call.  Returns a dict: ``aod_filled`` (device tensor), ``fire_rows`` / ``fire_cols`` (cluster
        centroids), ``sweeps`` = list of {``thresholds``, ``extents`` [T, n], ``threshold_index`` [n],
        ``regions`` int64 [n, 6] = area, bounding box (min_row, min_col, max_row + 1, max_col + 1), root; ``plume_masks`` =
        per cluster the bool mask INSIDE that bounding box, or None (``full_mask`` expands one to [H, W])}."""
        a = self._image(aod)
        if fill:
            a = self.fill_nearest(a)
        h, w = a.shape
        labels = self.cluster_fires((h, w), fire_rows, fire_cols)
        rows, cols = fire_cluster_centroids(labels)
        out = {"aod_filled": a, "fire_labels": labels, "fire_rows": rows, "fire_cols": cols, "sweeps": []}
        ranges = [np.abs(np.arange(0, tmax, step) - tmax) for step, tmax in zip(THRESHOLD_STEP_SIZES, THRESHOLD_MAX)]
        if len(rows) == 0:
            return out
        thr_all = np.concatenate(ranges)
        bits = self.mask_bits(a, thr_all)
        ext_all = self.extents_of_bits(bits, w, rows, cols, win)
        lo = 0
        for thr in ranges:
            ext = ext_all[lo:lo + len(thr)]
            index = find_threshold_index(ext)
            planes = [None if k is None else lo + k for k in index]
            masks, regions = self.fire_components(bits, w, planes, rows, cols, win)
            out["sweeps"].append({"thresholds": thr, "extents": ext, "threshold_index": index, "plume_masks": masks,
                                  "regions": regions})
            lo += len(thr)
        return out


_default: Optional[ThresholdSweep] = None


def _sweep() -> ThresholdSweep:
    global _default
    if _default is None:
        _default = ThresholdSweep()
    return _default


class MaskDict(dict):
    """What ``generate_mask_dict`` returns: a dict {threshold: bool mask [H, W]} whose masks stay on the device as bit
    planes until somebody reads them.  ``masks[t]`` (and values() / items()) copies that one plane to the host and
    unpacks it; ``find_plume_extents`` recognises the object and works on the resident planes, so the reference's call
    sequence generate_mask_dict -> find_plume_extents -> find_threshold_index moves no mask across PCIe at all, and
    ``extract_plume_roi`` (:243-303) only fetches the few planes it indexes.  Assigning to an entry turns it into an
    ordinary host entry (and find_plume_extents then takes the host path)."""

    def __init__(self, thresholds, bits: torch.Tensor, width: int):
        super().__init__()
        self._bits, self._width, self._plane, self._host_set = bits, width, {}, False
        for i, t in enumerate(thresholds):                   # equal thresholds: one key, like the reference's dict
            dict.__setitem__(self, t, None)
            self._plane[t] = i

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if v is None:
            v = ThresholdSweep.unpack_bits(self._bits[self._plane[key]:self._plane[key] + 1], self._width)[0]
            dict.__setitem__(self, key, v)
        return v

    def __setitem__(self, key, value):
        self._host_set = True
        dict.__setitem__(self, key, value)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def __iter__(self):                 # an overridden __iter__ also makes dict(d) / d2.update(d) go through __getitem__
        return super().__iter__()

    def copy(self):
        return dict(self.items())

    def pop(self, key, *default):
        if key in self:
            v = self[key]
            dict.pop(self, key)
            return v
        if default:
            return default[0]
        raise KeyError(key)

    def values(self):
        return [self[k] for k in self]

    def items(self):
        return [(k, self[k]) for k in self]

    def device_planes(self) -> Optional[torch.Tensor]:
        """int32 [len(self), H, ceil(W / 32)] in key order, or None once an entry was assigned from the host."""
        if self._host_set or any(k not in self._plane for k in self):
            return None
        return self._bits[torch.tensor([self._plane[k] for k in self], dtype=torch.long, device=self._bits.device)]


def generate_mask_dict(aod, threshold_range) -> Dict[float, np.ndarray]:
    """{threshold: bool mask [H, W]} -- aod > t with singleton pixels removed (erosion then dilation); a MaskDict."""
    threshold_range = list(threshold_range)
    return MaskDict(threshold_range, _sweep().mask_bits(aod, threshold_range), np.shape(aod)[1])


def find_plume_extents(masks_dict, fire_rows, fire_cols) -> np.ndarray:
    """[len(masks_dict), len(fires)]: per threshold (dict order) and fire the pixel count of the labelled region
    nearest to the fire within its 31 x 31 window, 0 where there is none."""
    s = _sweep()
    planes = masks_dict.device_planes() if isinstance(masks_dict, MaskDict) else None
    if planes is not None:
        return s.extents_of_bits(planes.to(s.device), masks_dict._width, fire_rows, fire_cols)
    stack = np.stack([np.asarray(masks_dict[k]) for k in masks_dict]) != 0
    return s.extents_of_bits(s.pack_bits_host(stack), stack.shape[2], fire_rows, fire_cols)


def plume_masks(masks_dict, threshold_index_for_fires, fire_rows, fire_cols):
    """For every fire with a threshold index (``find_threshold_index``'s result; None = no plume) the bool mask [H, W] of
    the labelled region nearest to the fire in that threshold's mask -- what ``find_plume_mask`` (:306-331) computes per
    fire by relabelling the whole mask: ``labelled_mask == extract_label(labelled_mask, r, c)``.  None where there is no
    index or the fire's window holds no region.  Also returns the regions' (area, bbox) rows."""
    s = _sweep()
    keys = list(masks_dict)
    planes = masks_dict.device_planes() if isinstance(masks_dict, MaskDict) else None
    if planes is not None:
        w = masks_dict._width
    else:
        stack = np.stack([np.asarray(masks_dict[k]) for k in keys]) != 0
        planes, w = s.pack_bits_host(stack), stack.shape[2]
    crops, stats = s.fire_components(planes.to(s.device), w, threshold_index_for_fires, fire_rows, fire_cols)
    shape = (planes.shape[1], w)
    return [None if c is None else s.full_mask(c, stats[f], shape) for f, c in enumerate(crops)], stats


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


def cluster_fires(aod, fire_rows, fire_cols) -> np.ndarray:
    """Label image of the fire clusters (8-connected, at least 3 pixels), shaped like ``aod``."""
    return _sweep().cluster_fires(np.shape(aod), fire_rows, fire_cols)


def interpolate_aod_nearest(aod) -> np.ndarray:
    """The AOD grid with its NULL_VALUE pixels filled from the nearest valid pixel; float64 like scipy's result
    (values are copied: a float32 image is widened exactly)."""
    return _sweep().fill_nearest(aod).cpu().numpy().astype(np.float64)


def fire_cluster_centroids(fire_labels):
    """``[r.centroid for r in regionprops(fire_labels)]`` as integer (rows, cols) arrays, labels ascending --
    what ``identify`` feeds to the sweep (a few dozen pixels: host numpy)."""
    fire_labels = np.asarray(fire_labels)
    r, c = np.nonzero(fire_labels)
    lab = fire_labels[r, c]
    ids, inv, cnt = np.unique(lab, return_inverse=True, return_counts=True)
    rows = np.bincount(inv, weights=r, minlength=len(ids)) / np.maximum(cnt, 1)
    cols = np.bincount(inv, weights=c, minlength=len(ids)) / np.maximum(cnt, 1)
    return rows.astype(int), cols.astype(int)
