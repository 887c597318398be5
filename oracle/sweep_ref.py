"""TEST INFRASTRUCTURE -- CPU restatement (numpy / scipy.ndimage) of the reference's threshold sweep, SURVEY.md
section 8(f) rank 2: ``plume_identifier_gaussian_profile.py:142-240``.

PARITY PARTLY PINNED.  The reference's control flow (mask generation loop, per-fire window search with its
distance matrix, size-ratio logic) is pinned: ``scripts/make_sweep_golden.py`` runs the reference's own functions,
compiled unmodified from the reference file, and ``tests/test_sweep_oracle.py`` checks this module against those
outputs.  The three scikit-image primitives they call (``label``, ``binary_erosion``, ``binary_dilation``) are NOT
pinned: scikit-image is not installed in this image and the reference pins no version (requirements.txt), so the
golden run binds them to scipy.ndimage with scikit-image's documented defaults (cross footprint; erosion sees set
pixels beyond the border, dilation unset ones; 8-connected labelling).  This module states the same primitives.

Only ``tests/``, ``smoke()`` and cpu_baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np
import scipy.ndimage as ndi

P_ID_WIN_SIZE = 15   # gaussian_profile.py:37


def threshold_masks_ref(aod: np.ndarray, thresholds) -> np.ndarray:
    """gaussian_profile.py:142-154 for every threshold: bool [T, H, W] = dilate(erode(aod > t)), cross footprint."""
    out = []
    for t in thresholds:
        m = aod > t
        p = np.pad(m, 1, constant_values=True)                       # erosion: beyond the border counts as set
        e = p[1:-1, 1:-1] & p[:-2, 1:-1] & p[2:, 1:-1] & p[1:-1, :-2] & p[1:-1, 2:]
        p = np.pad(e, 1, constant_values=False)                      # dilation: beyond the border counts as unset
        out.append(p[1:-1, 1:-1] | p[:-2, 1:-1] | p[2:, 1:-1] | p[1:-1, :-2] | p[1:-1, 2:])
    return np.stack(out)


def label_ref(mask: np.ndarray) -> np.ndarray:
    """8-connected components (skimage.measure.label default for 2-D, gaussian_profile.py:172) in canonical form:
    every pixel of a component carries 1 + the smallest row-major index of the component, background 0."""
    lab, n = ndi.label(mask, structure=np.ones((3, 3), dtype=int))
    if n == 0:
        return np.zeros(mask.shape, dtype=np.int64)
    first = ndi.minimum(np.arange(mask.size).reshape(mask.shape), lab, index=np.arange(1, n + 1)).astype(np.int64)
    canon = np.concatenate([[0], first + 1])
    return canon[lab]


def extract_label_ref(labelled: np.ndarray, r: int, c: int, win: int = P_ID_WIN_SIZE):
    """gaussian_profile.py:182-202: label of the labelled pixel nearest to (r, c) inside the (2 win + 1)^2 window
    (Euclidean pixel distance, first in row-major window order on ties), None if the window holds no label."""
    sub = labelled[r - win:r + win + 1, c - win:c + win + 1]
    if sub.shape != (2 * win + 1, 2 * win + 1):
        raise ValueError("fire closer than the window to the image edge (the reference filters these out)")
    nz = sub != 0
    if not nz.any():
        return None
    dy, dx = np.meshgrid(np.arange(-win, win + 1), np.arange(-win, win + 1), indexing="ij")
    d2 = (dy * dy + dx * dx)[nz]
    return sub[nz][np.argmin(d2)]


def find_plume_extents_ref(masks: np.ndarray, fire_rows, fire_cols, win: int = P_ID_WIN_SIZE) -> np.ndarray:
    """gaussian_profile.py:157-179: float64 [T, n_fires], the pixel count of the component nearest to each fire
    under each threshold (0 where the fire's window holds none)."""
    ext = np.zeros((len(masks), len(fire_rows)))
    for ti, m in enumerate(masks):
        lab = label_ref(m)
        sizes = np.bincount(lab.ravel())
        for fi, (r, c) in enumerate(zip(fire_rows, fire_cols)):
            l = extract_label_ref(lab, int(r), int(c), win)
            if l is not None:
                ext[ti, fi] = sizes[l]
    return ext


def find_threshold_index_ref(extents: np.ndarray):
    """gaussian_profile.py:204-240: per fire (column) the index of the largest ratio extents[i+1] / extents[i],
    None when there is no plume (all ratios undefined, or the maximum directly follows an undefined ratio)."""
    best = []
    for col in np.asarray(extents, dtype=np.float64).T:
        prev, nxt = col[:-1], col[1:]
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = nxt / prev
        ratio[prev == 0] = np.nan
        if np.all(np.isnan(ratio)):
            best.append(None)
            continue
        k = int(np.nanargmax(ratio))
        if np.any(np.isnan(ratio)) and k == np.where(np.isnan(ratio))[0][-1] + 1:
            best.append(None)
            continue
        best.append(None if k == ratio.size else k)
    return best


def cluster_fires_ref(shape, fire_rows, fire_cols, min_size: int = 3) -> np.ndarray:
    """gaussian_profile.py:126-139: the fire pixels as an 8-connected label image (labels 1..n in raster order of each
    cluster's first pixel, scikit-image's numbering), clusters of fewer than ``min_size`` pixels set to 0 with the
    other labels unchanged (``remove_small_objects`` on a label image)."""
    grid = np.zeros(shape, dtype=bool)
    grid[np.asarray(fire_rows, dtype=np.int64), np.asarray(fire_cols, dtype=np.int64)] = True
    lab, _ = ndi.label(grid, structure=np.ones((3, 3), dtype=int))
    lab = lab.astype(np.int64)
    small = np.bincount(lab.ravel()) < min_size
    lab[small[lab]] = 0
    return lab


def cluster_centroids_ref(labels: np.ndarray):
    """gaussian_profile.py:474-477: ``[r.centroid for r in regionprops(labels)]`` truncated to int -- per remaining
    label in ascending order the mean row and mean column of its pixels."""
    ids = np.unique(labels[labels > 0])
    rows = np.array([np.nonzero(labels == i)[0].mean() for i in ids]).astype(int)
    cols = np.array([np.nonzero(labels == i)[1].mean() for i in ids]).astype(int)
    return rows, cols


NULL_VALUE = -999   # gaussian_profile.py:41


def interpolate_aod_nearest_ref(aod: np.ndarray, null_value=NULL_VALUE, return_unique: bool = False):
    """gaussian_profile.py:451-461 by brute force: every pixel takes the value of the nearest pixel != null_value
    (Euclidean distance in pixel units; valid pixels are their own nearest).  scipy's NearestNDInterpolator answers a
    query that has several equidistant nearest points with whichever its kd-tree visits first; this restatement takes
    the first in row-major order (np.argmin), and ``return_unique`` also returns the mask of pixels whose nearest
    valid pixel is unique -- where both must agree (tests/test_sweep_oracle.py pins exactly that against scipy)."""
    aod = np.asarray(aod)
    good = aod != null_value
    if not good.any():
        raise ValueError("no valid pixel")
    gy, gx = np.nonzero(good)                                   # row-major order
    vals = aod[gy, gx]
    out = aod.copy()
    unique = np.ones(aod.shape, dtype=bool)
    by, bx = np.nonzero(~good)
    for lo in range(0, len(by), 2048):
        y, x = by[lo:lo + 2048, None].astype(np.int64), bx[lo:lo + 2048, None].astype(np.int64)
        d2 = (y - gy[None, :]) ** 2 + (x - gx[None, :]) ** 2
        k = np.argmin(d2, axis=1)
        out[by[lo:lo + 2048], bx[lo:lo + 2048]] = vals[k]
        unique[by[lo:lo + 2048], bx[lo:lo + 2048]] = (d2 == d2.min(axis=1, keepdims=True)).sum(axis=1) == 1
    return (out, unique) if return_unique else out


def plume_mask_ref(mask: np.ndarray, r: int, c: int, win: int = P_ID_WIN_SIZE):
    """gaussian_profile.py:306-331 (find_plume_mask up to assess_plume): label the mask, take the label nearest to the
    fire, return ``labelled == label`` (None when the window holds no label)."""
    lab = label_ref(mask)
    l = extract_label_ref(lab, int(r), int(c), win)
    return None if l is None else lab == l
