"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference's fire -> pixel geolocation, SURVEY.md section
8(f) rank 3.

PARITY PINNED against the reference's own functions (``plume_identifier_gaussian_profile.py:46-123``), compiled
unmodified from the reference file by ``scripts/make_fire_golden.py`` (fixtures ``tests/golden/fire_cases.npz``,
test ``tests/test_fire_oracle.py``).  Only ``tests/``, ``smoke()`` and cpu_baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np

P_ID_WIN_SIZE = 15          # gaussian_profile.py:37
HALF_BOX_DEG = 0.05         # gaussian_profile.py:97-98
EARTH_RADIUS_KM = 6367      # gaussian_profile.py:81


def haversine_ref(lon1, lat1, lon2, lat2):
    """gaussian_profile.py:65-82, float64."""
    lon1, lat1, lon2, lat2 = (np.radians(np.asarray(v, dtype=np.float64)) for v in (lon1, lat1, lon2, lat2))
    a = np.sin((lat2 - lat1) / 2.0) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin((lon2 - lon1) / 2.0) ** 2
    return EARTH_RADIUS_KM * 2 * np.arcsin(np.sqrt(a))


def nearest_pixel_ref(fire_lat, fire_lon, lats, lons):
    """The search of gaussian_profile.py:94-106 for every fire: among the pixels strictly inside the +-0.05 degree
    box around the fire, the first (row-major) one with the smallest haversine distance.  Returns int64 [n, 2]
    (row, col), -1 where the box holds no pixel (the reference's bare ``except`` skips those fires)."""
    lats = np.asarray(lats, dtype=np.float64)
    lons = np.asarray(lons, dtype=np.float64)
    w = lats.shape[1]
    out = np.full((len(fire_lat), 2), -1, dtype=np.int64)
    for i, (fl, fo) in enumerate(zip(fire_lat, fire_lon)):
        m = (lats > fl - HALF_BOX_DEG) & (lats < fl + HALF_BOX_DEG) & (lons > fo - HALF_BOX_DEG) & (lons < fo + HALF_BOX_DEG)
        idx = np.flatnonzero(m.ravel())
        if idx.size == 0:
            continue
        k = idx[np.argmin(haversine_ref(fo, fl, lons.ravel()[idx], lats.ravel()[idx]))]
        out[i] = (k // w, k % w)
    return out


def edge_filter_ref(rc, shape, win: int = P_ID_WIN_SIZE) -> np.ndarray:
    """gaussian_profile.py:108-114: keep fires at least win+1 pixels from the top/left and win+1 from the
    bottom/right image edge (the reference's exact, slightly asymmetric comparisons)."""
    r, c = rc[:, 0], rc[:, 1]
    ok = r >= 0
    ok &= ~((r < win + 1) | (r > shape[0] - win - 1))
    ok &= ~((c < win + 1) | (c > shape[1] - win - 1))
    return ok


def locate_fire_in_image_ref(fire_lat, fire_lon, lats, lons, win: int = P_ID_WIN_SIZE):
    """gaussian_profile.py:85-123 -> (fire_rows, fire_cols) lists of the fires that were found and kept."""
    rc = nearest_pixel_ref(fire_lat, fire_lon, lats, lons)
    ok = edge_filter_ref(rc, np.shape(lats), win)
    return rc[ok, 0].tolist(), rc[ok, 1].tolist()


def subset_fires_to_image_ref(lat, lon, fire_lat, fire_lon, fire_dt, date_to_find) -> np.ndarray:
    """gaussian_profile.py:46-54 -> indices of the fires of `date_to_find` strictly inside the image's extent."""
    fire_lat, fire_lon = np.asarray(fire_lat), np.asarray(fire_lon)
    m = np.asarray(fire_dt) == date_to_find
    m &= (fire_lat > np.min(lat)) & (fire_lat < np.max(lat)) & (fire_lon > np.min(lon)) & (fire_lon < np.max(lon))
    return np.flatnonzero(m)
