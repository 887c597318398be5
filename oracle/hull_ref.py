"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference's hull -> mask code, SURVEY.md section 8(f) rank 1.

PARITY PINNED: every function here is checked against the reference's own function, compiled unmodified from
``/root/reference/src/features/plume_selector.py`` by ``scripts/make_hull_golden.py`` (fixtures in
``tests/golden/hull_cases.npz``, test ``tests/test_hull_oracle.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and the cpu_baseline legs of the bench scripts may import this module.

The reference decides "pixel inside plume hull" with a Delaunay triangulation of the hull vertices and
``find_simplex(p) >= 0`` (plume_selector.py:88-98).  The union of the triangles of a point set's Delaunay
triangulation is its convex hull, and ``find_simplex`` accepts points on the boundary (barycentric coordinates
>= -eps), so for the integer pixel coordinates the reference uses (hull vertices are pixel indices,
plume_identifier_gaussian_profile.py:283-289) the test is: inside or on the boundary of the convex polygon.  Here
that is evaluated exactly in int64: the point is on the left of (or on) every edge of the counter-clockwise hull.
"""
from __future__ import annotations

import numpy as np


def convex_polygon(hull_x, hull_y) -> np.ndarray:
    """Counter-clockwise convex hull (int64 [m, 2] as (x, y), collinear points dropped) of the given vertices.
    The reference never orders the hull itself -- Delaunay takes the bare point set (plume_selector.py:96-97)."""
    pts = np.unique(np.stack([np.asarray(hull_x), np.asarray(hull_y)], 1).astype(np.int64), axis=0)  # sorted by x, y
    if len(pts) < 3:
        return pts

    def half(points):
        out = []
        for p in points:
            while len(out) >= 2:
                (ax, ay), (bx, by) = out[-2], out[-1]
                if (bx - ax) * (p[1] - ay) - (by - ay) * (p[0] - ax) <= 0:
                    out.pop()
                else:
                    break
            out.append((int(p[0]), int(p[1])))
        return out

    lower, upper = half(pts), half(pts[::-1])
    return np.array(lower[:-1] + upper[:-1], dtype=np.int64)


def in_hull_ref(p, hull) -> np.ndarray:
    """plume_selector.py:88-98.  p: [n, 2] integer (x, y) points; hull: [m, 2] vertices in any order."""
    p = np.asarray(p, dtype=np.int64)
    poly = convex_polygon(np.asarray(hull)[:, 0], np.asarray(hull)[:, 1])
    if len(poly) < 3:
        raise ValueError("degenerate hull (the reference's Delaunay raises QhullError here)")
    inside = np.ones(len(p), dtype=bool)
    for i in range(len(poly)):
        ax, ay = poly[i]
        bx, by = poly[(i + 1) % len(poly)]
        inside &= (bx - ax) * (p[:, 1] - ay) - (by - ay) * (p[:, 0] - ax) >= 0
    return inside


def rasterize_ref(hulls, h: int, w: int, origin=(0, 0)) -> np.ndarray:
    """uint8 [h, w]: 1 where pixel (x, y) = (origin_x + col, origin_y + row) is inside any hull.  One hull is the
    mask plume_selector.py:101-116 builds (pixel grid -> in_hull); the union over a scene's plumes is the label."""
    oy, ox = origin
    yy, xx = np.meshgrid(np.arange(h) + oy, np.arange(w) + ox, indexing="ij")
    pts = np.stack([xx.ravel(), yy.ravel()], 1)
    mask = np.zeros(h * w, dtype=bool)
    for hx, hy in hulls:
        mask |= in_hull_ref(pts, np.stack([hx, hy], 1))
    return mask.reshape(h, w).astype(np.uint8)


def find_plume_aod_ref(plume_image: np.ndarray, hull_x, hull_y) -> np.ndarray:
    """plume_selector.py:101-116: the image values at the pixels inside the hull.  (The reference builds its
    coordinate grid from swapped axes, which is only self-consistent for square crops; this restatement is the
    square-crop behaviour, in row-major pixel order -- compare sorted.)"""
    h, w = plume_image.shape
    m = rasterize_ref([(hull_x, hull_y)], h, w).astype(bool)
    return plume_image[m]


def subset_plume_ref(aod_shape, hull_x, hull_y, buffer: int = 40):
    """plume_selector.py:53-85.  Returns ((y0, y1, x0, x1), shifted hull_x, shifted hull_y): the crop window of
    the AOD image around the hull's bounding box grown by `buffer` pixels (clipped at the image edges) and the hull
    in the crop's coordinates; None if the hull holds NaNs."""
    hull_x = np.asarray(hull_x, dtype=np.float64)
    hull_y = np.asarray(hull_y, dtype=np.float64)
    min_x, max_x, min_y, max_y = hull_x.min(), hull_x.max(), hull_y.min(), hull_y.max()
    if min_x - buffer < 0:
        min_x = 0
    else:
        hull_x = hull_x - min_x + buffer
        min_x = min_x - buffer
    if min_y - buffer < 0:
        min_y = 0
    else:
        hull_y = hull_y - min_y + buffer
        min_y = min_y - buffer
    max_x = aod_shape[1] if max_x + buffer > aod_shape[1] else max_x + buffer
    max_y = aod_shape[0] if max_y + buffer > aod_shape[0] else max_y + buffer
    if np.isnan([min_y, max_y, min_x, max_x]).any():
        return None
    return (int(min_y), int(max_y), int(min_x), int(max_x)), hull_x, hull_y


def remove_duplicated_plumes_ref(ids, lats, lons, datetimes) -> np.ndarray:
    """plume_selector.py:26-49.  Row mask of the hull table that survives: plumes are keyed by (id, datetime); per
    datetime, a plume whose centroid (mean hull lat / lon, rounded to 3 decimals) repeats an earlier plume's (in
    (id, datetime) order) is dropped with all its rows."""
    ids, lats, lons = np.asarray(ids), np.asarray(lats, dtype=np.float64), np.asarray(lons, dtype=np.float64)
    uniq_dt = list(dict.fromkeys(np.asarray(datetimes).tolist()))
    dt_idx = np.array([uniq_dt.index(d) for d in np.asarray(datetimes).tolist()])
    keys = sorted(set(zip(ids.tolist(), dt_idx.tolist())))
    seen, keep_keys = set(), set()
    for pid, di in keys:
        sel = (ids == pid) & (dt_idx == di)
        c = (di, float(np.round(lats[sel].mean(), 3)), float(np.round(lons[sel].mean(), 3)))
        if c not in seen:
            seen.add(c)
            keep_keys.add((pid, di))
    return np.array([(i, d) in keep_keys for i, d in zip(ids.tolist(), dt_idx.tolist())])
