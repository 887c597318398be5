"""TEST INFRASTRUCTURE -- ctypes access to the plain-C oracle (oracle/geo_ref.c, built by oracle/Makefile).  Only
``tests/`` and the cpu_baseline legs of the bench scripts may import this module."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from oracle.hull_ref import convex_polygon

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libgeo_ref.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "geo_ref.c")):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def rasterize(hulls, h: int, w: int, origin=(0, 0)) -> np.ndarray:
    mask = np.zeros((h, w), dtype=np.uint8)
    for hx, hy in hulls:
        poly = np.ascontiguousarray(convex_polygon(hx, hy), dtype=np.int64)
        vx, vy = np.ascontiguousarray(poly[:, 0]), np.ascontiguousarray(poly[:, 1])
        lib().geo_rasterize_hull(_p(vx, ctypes.c_longlong), _p(vy, ctypes.c_longlong), len(poly), h, w,
                                 int(origin[0]), int(origin[1]), _p(mask, ctypes.c_ubyte))
    return mask


def nearest_pixels(fire_lat, fire_lon, lats, lons, half: float = 0.05) -> np.ndarray:
    lats = np.ascontiguousarray(lats, dtype=np.float64)
    lons = np.ascontiguousarray(lons, dtype=np.float64)
    fl = np.ascontiguousarray(fire_lat, dtype=np.float64)
    fo = np.ascontiguousarray(fire_lon, dtype=np.float64)
    out = np.empty((len(fl), 2), dtype=np.int64)
    lib().geo_locate_fires(_p(lats, ctypes.c_double), _p(lons, ctypes.c_double), lats.shape[0], lats.shape[1],
                           _p(fl, ctypes.c_double), _p(fo, ctypes.c_double), len(fl), ctypes.c_double(half),
                           _p(out, ctypes.c_longlong))
    return out


def threshold_masks(aod, thresholds) -> np.ndarray:
    aod = np.ascontiguousarray(aod, dtype=np.float32)
    h, w = aod.shape
    out = np.empty((len(thresholds), h, w), dtype=np.uint8)
    tmp = np.empty((h, w), dtype=np.uint8)
    for k, t in enumerate(thresholds):
        lib().geo_threshold_mask(_p(aod, ctypes.c_float), h, w, ctypes.c_double(float(t)), _p(tmp, ctypes.c_ubyte),
                                 _p(out[k], ctypes.c_ubyte))
    return out


def label8(mask):
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    h, w = mask.shape
    labels = np.empty((h, w), dtype=np.int32)
    sizes = np.empty((h, w), dtype=np.int32)
    lib().geo_label8(_p(mask, ctypes.c_ubyte), h, w, _p(labels, ctypes.c_int), _p(sizes, ctypes.c_int))
    return labels, sizes


def plume_extents(masks, fire_rows, fire_cols, win: int = 15) -> np.ndarray:
    rc = np.ascontiguousarray(np.stack([fire_rows, fire_cols], 1), dtype=np.int64)
    out = np.zeros((len(masks), len(rc)), dtype=np.float64)
    ext = np.empty(len(rc), dtype=np.int32)
    for k, m in enumerate(masks):
        labels, sizes = label8(m)
        lib().geo_fire_extents(_p(labels, ctypes.c_int), _p(sizes, ctypes.c_int), m.shape[0], m.shape[1],
                               _p(rc, ctypes.c_longlong), len(rc), int(win), _p(ext, ctypes.c_int))
        out[k] = ext
    return out
