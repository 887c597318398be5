"""Synthetic swath geometries for the resampler tests (shared by the CPU oracle tests and the GPU parity tests)."""
import numpy as np


def swath(h, w, lat0, lon0, step_km=1.0, rot_deg=12.0, seed=0, jitter=0.0):
    """A rotated, slightly curved lat/lon grid of h x w pixels about `step_km` apart (like a MAIAC tile / a swath)."""
    rng = np.random.default_rng(seed)
    r, c = np.meshgrid(np.arange(h) - h / 2.0, np.arange(w) - w / 2.0, indexing="ij")
    th = np.radians(rot_deg)
    east = (c * np.cos(th) - r * np.sin(th)) * step_km
    north = -(c * np.sin(th) + r * np.cos(th)) * step_km
    north = north + 0.0004 * east * east / max(step_km, 1e-9)       # bow-tie-like curvature
    if jitter:
        east = east + rng.normal(0, jitter * step_km, east.shape)
        north = north + rng.normal(0, jitter * step_km, north.shape)
    lat = lat0 + north / 111.2
    lon = lon0 + east / (111.2 * np.cos(np.radians(lat)))
    return lat, lon
