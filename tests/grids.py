"""Deterministic lat/lon test grids shared by scripts/make_fire_golden.py and the fire-geolocation tests (the golden
file stores fires and results only; the grids are regenerated)."""
import numpy as np


def sinusoidal_grid(h, w, lat0, lon0, km=1.0):
    """lat/lon (float64 [h, w]) of a km-spaced sinusoidal-projection tile whose top-left is (lat0, lon0): rows of
    constant latitude, longitude spacing growing with 1/cos(lat) -- what MAIAC tiles look like in lat/lon."""
    r = np.arange(h)[:, None]
    c = np.arange(w)[None, :]
    lat = lat0 - r * (km / 111.195) + 0 * c
    lon = lon0 + c * (km / 111.195) / np.cos(np.radians(lat))
    return lat.astype(np.float64), lon.astype(np.float64)


def rotated_grid(h, w, lat0, lon0, deg=12.0, km=0.75):
    """A swath-like grid rotated against the meridians."""
    r, c = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    a = np.radians(deg)
    dy = -(r * np.cos(a) + c * np.sin(a)) * km / 111.195
    dx = (c * np.cos(a) - r * np.sin(a)) * km / 111.195
    lat = lat0 + dy
    return lat.astype(np.float64), (lon0 + dx / np.cos(np.radians(lat))).astype(np.float64)


GRIDS = {"sinu": lambda: sinusoidal_grid(240, 200, -10.0, 120.0), "rot": lambda: rotated_grid(150, 260, 3.0, -60.0)}
