"""GPU: plume_locate_fires through the C ABI against the oracle and the golden vectors recorded from the
reference's locate_fire_in_image (index work: exact)."""
import os

import numpy as np
import pandas as pd
import pytest

from kcl_ltss_bioatm_b200 import fires
from oracle import fire_ref
from tests.grids import GRIDS, sinusoidal_grid

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "fire_cases.npz"))


@pytest.mark.parametrize("name", list(GRIDS))
def test_locate_fire_in_image_equals_reference_golden(name):
    lat, lon = GRIDS[name]()
    df = pd.DataFrame({"latitude": G[name + "_fire_lat"], "longitude": G[name + "_fire_lon"],
                       "date_time": G[name + "_fire_dt"]})
    sub = fires.subset_fires_to_image(lat, lon, df, "t0")
    rows_g, cols_g = fires.grid_indexes(lat)
    rows, cols = fires.locate_fire_in_image(sub, lat, lon, rows_g, cols_g)
    assert rows == G[name + "_rows"].tolist() and cols == G[name + "_cols"].tolist()
    loc = fires.FireLocator(lat, lon)
    rows, cols = loc.locate(sub.latitude.values, sub.longitude.values, win=-10 ** 6)
    assert rows == G[name + "_rows_nofilter"].tolist() and cols == G[name + "_cols_nofilter"].tolist()
    # every fire, including those outside the image and those the edge filter drops
    rc = loc.nearest_pixels(df.latitude.values, df.longitude.values)
    keep = fire_ref.edge_filter_ref(rc.astype(np.int64), lat.shape)
    assert np.array_equal(np.where(keep[:, None], rc, -1), G[name + "_per_fire"])


@pytest.mark.parametrize("h,w,n", [(1, 1, 5), (37, 53, 70), (300, 411, 257), (1200, 1200, 500)])
def test_nearest_pixels_equal_oracle(h, w, n):
    lat, lon = sinusoidal_grid(h, w, 35.0 + h * 0.001, -100.0)
    rng = np.random.default_rng(h + w + n)
    flat = rng.uniform(lat.min() - 0.1, lat.max() + 0.1, n)
    flon = rng.uniform(lon.min() - 0.1, lon.max() + 0.1, n)
    flat[: n // 4] = lat.ravel()[rng.integers(0, h * w, n // 4)]            # exactly on pixel centres
    flon[: n // 4] = lon.ravel()[rng.integers(0, h * w, n // 4)]
    got = fires.FireLocator(lat, lon).nearest_pixels(flat, flon)
    ref = fire_ref.nearest_pixel_ref(flat, flon, lat, lon)
    assert got.dtype == np.int32 and np.array_equal(got, ref)
    assert (ref[:, 0] >= 0).any() or h * w == 1


def test_ties_take_the_first_pixel_in_row_major_order():
    lat = np.array([[0.0, 0.0], [0.0, 0.0]])                # four pixels at the same place: all distances tie
    lon = np.array([[0.0, 0.0], [0.0, 0.0]])
    got = fires.FireLocator(lat, lon).nearest_pixels([0.01], [0.01])
    assert got.tolist() == [[0, 0]]
    lat = np.array([[0.02, 0.0], [0.0, 0.02]])              # symmetric pair (0,1) / (1,0): first wins
    lon = np.array([[0.03, 0.0], [0.0, 0.03]])
    assert fires.FireLocator(lat, lon).nearest_pixels([0.0], [0.0]).tolist() == [[0, 1]]


def test_no_fires():
    lat, lon = sinusoidal_grid(8, 8, 0.0, 0.0)
    assert fires.FireLocator(lat, lon).nearest_pixels([], []).shape == (0, 2)
