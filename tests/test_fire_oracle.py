"""CPU: pins oracle/fire_ref.py against outputs of the REFERENCE's own functions
(plume_identifier_gaussian_profile.py:46-123) recorded by scripts/make_fire_golden.py."""
import os

import numpy as np
import pytest

from oracle import fire_ref
from tests.grids import GRIDS

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "fire_cases.npz"))


def test_constant():
    assert int(G["p_id_win_size"]) == fire_ref.P_ID_WIN_SIZE


@pytest.mark.parametrize("name", list(GRIDS))
def test_locate_fire_in_image_matches_reference(name):
    lat, lon = GRIDS[name]()
    flat, flon, dt = G[name + "_fire_lat"], G[name + "_fire_lon"], G[name + "_fire_dt"]
    sub = fire_ref.subset_fires_to_image_ref(lat, lon, flat, flon, dt, "t0")
    assert np.array_equal(sub, G[name + "_subset_index"])
    rows, cols = fire_ref.locate_fire_in_image_ref(flat[sub], flon[sub], lat, lon)
    assert rows == G[name + "_rows"].tolist() and cols == G[name + "_cols"].tolist()
    rows, cols = fire_ref.locate_fire_in_image_ref(flat[sub], flon[sub], lat, lon, win=-10 ** 6)
    assert rows == G[name + "_rows_nofilter"].tolist() and cols == G[name + "_cols_nofilter"].tolist()


@pytest.mark.parametrize("name", list(GRIDS))
def test_per_fire_results_including_skipped_fires(name):
    lat, lon = GRIDS[name]()
    flat, flon = G[name + "_fire_lat"], G[name + "_fire_lon"]
    rc = fire_ref.nearest_pixel_ref(flat, flon, lat, lon)
    ok = fire_ref.edge_filter_ref(rc, lat.shape)
    got = np.where(ok[:, None], rc, -1)
    assert np.array_equal(got, G[name + "_per_fire"])
    assert (G[name + "_per_fire"][:, 0] < 0).sum() > 20          # the fixture does exercise skipped fires


@pytest.mark.parametrize("name", list(GRIDS))
def test_haversine_matches_reference_bitwise(name):
    lat, lon = GRIDS[name]()
    flat, flon = G[name + "_fire_lat"][:32], G[name + "_fire_lon"][:32]
    r, c = G[name + "_gen_rc"][:, 0], G[name + "_gen_rc"][:, 1]
    d = fire_ref.haversine_ref(flon, flat, lon[r, c], lat[r, c])
    assert np.array_equal(d, G[name + "_haversine32"])           # same numpy operations in the same order
