"""CPU: host side of the fire-geolocation path (kcl_ltss_bioatm_b200/fires.py) against the golden vectors recorded
from the reference's functions.  The search itself needs the GPU (tests/test_gpu_fires.py)."""
import os

import numpy as np
import pandas as pd
import pytest

from kcl_ltss_bioatm_b200 import fires
from tests.grids import GRIDS

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "fire_cases.npz"))


def test_constants_match_reference():
    assert fires.P_ID_WIN_SIZE == int(G["p_id_win_size"])


@pytest.mark.parametrize("name", list(GRIDS))
def test_subset_fires_to_image_matches_reference(name):
    lat, lon = GRIDS[name]()
    df = pd.DataFrame({"latitude": G[name + "_fire_lat"], "longitude": G[name + "_fire_lon"],
                       "date_time": G[name + "_fire_dt"]})
    sub = fires.subset_fires_to_image(lat, lon, df, "t0")
    assert np.array_equal(sub.index.values, G[name + "_subset_index"])


@pytest.mark.parametrize("name", list(GRIDS))
def test_haversine_matches_reference_bitwise(name):
    lat, lon = GRIDS[name]()
    r, c = G[name + "_gen_rc"][:, 0], G[name + "_gen_rc"][:, 1]
    d = fires.haversine(G[name + "_fire_lon"][:32], G[name + "_fire_lat"][:32], lon[r, c], lat[r, c])
    assert np.array_equal(d, G[name + "_haversine32"])


def test_grid_indexes():
    rows, cols = fires.grid_indexes(np.zeros((3, 5)))
    assert rows.shape == cols.shape == (3, 5) and rows[2, 1] == 2 and cols[2, 1] == 1


def test_reference_named_module_reexports():
    import src.features.plume_identifier_gaussian_profile as ref_named

    assert ref_named.locate_fire_in_image is fires.locate_fire_in_image
    assert ref_named.haversine is fires.haversine and ref_named.P_ID_WIN_SIZE == 15


def test_no_cpu_path():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    df = pd.DataFrame({"latitude": [0.0], "longitude": [0.0]})
    with pytest.raises(Exception):
        fires.locate_fire_in_image(df, np.zeros((4, 4)), np.zeros((4, 4)))
