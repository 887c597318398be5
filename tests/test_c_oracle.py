"""CPU: the plain-C oracle (oracle/geo_ref.c) against the numpy oracles and the golden vectors recorded from the
reference's own functions -- two independent restatements must agree exactly."""
import os

import numpy as np
import pytest

from oracle import c_ref, fire_ref, hull_ref, sweep_ref
from tests.grids import GRIDS
from tests.sweep_data import synthetic_aod

GD = os.path.join(os.path.dirname(__file__), "golden")
GH, GF, GS = (np.load(os.path.join(GD, n)) for n in ("hull_cases.npz", "fire_cases.npz", "sweep_cases.npz"))


@pytest.mark.parametrize("i", range(int(GH["n_mask_cases"])))
def test_rasterize_equals_reference_golden(i):
    k = f"c{i}"
    h, w = (int(v) for v in GH[k + "_hw"])
    assert np.array_equal(c_ref.rasterize([(GH[k + "_hull_x"], GH[k + "_hull_y"])], h, w), GH[k + "_mask"])


def test_rasterize_window_origin_equals_numpy_oracle():
    hulls = [(GH["c0_hull_x"], GH["c0_hull_y"]), (GH["c1_hull_x"] + 20, GH["c1_hull_y"] + 5)]
    assert np.array_equal(c_ref.rasterize(hulls, 50, 70, origin=(7, 11)), hull_ref.rasterize_ref(hulls, 50, 70, origin=(7, 11)))


@pytest.mark.parametrize("name", list(GRIDS))
def test_nearest_pixels_equal_reference_golden(name):
    lat, lon = GRIDS[name]()
    rc = c_ref.nearest_pixels(GF[name + "_fire_lat"], GF[name + "_fire_lon"], lat, lon)
    assert np.array_equal(rc, fire_ref.nearest_pixel_ref(GF[name + "_fire_lat"], GF[name + "_fire_lon"], lat, lon))
    keep = fire_ref.edge_filter_ref(rc, lat.shape)
    assert np.array_equal(np.where(keep[:, None], rc, -1), GF[name + "_per_fire"])


@pytest.mark.parametrize("i", range(int(GS["n_cases"])))
def test_sweep_equals_reference_golden(i):
    k = f"c{i}"
    h, w, seed = (int(v) for v in GS[k + "_hws"])
    aod, fires = synthetic_aod(h, w, seed)
    thr = GS[k + "_thr"]
    masks = np.unpackbits(GS[k + "_masks"])[: len(thr) * h * w].reshape(len(thr), h, w)
    got = c_ref.threshold_masks(aod, thr)
    assert np.array_equal(got, masks)
    assert np.array_equal(c_ref.plume_extents(got, fires[:, 0], fires[:, 1]), GS[k + "_extents"])


def test_label8_canonical_labels_equal_numpy_oracle():
    rng = np.random.default_rng(4)
    m = rng.random((90, 130)) < 0.58
    labels, sizes = c_ref.label8(m)
    ref = sweep_ref.label_ref(m)
    assert np.array_equal(labels.astype(np.int64) + 1, ref)
    assert sizes.sum() == m.sum()
