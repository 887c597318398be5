"""CPU: pins oracle/hull_ref.py against outputs of the REFERENCE's own functions (plume_selector.py:26-116),
recorded by scripts/make_hull_golden.py into tests/golden/hull_cases.npz."""
import os

import numpy as np
import pytest

from oracle import hull_ref

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hull_cases.npz"))


def pattern_image(h, w):
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    return (((yy * 31 + xx * 17) % 1000) / 1000.0).astype(np.float32)


@pytest.mark.parametrize("i", range(int(G["n_mask_cases"])))
def test_mask_matches_reference_in_hull(i):
    k = f"c{i}"
    h, w = G[k + "_hw"]
    hx, hy = G[k + "_hull_x"], G[k + "_hull_y"]
    mask = hull_ref.rasterize_ref([(hx, hy)], int(h), int(w))
    assert mask.dtype == np.uint8 and np.array_equal(mask, G[k + "_mask"])     # bit-exact, boundary pixels included
    assert mask.sum() > 0
    aod = hull_ref.find_plume_aod_ref(pattern_image(h, w), hx, hy)
    assert np.array_equal(np.sort(aod), G[k + "_aod_sorted"])


def test_hull_vertex_order_does_not_matter():
    hx, hy = G["c0_hull_x"], G["c0_hull_y"]
    perm = np.random.default_rng(0).permutation(len(hx))
    a = hull_ref.rasterize_ref([(hx, hy)], 64, 64)
    b = hull_ref.rasterize_ref([(hx[perm], hy[perm])], 64, 64)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("i", range(int(G["n_subset_cases"])))
def test_subset_plume_matches_reference(i):
    k = f"s{i}"
    aod = pattern_image(300, 200)
    (y0, y1, x0, x1), sx, sy = hull_ref.subset_plume_ref(aod.shape, G[k + "_hull_x"], G[k + "_hull_y"])
    crop = aod[y0:y1, x0:x1]
    assert tuple(crop.shape) == tuple(G[k + "_crop_shape"])
    assert crop.astype(np.float64).sum() == float(G[k + "_crop_sum"])
    assert crop[0, 0] == G[k + "_crop_corner"][0] and crop[-1, -1] == G[k + "_crop_corner"][1]
    assert np.array_equal(sx, G[k + "_shift_x"]) and np.array_equal(sy, G[k + "_shift_y"])


def test_subset_plume_nan_hull_is_rejected():
    assert hull_ref.subset_plume_ref((100, 100), [1.0, np.nan, 3.0], [1.0, 2.0, 3.0]) is None


def test_remove_duplicated_plumes_matches_reference():
    keep = hull_ref.remove_duplicated_plumes_ref(G["dedup_in_id"], G["dedup_in_lat"], G["dedup_in_lon"], G["dedup_in_dt"])
    assert np.array_equal(G["dedup_in_id"][keep], G["dedup_keep_id"])
    assert np.array_equal(G["dedup_in_dt"][keep], G["dedup_keep_dt"])
    assert 0 < keep.sum() < len(keep)


def test_degenerate_hull_raises_like_qhull():
    with pytest.raises(ValueError):
        hull_ref.in_hull_ref(np.zeros((1, 2)), np.array([[0, 0], [1, 1], [2, 2]]))
