"""CPU: the host side of the label-geometry path (kcl_ltss_bioatm_b200/labels.py) against the golden vectors
recorded from the reference's functions and against the oracle's polygon ordering."""
import os

import numpy as np
import pandas as pd
import pytest

from kcl_ltss_bioatm_b200 import labels
from oracle import hull_ref

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hull_cases.npz"))


def pattern_image(h, w):
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    return (((yy * 31 + xx * 17) % 1000) / 1000.0).astype(np.float32)


@pytest.mark.parametrize("i", range(int(G["n_mask_cases"])))
def test_convex_polygon_is_ccw_and_equals_oracle(i):
    hx, hy = G[f"c{i}_hull_x"], G[f"c{i}_hull_y"]
    poly = labels.convex_polygon(hx, hy)
    assert poly.dtype == np.int32
    assert np.array_equal(poly, hull_ref.convex_polygon(hx, hy))
    x, y = poly[:, 0].astype(np.int64), poly[:, 1].astype(np.int64)
    assert (x * np.roll(y, -1) - np.roll(x, -1) * y).sum() > 0                     # positive area = counter-clockwise
    perm = np.random.default_rng(i).permutation(len(hx))
    assert np.array_equal(labels.convex_polygon(hx[perm], hy[perm]), poly)


def test_convex_polygon_rejects_bad_hulls():
    with pytest.raises(ValueError):
        labels.convex_polygon([0, 1, 2], [0, 1, 2])             # collinear (the reference's Delaunay raises QhullError)
    with pytest.raises(ValueError):
        labels.convex_polygon([0, 1], [0, 1])
    with pytest.raises(ValueError):
        labels.convex_polygon([0, 1.5, 2], [0, 1, 0])           # not pixel coordinates
    with pytest.raises(ValueError):
        labels.convex_polygon([0, np.nan, 2], [0, 1, 0])


@pytest.mark.parametrize("i", range(int(G["n_subset_cases"])))
def test_subset_plume_matches_reference(i):
    k = f"s{i}"
    aod = pattern_image(300, 200)
    df = pd.DataFrame({"hull_x": G[k + "_hull_x"], "hull_y": G[k + "_hull_y"]})
    crop, sx, sy = labels.subset_plume(aod, df)
    assert tuple(crop.shape) == tuple(G[k + "_crop_shape"])
    assert crop.astype(np.float64).sum() == float(G[k + "_crop_sum"])
    assert np.array_equal(sx, G[k + "_shift_x"]) and np.array_equal(sy, G[k + "_shift_y"])


def test_subset_plume_nan():
    df = pd.DataFrame({"hull_x": [1.0, np.nan], "hull_y": [1.0, 2.0]})
    assert labels.subset_plume(pattern_image(50, 50), df) == (None, None, None)


def test_remove_duplicated_plumes_matches_reference():
    df = pd.DataFrame({"id": G["dedup_in_id"], "hull_lats": G["dedup_in_lat"], "hull_lons": G["dedup_in_lon"],
                       "hull_x": 1.0, "hull_y": 2.0, "datetime": G["dedup_in_dt"]})
    out = labels.remove_duplicated_plumes(df)
    assert np.array_equal(out["id"].to_numpy(), G["dedup_keep_id"])
    assert np.array_equal(out["datetime"].to_numpy().astype("U32"), G["dedup_keep_dt"])
    assert list(out.columns) == list(df.columns)


def test_pack_polygons_layout():
    hulls = [(G["c0_hull_x"], G["c0_hull_y"]), (G["c1_hull_x"], G["c1_hull_y"])]
    verts, offs, bbox = labels.pack_polygons(hulls, "cpu")
    assert verts.dtype == offs.dtype == bbox.dtype
    assert offs.tolist()[0] == 0 and offs.tolist()[-1] == verts.shape[0] and bbox.shape == (2, 4)
    p0 = verts[: offs[1]].numpy()
    assert bbox[0].tolist() == [p0[:, 0].min(), p0[:, 1].min(), p0[:, 0].max(), p0[:, 1].max()]
    v, o, b = labels.pack_polygons([], "cpu")
    assert v.shape == (0, 2) and o.tolist() == [0] and b.shape == (0, 4)


def test_reference_named_module_reexports():
    import src.features.plume_selector as ps

    assert ps.in_hull is labels.in_hull and ps.subset_plume is labels.subset_plume
    assert ps.find_plume_aod is labels.find_plume_aod and ps.remove_duplicated_plumes is labels.remove_duplicated_plumes
