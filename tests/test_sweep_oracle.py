"""CPU: pins oracle/sweep_ref.py against outputs of the REFERENCE's functions
(plume_identifier_gaussian_profile.py:142-240, scikit-image primitives replaced by scipy.ndimage stand-ins --
see scripts/make_sweep_golden.py) recorded in tests/golden/sweep_cases.npz."""
import os

import numpy as np
import pytest

from oracle import sweep_ref
from tests.sweep_data import synthetic_aod

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sweep_cases.npz"))
N = int(G["n_cases"])


def case(i):
    k = f"c{i}"
    h, w, seed = (int(v) for v in G[k + "_hws"])
    aod, fires = synthetic_aod(h, w, seed)
    assert np.array_equal(fires, G[k + "_fires"])
    thr = G[k + "_thr"]
    masks = np.unpackbits(G[k + "_masks"])[: len(thr) * h * w].reshape(len(thr), h, w).astype(bool)
    return aod, fires, thr, masks, G[k + "_extents"], G[k + "_index"]


def test_constants():
    assert int(G["p_id_win_size"]) == sweep_ref.P_ID_WIN_SIZE
    w = sweep_ref.P_ID_WIN_SIZE
    dy, dx = np.meshgrid(np.arange(-w, w + 1), np.arange(-w, w + 1), indexing="ij")
    assert np.array_equal(np.sqrt(dx ** 2 + dy ** 2), G["distance_matrix"])


@pytest.mark.parametrize("i", range(N))
def test_masks_extents_and_threshold_index_match_reference(i):
    aod, fires, thr, masks, extents, index = case(i)
    got = sweep_ref.threshold_masks_ref(aod, thr)
    assert np.array_equal(got, masks)
    ext = sweep_ref.find_plume_extents_ref(got, fires[:, 0], fires[:, 1])
    assert np.array_equal(ext, extents)
    idx = [-1 if v is None else v for v in sweep_ref.find_threshold_index_ref(ext)]
    assert idx == index.tolist()
    assert (extents > 0).any() and (extents == 0).any()


def test_find_threshold_index_tables():
    idx = [-1 if v is None else v for v in sweep_ref.find_threshold_index_ref(G["tables"])]
    assert idx == G["tables_index"].tolist()


def test_label_canonical_form():
    m = np.array([[1, 0, 0, 1], [0, 1, 0, 1], [0, 0, 0, 0], [1, 1, 0, 1]], dtype=bool)
    lab = sweep_ref.label_ref(m)
    assert lab.tolist() == [[1, 0, 0, 4], [0, 1, 0, 4], [0, 0, 0, 0], [13, 13, 0, 16]]   # diagonal joins (8-conn.)


GC = np.load(os.path.join(os.path.dirname(__file__), "golden", "cluster_cases.npz"))


@pytest.mark.parametrize("i", range(int(GC["n_cases"])))
def test_cluster_fires_oracle_equals_reference_golden(i):
    """oracle cluster_fires against the reference's own function (scripts/make_cluster_golden.py)."""
    from tests.sweep_data import synthetic_fire_pixels
    h, w, seed = (int(v) for v in GC[f"c{i}_hws"])
    rows, cols = synthetic_fire_pixels(h, w, seed)
    got = sweep_ref.cluster_fires_ref((h, w), rows, cols)
    assert got.dtype == np.int64 and np.array_equal(got, GC[f"c{i}_labels"])


def test_cluster_centroids_host_equals_oracle():
    from kcl_ltss_bioatm_b200 import sweep
    for i in range(int(GC["n_cases"])):
        lab = GC[f"c{i}_labels"].astype(np.int64)
        r, c = sweep.fire_cluster_centroids(lab)
        rr, cc = sweep_ref.cluster_centroids_ref(lab)
        assert np.array_equal(r, rr) and np.array_equal(c, cc) and r.dtype == rr.dtype


GF = np.load(os.path.join(os.path.dirname(__file__), "golden", "fill_cases.npz"))


@pytest.mark.parametrize("i", range(int(GF["n_cases"])))
def test_nearest_fill_oracle_equals_reference_golden(i):
    """oracle interpolate_aod_nearest against the reference's own function run with scipy (scripts/make_fill_golden.py):
    identical wherever the nearest valid pixel is unique; at distance ties scipy's kd-tree order decides, so there the
    golden value must be one of the tied candidates' values (the oracle takes the first in row-major order)."""
    from tests.sweep_data import synthetic_null_aod
    h, w, seed = (int(v) for v in GF[f"c{i}_hws"])
    aod = synthetic_null_aod(h, w, seed, np.dtype(str(GF[f"c{i}_dtype"])))
    got, unique = sweep_ref.interpolate_aod_nearest_ref(aod, return_unique=True)
    gold = GF[f"c{i}_filled"]
    assert int(GF["null_value"]) == sweep_ref.NULL_VALUE
    assert np.array_equal(got[unique].astype(np.float64), gold[unique])
    assert 0 < (~unique).sum() < 0.5 * unique.size and not (got == sweep_ref.NULL_VALUE).any()
    good = aod != sweep_ref.NULL_VALUE
    gy, gx = np.nonzero(good)
    for y, x in zip(*np.nonzero(~unique)):
        d2 = (gy - y) ** 2 + (gx - x) ** 2
        tied = aod[gy[d2 == d2.min()], gx[d2 == d2.min()]].astype(np.float64)
        assert gold[y, x] in tied and np.float64(got[y, x]) == tied[0]


def test_oracle_follows_scikit_image_documented_semantics():
    """The three scikit-image primitives are not installed here; tests/skimage_cases.py writes their documented behaviour
    down (docstring examples, border rules) and the oracle's restatements must reproduce it."""
    from tests import skimage_cases as C

    def rank_labels(mask):                      # label_ref's canonical labels -> scikit-image's 1..n raster-order numbers
        lab = sweep_ref.label_ref(mask)
        ids = np.unique(lab[lab > 0])
        out = np.zeros_like(lab)
        for k, v in enumerate(ids):
            out[lab == v] = k + 1
        return out

    assert np.array_equal(rank_labels(C.LABEL_EYE), C.LABEL_EYE_EXPECT)
    assert np.array_equal(rank_labels(C.LABEL_ORDER), C.LABEL_ORDER_EXPECT)
    # remove_small_objects(label(a), min_size, connectivity=2) as cluster_fires uses it, on the docstring's array
    for min_size, expect in ((7, C.RSO_MIN7_CONN2), (8, C.RSO_MIN8_CONN2)):
        r, c = np.nonzero(C.RSO)
        assert np.array_equal(sweep_ref.cluster_fires_ref(C.RSO.shape, r, c, min_size) > 0, expect)
    # border rules and the erosion -> dilation pair of generate_mask_dict
    for src, expect in ((C.ERODE_FULL, C.ERODE_FULL_EXPECT), (C.OPEN_IN, C.OPEN_EXPECT), (C.DILATE_DOT, np.zeros_like(C.DILATE_DOT))):
        assert np.array_equal(sweep_ref.threshold_masks_ref(src.astype(np.float32), [0.5])[0], expect)
    hole = sweep_ref.threshold_masks_ref(C.ERODE_HOLE.astype(np.float32), [0.5])[0]
    assert np.array_equal(hole, ndi_dilate_cross(C.ERODE_HOLE_EXPECT))


def ndi_dilate_cross(m):
    p = np.pad(m, 1, constant_values=False)
    return p[1:-1, 1:-1] | p[:-2, 1:-1] | p[2:, 1:-1] | p[1:-1, :-2] | p[1:-1, 2:]
