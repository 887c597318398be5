"""GPU: threshold masks, connected components and plume extents through the C ABI against the oracle and the
golden vectors recorded from the reference's functions (boolean / integer work: exact)."""
import os

import numpy as np
import pytest
import torch

from kcl_ltss_bioatm_b200 import sweep
from oracle import sweep_ref
from tests.sweep_data import synthetic_aod

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sweep_cases.npz"))
N = int(G["n_cases"])


@pytest.fixture(scope="module")
def sw():
    return sweep.ThresholdSweep("cuda:0")


def case(i):
    k = f"c{i}"
    h, w, seed = (int(v) for v in G[k + "_hws"])
    aod, fires = synthetic_aod(h, w, seed)
    thr = G[k + "_thr"]
    masks = np.unpackbits(G[k + "_masks"])[: len(thr) * h * w].reshape(len(thr), h, w).astype(bool)
    return aod, fires, thr, masks, G[k + "_extents"], G[k + "_index"]


@pytest.mark.parametrize("i", range(N))
def test_sweep_equals_reference_golden(sw, i):
    aod, fires, thr, masks, extents, index = case(i)
    m = sw.masks(aod, thr)
    assert np.array_equal(m.cpu().numpy().astype(bool), masks)
    assert np.array_equal(sw.unpack_bits(sw.mask_bits(aod, thr), aod.shape[1]), masks)      # bit planes
    ext = sw.extents(aod, thr, fires[:, 0], fires[:, 1])                                    # fused bit-plane path
    assert ext.dtype == np.float64 and np.array_equal(ext, extents)
    assert np.array_equal(sw.extents_dense(m, fires[:, 0], fires[:, 1]), extents)           # dense label planes
    assert np.array_equal(sw.extents_of_masks(m, fires[:, 0], fires[:, 1]), extents)        # byte masks -> bit planes
    idx = [-1 if v is None else v for v in sweep.find_threshold_index(ext)]
    assert idx == index.tolist()


def test_reference_named_functions(sw):
    aod, fires, thr, masks, extents, _ = case(0)
    d = sweep.generate_mask_dict(aod, thr)
    assert isinstance(d, dict) and len(d) == len(thr) and list(d.keys()) == list(thr) and list(d) == list(thr)
    assert np.array_equal(sweep.find_plume_extents(d, fires[:, 0], fires[:, 1]), extents)      # resident bit planes
    assert all(v is None for v in dict.values(d))                                              # ... nothing was fetched
    assert all(np.array_equal(d[t], masks[k]) and d[t].dtype == bool for k, t in enumerate(thr))
    assert np.array_equal(np.stack(d.values()), masks) and [k for k, _ in d.items()] == list(thr)
    assert np.array_equal(sweep.find_plume_extents(d, fires[:, 0], fires[:, 1]), extents)
    plain = dict(d.items())                                                                    # an ordinary dict of host masks
    assert np.array_equal(sweep.find_plume_extents(plain, fires[:, 0], fires[:, 1]), extents)
    d[thr[3]] = np.zeros_like(masks[3])                                                        # a host edit is honoured
    edited = sweep.find_plume_extents(d, fires[:, 0], fires[:, 1])
    assert not edited[3].any() and np.array_equal(np.delete(edited, 3, 0), np.delete(extents, 3, 0))
    dup = sweep.generate_mask_dict(aod, [0.3, 0.1, 0.3])                                       # equal thresholds: one key
    assert list(dup) == [0.3, 0.1] and sweep.find_plume_extents(dup, fires[:, 0], fires[:, 1]).shape == (2, len(fires))


@pytest.mark.parametrize("h,w,density", [(1, 1, 1.0), (7, 300, 0.5), (64, 64, 0.62), (257, 129, 0.4), (600, 800, 0.55)])
def test_components_equal_oracle_on_random_masks(sw, h, w, density):
    """Percolation-like random masks: long winding components, many merges across block boundaries."""
    rng = np.random.default_rng(h * 1000 + w)
    masks = (rng.random((3, h, w)) < density)
    masks[1] = ~masks[1] if h > 1 else masks[1]
    if h >= 8 and w >= 8:                       # a spiral: the worst case for label propagation
        sp = np.zeros((h, w), dtype=bool)
        y0, x0, y1, x1 = 0, 0, h - 1, w - 1
        while y1 - y0 > 3 and x1 - x0 > 3:
            sp[y0, x0:x1 + 1] = True
            sp[y0:y1 + 1, x1] = True
            sp[y1, x0 + 2:x1 + 1] = True
            sp[y0 + 2:y1 + 1, x0 + 2] = True
            y0, x0, y1, x1 = y0 + 2, x0 + 2, y1 - 2, x1 - 2
        masks[2] = sp
    labels, sizes = sw.label(torch.from_numpy(masks.astype(np.uint8)).cuda())
    labels, sizes = labels.cpu().numpy(), sizes.cpu().numpy()
    for t in range(3):
        ref = sweep_ref.label_ref(masks[t])               # 0 background, else 1 + smallest index of the component
        assert np.array_equal(labels[t].astype(np.int64) + 1, ref)
        cnt = np.bincount(ref.ravel(), minlength=h * w + 1)[1:].reshape(h, w)
        roots = (ref.ravel() == np.arange(1, h * w + 1)).reshape(h, w)
        assert np.array_equal(sizes[t], np.where(roots, cnt, 0))


def test_masks_border_rules_and_float64_compare(sw):
    aod = np.full((6, 7), 1.0, dtype=np.float32)          # everything set: erosion must not eat the border
    assert sw.masks(aod, [0.5]).cpu().numpy().all()
    aod[:] = 0.0
    aod[2, 3] = 1.0                                       # a singleton disappears
    assert not sw.masks(aod, [0.5]).cpu().numpy().any()
    x = np.float32(0.48)                                  # float32(0.48) > 0.48 in float64, as numpy compares
    aod[:] = x
    assert sw.masks(aod, [0.48]).cpu().numpy().all() == bool(np.float64(x) > 0.48)
    ref = sweep_ref.threshold_masks_ref(aod, [0.48])
    assert np.array_equal(sw.masks(aod, [0.48]).cpu().numpy().astype(bool), ref)


def test_fire_too_close_to_edge_is_rejected(sw):
    aod, _, thr, _, _, _ = case(0)
    with pytest.raises(ValueError):
        sw.extents(aod, thr, [3], [40])


def spiral(h, w):
    sp = np.zeros((h, w), dtype=bool)
    y0, x0, y1, x1 = 0, 0, h - 1, w - 1
    while y1 - y0 > 3 and x1 - x0 > 3:
        sp[y0, x0:x1 + 1] = True
        sp[y0:y1 + 1, x1] = True
        sp[y1, x0 + 2:x1 + 1] = True
        sp[y0 + 2:y1 + 1, x0 + 2] = True
        y0, x0, y1, x1 = y0 + 2, x0 + 2, y1 - 2, x1 - 2
    return sp


@pytest.mark.parametrize("h,w,density", [(1, 1, 1.0), (7, 300, 0.5), (64, 64, 0.62), (97, 129, 0.4), (257, 333, 0.55),
                                         (600, 800, 0.6)])
def test_bit_plane_extents_equal_oracle_on_random_masks(sw, h, w, density):
    """The run-based union-find on percolation-like masks, a spiral, a full and an empty plane."""
    rng = np.random.default_rng(h * 1000 + w)
    masks = rng.random((5, h, w)) < density
    masks[1] = ~masks[1] if h > 1 else masks[1]
    masks[3] = True
    masks[4] = False
    if h >= 8 and w >= 8:
        masks[2] = spiral(h, w)
    win = min(sweep_ref.P_ID_WIN_SIZE, (h - 1) // 2, (w - 1) // 2)
    rows = rng.integers(win, h - win, 40)
    cols = rng.integers(win, w - win, 40)
    ref = sweep_ref.find_plume_extents_ref(masks, rows, cols, win)
    dev = torch.from_numpy(masks.astype(np.uint8)).cuda()
    for _ in range(3):                                                     # atomics: every run must agree
        assert np.array_equal(sw.extents_of_masks(dev, rows, cols, win), ref)
    assert np.array_equal(sw.extents_of_bits(sw.pack_bits_host(masks), w, rows, cols, win), ref)
    assert np.array_equal(sw.extents_dense(dev, rows, cols, win), ref)


@pytest.mark.parametrize("h,w", [(1, 1), (2, 31), (3, 32), (5, 33), (7, 300), (40, 64), (33, 65), (61, 97), (300, 1200)])
def test_mask_bits_equal_oracle_on_noisy_images(sw, h, w):
    rng = np.random.default_rng(h * 131 + w)
    aod = rng.random((h, w)).astype(np.float32)
    aod[rng.random((h, w)) < 0.02] = np.nan
    thr = np.concatenate([[0.1, 0.25, 0.5, float(np.float32(0.3)), 0.3, 0.9, -1.0, 2.0, 0.25, np.nan, np.inf, -np.inf],
                          rng.random(29)])                                   # two chunks, unsorted, duplicates, non-finite
    with np.errstate(invalid="ignore"):
        ref = sweep_ref.threshold_masks_ref(aod, thr)
    assert np.array_equal(sw.unpack_bits(sw.mask_bits(aod, thr), w), ref)
    assert np.array_equal(sw.masks(aod, thr[:40]).cpu().numpy().astype(bool), ref[:40])
    blobs = (rng.random((h, w)) < 0.8).astype(np.float32)
    assert np.array_equal(sw.unpack_bits(sw.mask_bits(blobs, [0.5]), w), sweep_ref.threshold_masks_ref(blobs, [0.5]))


def test_threshold_rounding_decides_like_float64(sw):
    x = np.float32(0.48)
    for t in (0.48, float(x), float(np.nextafter(x, np.float32(1))), float(np.nextafter(x, np.float32(0))), 1e-50, -1e-50, 1e300):
        aod = np.full((6, 7), x, dtype=np.float32)
        want = bool(np.float64(x) > t)
        assert sw.unpack_bits(sw.mask_bits(aod, [t]), 7).all() == want
        assert bool(sw.masks(aod, [t]).cpu().numpy().all()) == want


def test_full_size_timestamp_bit_planes_agree_with_dense_planes(sw):
    """1200 x 1200, the reference's three sweeps (75 thresholds) in ONE call against the dense-plane path per sweep."""
    aod, _ = synthetic_aod(1200, 1200, 5)
    rng = np.random.default_rng(3)
    rows, cols = rng.integers(16, 1184, 64), rng.integers(16, 1184, 64)
    sweeps = [np.abs(np.arange(0, tmax, step) - tmax) for step, tmax in [(0.02, 0.5), (0.03, 0.75), (0.04, 1)]]
    fused = sw.extents(aod, np.concatenate(sweeps), rows, cols)
    dense = np.concatenate([sw.extents_dense(sw.masks(aod, t), rows, cols) for t in sweeps])
    assert fused.shape == (75, 64) and np.array_equal(fused, dense)
    assert np.array_equal(np.concatenate([sw.extents(aod, t, rows, cols) for t in sweeps]), fused)
    assert fused.max() > 1000 and (fused == 0).any()


GC = np.load(os.path.join(os.path.dirname(__file__), "golden", "cluster_cases.npz"))


@pytest.mark.parametrize("i", range(int(GC["n_cases"])))
def test_cluster_fires_equals_reference_golden(sw, i):
    from tests.sweep_data import synthetic_fire_pixels
    h, w, seed = (int(v) for v in GC[f"c{i}_hws"])
    rows, cols = synthetic_fire_pixels(h, w, seed)
    got = sw.cluster_fires((h, w), rows, cols)
    assert got.dtype == np.int64 and np.array_equal(got, GC[f"c{i}_labels"])
    assert np.array_equal(got, sweep_ref.cluster_fires_ref((h, w), rows, cols))
    assert not sw.cluster_fires((h, w), [], []).any()
    r, c = sweep.fire_cluster_centroids(got)
    rr, cc = sweep_ref.cluster_centroids_ref(got)
    assert np.array_equal(r, rr) and np.array_equal(c, cc)


def test_float64_image_is_compared_in_float64(sw):
    """MAIAC AOD as the reference reads it: int16 * 0.001 in float64 (tools.py:88), thresholds from np.arange -- the
    image values sit exactly on or next to the thresholds, where a float32 copy of the image decides differently."""
    rng = np.random.default_rng(11)
    raw = rng.integers(0, 1000, (90, 140)).astype(np.int16)
    raw[20:60, 30:100] = rng.integers(470, 490, (40, 70))                 # a plateau straddling 0.48
    aod = raw * 0.001
    assert aod.dtype == np.float64
    thr = np.abs(np.arange(0, 0.5, 0.02) - 0.5)                           # gaussian_profile.py:492
    ref = sweep_ref.threshold_masks_ref(aod, thr)
    assert not np.array_equal(sweep_ref.threshold_masks_ref(aod.astype(np.float32), thr), ref)   # float32 would differ
    assert np.array_equal(sw.unpack_bits(sw.mask_bits(aod, thr), 140), ref)
    assert np.array_equal(sw.masks(aod, thr).cpu().numpy().astype(bool), ref)
    rows, cols = rng.integers(16, 74, 9), rng.integers(16, 124, 9)
    assert np.array_equal(sw.extents(aod, thr, rows, cols), sweep_ref.find_plume_extents_ref(ref, rows, cols))
    d = sweep.generate_mask_dict(aod, thr)
    assert all(np.array_equal(d[t], ref[k]) for k, t in enumerate(thr))


GF = np.load(os.path.join(os.path.dirname(__file__), "golden", "fill_cases.npz"))


@pytest.mark.parametrize("i", range(int(GF["n_cases"])))
def test_nearest_fill_equals_oracle_and_reference_golden(sw, i):
    from tests.sweep_data import synthetic_null_aod
    h, w, seed = (int(v) for v in GF[f"c{i}_hws"])
    aod = synthetic_null_aod(h, w, seed, np.dtype(str(GF[f"c{i}_dtype"])))
    ref, unique = sweep_ref.interpolate_aod_nearest_ref(aod, return_unique=True)
    got = sw.fill_nearest(aod)
    assert got.dtype == (torch.float64 if aod.dtype == np.float64 else torch.float32)
    assert np.array_equal(got.cpu().numpy(), ref)                                   # every pixel, ties included
    named = sweep.interpolate_aod_nearest(aod)
    assert named.dtype == np.float64 and np.array_equal(named[unique], GF[f"c{i}_filled"][unique])


def test_nearest_fill_edge_cases(sw):
    aod = np.full((5, 70), -999.0)
    with pytest.raises(ValueError):
        sw.fill_nearest(aod)
    aod[3, 40] = 0.25                                                     # a single valid pixel fills everything
    assert (sw.fill_nearest(aod).cpu().numpy() == 0.25).all()
    aod = np.arange(12, dtype=np.float32).reshape(3, 4)                   # nothing to fill
    assert np.array_equal(sw.fill_nearest(aod).cpu().numpy(), aod)
    aod = np.full((1, 1), 0.5)
    assert sw.fill_nearest(aod).item() == 0.5
    aod = np.full((4, 33), -999.0)
    aod[0, 0], aod[3, 32] = 1.0, np.nan                                   # NaN != -999: a valid pixel, as in numpy
    out = sw.fill_nearest(aod).cpu().numpy()
    assert np.array_equal(out, sweep_ref.interpolate_aod_nearest_ref(aod), equal_nan=True) and np.isnan(out).any()


def test_nearest_fill_full_size_properties(sw):
    """1200 x 1200 (too large for the brute-force oracle): valid pixels unchanged, idempotent, every filled value is
    the value of a valid pixel at the distance scipy's exact Euclidean distance transform reports."""
    import scipy.ndimage as ndi
    from tests.sweep_data import synthetic_null_aod
    aod = synthetic_null_aod(1200, 1200, 9)
    good = aod != -999
    out = sw.fill_nearest(aod)
    o = out.cpu().numpy()
    assert np.array_equal(o[good], aod[good]) and not (o == -999).any()
    assert np.array_equal(sw.fill_nearest(out).cpu().numpy(), o)
    dist, idx = ndi.distance_transform_edt(~good, return_indices=True)
    same = o == aod[idx[0], idx[1]]                                       # differs only where ties were broken differently
    assert same.mean() > 0.9
    yy, xx = np.nonzero(~same)
    for y, x in list(zip(yy, xx))[:200]:                                  # at a tie: some valid pixel at exactly that distance holds our value
        r = int(np.ceil(dist[y, x]))
        y0, x0 = max(0, y - r), max(0, x - r)
        win = aod[y0:y + r + 1, x0:x + r + 1]
        wy, wx = np.nonzero(win == o[y, x])
        assert (np.round(dist[y, x] ** 2) == (wy + y0 - y) ** 2 + (wx + x0 - x) ** 2).any()


@pytest.mark.parametrize("h,w,density", [(7, 300, 0.5), (64, 64, 0.62), (97, 129, 0.4), (257, 333, 0.55), (600, 800, 0.6)])
def test_fire_components_equal_oracle_on_random_masks(sw, h, w, density):
    """find_plume_mask's label / extract_label / == without a label plane: component masks and bounding boxes."""
    rng = np.random.default_rng(h * 7 + w)
    masks = rng.random((4, h, w)) < density
    masks[1] = ~masks[1]
    masks[3] = False
    if h >= 8:
        masks[2] = spiral(h, w)
    win = min(sweep_ref.P_ID_WIN_SIZE, (h - 1) // 2, (w - 1) // 2)
    n = 24
    rows, cols = rng.integers(win, h - win, n), rng.integers(win, w - win, n)
    planes = [None if p < 0 else int(p) for p in rng.integers(-1, 4, n)]
    bits = sw.pack_bits_host(masks)
    crops, stats = sw.fire_components(bits, w, planes, rows, cols, win)
    for f, p in enumerate(planes):
        ref = None if p is None else sweep_ref.plume_mask_ref(masks[p], rows[f], cols[f], win)
        if ref is None:
            assert crops[f] is None and stats[f, 0] == 0 and stats[f, 5] == -1
        else:
            ys, xs = np.nonzero(ref)
            assert stats[f, :5].tolist() == [ref.sum(), ys.min(), xs.min(), ys.max() + 1, xs.max() + 1]
            assert np.array_equal(crops[f], ref[ys.min():ys.max() + 1, xs.min():xs.max() + 1])
            assert np.array_equal(sw.full_mask(crops[f], stats[f], ref.shape), ref)


@pytest.mark.parametrize("i", range(N))
def test_plume_masks_follow_the_reference_call_sequence(sw, i):
    """generate_mask_dict -> find_plume_extents -> find_threshold_index -> plume_masks on the golden cases: the mask
    of every fire with a threshold index equals label / extract_label / == of the oracle on the reference's own mask."""
    aod, fires, thr, masks, extents, index = case(i)
    d = sweep.generate_mask_dict(aod, thr)
    idx = sweep.find_threshold_index(sweep.find_plume_extents(d, fires[:, 0], fires[:, 1]))
    assert [-1 if v is None else v for v in idx] == index.tolist()
    got, stats = sweep.plume_masks(d, idx, fires[:, 0], fires[:, 1])
    for f, k in enumerate(idx):
        ref = None if k is None else sweep_ref.plume_mask_ref(masks[k], fires[f, 0], fires[f, 1])
        if ref is None:
            assert got[f] is None
        else:
            assert np.array_equal(got[f], ref) and stats[f, 0] == ref.sum() == extents[k, f]
    plain = {t: masks[k] for k, t in enumerate(thr)}                                  # an ordinary dict of host masks
    again, _ = sweep.plume_masks(plain, idx, fires[:, 0], fires[:, 1])
    assert all((a is None and b is None) or np.array_equal(a, b) for a, b in zip(got, again))


def test_kernels_follow_scikit_image_documented_semantics(sw):
    """tests/skimage_cases.py (docstring examples and border rules of the absent scikit-image) through the kernels."""
    from tests import skimage_cases as C
    for src, expect in ((C.ERODE_FULL, C.ERODE_FULL_EXPECT), (C.OPEN_IN, C.OPEN_EXPECT), (C.DILATE_DOT, np.zeros_like(C.DILATE_DOT))):
        for dt in (np.float32, np.float64):
            assert np.array_equal(sw.unpack_bits(sw.mask_bits(src.astype(dt), [0.5]), src.shape[1])[0], expect)
            assert np.array_equal(sw.masks(src.astype(dt), [0.5]).cpu().numpy()[0].astype(bool), expect)
    for mask, expect in ((C.LABEL_EYE, C.LABEL_EYE_EXPECT), (C.LABEL_ORDER, C.LABEL_ORDER_EXPECT)):
        labels, sizes = sw.label(torch.from_numpy(mask.astype(np.uint8))[None].cuda())
        lab = labels[0].cpu().numpy()
        ids = np.unique(lab[lab >= 0])                          # canonical labels ascend in raster order of first pixels
        got = np.zeros(mask.shape, dtype=np.int64)
        for k, v in enumerate(ids):
            got[lab == v] = k + 1
        assert np.array_equal(got, expect)
        assert sizes[0].cpu().numpy().sum() == mask.sum()
    r, c = np.nonzero(C.RSO)
    assert np.array_equal(sw.cluster_fires(C.RSO.shape, r, c, 7) > 0, C.RSO_MIN7_CONN2)
    assert np.array_equal(sw.cluster_fires(C.RSO.shape, r, c, 8) > 0, C.RSO_MIN8_CONN2)
    # 8-connectivity through the bit-plane path: the eye is one component of three pixels
    ext = sw.extents_of_bits(sw.pack_bits_host(C.LABEL_EYE[None]), 3, [1], [1], win=1)
    assert ext.tolist() == [[3.0]]


def test_timestamp_front_half_equals_the_oracle_step_by_step(sw):
    """ThresholdSweep.timestamp = main :611-613 + identify :478-499 up to find_plume_mask's labelling, against the same
    steps made of oracle functions (fill, cluster_fires + centroids, three sweeps, threshold index, plume masks)."""
    from tests.sweep_data import synthetic_fire_pixels
    h, w = 150, 170
    aod, _ = synthetic_aod(h, w, 2)
    aod = (np.round(aod.astype(np.float64) * 1000) * 0.001)                 # MAIAC-like: multiples of 0.001 in float64
    rng = np.random.default_rng(5)
    aod[rng.random((h, w)) < 0.03] = sweep.NULL_VALUE
    aod[60:70, 80:100] = sweep.NULL_VALUE
    fr, fc = synthetic_fire_pixels(h, w, 3)
    keep = (fr > 20) & (fr < h - 21) & (fc > 20) & (fc < w - 21)            # centroids must keep the window inside
    fr, fc = fr[keep], fc[keep]
    got = sw.timestamp(aod, fr, fc)
    filled = sweep_ref.interpolate_aod_nearest_ref(aod)
    assert np.array_equal(got["aod_filled"].cpu().numpy(), filled)
    lab = sweep_ref.cluster_fires_ref((h, w), fr, fc)
    rows, cols = sweep_ref.cluster_centroids_ref(lab)
    assert np.array_equal(got["fire_labels"], lab) and np.array_equal(got["fire_rows"], rows) and np.array_equal(got["fire_cols"], cols)
    assert len(rows) >= 3 and len(got["sweeps"]) == 3
    n_masks = 0
    for sweep_out, (step, tmax) in zip(got["sweeps"], zip(sweep.THRESHOLD_STEP_SIZES, sweep.THRESHOLD_MAX)):
        thr = np.abs(np.arange(0, tmax, step) - tmax)
        masks = sweep_ref.threshold_masks_ref(filled, thr)
        ext = sweep_ref.find_plume_extents_ref(masks, rows, cols)
        idx = sweep_ref.find_threshold_index_ref(ext)
        assert np.array_equal(sweep_out["thresholds"], thr) and np.array_equal(sweep_out["extents"], ext)
        assert sweep_out["threshold_index"] == idx
        for f, k in enumerate(idx):
            ref = None if k is None else sweep_ref.plume_mask_ref(masks[k], rows[f], cols[f])
            if ref is None:
                assert sweep_out["plume_masks"][f] is None
            else:
                n_masks += 1
                full = sw.full_mask(sweep_out["plume_masks"][f], sweep_out["regions"][f], ref.shape)
                assert np.array_equal(full, ref) and sweep_out["regions"][f, 0] == ref.sum()
    assert n_masks >= 1
    assert sw.timestamp(aod, [], [])["sweeps"] == []
