"""CPU: the per-thread bodies of the bit-plane sweep kernels (csrc/sweep_bits.cuh), compiled for the host by g++
(tests/emu/sweep_bits_emu.cpp) and run sequentially, against the oracle and the reference-generated golden vectors.
Exact (boolean / integer work).  This checks the bit logic the CUDA kernels share, not the kernels: those are checked
on the GPU (tests/test_gpu_sweep.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import sweep_ref
from tests.sweep_data import synthetic_aod

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "sweep_cases.npz"))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libsweep_bits_emu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", so,
                           os.path.join(HERE, "emu", "sweep_bits_emu.cpp")])
    lib = ctypes.CDLL(so)
    lib.emu_ent_count.restype = ctypes.c_longlong
    return lib


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def unpack(bits, w):
    t, h, segs = bits.shape
    return np.unpackbits(bits.view(np.uint8).reshape(t, h, segs * 4), axis=2, bitorder="little")[:, :, :w].astype(bool)


def mask_bits(emu, aod, thr):
    aod = np.ascontiguousarray(aod, dtype=np.float32)
    thr = np.ascontiguousarray(thr, dtype=np.float64)
    h, w = aod.shape
    out = []
    for strip_rows in (16, 8, 5):                                 # the kernel takes 16- or 8-row strips; 5: ragged
        bits = np.full((len(thr), h, (w + 31) // 32), 0xDEADBEEF, dtype=np.uint32)
        emu.emu_mask_bits(P(aod), h, w, P(thr), len(thr), P(bits), strip_rows)
        out.append(bits)
    assert np.array_equal(out[0], out[1]) and np.array_equal(out[0], out[2])
    return out[0]


def bits_extents(emu, bits, w, rows, cols, win=sweep_ref.P_ID_WIN_SIZE, seed=0):
    t, h, _ = bits.shape
    rc = np.ascontiguousarray(np.stack([rows, cols], 1), dtype=np.int32)
    ent = np.full((emu.emu_ent_count(h, w, t), 2), 0x5A5A5A5A, dtype=np.int32)     # garbage: init must cover what is read
    out = np.full((t, len(rc)), -7, dtype=np.int32)
    emu.emu_bits_extents(P(bits), t, h, w, P(rc), len(rc), win, P(ent), P(out), seed)
    bits_extents.last_entries = ent
    return out


def fire_components(emu, bits, w, rows, cols, planes, win=sweep_ref.P_ID_WIN_SIZE):
    """after bits_extents(...) on the same bits: (component masks bool [n, H, W], stats [n, 8])"""
    t, h, segs = bits.shape
    rc = np.ascontiguousarray(np.stack([rows, cols], 1), dtype=np.int32)
    planes = np.ascontiguousarray(planes, dtype=np.int32)
    comp = np.full((len(rc), h, segs), 0xDEADBEEF, dtype=np.uint32)
    stats = np.full((len(rc), 8), -7, dtype=np.int32)
    emu.emu_fire_components(P(bits), t, h, w, P(rc), P(planes), len(rc), win, P(bits_extents.last_entries), P(comp), P(stats))
    return unpack(comp, w), stats


def check_components(emu, bits, masks, w, rows, cols, win):
    rng = np.random.default_rng(len(rows))
    planes = rng.integers(-1, bits.shape[0], len(rows))
    got, stats = fire_components(emu, bits, w, rows, cols, planes, win)
    for f, p in enumerate(planes):
        ref = None if p < 0 else sweep_ref.plume_mask_ref(masks[p], rows[f], cols[f], win)
        if ref is None:
            assert not got[f].any() and stats[f, 0] == 0 and stats[f, 5] == -1
        else:
            ys, xs = np.nonzero(ref)
            assert np.array_equal(got[f], ref)
            assert stats[f, :5].tolist() == [ref.sum(), ys.min(), xs.min(), ys.max() + 1, xs.max() + 1]


@pytest.mark.parametrize("i", range(int(G["n_cases"])))
def test_emulated_kernels_equal_reference_golden(emu, i):
    k = f"c{i}"
    h, w, seed = (int(v) for v in G[k + "_hws"])
    aod, fires = synthetic_aod(h, w, seed)
    thr = G[k + "_thr"]
    masks = np.unpackbits(G[k + "_masks"])[: len(thr) * h * w].reshape(len(thr), h, w).astype(bool)
    bits = mask_bits(emu, aod, thr)
    assert np.array_equal(unpack(bits, w), masks)
    assert not (bits.reshape(-1, bits.shape[2])[:, -1] >> np.uint32((w - 1) % 32 + 1)).any() or w % 32 == 0
    for s in (0, 5):
        assert np.array_equal(bits_extents(emu, bits, w, fires[:, 0], fires[:, 1], seed=s), G[k + "_extents"])


@pytest.mark.parametrize("h,w", [(1, 1), (2, 31), (3, 32), (5, 33), (7, 300), (40, 64), (33, 65), (61, 97)])
def test_mask_bits_equal_oracle_on_noisy_images(emu, h, w):
    rng = np.random.default_rng(h * 131 + w)
    aod = rng.random((h, w)).astype(np.float32)
    aod[rng.random((h, w)) < 0.02] = np.nan
    thr = np.array([0.1, 0.25, 0.5, float(np.float32(0.3)), 0.3, 0.9, -1.0, 2.0, 0.25, np.nan, np.inf, -np.inf])
    with np.errstate(invalid="ignore"):
        assert np.array_equal(unpack(mask_bits(emu, aod, thr), w), sweep_ref.threshold_masks_ref(aod, thr))
    many = np.concatenate([thr[:9], rng.random(32), [0.5, 0.5]])            # two chunks, unsorted, duplicates
    assert np.array_equal(unpack(mask_bits(emu, aod, many), w), sweep_ref.threshold_masks_ref(aod, many))
    blobs = (rng.random((h, w)) < 0.8).astype(np.float32)                    # large blobs: erosion leaves something
    assert np.array_equal(unpack(mask_bits(emu, blobs, [0.5]), w), sweep_ref.threshold_masks_ref(blobs, [0.5]))
    ones = np.ones((h, w), dtype=np.float32)
    assert unpack(mask_bits(emu, ones, [0.5]), w).all()


def test_float64_threshold_decision(emu):
    x = np.float32(0.48)
    for t in (0.48, float(x), float(np.nextafter(x, np.float32(1))), float(np.nextafter(x, np.float32(0))), 1e-50, -1e-50, 1e300):
        aod = np.full((6, 7), x, dtype=np.float32)
        assert unpack(mask_bits(emu, aod, [t]), 7).all() == bool(np.float64(x) > t)


@pytest.mark.parametrize("h,w,density", [(1, 1, 1.0), (7, 300, 0.5), (64, 64, 0.62), (97, 129, 0.4), (200, 333, 0.55)])
def test_extents_equal_oracle_on_random_masks(emu, h, w, density):
    rng = np.random.default_rng(h * 1000 + w)
    masks = rng.random((4, h, w)) < density
    masks[1] = ~masks[1] if h > 1 else masks[1]
    masks[3] = True
    if h >= 8 and w >= 8:
        sp = np.zeros((h, w), dtype=bool)
        y0, x0, y1, x1 = 0, 0, h - 1, w - 1
        while y1 - y0 > 3 and x1 - x0 > 3:
            sp[y0, x0:x1 + 1] = True
            sp[y0:y1 + 1, x1] = True
            sp[y1, x0 + 2:x1 + 1] = True
            sp[y0 + 2:y1 + 1, x0 + 2] = True
            y0, x0, y1, x1 = y0 + 2, x0 + 2, y1 - 2, x1 - 2
        masks[2] = sp
    bits = np.zeros((4, h, (w + 31) // 32), dtype=np.uint32)
    m8 = np.ascontiguousarray(masks.astype(np.uint8))
    emu.emu_pack_bits(P(m8), 4, h, w, P(bits))
    assert np.array_equal(unpack(bits, w), masks)
    win = min(sweep_ref.P_ID_WIN_SIZE, (h - 1) // 2, (w - 1) // 2)
    n = 12
    rows = rng.integers(win, h - win, n)
    cols = rng.integers(win, w - win, n)
    ref = sweep_ref.find_plume_extents_ref(masks, rows, cols, win)
    for s in (0, 3, 11):
        assert np.array_equal(bits_extents(emu, bits, w, rows, cols, win, seed=s), ref)
    check_components(emu, bits, masks, w, rows, cols, win)


def test_property_random_small_images(emu):
    """Property test over many small random shapes (widths around the 32- and 64-column strip boundaries, heights
    around the strip heights): masks, extents and component masks of the emulated kernel bodies equal the oracle."""
    rng = np.random.default_rng(2024)
    widths = [1, 2, 31, 32, 33, 63, 64, 65, 95, 96, 97, 128, 130]
    heights = [1, 2, 3, 7, 8, 9, 15, 16, 17, 33]
    for trial in range(60):
        h, w = int(rng.choice(heights)), int(rng.choice(widths))
        kind = trial % 4
        if kind == 0:                                             # smooth blobs + noise
            yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
            aod = (0.5 + 0.4 * np.sin(yy / 3.0 + trial) * np.cos(xx / 5.0)).astype(np.float32)
            aod += (rng.random((h, w)) < 0.05) * rng.random((h, w)).astype(np.float32)
        elif kind == 1:                                           # dense noise
            aod = rng.random((h, w)).astype(np.float32)
        elif kind == 2:                                           # long horizontal / vertical bars crossing word boundaries
            aod = np.zeros((h, w), dtype=np.float32)
            aod[::3, :] = 1.0
            aod[:, ::7] = 1.0
            aod[rng.random((h, w)) < 0.1] = 0.0
        else:                                                     # everything set except a few holes
            aod = np.ones((h, w), dtype=np.float32)
            aod[rng.random((h, w)) < 0.08] = 0.0
        thr = np.sort(rng.random(int(rng.integers(1, 6))))[::-1].copy()
        ref_masks = sweep_ref.threshold_masks_ref(aod, thr)
        bits = mask_bits(emu, aod, thr)
        assert np.array_equal(unpack(bits, w), ref_masks), (trial, h, w)
        raw = aod > 0.5                                           # un-opened masks: thin structures, diagonal contacts
        stack = np.concatenate([ref_masks, raw[None], np.eye(h, w, dtype=bool)[None]])
        pb = np.zeros((len(stack), h, (w + 31) // 32), dtype=np.uint32)
        m8 = np.ascontiguousarray(stack.astype(np.uint8))
        emu.emu_pack_bits(P(m8), len(stack), h, w, P(pb))
        win = int(min(3, (h - 1) // 2, (w - 1) // 2))
        n = 6
        rows = rng.integers(win, h - win, n)
        cols = rng.integers(win, w - win, n)
        ref = sweep_ref.find_plume_extents_ref(stack, rows, cols, win)
        assert np.array_equal(bits_extents(emu, pb, w, rows, cols, win, seed=trial), ref), (trial, h, w)
        check_components(emu, pb, stack, w, rows, cols, win)
