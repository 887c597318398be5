"""Deterministic synthetic AOD fields shared by scripts/make_sweep_golden.py and the sweep tests (the golden
file stores results only)."""
import numpy as np


def synthetic_aod(h, w, seed):
    """Smooth background + elongated plumes + salt noise (singletons the erosion must remove); float32."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    aod = 0.05 + 0.03 * np.sin(yy / 17.0) * np.cos(xx / 23.0)
    fires = []
    for _ in range(6):
        cy, cx = rng.uniform(20, h - 20), rng.uniform(20, w - 20)
        th = rng.uniform(0, np.pi)
        ln, wd = rng.uniform(15, 45), rng.uniform(3, 8)
        u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
        v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        aod += rng.uniform(0.3, 1.1) * np.exp(-0.5 * ((u / ln) ** 2 + (v / wd) ** 2))
        fires.append((int(np.clip(cy, 16, h - 17)), int(np.clip(cx, 16, w - 17))))
    salt = rng.random((h, w)) < 0.01
    aod[salt] += rng.uniform(0.2, 1.0, int(salt.sum()))
    # extra fires: far from any plume (no label in the window), and on a plume edge
    fires += [(16, 16), (h - 17, w - 17), (h // 2, w // 2)]
    return aod.astype(np.float32), np.array(fires, dtype=np.int64)


def synthetic_fire_pixels(h, w, seed):
    """Active-fire detections as (rows, cols): clumps of touching pixels (some below the 3-pixel minimum), diagonal
    chains, isolated pixels and duplicates (a pixel detected twice)."""
    rng = np.random.default_rng(1000 + seed)
    rows, cols = [], []
    for _ in range(max(2, h * w // 400)):
        cy, cx = int(rng.integers(0, h)), int(rng.integers(0, w))
        for _ in range(int(rng.integers(1, 9))):
            rows.append(int(np.clip(cy + rng.integers(-2, 3), 0, h - 1)))
            cols.append(int(np.clip(cx + rng.integers(-2, 3), 0, w - 1)))
    for k in range(min(h, w, 6)):                                   # a diagonal chain: 8-connected only
        rows.append(k)
        cols.append(k)
    rows += rows[:3]
    cols += cols[:3]
    return np.array(rows, dtype=np.int64), np.array(cols, dtype=np.int64)


def synthetic_null_aod(h, w, seed, dtype=np.float64):
    """AOD as the reference reads it (int16 * 0.001, float64) with NULL_VALUE (-999) gaps: cloud-like blobs, salt,
    a fully null band and a null border."""
    rng = np.random.default_rng(2000 + seed)
    aod = (rng.integers(0, 1500, (h, w)).astype(np.int16) * 0.001).astype(dtype)
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    null = rng.random((h, w)) < 0.05
    for _ in range(min(200, max(1, h * w // 1500))):
        cy, cx, r = rng.integers(0, h), rng.integers(0, w), rng.integers(2, min(40, max(3, min(h, w) // 3)))
        null |= (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
    if h > 6:
        null[h // 2:h // 2 + 2, :] = True
    null[0, :] = True
    null[:, -1] = True
    if null.all():
        null[h // 3, w // 3] = False
    aod[null] = -999
    return aod
