"""GPU: plume_rasterize_hulls through the C ABI against the oracle (bit-exact) and against the golden masks
recorded from the reference's Delaunay in_hull."""
import os

import numpy as np
import pytest
import torch

from kcl_ltss_bioatm_b200 import labels
from oracle import hull_ref

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hull_cases.npz"))


def pattern_image(h, w):
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    return (((yy * 31 + xx * 17) % 1000) / 1000.0).astype(np.float32)


def random_hulls(rng, n, h, w):
    out = []
    for _ in range(n):
        cy, cx = rng.uniform(-10, h + 10), rng.uniform(-10, w + 10)
        r = rng.uniform(3, 60)
        k = rng.integers(3, 40)
        ys = np.round(cy + rng.normal(0, r, k) * rng.uniform(0.1, 1))
        xs = np.round(cx + rng.normal(0, r, k))
        try:
            labels.convex_polygon(xs, ys)
        except ValueError:
            continue
        out.append((xs, ys))
    return out


@pytest.fixture(scope="module")
def rast():
    return labels.LabelRasterizer("cuda:0")


@pytest.mark.parametrize("i", range(int(G["n_mask_cases"])))
def test_scene_mask_equals_reference_golden(rast, i):
    k = f"c{i}"
    h, w = (int(v) for v in G[k + "_hw"])
    m = rast.scene_mask([(G[k + "_hull_x"], G[k + "_hull_y"])], h, w).cpu().numpy()
    assert np.array_equal(m, G[k + "_mask"])
    aod = labels.find_plume_aod(pattern_image(h, w), G[k + "_hull_x"], G[k + "_hull_y"])
    assert np.array_equal(np.sort(aod), G[k + "_aod_sorted"])
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    pts = np.stack([xx.ravel(), yy.ravel()], 1)[:: 7]
    hull = np.stack([G[k + "_hull_x"], G[k + "_hull_y"]], 1)
    assert np.array_equal(labels.in_hull(pts, hull), G[k + "_mask"].ravel()[:: 7].astype(bool))


@pytest.mark.parametrize("h,w,n", [(1, 1, 3), (33, 130, 40), (200, 257, 300), (517, 1031, 700), (1200, 1200, 64)])
def test_union_of_many_hulls_ragged_sizes(rast, h, w, n):
    hulls = random_hulls(np.random.default_rng(h * 7 + w), n, h, w)
    m = rast.scene_mask(hulls, h, w).cpu().numpy()
    ref = hull_ref.rasterize_ref(hulls, h, w)
    assert m.dtype == np.uint8 and np.array_equal(m, ref)
    assert set(np.unique(m)) <= {0, 1}


def test_no_hulls_gives_empty_mask(rast):
    assert int(rast.scene_mask([], 70, 90).sum()) == 0


def test_tile_windows_equal_scene_crops(rast):
    h, w, t = 700, 900, 256
    hulls = random_hulls(np.random.default_rng(5), 120, h, w)
    scene = rast.scene_mask(hulls, h + t, w + t).cpu().numpy()      # room for windows hanging over the edge
    ys, xs = [0, 13, 444, 600, 699], [0, 700, 321, 644, 899]
    tiles = rast.tile_masks(hulls, ys, xs, t).cpu().numpy()
    for k, (y, x) in enumerate(zip(ys, xs)):
        assert np.array_equal(tiles[k], scene[y:y + t, x:x + t])


def test_build_training_tiles_and_file_format(rast, tmp_path):
    h, w, c, t = 512, 768, 8, 256
    g = torch.Generator().manual_seed(3)
    scene = torch.randn(h, w, c, generator=g).to(torch.bfloat16).cuda()
    hulls = [(np.array([100, 200, 200, 100.0]), np.array([50, 50, 120, 120.0])),      # inside tile (0, 0)
             (np.array([600, 700, 650.0]), np.array([300, 300, 400.0]))]              # inside tile (256, 512)
    x, m, ys, xs = labels.build_training_tiles(scene, hulls, tile=t, rasterizer=rast)
    assert sorted(zip(ys.tolist(), xs.tolist())) == [(0, 0), (256, 512)]
    full = hull_ref.rasterize_ref(hulls, h, w)
    for k in range(len(ys)):
        y0, x0 = int(ys[k]), int(xs[k])
        assert np.array_equal(m[k].cpu().numpy(), full[y0:y0 + t, x0:x0 + t])
        assert torch.equal(x[k], scene[y0:y0 + t, x0:x0 + t])
    path = labels.write_training_tiles(str(tmp_path), "scene0", x, m)
    blob = torch.load(path)
    assert blob["x"].shape == (2, t, t, c) and blob["mask"].dtype == torch.uint8
