"""Generates tests/golden/hull_cases.npz from the REFERENCE's own functions (run in the build container only;
/root/reference does not exist on the GPU box, and no reference source is copied into this repo).

`src/features/plume_selector.py` cannot be imported as a module here (it pulls in TkAgg, pyproj and pyhdf, which
are absent), so the four pure functions this path restates are compiled from its syntax tree, unmodified, into a
namespace holding only numpy / pandas / scipy.spatial.Delaunay:

    in_hull                   plume_selector.py:88-98     Delaunay(hull).find_simplex(p) >= 0
    find_plume_aod            plume_selector.py:101-116   pixel grid -> in_hull -> values inside
    subset_plume              plume_selector.py:53-85     crop to the hull's bounding box +- 40 px, shift the hull
    remove_duplicated_plumes  plume_selector.py:26-49     drop plumes whose rounded centroid repeats

Cases: random convex hulls (integer pixel vertices, as written by plume_identifier_gaussian_profile.py:283-289)
inside images of several sizes, including hulls touching the image border, thin hulls, and repeated plumes.
"""
import ast
import os
import sys

import numpy as np
import pandas as pd
from scipy.spatial import ConvexHull, Delaunay

REF = "/root/reference/src/features/plume_selector.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "hull_cases.npz")
WANT = ("in_hull", "find_plume_aod", "subset_plume", "remove_duplicated_plumes")


def load_reference_functions():
    tree = ast.parse(open(REF).read(), REF)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANT]
    assert len(body) == len(WANT), [n.name for n in body]
    ns = {"np": np, "pd": pd, "Delaunay": Delaunay}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns


def random_hull(rng, h, w, kind):
    """Integer hull vertices the way the reference writes them: ConvexHull of the pixels of a blob."""
    if kind == "thin":
        cy, cx = rng.integers(10, h - 10), rng.integers(10, w - 10)
        ang = rng.uniform(0, np.pi)
        t = rng.uniform(-1, 1, 60) * rng.uniform(15, min(h, w) / 3)
        s = rng.uniform(-1, 1, 60) * rng.uniform(1.5, 4)
        ys = cy + t * np.sin(ang) + s * np.cos(ang)
        xs = cx + t * np.cos(ang) - s * np.sin(ang)
    elif kind == "edge":
        ys = rng.uniform(-5, 30, 80)
        xs = rng.uniform(w - 35, w + 5, 80)
    else:
        cy, cx = rng.uniform(0.2, 0.8) * h, rng.uniform(0.2, 0.8) * w
        r = rng.uniform(6, min(h, w) / 4)
        ys = cy + rng.normal(0, r / 2, 100)
        xs = cx + rng.normal(0, r / 2, 100)
    pts = np.unique(np.stack([np.clip(np.round(ys), 0, h - 1), np.clip(np.round(xs), 0, w - 1)], 1).astype(np.int64),
                    axis=0)
    hull = ConvexHull(pts)
    return pts[hull.vertices, 1].astype(np.float64), pts[hull.vertices, 0].astype(np.float64)  # hull_x, hull_y


def pattern_image(h, w):
    """Deterministic image the tests regenerate instead of storing it (values in [0, 1), all distinct nearby)."""
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    return (((yy * 31 + xx * 17) % 1000) / 1000.0).astype(np.float32)


def main():
    ref = load_reference_functions()
    rng = np.random.default_rng(20181018)
    out = {}
    n_case = 0
    for (h, w) in [(64, 64), (120, 120), (97, 97), (256, 256)]:   # find_plume_aod indexes [yy, xx]: square images
        for kind in ("blob", "thin", "edge", "blob", "thin"):
            hx, hy = random_hull(rng, h, w, kind)
            img = pattern_image(h, w)
            # the reference's mask over the full pixel grid (find_plume_aod builds exactly these coordinates)
            yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
            pts = np.stack([xx.ravel(), yy.ravel()], 1)
            inside = ref["in_hull"](pts, np.vstack((hx, hy)).T).reshape(h, w)
            aod_in = ref["find_plume_aod"](img, hx, hy)
            k = f"c{n_case}"
            out[k + "_hw"] = np.array([h, w])
            out[k + "_hull_x"], out[k + "_hull_y"] = hx, hy
            out[k + "_mask"] = inside.astype(np.uint8)
            out[k + "_aod_sorted"] = np.sort(aod_in)
            n_case += 1
    out["n_mask_cases"] = np.array(n_case)

    # subset_plume: crop window and shifted hull for hulls at various positions of a 300 x 200 image
    aod = pattern_image(300, 200)
    n_sub = 0
    for (x0, y0, x1, y1) in [(60, 70, 120, 150), (10, 20, 50, 60), (150, 250, 199, 299), (0, 0, 30, 30),
                             (45, 41, 90, 100), (39, 40, 41, 42)]:
        hx = np.array([x0, x1, x1, x0, (x0 + x1) // 2], dtype=np.float64)
        hy = np.array([y0, y0, y1, y1, (y0 + y1) // 2], dtype=np.float64)
        df = pd.DataFrame({"hull_x": hx, "hull_y": hy})
        crop, sx, sy = ref["subset_plume"](aod, df)
        k = f"s{n_sub}"
        out[k + "_hull_x"], out[k + "_hull_y"] = hx, hy
        out[k + "_crop_shape"] = np.array(crop.shape)
        out[k + "_crop_sum"] = np.array(crop.astype(np.float64).sum())
        out[k + "_crop_corner"] = np.array([crop[0, 0], crop[-1, -1]])
        out[k + "_shift_x"], out[k + "_shift_y"] = np.asarray(sx, dtype=np.float64), np.asarray(sy, dtype=np.float64)
        n_sub += 1
    out["n_subset_cases"] = np.array(n_sub)

    # remove_duplicated_plumes: ids 0..5 over two datetimes, plumes 1 and 4 repeat plume 0's centroid
    rows = []
    base = {0: (10.0, 20.0), 1: (10.0001, 20.0002), 2: (11.0, 20.0), 3: (10.0, 21.0), 4: (10.0, 20.0), 5: (12.5, 22.5)}
    for dt in ("2017-08-01 10:30", "2017-08-02 11:00"):
        for pid, (la, lo) in base.items():
            if dt.endswith("11:00") and pid == 4:
                la += 0.5
            for dv in (-0.01, 0.0, 0.01):
                rows.append({"id": float(pid), "hull_lats": la + dv, "hull_lons": lo - dv, "hull_x": 1.0, "hull_y": 2.0,
                             "datetime": dt})
    df = pd.DataFrame(rows)
    kept = ref["remove_duplicated_plumes"](df.copy())
    out["dedup_in_id"] = df["id"].values
    out["dedup_in_lat"], out["dedup_in_lon"] = df["hull_lats"].values, df["hull_lons"].values
    out["dedup_in_dt"] = df["datetime"].values.astype("U32")
    out["dedup_keep_id"] = kept["id"].values
    out["dedup_keep_dt"] = kept["datetime"].values.astype("U32")
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", n_case, "mask cases,", n_sub, "subset cases,",
          len(kept), "of", len(df), "rows kept by the de-duplication")


if __name__ == "__main__":
    sys.exit(main())
