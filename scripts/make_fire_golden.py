"""Generates tests/golden/fire_cases.npz from the REFERENCE's own fire-geolocation functions (build container
only; no reference source is copied).  ``plume_identifier_gaussian_profile.py`` cannot be imported (TkAgg, pyhdf,
scikit-image are absent), so the pure numpy / pandas functions of this path are compiled, unmodified, from its
syntax tree together with the module constant they read:

    subset_fires_to_image   gaussian_profile.py:46-54    fires of one timestamp inside the image's lat/lon extent
    grid_indexes            gaussian_profile.py:57-62    row / column index grids
    haversine               gaussian_profile.py:65-82    great-circle distance, km
    locate_fire_in_image    gaussian_profile.py:85-123   nearest pixel per fire within a +-0.05 degree box,
                                                         fires closer than P_ID_WIN_SIZE+1 to the edge dropped
    P_ID_WIN_SIZE = 15      gaussian_profile.py:37

Grids: MAIAC-like 1 km sinusoidal tiles re-projected to lat/lon (rows of constant latitude, longitude spacing
growing with 1/cos(lat)), plus a rotated swath-like grid.  Fires: inside, near every edge, outside the image, and
in gaps no pixel covers (the reference silently skips those: bare except, :120-121).
"""
import ast
import os
import sys
import warnings

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.grids import GRIDS  # noqa: E402

REF = "/root/reference/src/features/plume_identifier_gaussian_profile.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fire_cases.npz")
WANT = ("subset_fires_to_image", "grid_indexes", "haversine", "locate_fire_in_image")


def load_reference_functions():
    tree = ast.parse(open(REF).read(), REF)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANT]
    consts = [n for n in tree.body if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "P_ID_WIN_SIZE"]
    assert len(body) == len(WANT) and len(consts) == 1
    ns = {"np": np, "pd": pd}
    exec(compile(ast.Module(body=consts + body, type_ignores=[]), REF, "exec"), ns)
    return ns


def main():
    ref = load_reference_functions()
    rng = np.random.default_rng(20181019)
    out = {"p_id_win_size": np.array(ref["P_ID_WIN_SIZE"])}
    for name, make in GRIDS.items():
        lat, lon = make()
        h, w = lat.shape
        n = 160
        # pixel-centred fires with sub-pixel jitter, fires along the edges, fires outside, duplicates
        rr = rng.integers(0, h, n)
        cc = rng.integers(0, w, n)
        rr[:20] = rng.integers(0, 18, 20)
        cc[20:40] = rng.integers(w - 18, w, 20)
        flat = lat[rr, cc] + rng.normal(0, 0.004, n)
        flon = lon[rr, cc] + rng.normal(0, 0.004, n)
        flat[40:50] += 5.0                      # far outside: empty box -> skipped
        flon[50:55] -= 0.2
        flat[60], flon[60] = flat[61], flon[61]   # duplicate fire
        fires = pd.DataFrame({"latitude": flat, "longitude": flon,
                              "date_time": ["t0"] * (n - 10) + ["t1"] * 10})
        rows, cols = ref["grid_indexes"](lat)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sub = ref["subset_fires_to_image"](lat, lon, fires, "t0")
        fr, fc = ref["locate_fire_in_image"](sub, lat, lon, rows, cols)
        # also the unfiltered nearest pixel per fire (edge filter off) to pin the search itself
        ref["P_ID_WIN_SIZE"] = -10 ** 6
        fr_all, fc_all = ref["locate_fire_in_image"](sub, lat, lon, rows, cols)
        ref["P_ID_WIN_SIZE"] = int(out["p_id_win_size"])
        d = ref["haversine"](flon[:32], flat[:32], lon[rr[:32], cc[:32]], lat[rr[:32], cc[:32]])
        # one fire at a time over ALL fires (also the ones outside the image): pins which fires are skipped
        per_fire = np.full((n, 2), -1, dtype=np.int64)
        for i in range(n):
            r1, c1 = ref["locate_fire_in_image"](fires.iloc[i:i + 1], lat, lon, rows, cols)
            if r1:
                per_fire[i] = (r1[0], c1[0])
        out[name + "_per_fire"] = per_fire
        out[name + "_fire_lat"], out[name + "_fire_lon"] = flat, flon
        out[name + "_fire_dt"] = fires["date_time"].values.astype("U4")
        out[name + "_subset_index"] = sub.index.values
        out[name + "_rows"], out[name + "_cols"] = np.array(fr), np.array(fc)
        out[name + "_rows_nofilter"], out[name + "_cols_nofilter"] = np.array(fr_all), np.array(fc_all)
        out[name + "_haversine32"] = d
        out[name + "_gen_rc"] = np.stack([rr[:32], cc[:32]], 1)
        print(name, lat.shape, "fires", n, "subset", len(sub), "located", len(fr), "without edge filter", len(fr_all),
              "per-fire located", int((per_fire[:, 0] >= 0).sum()))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
