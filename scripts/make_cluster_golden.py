"""Generates tests/golden/cluster_cases.npz from the REFERENCE's ``cluster_fires``
(src/features/plume_identifier_gaussian_profile.py:126-139), compiled unmodified from the file's syntax tree (build
container only; no reference source is copied).

scikit-image is not installed here and the reference pins no version, so the two primitives the function calls are
bound to stand-ins with scikit-image's documented behaviour: ``label(img, connectivity=2)`` -> scipy.ndimage.label with
the full 3 x 3 structure (labels 1..n numbered in raster order of each component's first pixel, background 0);
``remove_small_objects(labels, min_size, connectivity)`` on an integer label image -> every label whose pixel count is
below min_size is set to 0, the other labels keep their numbers.  Control flow pinned, primitives not.
"""
import ast
import os
import sys

import numpy as np
import scipy.ndimage as ndi

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.sweep_data import synthetic_fire_pixels  # noqa: E402

REF = "/root/reference/src/features/plume_identifier_gaussian_profile.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cluster_cases.npz")


def remove_small_objects(labels, min_size=64, connectivity=1):
    out = labels.copy()
    sizes = np.bincount(labels.ravel())
    too_small = sizes < min_size
    out[too_small[labels]] = 0
    return out


def load_reference_function():
    tree = ast.parse(open(REF).read(), REF)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "cluster_fires"]
    assert len(body) == 1
    ns = {"np": np, "remove_small_objects": remove_small_objects,
          "label": lambda m, connectivity=None: ndi.label(m, structure=np.ones((3, 3), dtype=int))[0]}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns["cluster_fires"]


def main():
    cluster_fires = load_reference_function()
    out = {}
    cases = [(40, 50, 1), (96, 128, 2), (200, 131, 3), (300, 300, 4), (1, 9, 5), (64, 64, 6)]
    for n, (h, w, seed) in enumerate(cases):
        rows, cols = synthetic_fire_pixels(h, w, seed)
        lab = cluster_fires(np.zeros((h, w), dtype=np.float32), rows, cols)
        out[f"c{n}_hws"] = np.array([h, w, seed])
        out[f"c{n}_labels"] = lab.astype(np.int32)
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", [int(out[f"c{i}_labels"].max()) for i in range(len(cases))])


if __name__ == "__main__":
    sys.exit(main())
