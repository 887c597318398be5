"""Generates tests/golden/sweep_cases.npz from the REFERENCE's threshold-sweep functions (build container only; no
reference source is copied).  Compiled unmodified from the syntax tree of
``src/features/plume_identifier_gaussian_profile.py``:

    construct_dist_matrix, P_ID_WIN_SIZE, DISTANCE_MATRIX     :29-38
    generate_mask_dict        :142-154   per threshold: aod > t -> binary_erosion -> binary_dilation
    find_plume_extents        :157-179   per threshold: label(); per fire: size of the nearest labelled region
    extract_label             :182-202   nearest labelled pixel inside the 31 x 31 window around the fire
    find_threshold_index      :204-240   per fire: threshold index of the largest size ratio

scikit-image (the reference's provider of ``label``, ``binary_erosion``, ``binary_dilation``) is NOT installed in
this image and the reference pins no version (requirements.txt), so those three names are bound to scipy.ndimage
stand-ins with scikit-image's documented defaults: cross-shaped footprint; erosion treats pixels beyond the border
as set (border_value=True), dilation as unset; ``label`` with full (8-) connectivity, background 0.  Everything
else -- loops, windows, distance matrix, ratio logic -- is the reference's code.  Parity of this path is therefore
pinned to the reference's control flow but NOT to scikit-image's primitives.
"""
import ast
import os
import sys

import numpy as np
import scipy.ndimage as ndi

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.sweep_data import synthetic_aod  # noqa: E402

REF = "/root/reference/src/features/plume_identifier_gaussian_profile.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sweep_cases.npz")
WANT = ("construct_dist_matrix", "generate_mask_dict", "find_plume_extents", "extract_label", "find_threshold_index")
CONSTS = ("P_ID_WIN_SIZE", "DISTANCE_MATRIX")
CROSS = ndi.generate_binary_structure(2, 1)


def load_reference_functions():
    tree = ast.parse(open(REF).read(), REF)
    body = [n for n in tree.body
            if (isinstance(n, ast.FunctionDef) and n.name in WANT)
            or (isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") in CONSTS)]
    assert len(body) == len(WANT) + len(CONSTS), [getattr(n, "name", None) for n in body]
    ns = {"np": np,
          "label": lambda m: ndi.label(m, structure=np.ones((3, 3), dtype=int))[0],
          "binary_erosion": lambda m: ndi.binary_erosion(m, structure=CROSS, border_value=True),
          "binary_dilation": lambda m: ndi.binary_dilation(m, structure=CROSS)}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns


def main():
    ref = load_reference_functions()
    out = {"p_id_win_size": np.array(ref["P_ID_WIN_SIZE"]), "distance_matrix": ref["DISTANCE_MATRIX"]}
    n = 0
    for (h, w, seed) in [(96, 128, 1), (150, 150, 2), (200, 131, 3)]:
        aod, fires = synthetic_aod(h, w, seed)
        for step, tmax in [(0.02, 0.5), (0.03, 0.75), (0.04, 1)]:
            thr = np.abs(np.arange(0, tmax, step) - tmax)                     # gaussian_profile.py:492
            masks = ref["generate_mask_dict"](aod, thr)
            ext = ref["find_plume_extents"](masks, fires[:, 0], fires[:, 1])
            idx = ref["find_threshold_index"](ext)
            k = f"c{n}"
            out[k + "_hws"] = np.array([h, w, seed])
            out[k + "_thr"] = thr
            out[k + "_fires"] = fires
            out[k + "_masks"] = np.packbits(np.stack([masks[t] for t in thr]).astype(np.uint8), axis=None)
            out[k + "_extents"] = ext
            out[k + "_index"] = np.array([-1 if i is None else i for i in idx])
            n += 1
    # find_threshold_index on hand-made extent tables (all-zero, leading zeros, max at the ends)
    tables = np.array([[0, 0, 0, 0, 0], [0, 0, 10, 40, 45], [5, 6, 30, 31, 32], [5, 50, 51, 52, 53],
                       [5, 6, 7, 8, 80], [0, 3, 0, 9, 27], [7, 7, 7, 7, 7], [9, 3, 1, 0, 0]], dtype=np.float64).T
    with np.errstate(all="ignore"):
        out["tables"] = tables
        out["tables_index"] = np.array([-1 if i is None else i for i in ref["find_threshold_index"](tables)])
    out["n_cases"] = np.array(n)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", n, "sweeps; table indices", out["tables_index"].tolist())


if __name__ == "__main__":
    sys.exit(main())
