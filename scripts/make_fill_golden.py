"""Generates tests/golden/fill_cases.npz from the REFERENCE's ``interpolate_aod_nearest``
(src/features/plume_identifier_gaussian_profile.py:451-461), compiled unmodified from the file's syntax tree and run
with the real scipy.interpolate of this image (build container only; no reference source is copied).  Stored per
case: the filled image.  The scipy version is recorded: at pixels with several equidistant nearest valid pixels the
answer depends on scipy's kd-tree traversal, so the tests compare there only that the value is one of the tied ones.
"""
import ast
import os
import sys

import numpy as np
import scipy
from scipy import interpolate

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.sweep_data import synthetic_null_aod  # noqa: E402

REF = "/root/reference/src/features/plume_identifier_gaussian_profile.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fill_cases.npz")
CASES = [(40, 50, 1, "float64"), (64, 96, 2, "float64"), (33, 130, 3, "float32"), (7, 9, 4, "float64"), (120, 75, 5, "float64")]


def main():
    tree = ast.parse(open(REF).read(), REF)
    body = [n for n in tree.body
            if (isinstance(n, ast.FunctionDef) and n.name == "interpolate_aod_nearest")
            or (isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "NULL_VALUE")]
    assert len(body) == 2
    ns = {"np": np, "interpolate": interpolate}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    out = {"scipy_version": np.array(scipy.__version__), "null_value": np.array(ns["NULL_VALUE"])}
    for n, (h, w, seed, dt) in enumerate(CASES):
        aod = synthetic_null_aod(h, w, seed, np.dtype(dt))
        filled = ns["interpolate_aod_nearest"](aod)
        assert filled.dtype == np.float64 and filled.shape == aod.shape      # scipy returns float64 for any input
        out[f"c{n}_hws"] = np.array([h, w, seed])
        out[f"c{n}_dtype"] = np.array(dt)
        out[f"c{n}_filled"] = filled
    out["n_cases"] = np.array(len(CASES))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; scipy", scipy.__version__)


if __name__ == "__main__":
    sys.exit(main())
