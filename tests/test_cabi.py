"""CPU: the C-ABI shared library loads without a GPU and exports exactly what include/plume_b200.h
declares; the ctypes prototypes cover every declared entry point; argument errors are reported through
return codes + plume_last_error() (no compute is launched here)."""
import os
import re

from kcl_ltss_bioatm_b200 import lib as plib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "plume_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(plume_[a-zA-Z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    handle = plib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/plume_b200.h but not exported"


def test_ctypes_signatures_cover_the_header():
    assert sorted(plib.SIGNATURES) == declared_symbols()


def test_version_and_error_reporting_without_gpu():
    handle = plib.load()
    assert b"sm_100a" in handle.plume_version()
    # null pointers are rejected before anything touches the device
    rc = handle.plume_pad_channels(None, 8, None, 64, 16, None)
    assert rc != 0 and b"null" in handle.plume_last_error()
    rc = handle.plume_adam(None, None, None, None, 4, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, None)
    assert rc != 0
    # planning helpers are pure host code
    assert handle.plume_wgrad_splits(32, 256, 256, 9, 64, 64) >= 1
    assert handle.plume_wgrad_workspace_bytes(32, 16, 16, 9, 1024, 1024) == 0  # atomics: no workspace
    assert handle.plume_wgrad_splits(0, 1, 1, 9, 64, 64) == 0


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    monkeypatch.setattr(plib, "_lib", None)
    monkeypatch.setattr(plib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        plib.load()
    except plib.PlumeError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("load() must raise when the library is absent")
    finally:
        monkeypatch.undo()
        plib._lib = None
        plib.load()


def test_cuda_ops_refuse_to_run_without_gpu():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kcl_ltss_bioatm_b200.ops import CudaOps

    with pytest.raises(plib.PlumeError):
        CudaOps()


def test_label_generation_entry_points_plan_and_reject_without_gpu():
    """Workspace planning of the sweep / fill calls is pure host code; null pointers and undersized workspaces are
    rejected before anything touches the device."""
    handle = plib.load()
    h, w, t = 1200, 1200, 75
    segs = (w + 31) // 32

    def up(n):
        return (n + 255) // 256 * 256

    assert handle.plume_sweep_workspace_bytes(h, w, t) == up(t * h * 16 * segs * 8) + up(t * h * segs * 4)
    assert handle.plume_sweep_workspace_bytes(0, w, t) == 0 and handle.plume_sweep_workspace_bytes(h, w, 0) == 0
    assert handle.plume_fill_nearest_workspace_bytes(h, w) == up(h * segs * 4) + up(h * w * 4)
    for name, args in (("plume_sweep_extents", (None, h, w, None, t, None, 4, 15, None, 0, None, None)),
                       ("plume_sweep_extents_f64", (None, h, w, None, t, None, 4, 15, None, 0, None, None)),
                       ("plume_bits_extents", (None, t, h, w, None, 4, 15, None, 0, None, None)),
                       ("plume_threshold_mask_bits", (None, h, w, None, t, None, None)),
                       ("plume_threshold_mask_bits_f64", (None, h, w, None, t, None, None)),
                       ("plume_threshold_masks_f64", (None, h, w, None, t, None, None)),
                       ("plume_pack_mask_bits", (None, t, h, w, None, None)),
                       ("plume_fill_nearest", (None, h, w, -999.0, None, 0, None, None)),
                       ("plume_fill_nearest_f64", (None, h, w, -999.0, None, 0, None, None))):
        assert getattr(handle, name)(*args) != 0 and b"null" in handle.plume_last_error(), name
    # nothing to do is not an error
    assert handle.plume_sweep_extents(None, h, w, None, 0, None, 4, 15, None, 0, None, None) == 0
    assert handle.plume_fill_nearest(None, 0, 0, -999.0, None, 0, None, None) == 0
