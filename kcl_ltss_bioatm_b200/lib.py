"""ctypes binding of ``libplume_b200.so`` (C ABI declared in ``include/plume_b200.h``).

The library is the product: there is no Python/CPU fallback.  ``load()`` raises if the shared object
is missing, and every wrapper raises :class:`PlumeError` with ``plume_last_error()`` when an entry
point returns non-zero.  Pointers are taken from torch tensors (``data_ptr()``); the CUDA stream is
torch's current stream.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libplume_b200.so")
# experiments: an alternative build of the same library (A/B timing of kernel variants)
LIB_PATH = os.environ.get("PLUME_B200_LIB", LIB_PATH)

_P, _I, _F, _LL, _D = c_void_p, c_int, c_float, c_longlong, c_double

# name -> (restype, argtypes); mirrors include/plume_b200.h one to one
SIGNATURES = {
    "plume_version": (c_char_p, []),
    "plume_last_error": (c_char_p, []),
    "plume_debug_word": (_I, []),
    "plume_num_sms": (_I, []),
    "plume_debug_set_prof": (None, [_P]),
    "plume_conv3x3_fwd": (_I, [_P, _I, _P, _P, _P, _I, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P]),
    "plume_conv3x3_dgrad": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "plume_wgrad_splits": (_I, [_I, _I, _I, _I, _I, _I]),
    "plume_wgrad_workspace_bytes": (c_size_t, [_I, _I, _I, _I, _I, _I]),
    "plume_conv3x3_wgrad": (_I, [_P, _I, _P, _I, _P, _I, _P, c_size_t, _I, _I, _I, _I, _I, _P]),
    "plume_convT2x2_concat_fwd": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "plume_convT2x2_dgrad": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "plume_convT2x2_wgrad": (_I, [_P, _I, _P, _I, _P, _I, _P, c_size_t, _I, _I, _I, _I, _I, _P]),
    "plume_conv3x3_fwd_x3": (_I, [_P, _I, _P, _P, _P, _I, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P]),
    "plume_conv3x3_dgrad_x3": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "plume_conv3x3_wgrad_x3": (_I, [_P, _I, _P, _I, _P, _I, _P, c_size_t, _I, _I, _I, _I, _I, _P]),
    "plume_convT2x2_concat_fwd_x3": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "plume_convT2x2_dgrad_x3": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "plume_convT2x2_wgrad_x3": (_I, [_P, _I, _P, _I, _P, _I, _P, c_size_t, _I, _I, _I, _I, _I, _P]),
    "plume_pack_conv3x3": (_I, [_P, _P, _P, _I, _I, _P]),
    "plume_pack_convT2x2": (_I, [_P, _P, _P, _I, _I, _P]),
    "plume_pack_blocks": (_I, [_I, _I, _I]),
    "plume_pack_batch": (_I, [_P, _I, _I, _P]),
    "plume_pad_channels": (_I, [_P, _I, _P, _I, _LL, _P]),
    "plume_bn_finalize": (_I, [_P, _P, _LL, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _I, _P]),
    "plume_bn_fold_eval": (_I, [_P, _P, _P, _P, _P, _F, _P, _P, _I, _P]),
    "plume_scale_shift_act": (_I, [_P, _I, _P, _P, _I, _P, _I, _LL, _I, _P]),
    "plume_scale_shift_act_pool": (_I, [_P, _I, _P, _P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _P]),
    "plume_maxpool2x2_fwd": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P]),
    "plume_maxpool2x2_bwd": (_I, [_P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P]),
    "plume_bn_bwd_reduce": (_I, [_P, _I, _P, _I, _P, _P, _P, _P, _I, _P, _P, _LL, _I, _P]),
    "plume_bn_bwd_apply": (_I, [_P, _I, _P, _I, _P, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P, _P, _I, _LL, _I, _P]),
    "plume_relu_bwd": (_I, [_P, _I, _P, _I, _P, _I, _P, _LL, _I, _P]),
    "plume_channel_sum": (_I, [_P, _I, _P, _LL, _I, _P]),
    "plume_head_fwd": (_I, [_P, _I, _P, _P, _P, _P, _P, _LL, _I, _P]),
    "plume_head_loss": (_I, [_P, _LL, _F, _F, _F, _P, _P]),
    "plume_head_bwd": (_I, [_P, _I, _P, _P, _P, _P, _F, _F, _F, _F, _P, _I, _P, _P, _LL, _I, _P]),
    "plume_adam": (_I, [_P, _P, _P, _P, _LL, _D, _D, _D, _D, _I, _F, _P]),
    "plume_adam_dev": (_I, [_P, _P, _P, _P, _LL, _P, _P]),
    "plume_pad_channels_x3": (_I, [_P, _I, _P, _I, _LL, _P]),
    "plume_scale_shift_act_x3": (_I, [_P, _I, _P, _P, _I, _P, _I, _LL, _I, _P]),
    "plume_scale_shift_act_pool_x3": (_I, [_P, _I, _P, _P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _P]),
    "plume_maxpool2x2_fwd_x3": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P]),
    "plume_maxpool2x2_bwd_x3": (_I, [_P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P]),
    "plume_bn_bwd_reduce_x3": (_I, [_P, _I, _P, _I, _P, _P, _P, _P, _I, _P, _P, _LL, _I, _P]),
    "plume_bn_bwd_apply_x3": (_I, [_P, _I, _P, _I, _P, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P, _P, _I, _LL, _I, _P]),
    "plume_relu_bwd_x3": (_I, [_P, _I, _P, _I, _P, _I, _P, _LL, _I, _P]),
    "plume_channel_sum_x3": (_I, [_P, _I, _P, _LL, _I, _P]),
    "plume_head_fwd_x3": (_I, [_P, _I, _P, _P, _P, _P, _P, _LL, _I, _P]),
    "plume_head_bwd_x3": (_I, [_P, _I, _P, _P, _P, _P, _F, _F, _F, _F, _P, _I, _P, _P, _LL, _I, _P]),
    "plume_extract_tiles_x3": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _P, _I, _P]),
    "plume_extract_tiles": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _P, _I, _P]),
    "plume_stitch_threshold": (_I, [_P, _P, _P, _I, _I, _I, _F, _P, _P, _I, _I, _P]),
    "plume_rasterize_hulls": (_I, [_P, _P, _P, _I, _P, _P, _I, _I, _I, _P, _P]),
    "plume_threshold_masks": (_I, [_P, _I, _I, _P, _I, _P, _P]),
    "plume_threshold_masks_f64": (_I, [_P, _I, _I, _P, _I, _P, _P]),
    "plume_label_components": (_I, [_P, _I, _I, _I, _P, _P, _P]),
    "plume_fire_extents": (_I, [_P, _P, _I, _I, _I, _P, _I, _I, _P, _P]),
    "plume_sweep_workspace_bytes": (c_size_t, [_I, _I, _I]),
    "plume_fill_nearest_workspace_bytes": (c_size_t, [_I, _I]),
    "plume_fill_nearest": (_I, [_P, _I, _I, _F, _P, c_size_t, _P, _P]),
    "plume_fill_nearest_f64": (_I, [_P, _I, _I, _D, _P, c_size_t, _P, _P]),
    "plume_threshold_mask_bits": (_I, [_P, _I, _I, _P, _I, _P, _P]),
    "plume_threshold_mask_bits_f64": (_I, [_P, _I, _I, _P, _I, _P, _P]),
    "plume_pack_mask_bits": (_I, [_P, _I, _I, _I, _P, _P]),
    "plume_bits_extents": (_I, [_P, _I, _I, _I, _P, _I, _I, _P, c_size_t, _P, _P]),
    "plume_fire_components": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _P, c_size_t, _P, _P, _P]),
    "plume_sweep_extents": (_I, [_P, _I, _I, _P, _I, _P, _I, _I, _P, c_size_t, _P, _P]),
    "plume_sweep_extents_f64": (_I, [_P, _I, _I, _P, _I, _P, _I, _I, _P, c_size_t, _P, _P]),
    "plume_maxpool2x2_bwd_bn": (_I, [_P, _I, _P, _P, _I, _P, _I, _P, _I, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _P]),
    "plume_maxpool2x2_bwd_bn_x3": (_I, [_P, _I, _P, _P, _I, _P, _I, _P, _I, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _P]),
    "plume_head_bwd_bn": (_I, [_P, _I, _P, _P, _P, _P, _F, _F, _F, _F, _P, _I, _P, _P, _P, _I, _P, _P, _P, _P, _I, _P, _P,
                               _LL, _I, _P]),
    "plume_head_bwd_bn_x3": (_I, [_P, _I, _P, _P, _P, _P, _F, _F, _F, _F, _P, _I, _P, _P, _P, _I, _P, _P, _P, _P, _I, _P,
                                  _P, _LL, _I, _P]),
    "plume_cast_f32_bf16": (_I, [_P, _P, _LL, _P]),
    "plume_cast_bf16_f32": (_I, [_P, _P, _LL, _P]),
    "plume_set_sm_margin": (None, [_I]),
    "plume_get_sm_margin": (_I, []),
    "plume_set_deterministic": (None, [_I]),
    "plume_get_deterministic": (_I, []),
    "plume_utm_zone_histogram": (_I, [_P, _LL, _P, _P]),
    "plume_sinusoidal_grid_latlon": (_I, [_D, _D, _D, _D, _I, _I, _D, _P, _P, _P]),
    "plume_utm_forward": (_I, [_P, _P, _LL, _I, _P, _P, _P]),
    "plume_utm_inverse": (_I, [_P, _P, _LL, _I, _P, _P, _P]),
    "plume_resample_workspace_bytes": (c_size_t, [_I, _D, _D, _D, _D, _D]),
    "plume_resample_nearest_index": (_I, [_P, _P, _I, _I, _D, _D, _D, _D, _I, _I, _D, _P, c_size_t, _P, _P]),
    "plume_gather_fill": (_I, [_P, _I, _P, _LL, _D, _P, _P]),
    "plume_locate_fires_workspace_bytes": (ctypes.c_size_t, [_I]),
    "plume_locate_fires": (_I, [_P, _P, _I, _I, _P, _P, _I, ctypes.c_double, _P, ctypes.c_size_t, _P, _P]),
}


class PlumeError(RuntimeError):
    """An entry point of libplume_b200.so reported failure."""


_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library once and attach the prototypes.  No fallback: a missing build is fatal."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PlumeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C kcl_ltss_bioatm_b200/csrc` (there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().plume_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise PlumeError(f"{what} failed (code {rc}): {last_error()}")


def ptr(t) -> c_void_p:
    """Device (or host) pointer of a torch tensor, None -> NULL."""
    return c_void_p(0) if t is None else c_void_p(t.data_ptr())


def current_stream() -> c_void_p:
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)
