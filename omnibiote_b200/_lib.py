"""ctypes binding of ``libomnibiote_b200.so`` (the C ABI declared in ``include/omnibiote_b200.h``).

There is no CPU fallback: if the library is missing it is built with nvcc; if that fails, importing raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_ulonglong, c_void_p
from pathlib import Path

from . import build as _build

_LIB = None

vp, i32, i64, u64, f32 = c_void_p, c_int, c_longlong, c_ulonglong, c_float

# name -> (restype, argtypes). Must list every symbol declared in include/omnibiote_b200.h.
SIGNATURES = {
    "obt_last_error": (c_char_p, []),
    "obt_version": (i32, []),
    "obt_clear_descriptor_cache": (None, []),
    "obt_gemm_set_cta_group": (None, [i32]),
    "obt_gemm_workspace_elems": (i64, [i64, i64, i64]),
    "obt_gemm_bf16": (i32, [vp, vp, vp, i64, i64, i64, i64, i64, i64, i32, i32, i32, vp, i64, vp, i64, i32, f32, u64,
                            u64, vp, i64, vp, vp, i32, i32, i32, vp]),
    "obt_embed_fwd": (i32, [vp, vp, vp, i64, i32, i32, f32, u64, u64, vp, vp]),
    "obt_embed_bwd": (i32, [vp, vp, vp, vp, vp, i64, i32, i32, i32, f32, u64, u64, vp]),
    "obt_layernorm_fwd": (i32, [vp, vp, vp, vp, vp, vp, i64, i32, f32, f32, vp]),
    "obt_layernorm_bwd_workspace_rows": (i32, []),
    "obt_layernorm_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, i64, i32, f32, vp, f32, u64, u64, vp]),
    "obt_rope": (i32, [vp, vp, vp, i64, i32, i32, i32, i64, i32, vp]),
    "obt_dropout": (i32, [vp, vp, i64, f32, u64, u64, vp]),
    "obt_scale_div": (i32, [vp, vp, i64, f32, vp]),
    "obt_pool_splits": (i32, [i32]),
    "obt_pool": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
    "obt_pool_bwd": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "obt_attn_keep_mask": (i32, [vp, i32, i32, i32, f32, u64, u64, vp, vp, vp]),
    "obt_attn_simt_fwd": (i32, [vp, vp, vp, i64, vp, i64, i64, i64, vp, vp, vp, i64, vp, i32, i32, i32, i32, f32, f32,
                                vp, vp]),
    "obt_attn_simt_bwd": (i32, [vp, vp, vp, i64, vp, i64, i64, i64, vp, vp, vp, i64, vp, i64, vp, vp, vp, vp, vp, i64,
                                i32, i32, i32, i32, f32, f32, vp, vp]),
    "obt_attn_tc_fwd": (i32, [vp, i64, vp, i64, i64, i64, vp, vp, vp, i64, vp, i32, i32, i32, i32, f32, f32, vp, vp, vp]),
    "obt_attn_tile_meta": (i32, [vp, vp, i32, i32, vp, vp, vp]),
    "obt_attn_tc_bwd": (i32, [vp, i64, vp, i64, i64, i64, vp, vp, vp, i64, vp, i64, vp, vp, i32, vp, i64, i32, i32, i32,
                              i32, f32, f32, vp, vp, vp, vp, vp, vp, vp]),
    "obt_doc_mask_intervals": (i32, [vp, vp, vp, i32, i32, i64, i32, vp]),
    "obt_pad_mask_intervals": (i32, [vp, vp, vp, i32, i32, i64, vp]),
    "obt_mask_from_intervals": (i32, [vp, vp, vp, i32, i32, vp]),
    "obt_mask_compress": (i32, [vp, i64, i64, vp, vp, vp, i32, i32, vp]),
    "obt_mlm_mask": (i32, [vp, vp, vp, i64, f32, u64, u64, i64, i64, i64, vp, vp]),
    "obt_compact_rows": (i32, [vp, vp, i64, i32, vp, vp, vp, vp, vp]),
    "obt_gather_rows": (i32, [vp, i64, vp, vp, i64, i32, i32, vp]),
    "obt_scatter_rows": (i32, [vp, i64, vp, vp, i64, i64, i32, i32, vp]),
    "obt_ce_fwd": (i32, [vp, i64, vp, vp, vp, vp, vp, i64, i32, f32, vp]),
    "obt_ce_bwd": (i32, [vp, i64, vp, vp, vp, vp, f32, vp, i64, i32, i32, vp]),
    "obt_opt_chunk_elems": (i32, []),
    "obt_opt_meta_bytes": (i32, []),
    "obt_grad_norm": (i32, [vp, vp, vp, i32, f32, f32, vp, vp, vp]),
    "obt_adamw_step": (i32, [vp, vp, vp, i32, vp, vp, f32, f32, c_double, c_double, c_double, i32, i32, vp]),
}


def library_path() -> Path:
    return _build.LIB_PATH


def load(rebuild: bool = False):
    """Load (building first if needed) the CUDA library and attach argument types."""
    global _LIB
    if _LIB is not None and not rebuild:
        return _LIB
    if rebuild or _build.needs_build():
        if os.environ.get("OBT_NO_BUILD") and _build.LIB_PATH.exists():
            pass
        else:
            _build.build(force=rebuild)
    lib = ctypes.CDLL(str(_build.LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # raises AttributeError when the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


class ObtError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().obt_last_error()
        raise ObtError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
