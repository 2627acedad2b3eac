"""CPU: the C-ABI shared library builds for sm_100a, loads without a GPU and exports every symbol that
include/omnibiote_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "omnibiote_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(obt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path_entry_points():
    syms = declared_symbols()
    for must in ["obt_gemm_bf16", "obt_attn_tc_fwd", "obt_attn_tc_bwd", "obt_layernorm_fwd", "obt_layernorm_bwd",
                 "obt_embed_fwd", "obt_embed_bwd", "obt_rope", "obt_ce_fwd", "obt_ce_bwd", "obt_pool", "obt_adamw_step",
                 "obt_grad_norm", "obt_doc_mask_intervals", "obt_mlm_mask"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    from omnibiote_b200 import _lib, build
    lib = _lib.load()
    raw = ctypes.CDLL(str(build.LIB_PATH))
    for name in declared_symbols():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    for name in _lib.SIGNATURES:
        assert name in declared_symbols(), f"{name} bound in Python but not declared in the header"
    assert lib.obt_version() >= 100
    assert lib.obt_opt_meta_bytes() == 48


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md, 'What proves a Blackwell-native kernel')."""
    import shutil
    import subprocess
    from omnibiote_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(build.LIB_PATH)], capture_output=True, text=True).stdout
    for mnemonic in ["UTCHMMA", "LDTM", "UTMALDG", "UTCBAR"]:
        assert mnemonic in sass, mnemonic
    assert "HGMMA" not in sass


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "omnibiote_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(import|from)\s+[^\n]*oracle", src, flags=re.M), fn
            assert "sys.path" not in src, fn


def test_no_cpu_fallback():
    import torch
    from omnibiote_b200 import ops
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.gemm(a, a)
    with pytest.raises(RuntimeError):
        ops.layernorm_fwd(a, a[0])


def test_ctypes_signatures_match_the_header_prototypes():
    """Argument count and coarse type class (pointer / 64-bit int / 32-bit int / float / double) of every ctypes
    binding against the C prototype in include/omnibiote_b200.h: an ABI drift between the two corrupts arguments
    silently."""
    import ctypes as ct
    from omnibiote_b200 import _lib
    text = open(os.path.join(ROOT, "include", "omnibiote_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = dict(re.findall(r"\b(obt_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text))

    def c_class(arg):
        arg = arg.strip()
        if arg in ("void", ""):
            return None
        if "*" in arg or "cudaStream_t" in arg:
            return "ptr"
        if "unsigned long long" in arg or "long long" in arg:
            return "i64"
        if "double" in arg:
            return "f64"
        if "float" in arg:
            return "f32"
        if "int" in arg:
            return "i32"
        raise AssertionError(f"unclassified C argument: {arg!r}")

    def py_class(t):
        if t in (ct.c_void_p, ct.c_char_p):
            return "ptr"
        if t in (ct.c_longlong, ct.c_ulonglong):
            return "i64"
        if t is ct.c_int:
            return "i32"
        if t is ct.c_float:
            return "f32"
        if t is ct.c_double:
            return "f64"
        raise AssertionError(f"unclassified ctypes type: {t!r}")

    for name, (_res, args) in _lib.SIGNATURES.items():
        assert name in protos, name
        want = [c for c in (c_class(a) for a in protos[name].split(",")) if c is not None]
        got = [py_class(t) for t in args]
        assert got == want, (name, got, want)
