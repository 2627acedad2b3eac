"""Thin tensor-level wrappers over the C ABI (``include/omnibiote_b200.h``).

PyTorch is used for device memory, streams and autograd bookkeeping only; every computation below is a call into
``libomnibiote_b200.so``. There is no CPU or eager fallback: non-CUDA / non-bf16 inputs raise.
"""
from __future__ import annotations

import torch

from . import _lib

EPI_PLAIN, EPI_RESID, EPI_GELU, EPI_GELU_BWD, EPI_RESID_DROPOUT, EPI_ROPE, EPI_ROWMASK = 0, 1, 2, 3, 5, 7, 8
EPI_GELU_DG, EPI_MUL, EPI_DELTA = 9, 10, 11

# gelu_mode 0: one rounding (TorchScript-fused execution on CUDA); 1: a bf16 rounding per primitive (eager CPU run of
# the same expression, which is what the CPU oracle does). See SURVEY Appendix A.2.
GELU_MODE = 0

# attention kernel selection: "auto" = tensor-core kernel for head_dim 128, generic CUDA-core kernel otherwise;
# "tc" / "simt" force one (tests, cross-checks). Both are CUDA kernels of this library; neither is a fallback to torch.
import os as _os

ATTN_IMPL = _os.environ.get("OBT_ATTN_IMPL", "auto")
# delta = rowsum(dO * O) of the attention backward from the epilogue of the GEMM that produces dO (EPI_DELTA) instead of
# a separate memory-bound pass; OBT_FUSE_ATTN_DELTA=0 restores the stand-alone kernel (A/B runs, tests)
FUSE_ATTN_DELTA = _os.environ.get("OBT_FUSE_ATTN_DELTA", "1") != "0"
# attention backward: the dK/dV kernel hands its dS tiles to a score-free dQ kernel through a bf16 scratch instead of a
# second evaluation of the scores and exponentials in a stand-alone dQ kernel; OBT_ATTN_DS_HANDOVER=0 = stand-alone dQ
ATTN_DS_HANDOVER = _os.environ.get("OBT_ATTN_DS_HANDOVER", "1") != "0"

_workspaces: dict = {}


# Instrumentation used by bench.py: LAUNCHES counts C-ABI calls (each enqueues at least one kernel of this library);
# PROFILE_GEMM, when a list, receives (flops, start_event, end_event) for every GEMM launch.
LAUNCHES = 0
PROFILE_GEMM = None


def _stream() -> int:
    global LAUNCHES
    LAUNCHES += 1
    return torch.cuda.current_stream().cuda_stream


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def _req(t: torch.Tensor, name: str, dtype=torch.bfloat16) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"omnibiote_b200: {name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"omnibiote_b200: {name} must be {dtype}, got {t.dtype}")


def workspace(key: str, numel: int, dtype, device, zero: bool = False) -> torch.Tensor:
    """Cached scratch buffer (grown on demand). ``zero`` buffers are zero-filled when (re)allocated only."""
    k = (key, dtype, device)
    buf = _workspaces.get(k)
    if buf is None or buf.numel() < numel:
        buf = torch.zeros(numel, dtype=dtype, device=device) if zero else torch.empty(numel, dtype=dtype, device=device)
        _workspaces[k] = buf
    return buf


def philox_args(device, increment: int = 4):
    """Reserve a slice of the device's default Philox stream (so torch.manual_seed and checkpoint RNG replay work)."""
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    seed = gen.initial_seed() & 0xFFFFFFFFFFFFFFFF
    off = gen.get_offset()
    gen.set_offset(off + increment)
    return seed, off


def _mat(t: torch.Tensor, name: str):
    """2-D view with unit inner stride -> (tensor, leading dimension)."""
    _req(t, name)
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError(f"omnibiote_b200: {name} must be 2-D with unit inner stride, got {tuple(t.shape)} / {t.stride()}")
    return t, t.stride(0)


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor | None = None, *, a_mn: bool = False, b_mn: bool = False,
         epilogue: int = EPI_PLAIN, aux_in: torch.Tensor | None = None, aux_out: torch.Tensor | None = None,
         drop_p: float = 0.0, seed: int = 0, offset: int = 0, allow_splitk: bool = True,
         rope: tuple | None = None, delta: tuple | None = None) -> torch.Tensor:
    """out[M,N] = epilogue(op(a) @ op(b)^T).  a: [M,K] (or [K,M] if a_mn); b: [N,K] (or [K,N] if b_mn).
    rope = (cos_tab, sin_tab | None, T, head_dim, n_cols) with epilogue=EPI_ROPE: rotary on the first n_cols columns.
    delta = (delta_out fp32 [B, N/128, T], T) with epilogue=EPI_DELTA: per-head rowsum(out * aux_in)."""
    a, lda = _mat(a, "gemm A")
    b, ldb = _mat(b, "gemm B")
    M, K = (a.shape[1], a.shape[0]) if a_mn else a.shape
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else b.shape
    if K != Kb:
        raise RuntimeError(f"omnibiote_b200: gemm K mismatch {K} vs {Kb}")
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    out, ldd = _mat(out, "gemm D")
    if tuple(out.shape) != (M, N):
        raise RuntimeError(f"omnibiote_b200: gemm D shape {tuple(out.shape)} != {(M, N)}")
    ld_ai = ld_ao = 0
    if epilogue == EPI_ROWMASK:
        if aux_in is None or aux_in.dtype != torch.uint8 or aux_in.numel() != M or not aux_in.is_contiguous():
            raise RuntimeError("omnibiote_b200: EPI_ROWMASK needs aux_in = contiguous uint8 row mask [M]")
    elif aux_in is not None:
        aux_in, ld_ai = _mat(aux_in, "gemm aux_in")
    if aux_out is not None:
        aux_out, ld_ao = _mat(aux_out, "gemm aux_out")
    ws, ws_elems = None, 0
    rope_T = 0
    if epilogue == EPI_DELTA:
        if delta is None or aux_in is None:
            raise RuntimeError("omnibiote_b200: EPI_DELTA needs aux_in (y) and delta=(fp32 [B, N/128, T], T)")
        ws, rope_T = delta
        _req(ws, "delta", torch.float32)
        ws_elems = ws.numel()
    elif allow_splitk and K >= 1024 and M * N <= 8 * 1024 * 1024:
        ws_elems = 8 * M * N
        ws = workspace("splitk", ws_elems, torch.float32, a.device)
    rope_cos = rope_sin = None
    rope_d = rope_cols = 0
    if epilogue == EPI_ROPE:
        if rope is None:
            raise RuntimeError("omnibiote_b200: EPI_ROPE needs rope=(cos, sin, T, head_dim, n_cols)")
        rope_cos, rope_sin, rope_T, rope_d, rope_cols = rope
        _req(rope_cos, "cos table", torch.float32)
        if rope_cos.shape[0] < rope_T:
            raise RuntimeError("omnibiote_b200: rotary table shorter than the sequence")
    lib = _lib.load()
    prof = PROFILE_GEMM
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    rc = lib.obt_gemm_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, lda, ldb, ldd, int(a_mn), int(b_mn),
                           epilogue, _ptr(aux_in), ld_ai, _ptr(aux_out), ld_ao, GELU_MODE, float(drop_p), seed, offset,
                           _ptr(ws), ws_elems, _ptr(rope_cos), _ptr(rope_sin), int(rope_T), int(rope_d), int(rope_cols),
                           _stream())
    _lib.check(rc, "obt_gemm_bf16")
    if prof is not None:
        ev1.record()
        prof.append((2.0 * M * N * K, ev0, ev1, f"{M}x{N}x{K} {'T' if a_mn else 'N'}{'T' if b_mn else 'N'} epi{epilogue}"))
    return out


def embed_fwd(idx: torch.Tensor, wte: torch.Tensor, drop_p: float = 0.0, seed: int = 0, offset: int = 0) -> torch.Tensor:
    _req(wte, "wte")
    _req(idx, "idx", torch.int64)
    idx = idx.contiguous()
    V, C = wte.shape
    M = idx.numel()
    out = torch.empty((M, C), dtype=torch.bfloat16, device=wte.device)
    err = workspace("embed_err", 1, torch.int32, wte.device, zero=True)
    rc = _lib.load().obt_embed_fwd(idx.data_ptr(), wte.data_ptr(), out.data_ptr(), M, C, V, float(drop_p), seed, offset,
                                   err.data_ptr(), _stream())
    _lib.check(rc, "obt_embed_fwd")
    return out


def check_embedding_ids(device) -> None:
    """Host check (synchronises) of the flag obt_embed_fwd raises for token ids outside [0, vocab): the kernel reads
    row 0 for such ids instead of faulting like nn.Embedding's device assert, so callers poll this at logging cadence."""
    err = _workspaces.get(("embed_err", torch.int32, device))
    if err is not None and int(err.item()) != 0:
        err.zero_()
        raise RuntimeError("omnibiote_b200: token id outside the vocabulary reached the embedding "
                           "(corrupt shard or wrong vocab_size?)")


def embed_bwd(idx: torch.Tensor, dout: torch.Tensor, dwte: torch.Tensor, accumulate: bool, drop_p: float = 0.0,
              seed: int = 0, offset: int = 0) -> None:
    _req(dout, "dout")
    _req(dwte, "dwte")
    idx = idx.contiguous()
    V, C = dwte.shape
    M = idx.numel()
    scratch = workspace("embed_scratch", V * C, torch.float32, dwte.device, zero=True)
    touched = workspace("embed_touched", V, torch.int32, dwte.device, zero=True)
    rc = _lib.load().obt_embed_bwd(idx.data_ptr(), dout.data_ptr(), dwte.data_ptr(), scratch.data_ptr(),
                                   touched.data_ptr(), M, C, V, int(accumulate), float(drop_p), seed, offset, _stream())
    _lib.check(rc, "obt_embed_bwd")


def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, eps: float = 1e-5, readout_div: float | None = None):
    """Returns (y, z, mean, rstd); z = rb(y / readout_div) when readout_div is given (MuReadout input scaling)."""
    _req(x, "ln x")
    _req(gamma, "ln weight")
    M, C = x.shape
    y = torch.empty_like(x)
    z = torch.empty_like(x) if readout_div is not None else None
    mean = torch.empty(M, dtype=torch.float32, device=x.device)
    rstd = torch.empty(M, dtype=torch.float32, device=x.device)
    rc = _lib.load().obt_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), y.data_ptr(), _ptr(z), mean.data_ptr(),
                                       rstd.data_ptr(), M, C, eps, float(readout_div or 1.0), _stream())
    _lib.check(rc, "obt_layernorm_fwd")
    return y, z, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dres=None, dgamma=None, accumulate_dgamma=False, dy_div: float = 1.0,
                  drop: tuple | None = None):
    """Returns (dx, dgamma) or, with drop=(p, seed, offset), (dx, dgamma, dx_drop).
    dx = rb(dres + rb(ln_bwd(rb(dy / dy_div)))); dx_drop = dropout(dx, p, seed, offset) (same mask as ops.dropout and
    the GEMM's EPI_RESID_DROPOUT): the gradient entering the dropped residual branch that produced x."""
    _req(dy, "ln dy")
    M, C = x.shape
    lib = _lib.load()
    dx = torch.empty_like(x)
    if dgamma is None:
        dgamma = torch.empty(C, dtype=torch.bfloat16, device=x.device)
        accumulate_dgamma = False
    # words 0..1 = grid-sync counters of the fused dgamma reduction (zero on entry, re-zeroed by the kernel)
    ws = workspace("ln_bwd", 32 + lib.obt_layernorm_bwd_workspace_rows() * max(C, 2048), torch.float32, x.device, zero=True)
    dx_drop, p, seed, off = None, 0.0, 0, 0
    if drop is not None and drop[0] > 0.0:
        p, seed, off = drop
        dx_drop = torch.empty_like(x)
    rc = lib.obt_layernorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                               _ptr(dres), dx.data_ptr(), dgamma.data_ptr(), int(accumulate_dgamma), ws.data_ptr(), M, C,
                               float(dy_div), _ptr(dx_drop), float(p), seed, off, _stream())
    _lib.check(rc, "obt_layernorm_bwd")
    if drop is not None:
        return dx, dgamma, dx_drop
    return dx, dgamma


def rope_(qkv: torch.Tensor, cos_tab: torch.Tensor, sin_tab: torch.Tensor | None, T: int, C: int, head_dim: int,
          inverse: bool = False) -> torch.Tensor:
    """In-place rotary / cosine-scale on the q and k column ranges of qkv [M, 3C]."""
    qkv, ld = _mat(qkv, "qkv")
    _req(cos_tab, "cos table", torch.float32)
    if cos_tab.shape[0] < T:
        raise RuntimeError("omnibiote_b200: rotary table shorter than the sequence")
    rc = _lib.load().obt_rope(qkv.data_ptr(), cos_tab.data_ptr(), _ptr(sin_tab), qkv.shape[0], T, C, head_dim, ld,
                              int(inverse), _stream())
    _lib.check(rc, "obt_rope")
    return qkv


def dropout(x: torch.Tensor, p: float, seed: int, offset: int, out: torch.Tensor | None = None) -> torch.Tensor:
    _req(x, "dropout input")
    if not x.is_contiguous():
        raise RuntimeError("omnibiote_b200: dropout input must be contiguous")
    if out is None:
        out = torch.empty_like(x)
    rc = _lib.load().obt_dropout(x.data_ptr(), out.data_ptr(), x.numel(), float(p), seed, offset, _stream())
    _lib.check(rc, "obt_dropout")
    return out


def scale_div(x: torch.Tensor, div: float) -> torch.Tensor:
    _req(x, "scale_div input")
    x = x.contiguous()
    out = torch.empty_like(x)
    rc = _lib.load().obt_scale_div(x.data_ptr(), out.data_ptr(), x.numel(), float(div), _stream())
    _lib.check(rc, "obt_scale_div")
    return out


def pool(emb: torch.Tensor, mode: str) -> torch.Tensor:
    _req(emb, "emb")
    emb = emb.contiguous()
    B, T, C = emb.shape
    lib = _lib.load()
    out = torch.empty((B, C), dtype=torch.bfloat16, device=emb.device)
    ws = workspace("pool", B * lib.obt_pool_splits(T) * C, torch.float32, emb.device)
    rc = lib.obt_pool(emb.data_ptr(), out.data_ptr(), ws.data_ptr(), B, T, C, {"mean": 0, "max": 1}[mode], _stream())
    _lib.check(rc, "obt_pool")
    return out


def pool_bwd(emb, pooled, dout, mode: str) -> torch.Tensor:
    B, T, C = emb.shape
    demb = torch.empty_like(emb)
    rc = _lib.load().obt_pool_bwd(emb.data_ptr(), pooled.data_ptr(), dout.contiguous().data_ptr(), demb.data_ptr(), B,
                                  T, C, {"mean": 0, "max": 1}[mode], _stream())
    _lib.check(rc, "obt_pool_bwd")
    return demb


class MaskSpec:
    """Additive attention mask as the kernels consume it: a bf16 tensor plus (batch, head, query) strides."""

    __slots__ = ("tensor", "msb", "msh", "msq", "row_lo", "row_hi", "_meta")

    def tile_meta(self):
        """(qmeta, kmeta) of the interval form (obt_attn_tile_meta), computed on first use and shared by every layer
        and attention kernel of the micro-batch; (None, None) for dense / absent masks."""
        if self.tensor is not None or self.row_lo is None:
            return None, None
        if self._meta is None:
            B, T = self.row_lo.shape
            nT = (T + 127) // 128
            qmeta = torch.empty((B, nT, 4), dtype=torch.int32, device=self.row_lo.device)
            kmeta = torch.empty((B, nT, 4), dtype=torch.int32, device=self.row_lo.device)
            rc = _lib.load().obt_attn_tile_meta(self.row_lo.data_ptr(), self.row_hi.data_ptr(), B, T, qmeta.data_ptr(),
                                                kmeta.data_ptr(), _stream())
            _lib.check(rc, "obt_attn_tile_meta")
            self._meta = (qmeta, kmeta)
        return self._meta

    def __init__(self, attn_mask: torch.Tensor | None, B: int, H: int, T: int, row_lo=None, row_hi=None):
        self.tensor = None
        self._meta = None
        self.msb = self.msh = self.msq = 0
        # optional interval form (int32 [B,T] each): key j visible to query (b,i) iff lo <= j < hi
        self.row_lo, self.row_hi = row_lo, row_hi
        if attn_mask is None:
            return
        m = attn_mask
        if m.dim() != 4 or m.shape[0] != B or m.shape[-2] != T or m.shape[-1] != T or m.shape[1] not in (1, H):
            raise RuntimeError(f"omnibiote_b200: attn_mask must be (b, n_head, t, t), got {tuple(m.shape)}")
        if not m.is_cuda:
            raise RuntimeError("omnibiote_b200: attn_mask must be a CUDA tensor")
        if m.dtype != torch.bfloat16 or m.stride(-1) != 1:
            # keep head-expanded (stride-0) views cheap: convert the un-expanded slice only
            if m.stride(1) == 0 or m.shape[1] == 1:
                m = m[:, :1].to(torch.bfloat16).contiguous().expand(-1, H, -1, -1)
            else:
                m = m.to(torch.bfloat16).contiguous()
        self.tensor = m
        self.msb = m.stride(0)
        self.msh = m.stride(1) if m.shape[1] > 1 else 0
        self.msq = m.stride(2)


def keep_words(T: int) -> int:
    return (T + 31) // 32


def attn_keep_mask(B: int, H: int, T: int, drop_p: float, seed: int, offset: int, device,
                   mask: "MaskSpec | None" = None) -> torch.Tensor:
    """Attention-dropout keep bits [B,H,T,ceil(T/32)] (int32 words, bit layout: csrc/dropmask.cuh) for one layer and
    micro-batch, drawn once and read by the forward, dQ and dK/dV kernels. With an interval `mask`, words of a row that
    lie entirely outside its visible interval are stored as all-ones instead of being drawn."""
    keep = torch.empty((B, H, T, keep_words(T)), dtype=torch.int32, device=device)
    use_iv = mask is not None and mask.tensor is None and mask.row_lo is not None
    rc = _lib.load().obt_attn_keep_mask(keep.data_ptr(), B, H, T, float(drop_p), seed, offset,
                                        _ptr(mask.row_lo) if use_iv else 0, _ptr(mask.row_hi) if use_iv else 0,
                                        _stream())
    _lib.check(rc, "obt_attn_keep_mask")
    return keep


def keep_mask_to_bool(keep: torch.Tensor, T: int) -> torch.Tensor:
    """Unpack the keep bits into a bool [B,H,T,T] tensor (tests / debugging)."""
    e = torch.arange(32, device=keep.device)
    pos = 8 * (e & 3) + 7 - (e >> 2)
    bits = (keep.unsqueeze(-1) >> pos) & 1
    return bits.flatten(-2)[..., :T].bool()


def attention_fwd(qkv: torch.Tensor, B: int, T: int, H: int, d: int, scale: float, mask: MaskSpec, drop_p: float,
                  keep: torch.Tensor | None = None, impl: str = "auto"):
    """qkv [M,3C] (post-rotary) -> (y [M,C], lse [B,H,T,2]). keep: attn_keep_mask(...) when drop_p > 0."""
    if drop_p > 0.0 and keep is None:
        raise RuntimeError("omnibiote_b200: attention dropout needs the keep mask (ops.attn_keep_mask)")
    qkv, ld = _mat(qkv, "qkv")
    C = H * d
    M = B * T
    y = torch.empty((M, C), dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty((B, H, T, 2), dtype=torch.float32, device=qkv.device)
    esz = 2
    q, k, v = qkv.data_ptr(), qkv.data_ptr() + C * esz, qkv.data_ptr() + 2 * C * esz
    lib = _lib.load()
    use_iv = mask.tensor is None and mask.row_lo is not None
    if impl == "auto":
        impl = ATTN_IMPL
    if impl == "auto":
        dense_ok = mask.tensor is None or (T % 8 == 0 and mask.msq % 8 == 0 and mask.msb % 8 == 0 and mask.msh % 8 == 0
                                           and mask.tensor.data_ptr() % 16 == 0)
        impl = "tc" if (d == 128 and dense_ok) else "simt"
    if impl == "tc":
        qmeta = mask.tile_meta()[0] if use_iv else None
        rc = lib.obt_attn_tc_fwd(qkv.data_ptr(), ld, _ptr(mask.tensor), mask.msb, mask.msh, mask.msq,
                                 _ptr(mask.row_lo) if use_iv else 0, _ptr(mask.row_hi) if use_iv else 0, y.data_ptr(), C,
                                 lse.data_ptr(), B, H, T, d, scale, float(drop_p), _ptr(keep), _ptr(qmeta), _stream())
        _lib.check(rc, "obt_attn_tc_fwd")
        return y, lse
    rc = lib.obt_attn_simt_fwd(q, k, v, ld, _ptr(mask.tensor), mask.msb, mask.msh, mask.msq,
                               _ptr(mask.row_lo) if use_iv else 0, _ptr(mask.row_hi) if use_iv else 0, y.data_ptr(), C,
                               lse.data_ptr(), B, H, T, d, scale, float(drop_p), _ptr(keep), _stream())
    _lib.check(rc, "obt_attn_simt_fwd")
    return y, lse


def attention_bwd(qkv, y, dy, lse, B, T, H, d, scale, mask: MaskSpec, drop_p, keep=None, impl: str = "auto",
                  rope: tuple | None = None, delta: torch.Tensor | None = None):
    """Returns dqkv [M,3C]: the gradient w.r.t. the post-rotary q,k and v, or, with rope=(cos_tab, sin_tab | None),
    w.r.t. the PRE-rotary c_attn output (the rotary adjoint is applied in the kernels' epilogues).
    keep: the forward's keep mask. delta: fp32 [B,H,T] = rowsum(dy * y) per head when the GEMM that produced dy already
    computed it (ops.gemm(..., epilogue=EPI_DELTA)); tensor-core path only."""
    if drop_p > 0.0 and keep is None:
        raise RuntimeError("omnibiote_b200: attention dropout needs the keep mask (ops.attn_keep_mask)")
    qkv, ld = _mat(qkv, "qkv")
    dy, lddy = _mat(dy, "attention dy")
    C = H * d
    M = B * T
    dqkv = torch.empty((M, 3 * C), dtype=torch.bfloat16, device=qkv.device)
    delta_ready = delta is not None
    if delta is None:
        delta = torch.empty((B, H, T), dtype=torch.float32, device=qkv.device)
    elif tuple(delta.shape) != (B, H, T) or delta.dtype != torch.float32 or not delta.is_contiguous():
        raise RuntimeError("omnibiote_b200: precomputed delta must be a contiguous fp32 [B, H, T] tensor")
    esz = 2
    q, k, v = qkv.data_ptr(), qkv.data_ptr() + C * esz, qkv.data_ptr() + 2 * C * esz
    dq, dk, dv = dqkv.data_ptr(), dqkv.data_ptr() + C * esz, dqkv.data_ptr() + 2 * C * esz
    use_iv = mask.tensor is None and mask.row_lo is not None
    if impl == "auto":
        impl = ATTN_IMPL
    if impl == "auto":
        impl = "tc" if d == 128 else "simt"
    if impl == "tc":
        qmeta, kmeta = mask.tile_meta() if use_iv else (None, None)
        ds_scratch = None
        if ATTN_DS_HANDOVER:
            ds_scratch = workspace("attn_ds", B * H * T * ((T + 63) // 64 * 64), torch.bfloat16, qkv.device)
        rc = _lib.load().obt_attn_tc_bwd(qkv.data_ptr(), ld, _ptr(mask.tensor), mask.msb, mask.msh, mask.msq,
                                         _ptr(mask.row_lo) if use_iv else 0, _ptr(mask.row_hi) if use_iv else 0,
                                         y.data_ptr(), C, dy.data_ptr(), lddy, lse.data_ptr(), delta.data_ptr(),
                                         int(delta_ready), dqkv.data_ptr(), 3 * C, B, H, T, d, scale, float(drop_p),
                                         _ptr(keep),
                                         _ptr(rope[0]) if rope else 0, _ptr(rope[1]) if rope else 0, _ptr(qmeta),
                                         _ptr(kmeta), _ptr(ds_scratch), _stream())
        _lib.check(rc, "obt_attn_tc_bwd")
        return dqkv
    rc = _lib.load().obt_attn_simt_bwd(q, k, v, ld, _ptr(mask.tensor), mask.msb, mask.msh, mask.msq,
                                       _ptr(mask.row_lo) if use_iv else 0, _ptr(mask.row_hi) if use_iv else 0,
                                       y.data_ptr(), C, dy.data_ptr(), lddy, lse.data_ptr(), delta.data_ptr(), dq, dk,
                                       dv, 3 * C, B, H, T, d, scale, float(drop_p), _ptr(keep), _stream())
    _lib.check(rc, "obt_attn_simt_bwd")
    if rope is not None:
        rope_(dqkv, rope[0], rope[1], T, C, d, inverse=True)
    return dqkv


def compact_rows(row_mask: torch.Tensor, targets: torch.Tensor, cap: int):
    """Rows with mask != 0, in row order: (idx int32 [cap] (-1 padded), targets_c int64 [cap], valid_c uint8 [cap],
    meta int32 [2] = {count, overflow})."""
    row_mask = row_mask.reshape(-1).to(torch.uint8).contiguous()
    targets = targets.reshape(-1).contiguous()
    _req(targets, "targets", torch.int64)
    dev = row_mask.device
    idx = torch.empty(cap, dtype=torch.int32, device=dev)
    tgt = torch.empty(cap, dtype=torch.int64, device=dev)
    valid = torch.empty(cap, dtype=torch.uint8, device=dev)
    meta = torch.empty(2, dtype=torch.int32, device=dev)
    rc = _lib.load().obt_compact_rows(row_mask.data_ptr(), targets.data_ptr(), row_mask.numel(), int(cap), idx.data_ptr(),
                                      tgt.data_ptr(), valid.data_ptr(), meta.data_ptr(), _stream())
    _lib.check(rc, "obt_compact_rows")
    return idx, tgt, valid, meta


def gather_rows(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    src, lds = _mat(src, "gather source")
    out = torch.empty((idx.numel(), src.shape[1]), dtype=torch.bfloat16, device=src.device)
    rc = _lib.load().obt_gather_rows(src.data_ptr(), lds, idx.data_ptr(), out.data_ptr(), out.shape[1], idx.numel(),
                                     src.shape[1], _stream())
    _lib.check(rc, "obt_gather_rows")
    return out


def scatter_rows(src: torch.Tensor, idx: torch.Tensor, M: int) -> torch.Tensor:
    """out [M, C] = 0 except out[idx[s]] = src[s]."""
    src, lds = _mat(src, "scatter source")
    out = torch.empty((M, src.shape[1]), dtype=torch.bfloat16, device=src.device)
    rc = _lib.load().obt_scatter_rows(src.data_ptr(), lds, idx.data_ptr(), out.data_ptr(), out.shape[1], M, idx.numel(),
                                      src.shape[1], _stream())
    _lib.check(rc, "obt_scatter_rows")
    return out


def ce_fwd(logits: torch.Tensor, targets: torch.Tensor, row_mask: torch.Tensor | None, n_acc: float):
    """Returns (scalars[4] fp32 device: loss, count, dloss/dCE; lse [M]; tok_loss [M])."""
    logits, ld = _mat(logits, "logits")
    M, V = logits.shape
    dev = logits.device
    lse = torch.empty(M, dtype=torch.float32, device=dev)
    tok = torch.empty(M, dtype=torch.float32, device=dev)
    scalars = torch.empty(4, dtype=torch.float32, device=dev)
    if row_mask is not None:
        row_mask = row_mask.reshape(-1).to(torch.uint8).contiguous()
    targets = targets.reshape(-1).contiguous()
    rc = _lib.load().obt_ce_fwd(logits.data_ptr(), ld, targets.data_ptr(), _ptr(row_mask), lse.data_ptr(),
                                tok.data_ptr(), scalars.data_ptr(), M, V, float(n_acc), _stream())
    _lib.check(rc, "obt_ce_fwd")
    return scalars, lse, tok, row_mask, targets


def ce_bwd_(logits, targets, row_mask, lse, scalars, upstream: float = 1.0, unmasked_rows_zero: bool = False,
            upstream_dev: torch.Tensor | None = None):
    """In place: logits <- d loss / d logits for an incoming d loss = upstream * (upstream_dev or 1); upstream_dev is a
    bf16 device scalar (autograd's gradient of the bf16 loss), read by the kernel without a host synchronisation."""
    logits, ld = _mat(logits, "logits")
    M, V = logits.shape
    if upstream_dev is not None:
        _req(upstream_dev, "upstream gradient")
        if upstream_dev.numel() != 1:
            raise RuntimeError("omnibiote_b200: the upstream gradient of the loss must be a scalar")
    rc = _lib.load().obt_ce_bwd(logits.data_ptr(), ld, targets.data_ptr(), _ptr(row_mask), lse.data_ptr(),
                                scalars.data_ptr(), float(upstream), _ptr(upstream_dev), M, V,
                                int(unmasked_rows_zero and row_mask is not None), _stream())
    _lib.check(rc, "obt_ce_bwd")
    return logits


def doc_mask_intervals(ids: torch.Tensor, eos_token: int = 3, padding: bool = False):
    """Per-token visible key interval of the reference's create_attention_mask (train_encoder.py:25-57)."""
    _req(ids, "ids", torch.int64)
    ids = ids.contiguous()
    B, T = ids.shape
    lo = torch.empty((B, T), dtype=torch.int32, device=ids.device)
    hi = torch.empty((B, T), dtype=torch.int32, device=ids.device)
    rc = _lib.load().obt_doc_mask_intervals(ids.data_ptr(), lo.data_ptr(), hi.data_ptr(), B, T, int(eos_token),
                                            int(padding), _stream())
    _lib.check(rc, "obt_doc_mask_intervals")
    return lo, hi


def pad_mask_intervals(ids: torch.Tensor, pad_token: int = 1):
    """Interval form of evals/gue.py:15-21 pad_attn."""
    _req(ids, "ids", torch.int64)
    ids = ids.contiguous()
    B, T = ids.shape
    lo = torch.empty((B, T), dtype=torch.int32, device=ids.device)
    hi = torch.empty((B, T), dtype=torch.int32, device=ids.device)
    rc = _lib.load().obt_pad_mask_intervals(ids.data_ptr(), lo.data_ptr(), hi.data_ptr(), B, T, int(pad_token), _stream())
    _lib.check(rc, "obt_pad_mask_intervals")
    return lo, hi


def mask_from_intervals(lo: torch.Tensor, hi: torch.Tensor) -> torch.Tensor:
    """Dense additive bf16 (B,T,T) mask with values {0, -1e9}."""
    B, T = lo.shape
    mask = torch.empty((B, T, T), dtype=torch.bfloat16, device=lo.device)
    rc = _lib.load().obt_mask_from_intervals(lo.data_ptr(), hi.data_ptr(), mask.data_ptr(), B, T, _stream())
    _lib.check(rc, "obt_mask_from_intervals")
    return mask


def mask_compress(mask: torch.Tensor):
    """Dense additive bf16 mask (B,T,T) or (B,H,T,T) head-broadcast -> (lo, hi, not_interval_flag[int32 device])."""
    _req(mask, "mask")
    if mask.dim() == 4:
        mask = mask[:, 0]
    if mask.stride(-1) != 1:
        mask = mask.contiguous()
    B, T, _ = mask.shape
    lo = torch.empty((B, T), dtype=torch.int32, device=mask.device)
    hi = torch.empty((B, T), dtype=torch.int32, device=mask.device)
    flag = torch.zeros(1, dtype=torch.int32, device=mask.device)
    rc = _lib.load().obt_mask_compress(mask.data_ptr(), mask.stride(0), mask.stride(1), lo.data_ptr(), hi.data_ptr(),
                                       flag.data_ptr(), B, T, _stream())
    _lib.check(rc, "obt_mask_compress")
    return lo, hi, flag
