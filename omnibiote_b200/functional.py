"""Autograd glue for the encoder hot path: each ``torch.autograd.Function`` below owns one fused forward/backward
schedule of C-ABI kernel calls (``ops.py``). Torch only records the graph between them, so the evals'
``loss.backward()`` and stock optimizers keep working on the drop-in module (``model.py``).

Reference arithmetic followed (rounding points as in SURVEY Appendix D):
  Block            training/model.py:170-181  (pre-LN residual block; SelfAttention :98-152; MLP :162-168)
  embedding        training/model.py:241-242
  ln_f / MuReadout training/model.py:248-254 + mup MuReadout.forward
  MLM loss         training/train_encoder.py:301-305
"""
from __future__ import annotations

import contextlib

import torch

from . import ops

# When enabled (by the trainer), weight gradients are accumulated by the GEMM epilogue straight into the existing
# ``param.grad`` buffer (bf16 ``+=`` like autograd's accumulation) and autograd receives ``None`` for them.
_DIRECT_GRAD = False


@contextlib.contextmanager
def direct_grad_accumulation(enabled: bool = True):
    global _DIRECT_GRAD
    prev, _DIRECT_GRAD = _DIRECT_GRAD, enabled
    try:
        yield
    finally:
        _DIRECT_GRAD = prev


# Optional sink (parallel.FlatGradBuckets) told which parameters' gradients just became final, so their bucket can be
# all-reduced while the rest of the backward is still running.
_GRAD_SINK = None


def set_grad_sink(sink) -> None:
    global _GRAD_SINK
    _GRAD_SINK = sink


def _grads_final(params) -> None:
    if _GRAD_SINK is not None:
        _GRAD_SINK.notify(params)


# Dropout replays produced ahead of time by the LayerNorm backward of the NEXT stage (block l+1's ln_1, or ln_f in the
# head): block l's backward needs dropout(dx2, its mlp seeds) as the operand of its first two GEMMs, and the kernel that
# writes dx2 can emit it in the same pass (ops.layernorm_bwd(drop=...)). Keyed by the data pointer of dx2; the entry
# holds dx2 itself, so the address cannot be recycled while the entry exists. A miss (gradient accumulated by autograd,
# foreign caller, ...) falls back to the stand-alone dropout kernel. Cleared at the start of every forward.
_DROP_STASH: dict = {}


def clear_drop_stash() -> None:
    _DROP_STASH.clear()


def _stash_dropped(dx: torch.Tensor, dx_drop, seeds) -> None:
    if dx_drop is not None:
        _DROP_STASH[dx.data_ptr()] = (dx, dx_drop, seeds)


def _take_dropped(dx: torch.Tensor, seeds):
    ent = _DROP_STASH.pop(dx.data_ptr(), None)
    if ent is None:
        return None
    src, dropped, ent_seeds = ent
    if ent_seeds != seeds or src.shape != dx.shape or src._version != dx._version:
        return None
    return dropped


def _wgrad(dy2d: torch.Tensor, x2d: torch.Tensor, param):
    """dW[N,K] = dy[M,N]^T @ x[M,K] (both operands MN-major for the tensor cores, no transposes materialised)."""
    if _DIRECT_GRAD and param is not None and param.grad is not None:
        ops.gemm(dy2d, x2d, out=param.grad, a_mn=True, b_mn=True, epilogue=ops.EPI_RESID, aux_in=param.grad)
        return None
    return ops.gemm(dy2d, x2d, a_mn=True, b_mn=True)


def _vec_grad(g: torch.Tensor, param):
    if _DIRECT_GRAD and param is not None and param.grad is not None:
        return None  # already accumulated in place by the kernel
    return g


class EmbedFunction(torch.autograd.Function):
    """x0 = dropout(wte[idx])  (model.py:241-242). Output [M, C]."""

    @staticmethod
    def forward(ctx, idx, wte, p_drop, training):
        p = float(p_drop) if training else 0.0
        seed = off = 0
        if p > 0.0:
            seed, off = ops.philox_args(wte.device, 4)
        out = ops.embed_fwd(idx, wte, p, seed, off)
        ctx.save_for_backward(idx)
        ctx.meta = (p, seed, off, wte.shape)
        ctx.param = wte
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        p, seed, off, shape = ctx.meta
        w = ctx.param
        dout = dout.contiguous()
        if _DIRECT_GRAD and w.grad is not None:
            ops.embed_bwd(idx, dout, w.grad, True, p, seed, off)
            _grads_final((w,))
            return None, None, None, None
        dw = torch.empty(shape, dtype=torch.bfloat16, device=dout.device)
        ops.embed_bwd(idx, dout, dw, False, p, seed, off)
        return None, dw, None, None


class BlockFunction(torch.autograd.Function):
    """x -> x + attn(ln_1(x)) -> (+ mlp(ln_2(.)))  (model.py:170-181). x is [M, C] with M = B*T."""

    @staticmethod
    def forward(ctx, x, g1, w_qkv, w_o, g2, w_fc, w_pr, cos_tab, sin_tab, mask, B, T, H, p_drop, training, seeds=None,
                up_drop=None):
        """seeds: three (seed, offset) pairs (attention P, attention residual, MLP residual) drawn by the caller, or
        None to draw them here. up_drop = (p, seed, offset) of the dropout that produced `x` in the previous block
        (its MLP residual dropout): this block's backward then also emits the replayed gradient for that block."""
        M, C = x.shape
        d = C // H
        p = float(p_drop) if training else 0.0
        dev = x.device
        if p > 0.0:
            if seeds is None:
                seeds = [ops.philox_args(dev, 4) for _ in range(3)]  # attn P, attn resid, mlp resid
        else:
            seeds = [(0, 0)] * 3
        scale = 8.0 / C  # model.py:119,135: 8 / n_embd, independent of n_head

        h1, _, mean1, rstd1 = ops.layernorm_fwd(x, g1)
        # c_attn with the rotary embedding of q and k fused into the GEMM epilogue (model.py:102-108)
        qkv = ops.gemm(h1, w_qkv, epilogue=ops.EPI_ROPE, rope=(cos_tab, sin_tab, T, d, 2 * C))
        keep = ops.attn_keep_mask(B, H, T, p, *seeds[0], dev, mask) if p > 0.0 else None
        y, lse = ops.attention_fwd(qkv, B, T, H, d, scale, mask, p, keep)
        if p > 0.0:
            x1 = ops.gemm(y, w_o, epilogue=ops.EPI_RESID_DROPOUT, aux_in=x, drop_p=p, seed=seeds[1][0], offset=seeds[1][1])
        else:
            x1 = ops.gemm(y, w_o, epilogue=ops.EPI_RESID, aux_in=x)
        h2, _, mean2, rstd2 = ops.layernorm_fwd(x1, g2)
        # u = gelu'(c_fc output), not the pre-activation: one erf evaluation in the forward epilogue serves both, and the
        # backward's epilogue becomes a single multiply (GELU_MODE 1 keeps the eager-rounding pair 2 + 3)
        u = torch.empty((M, w_fc.shape[0]), dtype=torch.bfloat16, device=dev)
        fused_dg = ops.GELU_MODE == 0
        g = ops.gemm(h2, w_fc, epilogue=ops.EPI_GELU_DG if fused_dg else ops.EPI_GELU, aux_out=u)
        if p > 0.0:
            x2 = ops.gemm(g, w_pr, epilogue=ops.EPI_RESID_DROPOUT, aux_in=x1, drop_p=p, seed=seeds[2][0], offset=seeds[2][1])
        else:
            x2 = ops.gemm(g, w_pr, epilogue=ops.EPI_RESID, aux_in=x1)

        ctx.save_for_backward(x, g1, w_qkv, w_o, g2, w_fc, w_pr, cos_tab, sin_tab, mean1, rstd1, h1, qkv, y, lse, x1,
                              mean2, rstd2, h2, u, g)
        ctx.mask = mask
        ctx.keep = keep
        ctx.meta = (B, T, H, d, p, seeds, scale)
        ctx.up_drop = up_drop if (up_drop is not None and up_drop[0] > 0.0) else None
        ctx.fused_dg = fused_dg
        ctx.params = (g1, w_qkv, w_o, g2, w_fc, w_pr)
        return x2

    @staticmethod
    def backward(ctx, dx2):
        (x, g1, w_qkv, w_o, g2, w_fc, w_pr, cos_tab, sin_tab, mean1, rstd1, h1, qkv, y, lse, x1, mean2, rstd2, h2, u,
         g) = ctx.saved_tensors
        B, T, H, d, p, seeds, scale = ctx.meta
        pg1, pqkv, po, pg2, pfc, ppr = ctx.params
        C = x.shape[1]
        dx2 = dx2.contiguous()
        direct = _DIRECT_GRAD

        # ---- MLP branch: x2 = x1 + dropout(g @ Wpr^T)
        d_dd = dx2
        if p > 0.0:
            # usually already written by the LayerNorm backward that produced dx2 (next block's ln_1 or the head's ln_f)
            d_dd = _take_dropped(dx2, (p, *seeds[2]))
            if d_dd is None:
                d_dd = ops.dropout(dx2, p, *seeds[2])
        dw_pr = _wgrad(d_dd, g, ppr)
        du = ops.gemm(d_dd, w_pr, b_mn=True, epilogue=ops.EPI_MUL if ctx.fused_dg else ops.EPI_GELU_BWD, aux_in=u)
        dw_fc = _wgrad(du, h2, pfc)
        dh2 = ops.gemm(du, w_fc, b_mn=True)
        acc2 = direct and pg2.grad is not None
        # the same pass also writes dropout(dx1) with the attention-residual mask: the operand of the next two GEMMs
        dx1, dg2, d_a = ops.layernorm_bwd(dh2, x1, g2, mean2, rstd2, dres=dx2, dgamma=pg2.grad if acc2 else None,
                                          accumulate_dgamma=acc2, drop=(p, *seeds[1]))
        if d_a is None:
            d_a = dx1

        # ---- attention branch: x1 = x + dropout(y @ Wo^T)
        dw_o = _wgrad(d_a, y, po)
        # dy = d_a Wo; with the tensor-core attention (head_dim 128) the same GEMM also emits delta = rowsum(dy * y)
        # per head from the tile it holds, which the attention backward would otherwise compute in a pass of its own
        delta = None
        if d == 128 and ops.ATTN_IMPL in ("auto", "tc") and ops.FUSE_ATTN_DELTA:
            delta = torch.empty((B, H, T), dtype=torch.float32, device=x.device)
            dy = ops.gemm(d_a, w_o, b_mn=True, epilogue=ops.EPI_DELTA, aux_in=y, delta=(delta, T))
        else:
            dy = ops.gemm(d_a, w_o, b_mn=True)
        # rotary adjoint fused into the dQ / dK epilogues: dqkv is the gradient of c_attn's raw output
        dqkv = ops.attention_bwd(qkv, y, dy, lse, B, T, H, d, scale, ctx.mask, p, ctx.keep, rope=(cos_tab, sin_tab),
                                 delta=delta)
        dw_qkv = _wgrad(dqkv, h1, pqkv)
        dh1 = ops.gemm(dqkv, w_qkv, b_mn=True)
        acc1 = direct and pg1.grad is not None
        dx, dg1, dx_drop = ops.layernorm_bwd(dh1, x, g1, mean1, rstd1, dres=dx1, dgamma=pg1.grad if acc1 else None,
                                             accumulate_dgamma=acc1, drop=ctx.up_drop or (0.0, 0, 0))
        _stash_dropped(dx, dx_drop, ctx.up_drop)
        _grads_final(ctx.params)
        return (dx, None if acc1 else dg1, dw_qkv, dw_o, None if acc2 else dg2, dw_fc, dw_pr, None, None, None, None,
                None, None, None, None, None, None)


class LayerNormFunction(torch.autograd.Function):
    """F.layer_norm(x, (C,), weight, None, 1e-5) on [M, C]  (model.py:63-72)."""

    @staticmethod
    def forward(ctx, x, gamma):
        y, _, mean, rstd = ops.layernorm_fwd(x, gamma)
        ctx.save_for_backward(x, gamma, mean, rstd)
        ctx.param = gamma
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        pg = ctx.param
        acc = _DIRECT_GRAD and pg.grad is not None
        dx, dg = ops.layernorm_bwd(dy.contiguous(), x, gamma, mean, rstd, dgamma=pg.grad if acc else None,
                                   accumulate_dgamma=acc)
        return dx, (None if acc else dg)


class ReadoutFunction(torch.autograd.Function):
    """MuReadout: logits = Linear(output_mult * x / width_mult)  (mup MuReadout.forward; model.py:208,253)."""

    @staticmethod
    def forward(ctx, x, weight, div):
        z = ops.scale_div(x, div)
        logits = ops.gemm(z, weight)
        ctx.save_for_backward(z, weight)
        ctx.div = div
        ctx.param = weight
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        z, weight = ctx.saved_tensors
        dlogits = dlogits.contiguous()
        dw = _wgrad(dlogits, z, ctx.param)
        dz = ops.gemm(dlogits, weight, b_mn=True)
        return ops.scale_div(dz, ctx.div), dw, None


class HeadLossFunction(torch.autograd.Function):
    """ln_f -> MuReadout -> masked-LM cross-entropy in one schedule (model.py:248-254 + train_encoder.py:301-305).

    The logits (M x V bf16) are materialised once; the backward overwrites them in place with d loss / d logits
    (rows outside the MLM mask are exact zeros, as in the reference; the incoming d loss is read on the device, so
    ``(k * loss).backward()`` and loss scalers work) and the two head GEMMs consume that buffer. Returns the
    bf16-rounded scalar loss of the reference. The saved logits are consumed by the first backward: a second one
    through the same graph (retain_graph=True) raises.
    """

    @staticmethod
    def forward(ctx, x, gamma, weight, div, targets, loss_mask, n_acc, up_drop=None):
        emb, z, mean, rstd = ops.layernorm_fwd(x, gamma, readout_div=div)
        del emb
        # the full [M, V] head GEMM; rows outside the loss mask are stored as the zeros d loss / d logits needs there
        # (nothing ever reads their logits), so the CE kernels only touch the masked rows
        row_mask = loss_mask.reshape(-1).to(torch.uint8).contiguous()
        logits = ops.gemm(z, weight, epilogue=ops.EPI_ROWMASK, aux_in=row_mask)
        scalars, lse, tok, row_mask, tgt = ops.ce_fwd(logits, targets, row_mask, n_acc)
        ctx.save_for_backward(x, gamma, weight, mean, rstd, z, logits, lse, row_mask, tgt, scalars)
        ctx.div = div
        ctx.params = (gamma, weight)
        ctx.up_drop = up_drop if (up_drop is not None and up_drop[0] > 0.0) else None
        ctx.consumed = False
        ctx.mark_non_differentiable(scalars)
        loss = scalars[0].to(torch.bfloat16)
        return loss, scalars

    @staticmethod
    def backward(ctx, dloss, _dscalars):
        x, gamma, weight, mean, rstd, z, logits, lse, row_mask, tgt, scalars = ctx.saved_tensors
        if ctx.consumed:
            raise RuntimeError("omnibiote_b200: the MLM head's logits buffer was overwritten by the first backward; "
                               "a second backward through the same graph is not supported")
        ctx.consumed = True
        pg, pw = ctx.params
        # d loss / d logits for the incoming d loss (a bf16 device scalar), in place over the logits
        dlogits = ops.ce_bwd_(logits, tgt, row_mask, lse, scalars, 1.0, unmasked_rows_zero=True,
                              upstream_dev=dloss.to(torch.bfloat16))
        dw = _wgrad(dlogits, z, pw)
        dz = ops.gemm(dlogits, weight, b_mn=True)
        acc = _DIRECT_GRAD and pg.grad is not None
        dx, dg, dx_drop = ops.layernorm_bwd(dz, x, gamma, mean, rstd, dgamma=pg.grad if acc else None,
                                            accumulate_dgamma=acc, dy_div=ctx.div, drop=ctx.up_drop or (0.0, 0, 0))
        _stash_dropped(dx, dx_drop, ctx.up_drop)
        _grads_final(ctx.params)
        return dx, (None if acc else dg), dw, None, None, None, None, None


class HeadLossMaskedRowsFunction(torch.autograd.Function):
    """HeadLossFunction restricted to the rows inside the MLM mask (optional; the dense path is the default).

    d loss / d logits of train_encoder.py:301-305 is exactly zero outside the mask, so ln_f's rows
    inside the mask are compacted into a fixed-capacity buffer [cap, C] (no host synchronisation: unused slots are
    zero rows with a zero gradient), the head GEMM, the CE and both backward GEMMs run on cap rows instead of M, and
    d z is scattered back with zeros elsewhere. Loss and gradients equal the dense path's (same rows, same arithmetic;
    only the fp32 summation order of the weight gradient's reduction differs). `meta` = {count, overflow}: overflow
    != 0 means more masked rows than `cap` and the caller must redo the micro-batch with the dense path.
    """

    @staticmethod
    def forward(ctx, x, gamma, weight, div, targets, loss_mask, n_acc, cap, up_drop=None):
        emb, z, mean, rstd = ops.layernorm_fwd(x, gamma, readout_div=div)
        del emb
        idx, tgt_c, valid_c, meta = ops.compact_rows(loss_mask, targets, cap)
        zc = ops.gather_rows(z, idx)
        del z
        logits = ops.gemm(zc, weight)
        scalars, lse, tok, row_mask, tgt = ops.ce_fwd(logits, tgt_c, valid_c, n_acc)
        ctx.save_for_backward(x, gamma, weight, mean, rstd, zc, logits, idx, lse, row_mask, tgt, scalars)
        ctx.div = div
        ctx.params = (gamma, weight)
        ctx.up_drop = up_drop if (up_drop is not None and up_drop[0] > 0.0) else None
        ctx.consumed = False
        ctx.mark_non_differentiable(scalars, meta)
        loss = scalars[0].to(torch.bfloat16)
        return loss, scalars, meta

    @staticmethod
    def backward(ctx, dloss, _dscalars, _dmeta):
        x, gamma, weight, mean, rstd, zc, logits, idx, lse, row_mask, tgt, scalars = ctx.saved_tensors
        if ctx.consumed:
            raise RuntimeError("omnibiote_b200: the MLM head's logits buffer was overwritten by the first backward; "
                               "a second backward through the same graph is not supported")
        ctx.consumed = True
        pg, pw = ctx.params
        dlogits = ops.ce_bwd_(logits, tgt, row_mask, lse, scalars, 1.0, upstream_dev=dloss.to(torch.bfloat16))
        dw = _wgrad(dlogits, zc, pw)
        dzc = ops.gemm(dlogits, weight, b_mn=True)
        dz = ops.scatter_rows(dzc, idx, x.shape[0])
        acc = _DIRECT_GRAD and pg.grad is not None
        dx, dg, dx_drop = ops.layernorm_bwd(dz, x, gamma, mean, rstd, dgamma=pg.grad if acc else None,
                                            accumulate_dgamma=acc, dy_div=ctx.div, drop=ctx.up_drop or (0.0, 0, 0))
        _stash_dropped(dx, dx_drop, ctx.up_drop)
        _grads_final(ctx.params)
        return dx, (None if acc else dg), dw, None, None, None, None, None, None


def masked_rows_capacity(n_rows: int, mask_prob: float, sigmas: float = 8.0) -> int:
    """Rows to reserve for the masked-rows-only head: mean + `sigmas` standard deviations of Binomial(n_rows, p),
    rounded up to the GEMM's 256-row tile."""
    import math
    want = n_rows * mask_prob + sigmas * math.sqrt(max(n_rows * mask_prob * (1.0 - mask_prob), 1.0))
    return min(-(-n_rows // 256) * 256, int(-(-want // 256)) * 256)


class PoolFunction(torch.autograd.Function):
    """encode() pooling over the token axis: mean / max  (model.py:269-278)."""

    @staticmethod
    def forward(ctx, emb, mode):
        out = ops.pool(emb, mode)
        ctx.save_for_backward(emb, out)
        ctx.mode = mode
        return out

    @staticmethod
    def backward(ctx, dout):
        emb, out = ctx.saved_tensors
        return ops.pool_bwd(emb, out, dout, ctx.mode), None
