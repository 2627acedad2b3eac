"""Drop-in replacement for the reference ``training/model.py`` whose forward/backward run on hand-written sm_100a
kernels (``libomnibiote_b200.so``) instead of PyTorch library ops.

Same public surface as the reference (training/model.py:183-278):
``OmniBioTAConfig`` fields and defaults, ``OmniBioTA(config)``, ``forward(idx, attn_mask=None,
return_embeddings=False)``, ``encode(idx, method)``, ``get_num_params``, the submodule tree / ``state_dict`` key set
(``transformer.wte.weight``, ``transformer.h.{i}.ln_1.weight``, ``.attn.freqs_cis``, ``.attn.c_attn.weight``,
``.attn.c_proj.weight``, ``.ln_2.weight``, ``.mlp.c_fc.weight``, ``.mlp.c_proj.weight``, ``transformer.ln_f.weight``,
``lm_head.weight``) and error behaviour (``AssertionError`` for ``t > block_size``, bad pooling method,
``n_embd % n_head``).  Usage is identical: ``from omnibiote_b200.model import OmniBioTA, OmniBioTAConfig``.

The device path is bf16-on-CUDA only and fails loudly otherwise; there is no CPU or eager fallback.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn
from torch.utils.checkpoint import checkpoint

from . import functional as Fn
from . import ops
from .mup import MuReadout


def precompute_freqs_cis(dim: int, end: int, theta: float = 10000.0) -> torch.Tensor:
    """complex64 table exp(j * t * theta^(-2i/dim)), shape (end, dim/2) — same values as model.py:53-61."""
    exponent = torch.arange(0, dim, 2)[: dim // 2].float() / dim
    inv_freq = 1.0 / (theta ** exponent)
    angles = torch.outer(torch.arange(end, device=inv_freq.device), inv_freq).float()
    return torch.polar(torch.ones_like(angles), angles)


def _require_cuda_bf16(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda or t.dtype != torch.bfloat16:
        raise RuntimeError(
            f"omnibiote_b200: {what} is {t.dtype} on {t.device}; the B200 kernels need bf16 CUDA tensors "
            "(call model.to(torch.bfloat16).to('cuda')). There is no CPU / fp32 fallback.")


class LayerNorm(nn.Module):
    """LayerNorm with an optional bias (reference model.py:63-72); only bias=False is supported by the kernels."""

    def __init__(self, ndim, bias):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(ndim))
        self.bias = nn.Parameter(torch.zeros(ndim)) if bias else None

    def forward(self, input):
        if self.bias is not None:
            raise RuntimeError("omnibiote_b200: LayerNorm bias is not supported (the reference trains with bias=False)")
        _require_cuda_bf16(input, "LayerNorm input")
        shape = input.shape
        out = Fn.LayerNormFunction.apply(input.reshape(-1, shape[-1]), self.weight)
        return out.view(shape)


class SelfAttention(nn.Module):
    """Parameter/buffer container with the reference's names (model.py:74-96). Compute happens in Block."""

    def __init__(self, config):
        super().__init__()
        assert config.n_embd % config.n_head == 0
        self.c_attn = nn.Linear(config.n_embd, 3 * config.n_embd, bias=config.bias)
        self.c_proj = nn.Linear(config.n_embd, config.n_embd, bias=config.bias)
        self.attn_dropout = nn.Dropout(config.dropout, inplace=True)
        self.resid_dropout = nn.Dropout(config.dropout, inplace=True)
        self.n_head = config.n_head
        self.n_embd = config.n_embd
        self.dropout = config.dropout
        self.autoregressive = config.autoregressive
        self.flash = getattr(config, "flash", True)
        if self.autoregressive:
            raise NotImplementedError("omnibiote_b200: autoregressive=True is outside the hot path (never used by "
                                      "the reference's training or evals)")
        self.register_buffer("freqs_cis", precompute_freqs_cis(self.n_embd // self.n_head, config.block_size))
        self._rot_cache = None

    def rotary_tables(self):
        """fp32 (cos, sin|None) tables for the kernel. A complex buffer means a true rotation; a real one (what
        ``module.to(bfloat16)`` leaves behind) means cosine scaling only."""
        buf = self.freqs_cis
        key = (buf.data_ptr(), buf._version, buf.dtype, buf.device)
        if self._rot_cache is None or self._rot_cache[0] != key:
            if buf.is_complex():
                re_im = torch.view_as_real(buf)
                cos, sin = re_im[..., 0].float().contiguous(), re_im[..., 1].float().contiguous()
            else:
                cos, sin = buf.float().contiguous(), None
            self._rot_cache = (key, cos, sin)
        return self._rot_cache[1], self._rot_cache[2]

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_rot_cache"] = None
        return state


class MLP(nn.Module):
    """Parameter container (model.py:154-160)."""

    def __init__(self, config):
        super().__init__()
        self.c_fc = nn.Linear(config.n_embd, 4 * config.n_embd, bias=config.bias)
        self.c_proj = nn.Linear(4 * config.n_embd, config.n_embd, bias=config.bias)
        self.dropout = nn.Dropout(config.dropout, inplace=True)


class Block(nn.Module):
    """Pre-LN residual block (model.py:170-181) executed as one fused schedule of kernels."""

    def __init__(self, config):
        super().__init__()
        if config.bias:
            raise NotImplementedError("omnibiote_b200: bias=True is not supported (reference default is False)")
        self.ln_1 = LayerNorm(config.n_embd, bias=config.bias)
        self.attn = SelfAttention(config)
        self.ln_2 = LayerNorm(config.n_embd, bias=config.bias)
        self.mlp = MLP(config)

    def forward(self, x, attn_mask=None, seeds=None, up_drop=None):
        """seeds / up_drop: dropout bookkeeping supplied by OmniBioTA._trunk (see functional.BlockFunction.forward);
        a stand-alone call leaves them None and the block draws its own seeds."""
        _require_cuda_bf16(x, "Block input")
        B, T, C = x.shape
        H = self.attn.n_head
        cos, sin = self.attn.rotary_tables()
        if cos.device != x.device:
            raise RuntimeError("omnibiote_b200: freqs_cis buffer and activations are on different devices")
        mask = attn_mask if isinstance(attn_mask, ops.MaskSpec) else ops.MaskSpec(attn_mask, B, H, T)
        out = Fn.BlockFunction.apply(
            x.reshape(B * T, C), self.ln_1.weight, self.attn.c_attn.weight, self.attn.c_proj.weight, self.ln_2.weight,
            self.mlp.c_fc.weight, self.mlp.c_proj.weight, cos, sin, mask, B, T, H, self.attn.dropout, self.training,
            seeds, up_drop)
        return out.view(B, T, C)


@dataclass
class OmniBioTAConfig:
    block_size: int = 2048
    vocab_size: int = 2**16
    n_layer: int = 12
    n_head: int = 12
    n_embd: int = 1024
    dropout: float = 0.1
    bias: bool = False
    autoregressive: bool = False
    checkpoint_freq: int = 0


class OmniBioTA(nn.Module):
    def __init__(self, config):
        super().__init__()
        assert config.vocab_size is not None
        assert config.block_size is not None
        self.config = config

        self.transformer = nn.ModuleDict(dict(
            wte=nn.Embedding(config.vocab_size, config.n_embd),
            drop=nn.Dropout(config.dropout, inplace=True),
            h=nn.ModuleList([Block(config) for _ in range(config.n_layer)]),
            ln_f=LayerNorm(config.n_embd, bias=config.bias),
        ))
        self.lm_head = MuReadout(config.n_embd, config.vocab_size, bias=False)
        self._last_drop = None

        print("number of parameters: %.2fM" % (self.get_num_params() / 1e6,))

    def get_num_params(self, non_embedding=True):
        n_params = sum(p.numel() for p in self.parameters())
        if non_embedding:
            n_params -= self.transformer.wte.weight.numel()
        return n_params

    # ------------------------------------------------------------------------------------------------------------
    def _trunk(self, idx, attn_mask):
        """Embedding + all blocks; returns the pre-ln_f residual stream (b, t, C). ``self._last_drop`` is left holding
        the (p, seed, offset) of the last block's MLP residual dropout (None without dropout) for the fused head."""
        if idx.dim() != 2:
            raise RuntimeError(f"omnibiote_b200: idx must be (b, t), got {tuple(idx.shape)}")
        b, t = idx.size()
        assert t <= self.config.block_size, \
            f"Cannot forward sequence of length {t}, block size is only {self.config.block_size}"
        wte = self.transformer.wte.weight
        _require_cuda_bf16(wte, "model parameters")
        if not idx.is_cuda:
            raise RuntimeError("omnibiote_b200: idx must be a CUDA tensor (there is no CPU path)")
        if idx.dtype != torch.int64:
            idx = idx.long()
        C = wte.shape[1]
        x = Fn.EmbedFunction.apply(idx.reshape(-1), wte, self.transformer.drop.p, self.training).view(b, t, C)
        blocks = self.transformer.h
        if isinstance(attn_mask, ops.MaskSpec):
            mask = attn_mask  # already in kernel form (dense bias or per-row intervals)
        else:
            mask = ops.MaskSpec(attn_mask, b, blocks[0].attn.n_head, t) if len(blocks) else None
        ckpt = self.config.checkpoint_freq
        # Dropout seeds are drawn here, not inside the blocks, so that block l+1 knows the mask of block l's MLP
        # residual dropout: its LayerNorm backward then writes the replayed gradient block l needs in the same pass
        # (functional._DROP_STASH). Plain arguments, so activation-checkpoint recomputation replays them unchanged.
        Fn.clear_drop_stash()
        up_drop = None
        for i, block in enumerate(blocks):
            p = float(block.attn.dropout) if self.training else 0.0
            seeds = [ops.philox_args(x.device, 4) for _ in range(3)] if p > 0.0 else None
            if ckpt > 0 and i % ckpt == 0 and torch.is_grad_enabled():
                x = checkpoint(block, x, mask, seeds, up_drop, use_reentrant=False)
            else:
                x = block(x, attn_mask=mask, seeds=seeds, up_drop=up_drop)
            up_drop = (p, *seeds[2]) if p > 0.0 else None
        self._last_drop = up_drop
        return x

    def forward(self, idx, attn_mask=None, return_embeddings=False):
        """idx (b, t) int64; attn_mask None or additive (b, n_head, t, t). Returns logits (b, t, vocab) or, with
        ``return_embeddings``, ln_f token embeddings (b, t, n_embd)  (reference model.py:225-254)."""
        x = self._trunk(idx, attn_mask)
        emb = self.transformer.ln_f(x)
        if return_embeddings:
            return emb
        return self.lm_head(emb)

    def encode(self, idx, method="mean"):
        """Pool ln_f embeddings over the token axis without any padding mask (reference model.py:256-278)."""
        assert method in ["mean", "first", "last", "max", "all"], f"Unknown pooling method {method}"
        emb = self.forward(idx, return_embeddings=True)
        if method == "mean":
            return Fn.PoolFunction.apply(emb, "mean")
        elif method == "first":
            return emb[:, 0]
        elif method == "last":
            return emb[:, -1]
        elif method == "max":
            return Fn.PoolFunction.apply(emb, "max")
        elif method == "all":
            return emb

    # ------------------------------------------------------------------------------------------------------------
    def mlm_loss(self, masked_idx, targets, loss_mask, attn_mask=None, n_accum: int = 1, masked_rows_cap: int = 0):
        """Fused training-step front half: forward + head + masked-LM loss of train_encoder.py:296-305.

        Equivalent to ``ce = F.cross_entropy(model(masked_idx, attn_mask).view(-1, V), targets.view(-1),
        reduction="none") / n_accum; ce *= loss_mask.view(-1).float(); loss = ce.sum() / loss_mask.sum()`` with the
        reference's bf16 rounding points, but ln_f, the readout scaling, the head GEMM and the CE forward/backward
        run as one schedule. Returns (loss [bf16 scalar, differentiable], scalars [fp32: loss, n_masked, dCE]).

        masked_rows_cap > 0 selects the masked-rows-only head (functional.HeadLossMaskedRowsFunction): same loss and
        gradients, the head runs on at most that many rows; ``self.head_rows_meta`` (device int32 {count, overflow})
        tells whether the capacity sufficed.
        """
        x = self._trunk(masked_idx, attn_mask)
        b, t, C = x.shape
        if masked_rows_cap > 0:
            loss, scalars, meta = Fn.HeadLossMaskedRowsFunction.apply(
                x.reshape(b * t, C), self.transformer.ln_f.weight, self.lm_head.weight,
                float(self.lm_head.readout_div()), targets.reshape(-1), loss_mask.reshape(-1), float(n_accum),
                int(masked_rows_cap), self._last_drop)
            self.head_rows_meta = meta
            return loss, scalars
        loss, scalars = Fn.HeadLossFunction.apply(
            x.reshape(b * t, C), self.transformer.ln_f.weight, self.lm_head.weight, float(self.lm_head.readout_div()),
            targets.reshape(-1), loss_mask.reshape(-1), float(n_accum), self._last_drop)
        return loss, scalars
