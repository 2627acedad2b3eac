"""TEST INFRASTRUCTURE ONLY — CPU oracle: a functional restatement of the reference's hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module; the product (``omnibiote_b200/``) never does.

Each function restates, op for op and in the same order, what the reference computes, so that on CPU it produces
the reference's own numbers in either fp32 or bf16 (torch rounds after every op, which is exactly the reference's
rounding behaviour). Parity of this restatement is PINNED by ``tests/golden/*.pt``: outputs of the unmodified
reference ``training/model.py`` imported in the build container (generator: ``oracle/gen_golden.py``); see
``tests/test_oracle.py``.  The reference itself ships no tests or golden vectors.

Citations are into /root/reference (nyuolab/OmniBioTE).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

EOS_TOKEN, MASK_TOKEN, PAD_TOKEN = 3, 2, 1  # training/loader.py:4-6, training/train_encoder.py:20


# ---------------------------------------------------------------------------------------------------------------
# model (training/model.py)
# ---------------------------------------------------------------------------------------------------------------
def precompute_freqs_cis(dim: int, end: int, theta: float = 10000.0) -> torch.Tensor:
    """training/model.py:53-61."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
    t = torch.arange(end)
    freqs = torch.outer(t, freqs).float()
    return torch.polar(torch.ones_like(freqs), freqs)


def apply_rotary_emb(xq, xk, freqs_cis):
    """training/model.py:28-50. With a REAL (bf16) table this degenerates to cosine scaling."""
    xq_ = torch.view_as_complex(xq.float().reshape(*xq.shape[:-1], -1, 2))
    xk_ = torch.view_as_complex(xk.float().reshape(*xk.shape[:-1], -1, 2))
    f = freqs_cis[: xq_.shape[1]]
    f = f.view(1, xq_.shape[1], 1, xq_.shape[-1])
    xq_out = torch.view_as_real(xq_ * f).flatten(3)
    xk_out = torch.view_as_real(xk_ * f).flatten(3)
    return xq_out.type_as(xq), xk_out.type_as(xk)


def fused_gelu(x):
    """training/model.py:23-25 (constant 1.41421, not sqrt(2)); eager evaluation = the un-fused TorchScript run."""
    return x * 0.5 * (1.0 + torch.erf(x / 1.41421))


def layer_norm(x, weight):
    """training/model.py:72."""
    return F.layer_norm(x, weight.shape, weight, None, 1e-5)


def self_attention(x, p, prefix, n_head, attn_mask, dropout_p=0.0):
    """training/model.py:98-152 (flash branch)."""
    B, T, C = x.shape
    q, k, v = F.linear(x, p[prefix + "c_attn.weight"]).split(C, dim=2)
    k = k.view(B, T, n_head, C // n_head)
    q = q.view(B, T, n_head, C // n_head)
    v = v.view(B, T, n_head, C // n_head)
    q, k = apply_rotary_emb(q, k, p[prefix + "freqs_cis"])
    k, q, v = k.transpose(1, 2), q.transpose(1, 2), v.transpose(1, 2)
    y = F.scaled_dot_product_attention(q, k, v, scale=8 / C, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=False)
    y = y.transpose(1, 2).contiguous().view(B, T, C)
    return F.linear(y, p[prefix + "c_proj.weight"])


def mlp(x, p, prefix):
    """training/model.py:162-168."""
    return F.linear(fused_gelu(F.linear(x, p[prefix + "c_fc.weight"])), p[prefix + "c_proj.weight"])


def block(x, p, i, n_head, attn_mask):
    """training/model.py:178-181."""
    pre = f"transformer.h.{i}."
    x = x + self_attention(layer_norm(x, p[pre + "ln_1.weight"]), p, pre + "attn.", n_head, attn_mask)
    x = x + mlp(layer_norm(x, p[pre + "ln_2.weight"]), p, pre + "mlp.")
    return x


def forward(p: dict, n_layer: int, n_head: int, idx, attn_mask=None, return_embeddings=False, readout_width_mult=1.0,
            block_size=None):
    """training/model.py:225-254 with dropout = 0 (eval / parity runs). ``p`` is a reference state_dict."""
    _, t = idx.size()
    if block_size is not None:
        assert t <= block_size, f"Cannot forward sequence of length {t}, block size is only {block_size}"
    x = F.embedding(idx, p["transformer.wte.weight"])
    for i in range(n_layer):
        x = block(x, p, i, n_head, attn_mask)
    emb = layer_norm(x, p["transformer.ln_f.weight"])
    if return_embeddings:
        return emb
    # mup MuReadout.forward: Linear(output_mult * x / width_mult), output_mult = 1.0
    return F.linear(1.0 * emb / readout_width_mult, p["lm_head.weight"])


def encode(p, n_layer, n_head, idx, method="mean"):
    """training/model.py:256-278."""
    assert method in ["mean", "first", "last", "max", "all"], f"Unknown pooling method {method}"
    emb = forward(p, n_layer, n_head, idx, return_embeddings=True)
    if method == "mean":
        return emb.mean(dim=1)
    if method == "first":
        return emb[:, 0]
    if method == "last":
        return emb[:, -1]
    if method == "max":
        return emb.max(dim=1)[0]
    return emb


# ---------------------------------------------------------------------------------------------------------------
# attention-mask builders (inputs of the hot path)
# ---------------------------------------------------------------------------------------------------------------
def create_attention_mask(attn_mask, input_ids, eos_token=EOS_TOKEN, padding=False):
    """training/train_encoder.py:31-57, restated without TorchScript (same loop, same quirk)."""
    if not padding:
        temp = torch.ones(input_ids.size(0), input_ids.size(1) + 1, dtype=input_ids.dtype)
        temp[:, :-1] = input_ids
        temp[:, -1] = eos_token
        input_ids = temp
    eos_positions = (input_ids == eos_token).nonzero()
    attn_mask.fill_(-1e9)
    prev_index, prev_batch_idx = 0, 0
    for i in range(len(eos_positions)):
        r, c = int(eos_positions[i][0]), int(eos_positions[i][1])
        if r == prev_batch_idx:
            attn_mask[prev_batch_idx, prev_index:c + 1, prev_index:c + 1] = 0
            prev_index = c + 1
        else:
            prev_batch_idx = r
            prev_index = 0
            attn_mask[prev_batch_idx, prev_index:c + 1, prev_index:c + 1] = 0
    for i in range(len(input_ids)):
        if not torch.any(eos_positions[:, 0] == i):
            attn_mask[i, :, :] = 0
    return attn_mask


def pad_attn(attn_mask, x, pad_token=PAD_TOKEN):
    """evals/gue.py:15-21."""
    attn_mask.fill_(0)
    for i in range(x.shape[0]):
        pads = (x[i] == pad_token).nonzero()
        if len(pads) > 0:
            first_pad = int(pads[0])
            attn_mask[i, first_pad + 1:, :] = -1e9
            attn_mask[i, :, first_pad + 1:] = -1e9
    return attn_mask


def mask_intervals(mask3d: torch.Tensor):
    """Per-row [lo, hi) of zero entries of a dense (B,T,T) additive mask; (0,0) for fully-masked rows."""
    z = (mask3d == 0)
    B, T, _ = z.shape
    lo = torch.zeros(B, T, dtype=torch.int32)
    hi = torch.zeros(B, T, dtype=torch.int32)
    for b in range(B):
        for i in range(T):
            nz = z[b, i].nonzero()
            if len(nz):
                lo[b, i], hi[b, i] = int(nz[0]), int(nz[-1]) + 1
    return lo, hi


# ---------------------------------------------------------------------------------------------------------------
# training step (training/train_encoder.py:270-318)
# ---------------------------------------------------------------------------------------------------------------
def mlm_mask(input_ids: torch.Tensor, rng: np.random.RandomState, mask_prob=0.15):
    """train_encoder.py:273-279: Bernoulli(0.15) & != PAD & != EOS; masked inputs get MASK_TOKEN (no 80/10/10)."""
    m = torch.as_tensor(rng.binomial(1, mask_prob, tuple(input_ids.shape)), dtype=torch.bool)
    m = m & (input_ids != PAD_TOKEN) & (input_ids != EOS_TOKEN)
    return m, input_ids.masked_fill(m, MASK_TOKEN)


def mlm_loss(logits, targets, mask, n_accum=1):
    """train_encoder.py:301-305."""
    loss = F.cross_entropy(logits.view(-1, logits.size(-1)), targets.view(-1), reduction="none") / n_accum
    loss *= mask.view(-1).float()
    return loss.sum() / mask.view(-1).sum()


def clip_grad_norm(grads, max_norm=1.0):
    """torch.nn.utils.clip_grad_norm_ (train_encoder.py:316), fp32 restatement."""
    total = torch.sqrt(sum((g.float() ** 2).sum() for g in grads))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return total, coef


def adamw_step(p, g, m, v, step, lr, wd, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.AdamW single-tensor step (train_encoder.py:317), evaluated in the tensors' own dtype so that bf16
    state rounds after every primitive like the reference's foreach implementation."""
    p = p * (1 - lr * wd)
    m = torch.lerp(m, g, 1 - beta1)
    v = (v * beta2).addcmul(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2_sqrt = math.sqrt(1 - beta2 ** step)
    denom = (v.sqrt() / bc2_sqrt).add(eps)
    p = p.addcdiv(m, denom, value=-(lr / bc1))
    return p, m, v


def mu_lr_wd(name: str, shape, lr, wd, n_embd, base_n_embd=24):
    """mup.MuAdamW group scaling for the OmniBioTA parameter set (train_encoder.py:158-166,199): block matrices have
    two width-dependent dims -> lr / width_mult, wd * width_mult; wte, LayerNorm gains and lm_head are vector-like."""
    wm = n_embd / base_n_embd
    matrix_like = len(shape) == 2 and ("c_attn" in name or "c_proj" in name or "c_fc" in name)
    return (lr / wm, wd * wm) if matrix_like else (lr, wd)
