"""GPU: the code paths the benchmark configuration (dropout 0.1, device MLM masking) actually runs.

* the three dropout implementations (GEMM epilogue 5, embedding gather/scatter, LayerNorm-backward replay) must agree
  bit for bit with the stand-alone kernel on WHICH elements are kept, and on the kept values;
* one full block forward + backward at p = 0.1 against an fp32 torch restatement of model.py:98-181 that applies the
  extracted masks (same approach as test_attn_dropout_against_torch_with_the_same_keep_mask);
* obt_mlm_mask against the properties of train_encoder.py:273-279;
* a non-unit upstream gradient through the fused head (`(3 * loss).backward()`).
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _ops():
    from omnibiote_b200 import ops
    return ops


def _keep_pattern(ops, shape, p, seed, off):
    """keep mask of the library's element-wise dropout stream on a contiguous tensor of `shape`."""
    ones = torch.ones(shape, dtype=BF, device="cuda")
    return ops.dropout(ones, p, seed, off) != 0


def _scaled(x_bf16, p):
    """rb(x * 1/(1-p)) with the kernel's fp32 scale"""
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return (x_bf16.float() * float(scale)).to(BF)


@pytest.mark.parametrize("M,N,K", [(640, 512, 256), (32768, 1024, 1024)])
def test_gemm_resid_dropout_epilogue_uses_the_dropout_kernels_mask(M, N, K):
    ops = _ops()
    p, seed, off = 0.1, 0x1234ABCD, 8
    g = torch.Generator(device="cuda").manual_seed(3)
    a = (torch.rand(M, K, generator=g, device="cuda") + 0.5).to(BF)   # strictly positive products: acc != 0
    b = (torch.rand(N, K, generator=g, device="cuda") + 0.5).to(BF)
    zero = torch.zeros(M, N, dtype=BF, device="cuda")
    plain = ops.gemm(a, b)                                             # rb(acc)
    assert float(plain.abs().min()) > 0
    got = ops.gemm(a, b, epilogue=ops.EPI_RESID_DROPOUT, aux_in=zero, drop_p=p, seed=seed, offset=off)
    keep = _keep_pattern(ops, (M, N), p, seed, off)
    assert torch.equal(got != 0, keep), "GEMM epilogue and dropout kernel disagree on the kept elements"
    n = M * N
    assert abs(float(keep.float().mean()) - (1 - p)) < 4 * math.sqrt(p * (1 - p) / n) + 1e-4
    want = torch.where(keep, _scaled(plain, p), torch.zeros_like(plain))
    assert torch.equal(got, want)                                      # kept values = rb(rb(acc) / (1 - p))
    # with a residual: D = rb(resid + dropped)
    resid = torch.randn(M, N, generator=g, device="cuda").to(BF)
    got2 = ops.gemm(a, b, epilogue=ops.EPI_RESID_DROPOUT, aux_in=resid, drop_p=p, seed=seed, offset=off)
    assert torch.equal(got2, (resid.float() + want.float()).to(BF))
    # the backward's replay of the same (seed, offset) on an arbitrary gradient selects the same elements
    dy = torch.randn(M, N, generator=g, device="cuda").to(BF)
    assert torch.equal(ops.dropout(dy, p, seed, off), torch.where(keep, _scaled(dy, p), torch.zeros_like(dy)))


def test_embedding_dropout_forward_and_backward_share_the_mask():
    ops = _ops()
    V, C, M = 4096, 256, 4096
    p, seed, off = 0.1, 77, 4
    wte = (torch.rand(V, C, device="cuda") + 0.5).to(BF)
    idx = torch.randperm(V, device="cuda")[:M]                         # distinct rows: no accumulation in the backward
    keep = _keep_pattern(ops, (M, C), p, seed, off)
    out = ops.embed_fwd(idx, wte, p, seed, off)
    assert torch.equal(out != 0, keep)
    assert torch.equal(out, torch.where(keep, _scaled(wte[idx], p), torch.zeros_like(out)))
    assert abs(float(keep.float().mean()) - (1 - p)) < 4 * math.sqrt(p * (1 - p) / keep.numel()) + 1e-4
    dout = (torch.rand(M, C, device="cuda") + 0.5).to(BF)
    dw = torch.empty(V, C, dtype=BF, device="cuda")
    ops.embed_bwd(idx, dout, dw, False, p, seed, off)
    assert torch.equal(dw[idx] != 0, keep), "embedding backward drops different elements than the forward"
    assert torch.equal(dw[idx], torch.where(keep, _scaled(dout, p), torch.zeros_like(dout)))


@pytest.mark.parametrize("M,C", [(37, 256), (4096, 1024), (300, 2048)])
def test_layernorm_backward_fused_dropout_replay_and_dgamma(M, C):
    ops = _ops()
    p, seed, off = 0.1, 991, 12
    x = (torch.randn(M, C, device="cuda") * 2 + 0.3).to(BF)
    gam = (1 + 0.1 * torch.randn(C, device="cuda")).to(BF)
    _, _, mean, rstd = ops.layernorm_fwd(x, gam)
    dy = torch.randn(M, C, device="cuda").to(BF)
    dres = torch.randn(M, C, device="cuda").to(BF)
    dx0, dg0 = ops.layernorm_bwd(dy, x, gam, mean, rstd, dres=dres)
    dx1, dg1, dxd = ops.layernorm_bwd(dy, x, gam, mean, rstd, dres=dres, drop=(p, seed, off))
    assert torch.equal(dx0, dx1) and torch.equal(dg0, dg1)
    assert torch.equal(dxd, ops.dropout(dx1, p, seed, off))            # bit-exact replay of the stand-alone kernel
    # the in-kernel dgamma reduction is deterministic and re-arms itself: repeated launches give identical bits
    for _ in range(3):
        dx2, dg2 = ops.layernorm_bwd(dy, x, gam, mean, rstd, dres=dres)
        assert torch.equal(dg2, dg0) and torch.equal(dx2, dx0)
    xr, gr = x.float().requires_grad_(True), gam.float().requires_grad_(True)
    F.layer_norm(xr, (C,), gr, None, 1e-5).backward(dy.float())
    assert rel_err(dg0, gr.grad) < 4e-3 and rel_err(dx0, dres.float() + xr.grad) < 4e-3


def _block_reference(x, g1, wqkv, wo, g2, wfc, wpr, cos, B, T, H, mask_add, keepP, keepA, keepM, p):
    """fp32 restatement of model.py:98-181 (bf16 model: cosine-only rotary) with explicit dropout masks."""
    C = x.shape[1]
    d = C // H
    inv = 1.0 / (1.0 - p)
    h1 = F.layer_norm(x, (C,), g1, None, 1e-5)
    q, k, v = (h1 @ wqkv.t()).split(C, dim=1)
    cs = cos[:T].repeat_interleave(2, dim=1).view(1, T, 1, d)           # pairs (2i, 2i+1) share cos[t, i]
    q = (q.view(B, T, H, d) * cs).transpose(1, 2)
    k = (k.view(B, T, H, d) * cs).transpose(1, 2)
    v = v.view(B, T, H, d).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (8.0 / C) + mask_add
    P = torch.softmax(s, dim=-1) * keepP * inv
    y = (P @ v).transpose(1, 2).reshape(B * T, C)
    x1 = x + (y @ wo.t()) * keepA * inv
    h2 = F.layer_norm(x1, (C,), g2, None, 1e-5)
    u = h2 @ wfc.t()
    g = u * 0.5 * (1.0 + torch.erf(u / 1.41421))
    return x1 + (g @ wpr.t()) * keepM * inv


def test_block_forward_backward_with_dropout_against_fp32_restatement():
    """BlockFunction at p = 0.1 (EPI_RESID_DROPOUT forward, fused LayerNorm-backward replays, attention keep bits)."""
    from omnibiote_b200 import functional as Fn
    ops = _ops()
    B, T, H, d, p = 2, 256, 2, 128, 0.1
    C = H * d
    gen = torch.Generator(device="cuda").manual_seed(11)
    rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=gen, device="cuda") * scale).to(BF)
    x = rnd(B * T, C)
    g1, g2 = (1 + rnd(C, scale=0.1).float()).to(BF), (1 + rnd(C, scale=0.1).float()).to(BF)
    wqkv, wo = rnd(3 * C, C, scale=C ** -0.5), rnd(C, C, scale=C ** -0.5)
    wfc, wpr = rnd(4 * C, C, scale=C ** -0.5), rnd(C, 4 * C, scale=(4 * C) ** -0.5)
    ang = torch.outer(torch.arange(T, device="cuda").float(),
                      1.0 / (10000 ** (torch.arange(0, d, 2, device="cuda").float() / d)))
    cos = torch.cos(ang).to(BF).float().contiguous()
    ids = torch.randint(20, 500, (B, T), device="cuda")
    ids[0, 70] = ids[0, 200] = ids[1, 33] = ids[1, 180] = 3
    lo, hi = ops.doc_mask_intervals(ids, 3, False)
    mask_add = ops.mask_from_intervals(lo, hi).float().unsqueeze(1)     # (B,1,T,T) of {0, -1e9}
    spec = ops.MaskSpec(None, B, H, T, lo, hi)
    seeds = [(4242, 0), (4242, 4), (4242, 8)]
    up_drop = (p, 4242, 12)                                            # pretend a previous block dropped x

    params = [t.clone().requires_grad_(True) for t in (x, g1, wqkv, wo, g2, wfc, wpr)]
    Fn.clear_drop_stash()
    out = Fn.BlockFunction.apply(*params, cos, None, spec, B, T, H, p, True, seeds, up_drop)
    dout = rnd(B * T, C)
    out.backward(dout)
    # the replayed gradient for the previous block was produced by this block's last LayerNorm backward
    assert len(Fn._DROP_STASH) == 1
    (src, dropped, seeds_), = Fn._DROP_STASH.values()
    assert seeds_ == up_drop and torch.equal(src, params[0].grad)
    assert torch.equal(dropped, ops.dropout(src, *up_drop))
    assert Fn._take_dropped(src, up_drop) is dropped and Fn._take_dropped(src, up_drop) is None

    keepP = ops.keep_mask_to_bool(ops.attn_keep_mask(B, H, T, p, *seeds[0], x.device), T).float()
    keepA = _keep_pattern(ops, (B * T, C), p, *seeds[1]).float()
    keepM = _keep_pattern(ops, (B * T, C), p, *seeds[2]).float()
    ref_in = [t.detach().float().requires_grad_(True) for t in (x, g1, wqkv, wo, g2, wfc, wpr)]
    ref = _block_reference(*ref_in, cos, B, T, H, mask_add, keepP, keepA, keepM, p)
    ref.backward(dout.float())
    assert rel_err(out, ref) < 1e-2, rel_err(out, ref)
    names = ["dx", "dg1", "dWqkv", "dWo", "dg2", "dWfc", "dWpr"]
    report = {n: rel_err(a.grad, b.grad) for n, a, b in zip(names, params, ref_in)}
    print({k: f"{v:.2e}" for k, v in report.items()})
    for n, v in report.items():
        assert v < 2e-2, (n, v)   # bf16 kernels vs fp32 math on the same masks (reference noise floor: SURVEY C.2)

    # the stash path and the stand-alone-dropout fallback give identical gradients for an upstream block
    params2 = [t.clone().requires_grad_(True) for t in (x, g1, wqkv, wo, g2, wfc, wpr)]
    Fn.clear_drop_stash()
    out2 = Fn.BlockFunction.apply(*params2, cos, None, spec, B, T, H, p, True, seeds, None)
    out2.backward(dout)
    for a, b in zip(params, params2):
        assert torch.equal(a.grad, b.grad)


def test_two_blocks_with_dropout_stash_equals_fallback():
    """Block l+1 hands block l its replayed gradient through functional._DROP_STASH; disabling the hand-over (each
    block replays with the stand-alone kernel) must give bit-identical gradients."""
    from conftest import make_model
    from omnibiote_b200 import functional as Fn
    model = make_model(3, 2, 256, dropout=0.1).cuda().train()
    ids = torch.randint(20, 512, (2, 256), device="cuda")
    lm = torch.rand(2, 256, device="cuda") < 0.15
    lm[0, 0] = True
    grads = []
    for use_stash in (True, False):
        model.zero_grad(set_to_none=True)
        torch.manual_seed(5)                                           # same dropout seeds for both runs
        orig = Fn._stash_dropped
        if not use_stash:
            Fn._stash_dropped = lambda *a, **k: None
        try:
            loss, _ = model.mlm_loss(ids.masked_fill(lm, 2), ids, lm, n_accum=1)
            loss.backward()
        finally:
            Fn._stash_dropped = orig
        grads.append({n: q.grad.clone() for n, q in model.named_parameters()})
        assert len(Fn._DROP_STASH) == 0                                # every hand-over was consumed
    for n in grads[0]:
        assert torch.equal(grads[0][n], grads[1][n]), n


def test_mlm_mask_properties():
    """train_encoder.py:273-279: mask = Bernoulli(0.15) & != PAD & != EOS; masked_fill(mask, MASK_TOKEN)."""
    from omnibiote_b200 import train
    g = torch.Generator(device="cuda").manual_seed(1)
    ids = torch.randint(20, 65536, (64, 1024), generator=g, device="cuda")
    ids[:, 100::97] = train.EOS_TOKEN
    ids[10:20, 900:] = train.PAD_TOKEN
    ids[3, :] = train.PAD_TOKEN
    counters = torch.zeros(2, dtype=torch.float32, device="cuda")
    torch.manual_seed(123)
    masked, mask = train.mlm_mask(ids, 0.15, counters=counters)
    m = mask.bool()
    special = (ids == train.PAD_TOKEN) | (ids == train.EOS_TOKEN)
    assert not bool((m & special).any())                               # PAD / EOS are never masked
    assert bool((masked[m] == train.MASK_TOKEN).all())                 # masked positions hold MASK
    assert torch.equal(masked[~m], ids[~m])                            # everything else is untouched
    n = int((~special).sum())
    rate = float(m.sum()) / n
    assert abs(rate - 0.15) < 4 * math.sqrt(0.15 * 0.85 / n), rate
    assert counters.tolist() == [float(m.sum()), float((ids != train.PAD_TOKEN).sum())]
    # replay: same (seed, offset) -> same mask; the next draw differs
    torch.manual_seed(123)
    masked2, mask2 = train.mlm_mask(ids, 0.15)
    assert torch.equal(mask2, mask) and torch.equal(masked2, masked)
    _, mask3 = train.mlm_mask(ids, 0.15)
    assert not torch.equal(mask3, mask)
    # prob 0 / 1 and a ragged length
    assert int(train.mlm_mask(ids, 0.0)[1].sum()) == 0
    assert torch.equal(train.mlm_mask(ids, 1.0)[1].bool(), ~special)
    odd = ids.reshape(-1)[:1027].clone().reshape(1, -1)
    mo, ko = train.mlm_mask(odd, 0.5)
    assert torch.equal(mo[ko.bool()], torch.full_like(mo[ko.bool()], train.MASK_TOKEN))
    with pytest.raises(RuntimeError):
        train.mlm_mask(ids.cpu(), 0.15)                                # no CPU path


@pytest.mark.parametrize("path", ["fused", "masked_rows"])
def test_head_loss_honours_the_upstream_gradient(golden, path):
    """(3 * loss).backward() through mlm_loss: every parameter gradient is 3x that of loss.backward() (up to the bf16
    rounding of g = rb(rb(3 / count) / n_acc) vs 3 * rb(rb(1 / count) / n_acc)); a loss / 2 likewise."""
    from test_model_gpu import build, m4
    c = golden("bf16_h2")
    H = c["cfg"]["n_head"]
    ids, masked, lm = c["ids"].cuda(), c["ids_masked"].cuda(), c["mlm_mask"].cuda()
    mask = m4(c["mask_doc"], H)
    cap = (-(-int(lm.sum()) // 8) * 8 + 8) if path == "masked_rows" else 0

    def grads(factor):
        model = build(c).train()
        loss, _ = model.mlm_loss(masked, ids, lm, attn_mask=mask, n_accum=2, masked_rows_cap=cap)
        (loss * factor).backward()
        return {n: q.grad.float() for n, q in model.named_parameters()}

    g1 = grads(1.0)
    for factor in (3.0, 0.5):
        gk = grads(factor)
        for n in g1:
            assert rel_err(gk[n], factor * g1[n]) < 1.5e-2, (n, factor, rel_err(gk[n], factor * g1[n]))
    # a second backward through the same graph must fail loudly, not reuse the overwritten logits
    model = build(c).train()
    loss, _ = model.mlm_loss(masked, ids, lm, attn_mask=mask, n_accum=2, masked_rows_cap=cap)
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError):
        loss.backward()


def test_clip_coefficient_propagates_nan_and_skip_flag_skips():
    from omnibiote_b200.optim import FusedAdamW
    p = torch.nn.Parameter(torch.ones(4096, dtype=BF, device="cuda"))
    opt = FusedAdamW([p], lr=1e-2, weight_decay=0.0)
    p.grad = torch.full_like(p, 0.5)
    skip = torch.ones(1, dtype=torch.int32, device="cuda")
    opt.clip_and_step(1.0, zero_grad=True, skip_flag=skip)
    assert torch.equal(p.detach(), torch.ones_like(p)) and float(p.grad.abs().max()) == 0.0   # skipped, grads cleared
    assert float(opt.state[p]["exp_avg"].abs().max()) == 0.0
    p.grad = torch.full_like(p, 0.5)
    skip.zero_()
    opt.clip_and_step(1.0, zero_grad=True, skip_flag=skip)
    assert float((p.detach().float() - 1.0).abs().max()) > 0                                  # applied
    p.grad = torch.full_like(p, float("nan"))
    norm = opt.clip_and_step(1.0)
    assert math.isnan(float(norm[1])) and bool(torch.isnan(p.detach().float()).all())         # torch.clamp semantics
