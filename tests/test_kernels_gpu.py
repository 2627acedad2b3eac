"""GPU: memory-bound kernels and the generic attention kernel through the C ABI, each against the torch expression
the reference evaluates (bf16 inputs, fp32 math, one rounding)."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err, max_abs

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _ops():
    from omnibiote_b200 import ops
    return ops


def test_embedding_fwd_bwd_bit_exact():
    ops = _ops()
    V, C, M = 512, 256, 1000
    wte = torch.randn(V, C, device="cuda").to(BF)
    idx = torch.randint(0, V, (M,), device="cuda")
    idx[:300] = 2  # hot row (MASK token receives ~15 % of all rows)
    out = ops.embed_fwd(idx, wte)
    assert torch.equal(out, wte[idx])  # index selection: bit-exact
    dout = (torch.randn(M, C, device="cuda") * 0.1).to(BF)
    dw = torch.empty(V, C, dtype=BF, device="cuda")
    ops.embed_bwd(idx, dout, dw, accumulate=False)
    ref = torch.zeros(V, C, device="cuda").index_add_(0, idx, dout.float())
    assert rel_err(dw, ref) < 3e-3
    untouched = torch.ones(V, dtype=torch.bool, device="cuda")
    untouched[idx] = False
    assert float(dw[untouched].abs().max()) == 0.0
    # accumulate into an existing gradient (second micro-batch), scratch must have been re-zeroed
    dw2 = dw.clone()
    ops.embed_bwd(idx, dout, dw2, accumulate=True)
    assert rel_err(dw2, dw.float() + ref.to(BF).float()) < 3e-3


@pytest.mark.parametrize("M,C", [(37, 256), (1024, 1024), (130, 2048)])
def test_layernorm_fwd_bwd(M, C):
    ops = _ops()
    x = (torch.randn(M, C, device="cuda") * 2 + 0.3).to(BF)
    g = (1 + 0.1 * torch.randn(C, device="cuda")).to(BF)
    y, z, mean, rstd = ops.layernorm_fwd(x, g, readout_div=42.666668)
    ref = F.layer_norm(x.float(), (C,), g.float(), None, 1e-5)
    assert max_abs(y, ref) <= float(ref.abs().max()) * 2 ** -8 + 1e-6
    assert torch.equal(z, (y.float() / 42.666668).to(BF))
    dy = torch.randn(M, C, device="cuda").to(BF)
    dres = torch.randn(M, C, device="cuda").to(BF)
    xr = x.float().requires_grad_(True)
    gr = g.float().requires_grad_(True)
    F.layer_norm(xr, (C,), gr, None, 1e-5).backward(dy.float())
    dx, dg = ops.layernorm_bwd(dy, x, g, mean, rstd, dres=dres)
    assert rel_err(dx, dres.float() + xr.grad) < 4e-3
    assert rel_err(dg, gr.grad) < 4e-3
    dg2 = dg.clone()
    ops.layernorm_bwd(dy, x, g, mean, rstd, dgamma=dg2, accumulate_dgamma=True)
    assert rel_err(dg2, 2 * gr.grad) < 6e-3


def test_rope_real_table_is_cosine_scaling_and_complex_is_rotation():
    ops = _ops()
    B, T, H, d = 2, 24, 2, 128
    C = H * d
    qkv = torch.randn(B * T, 3 * C, device="cuda").to(BF)
    ang = torch.outer(torch.arange(T, device="cuda").float(),
                      1.0 / (10000 ** (torch.arange(0, d, 2, device="cuda").float() / d)))
    cos_bf = torch.cos(ang).to(BF)  # what module.to(bfloat16) leaves of the complex buffer
    out = ops.rope_(qkv.clone(), cos_bf.float().contiguous(), None, T, C, d)
    q = qkv[:, :C].view(B, T, H, d // 2, 2).float() * cos_bf.float().view(1, T, 1, d // 2, 1)
    assert torch.equal(out[:, :C], q.reshape(B * T, C).to(BF))
    assert torch.equal(out[:, 2 * C:], qkv[:, 2 * C:])  # v untouched
    # complex table: rotation and its adjoint
    cos, sin = torch.cos(ang).contiguous(), torch.sin(ang).contiguous()
    rot = ops.rope_(qkv.clone(), cos, sin, T, C, d)
    k = torch.view_as_complex(qkv[:, C:2 * C].float().reshape(B, T, H, d // 2, 2))
    kr = torch.view_as_real(k * torch.polar(torch.ones_like(ang), ang).view(1, T, 1, d // 2)).reshape(B * T, C)
    assert max_abs(rot[:, C:2 * C], kr) <= float(kr.abs().max()) * 2 ** -8
    back = ops.rope_(rot.clone(), cos, sin, T, C, d, inverse=True)
    assert rel_err(back[:, :2 * C], qkv[:, :2 * C]) < 8e-3


@pytest.mark.parametrize("complex_table", [False, True])
def test_rope_fused_into_gemm_epilogue_and_attention_backward(complex_table):
    """The fused paths (c_attn GEMM epilogue; dQ / dK epilogues of the attention backward) against the stand-alone
    rotary kernel applied to the un-fused results."""
    ops = _ops()
    B, T, H, d = 2, 256, 2, 128
    C = H * d
    ang = torch.outer(torch.arange(T, device="cuda").float(),
                      1.0 / (10000 ** (torch.arange(0, d, 2, device="cuda").float() / d)))
    if complex_table:
        cos, sin = torch.cos(ang).contiguous(), torch.sin(ang).contiguous()
    else:
        cos, sin = torch.cos(ang).to(BF).float().contiguous(), None
    g = torch.Generator(device="cuda").manual_seed(5)
    x = (torch.randn(B * T, C, generator=g, device="cuda") * 0.5).to(BF)
    w = (torch.randn(3 * C, C, generator=g, device="cuda") * 0.05).to(BF)
    plain = ops.gemm(x, w)
    want = ops.rope_(plain.clone(), cos, sin, T, C, d)
    fused = ops.gemm(x, w, epilogue=ops.EPI_ROPE, rope=(cos, sin, T, d, 2 * C))
    same = (lambda a, b: rel_err(a, b) < 1e-3) if complex_table else torch.equal  # fma contraction may differ
    assert same(fused, want)                   # same roundings: rb(acc), then rotary in fp32, then rb
    # backward: rotary adjoint inside the dQ / dK epilogues
    qkv = want
    spec = ops.MaskSpec(None, B, H, T)
    y, lse = ops.attention_fwd(qkv, B, T, H, d, 8.0 / C, spec, 0.0, None, impl="tc")
    dy = torch.randn(B * T, C, generator=g, device="cuda").to(BF)
    g0 = ops.attention_bwd(qkv, y, dy, lse, B, T, H, d, 8.0 / C, spec, 0.0, None, impl="tc")
    want_g = ops.rope_(g0.clone(), cos, sin, T, C, d, inverse=True)
    g1 = ops.attention_bwd(qkv, y, dy, lse, B, T, H, d, 8.0 / C, spec, 0.0, None, impl="tc", rope=(cos, sin))
    assert same(g1, want_g)
    g2 = ops.attention_bwd(qkv, y, dy, lse, B, T, H, d, 8.0 / C, spec, 0.0, None, impl="simt", rope=(cos, sin))
    assert rel_err(g2, want_g) < 1.5e-2


def _sdpa_ref(qkv, B, T, H, d, scale, mask4):
    C = H * d
    q, k, v = [t.view(B, T, H, d).transpose(1, 2).float() for t in qkv.float().split(C, dim=1)]
    s = (q @ k.transpose(-1, -2)) * scale
    if mask4 is not None:
        s = s + mask4.float()
    p = torch.softmax(s, dim=-1)
    return (p @ v).transpose(1, 2).reshape(B * T, C), p


@pytest.mark.parametrize("d,T", [(64, 40), (128, 72), (16, 33)])
@pytest.mark.parametrize("masked", [False, True])
def test_attention_generic_fwd_bwd(d, T, masked):
    ops = _ops()
    B, H = 2, 2
    C = H * d
    scale = 8.0 / C
    qkv = (torch.randn(B * T, 3 * C, device="cuda") * 1.5).to(BF)
    mask4 = None
    if masked:
        m3 = torch.full((B, T, T), -1e9, device="cuda")
        m3[0, :T // 2, :T // 2] = 0
        m3[0, T // 2:, T // 2:] = 0
        m3[1, :T - 5, :T - 5] = 0  # last 5 rows fully masked -> uniform attention over all keys
        mask4 = m3.to(BF).unsqueeze(1).expand(-1, H, -1, -1)
    spec = ops.MaskSpec(mask4, B, H, T)
    y, lse = ops.attention_fwd(qkv, B, T, H, d, scale, spec, 0.0, None, impl="simt")
    qr = qkv.float().requires_grad_(True)
    ref, _ = _sdpa_ref(qr, B, T, H, d, scale, mask4)
    assert rel_err(y, ref) < 5e-3
    if masked:  # fully-masked rows: mean of v over ALL keys
        v = qkv[:, 2 * C:].view(B, T, C).float()
        assert rel_err(y.view(B, T, C)[1, T - 5:], v[1].mean(0, keepdim=True).expand(5, -1)) < 5e-3
    dy = torch.randn(B * T, C, device="cuda").to(BF)
    if masked:
        dy.view(B, T, C)[1, T - 5:] = 0  # fully-masked rows never carry gradient in the reference
    ref.backward(dy.float())
    dqkv = ops.attention_bwd(qkv, y, dy, lse, B, T, H, d, scale, spec, 0.0, None, impl="simt")
    assert rel_err(dqkv, qr.grad) < 8e-3


def test_attention_interval_mask_equals_dense_mask():
    ops = _ops()
    B, T, H, d = 2, 64, 2, 64
    C = H * d
    ids = torch.randint(20, 300, (B, T), device="cuda")
    ids[0, 10] = ids[0, 30] = ids[1, 7] = ids[1, 20] = ids[1, 50] = 3
    lo, hi = ops.doc_mask_intervals(ids, 3, False)
    dense = ops.mask_from_intervals(lo, hi)
    qkv = torch.randn(B * T, 3 * C, device="cuda").to(BF)
    y1, _ = ops.attention_fwd(qkv, B, T, H, d, 8.0 / C, ops.MaskSpec(dense.unsqueeze(1).expand(-1, H, -1, -1), B, H, T),
                              0.0, None, impl="simt")
    y2, _ = ops.attention_fwd(qkv, B, T, H, d, 8.0 / C, ops.MaskSpec(None, B, H, T, lo, hi), 0.0, None, impl="simt")
    assert torch.equal(y1, y2)


def test_mask_builders_bit_exact_vs_oracle(golden):
    import omnibiota_oracle as orc
    ops = _ops()
    for name in ["bf16_h2", "bf16_h1"]:
        c = golden(name)
        for key, ids, padding in [("mask_doc", c["ids"], False), ("mask_docpad", c["ids_pad"], True)]:
            lo, hi = ops.doc_mask_intervals(ids.cuda(), orc.EOS_TOKEN, padding)
            dense = ops.mask_from_intervals(lo, hi)
            assert torch.equal(dense.cpu(), c[key]), (name, key)
            lo2, hi2, flag = ops.mask_compress(c[key].cuda())
            assert int(flag.item()) == 0
            assert torch.equal(ops.mask_from_intervals(lo2, hi2).cpu(), c[key])
        lo, hi = ops.pad_mask_intervals(c["ids_pad"].cuda(), orc.PAD_TOKEN)
        assert torch.equal(ops.mask_from_intervals(lo, hi).cpu(), c["mask_pad"]), name
    # a non-interval mask must be flagged
    m = torch.full((1, 16, 16), -1e9, device="cuda").to(BF)
    m[0, 3, 2] = 0
    m[0, 3, 9] = 0
    assert int(ops.mask_compress(m)[2].item()) == 1


def test_pooling_mean_and_max():
    ops = _ops()
    B, T, C = 3, 150, 256
    emb = torch.randn(B, T, C, device="cuda").to(BF)
    assert torch.equal(ops.pool(emb, "max"), emb.max(dim=1)[0])  # selection: bit-exact
    assert max_abs(ops.pool(emb, "mean"), emb.float().mean(dim=1)) <= 2 ** -8
    assert torch.equal(ops.pool(emb, "mean"), emb.mean(dim=1)) or max_abs(ops.pool(emb, "mean"), emb.mean(dim=1)) <= 2 ** -9


def test_scale_div_and_dropout_statistics():
    ops = _ops()
    x = torch.randn(4096, 256, device="cuda").to(BF)
    assert torch.equal(ops.scale_div(x, 42.666668), (x.float() / 42.666668).to(BF))
    p = 0.1
    y = ops.dropout(x, p, 1234, 0)
    kept = (y != 0) | (x == 0)
    assert abs(float(kept.float().mean()) - (1 - p)) < 5e-3
    assert torch.equal(y[kept], (x.float() / (1 - p)).to(BF)[kept])
    assert torch.equal(ops.dropout(x, p, 1234, 0), y)        # same (seed, offset) -> same mask (backward replay)
    assert not torch.equal(ops.dropout(x, p, 1234, 4), y)    # different offset -> different mask


def test_cross_entropy_mlm_fwd_bwd():
    ops = _ops()
    M, V, n_acc = 300, 1024, 2
    logits = (torch.randn(M, V, device="cuda") * 2).to(BF)
    y = torch.randint(0, V, (M,), device="cuda")
    m = torch.rand(M, device="cuda") < 0.15
    m[0] = True
    lr = logits.float().requires_grad_(True)
    ce = F.cross_entropy(lr, y, reduction="none")
    ref_loss = ((ce / n_acc) * m.float()).sum() / m.sum()
    ref_loss.backward()
    scalars, lse, tok, rm, tgt = ops.ce_fwd(logits, y, m, n_acc)
    assert abs(float(scalars[0]) - float(ref_loss)) <= 2 ** -7 * float(ref_loss)  # loss is bf16-rounded twice
    assert int(scalars[1]) == int(m.sum())
    ops.ce_bwd_(logits, tgt, rm, lse, scalars)
    assert float(logits[~m].abs().max()) == 0.0   # exact zeros outside the MLM mask
    assert rel_err(logits[m], lr.grad[m]) < 2e-2  # g = rb(rb(1/count)/n_acc) carries up to 2 bf16 roundings


def test_fused_adamw_matches_torch_bf16_foreach(golden):
    from omnibiote_b200.optim import FusedAdamW
    a = golden("adamw_bf16")
    p = torch.nn.Parameter(a["p0"].clone().cuda())
    opt = FusedAdamW([p], lr=a["lr"], weight_decay=a["wd"], betas=(0.9, 0.999), eps=1e-8)
    for s in range(3):
        p.grad = a["g"][s].clone().cuda()
        opt.step()
        st = opt.state[p]
        assert max_abs(st["exp_avg"], a["m"][s]) <= 2 ** -7 * float(a["m"][s].abs().max())  # <= 1 bf16 ulp
        assert max_abs(st["exp_avg_sq"], a["v"][s]) <= 2 ** -7 * float(a["v"][s].abs().max())
        assert max_abs(p, a["p"][s]) <= 2 ** -7 * float(a["p"][s].abs().max())


def test_row_compaction_gather_scatter():
    ops = _ops()
    torch.manual_seed(0)
    M, C, cap = 5000, 256, 1024
    mask = torch.rand(M, device="cuda") < 0.15
    tgt = torch.randint(0, 1000, (M,), device="cuda")
    idx, tgt_c, valid_c, meta = ops.compact_rows(mask, tgt, cap)
    want = mask.nonzero().flatten()
    n = want.numel()
    assert meta.tolist() == [n, 0]
    assert torch.equal(idx[:n].long(), want) and bool((idx[n:] == -1).all())          # row order preserved, -1 padded
    assert torch.equal(tgt_c[:n], tgt[want]) and torch.equal(valid_c.bool(), torch.arange(cap, device="cuda") < n)
    x = torch.randn(M, C, device="cuda").to(BF)
    g = ops.gather_rows(x, idx)
    assert torch.equal(g[:n], x[want]) and float(g[n:].abs().max()) == 0.0
    back = ops.scatter_rows(g, idx, M)
    assert torch.equal(back[want], x[want]) and float(back[~mask].abs().max()) == 0.0
    # capacity too small: flagged, never written out of bounds
    idx2, _, _, meta2 = ops.compact_rows(mask, tgt, 64)
    assert meta2.tolist() == [n, 1] and torch.equal(idx2.long(), want[:64])
