"""GPU: tcgen05 attention forward (head_dim 128) against an fp32 torch restatement of
F.scaled_dot_product_attention(q,k,v,attn_mask,scale=8/n_embd) and against the generic CUDA-core kernel."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(autouse=True, params=[False, True], ids=["dq_standalone", "dq_from_ds_handover"])
def _backward_variant(request, monkeypatch):
    """every test of this module runs with both backward schedules: stand-alone dQ kernel (re-evaluates the scores), and
    the dK/dV kernel handing its dS tiles to the score-free dQ kernel (ops.ATTN_DS_HANDOVER, the default)"""
    from omnibiote_b200 import ops
    monkeypatch.setattr(ops, "ATTN_DS_HANDOVER", request.param)
    yield


def _ref(qkv, B, T, H, d, scale, mask4):
    C = H * d
    q, k, v = [t.view(B, T, H, d).transpose(1, 2).float() for t in qkv.float().split(C, dim=1)]
    s = (q @ k.transpose(-1, -2)) * scale
    if mask4 is not None:
        s = s + mask4.float()
    p = torch.softmax(s, dim=-1)
    return (p @ v).transpose(1, 2).reshape(B * T, C)


def _doc_ids(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(20, 1000, (B, T), generator=g)
    for b in range(B):
        pos = 0
        while True:
            pos += int(torch.randint(5, max(6, T // 3), (1,), generator=g))
            if pos >= T:
                break
            ids[b, pos] = 3
    return ids.cuda()


@pytest.mark.parametrize("B,T,H", [(1, 128, 1), (2, 256, 2), (1, 200, 2), (2, 1024, 2), (1, 77, 1)])
@pytest.mark.parametrize("mode", ["none", "interval", "dense"])
def test_attn_tc_forward(B, T, H, mode):
    from omnibiote_b200 import ops
    d = 128
    C = H * d
    scale = 8.0 / C
    torch.manual_seed(B * 1000 + T)
    qkv = (torch.randn(B * T, 3 * C, device="cuda") * 1.5).to(BF)
    mask4, spec = None, ops.MaskSpec(None, B, H, T)
    if mode != "none":
        ids = _doc_ids(B, T, T)
        lo, hi = ops.doc_mask_intervals(ids, 3, True)  # padding=True: tokens after the last EOS are fully masked
        if T % 8 == 0:
            dense = ops.mask_from_intervals(lo, hi)
        else:
            j = torch.arange(T, device="cuda").view(1, 1, T)
            dense = torch.where((j >= lo.unsqueeze(-1)) & (j < hi.unsqueeze(-1)), 0.0, -1e9).to(BF)
        mask4 = dense.unsqueeze(1).expand(-1, H, -1, -1)
        if mode == "interval":
            spec = ops.MaskSpec(None, B, H, T, lo, hi)
        else:
            if T % 8:
                pytest.skip("dense bias path needs 16-byte aligned mask rows")
            spec = ops.MaskSpec(mask4, B, H, T)
    y, lse = ops.attention_fwd(qkv, B, T, H, d, scale, spec, 0.0, None, impl="tc")
    torch.cuda.synchronize()
    ref = _ref(qkv, B, T, H, d, scale, mask4)
    err = rel_err(y, ref)
    assert err < 8e-3, (mode, B, T, H, err)
    y2, lse2 = ops.attention_fwd(qkv, B, T, H, d, scale, spec, 0.0, None, impl="simt")
    assert rel_err(y, y2) < 8e-3
    # log-sum-exp bookkeeping agrees with the generic kernel (needed by the backward)
    tot, tot2 = lse[..., 0] + lse[..., 1], lse2[..., 0] + lse2[..., 1]
    fin = tot2.abs() < 1e6  # rows with a -1e9 row max keep (max, logsum) apart; compare the logsum part there
    assert float((tot[fin] - tot2[fin]).abs().max()) < 2e-2


@pytest.mark.parametrize("B,T,H", [(1, 128, 1), (2, 256, 2), (1, 200, 2), (2, 512, 2)])
@pytest.mark.parametrize("mode", ["none", "interval", "dense"])
def test_attn_tc_backward(B, T, H, mode):
    from omnibiote_b200 import ops
    d = 128
    C = H * d
    scale = 8.0 / C
    torch.manual_seed(7 + B * 1000 + T)
    qkv = (torch.randn(B * T, 3 * C, device="cuda") * 1.5).to(BF)
    mask4, spec = None, ops.MaskSpec(None, B, H, T)
    live = torch.ones(B, T, dtype=torch.bool, device="cuda")
    if mode != "none":
        ids = _doc_ids(B, T, T + 1)
        lo, hi = ops.doc_mask_intervals(ids, 3, True)
        live = hi > lo  # fully-masked rows never carry gradient in the reference's usage
        j = torch.arange(T, device="cuda").view(1, 1, T)
        dense = torch.where((j >= lo.unsqueeze(-1)) & (j < hi.unsqueeze(-1)), 0.0, -1e9).to(BF)
        mask4 = dense.unsqueeze(1).expand(-1, H, -1, -1)
        spec = ops.MaskSpec(None, B, H, T, lo, hi) if mode == "interval" else ops.MaskSpec(mask4, B, H, T)
    y, lse = ops.attention_fwd(qkv, B, T, H, d, scale, spec, 0.0, None, impl="tc" if (mode != "dense" or T % 8 == 0) else "simt")
    dy = torch.randn(B * T, C, device="cuda").to(BF)
    dy = (dy.view(B, T, C) * live.unsqueeze(-1)).reshape(B * T, C).contiguous()
    qr = qkv.float().requires_grad_(True)
    _ref(qr, B, T, H, d, scale, mask4).backward(dy.float())
    dqkv = ops.attention_bwd(qkv, y, dy, lse, B, T, H, d, scale, spec, 0.0, None, impl="tc")
    torch.cuda.synchronize()
    for name, sl in [("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))]:
        err = rel_err(dqkv[:, sl], qr.grad[:, sl])
        assert err < 1.5e-2, (mode, B, T, H, name, err)
    dqkv2 = ops.attention_bwd(qkv, y, dy, lse, B, T, H, d, scale, spec, 0.0, None, impl="simt")
    assert rel_err(dqkv, dqkv2) < 1.5e-2


def _ref_drop(qkv, B, T, H, d, scale, mask4, keep_bool, p):
    """fp32 torch restatement of SDPA with an explicit dropout keep mask (dropout on the softmax output)."""
    C = H * d
    q, k, v = [t.view(B, T, H, d).transpose(1, 2).float() for t in qkv.float().split(C, dim=1)]
    s = (q @ k.transpose(-1, -2)) * scale
    if mask4 is not None:
        s = s + mask4.float()
    pr = torch.softmax(s, dim=-1) * keep_bool.float() / (1.0 - p)
    return (pr @ v).transpose(1, 2).reshape(B * T, C)


def test_keep_mask_statistics_and_determinism():
    from omnibiote_b200 import ops
    B, H, T, p = 2, 4, 1024, 0.1
    keep = ops.attn_keep_mask(B, H, T, p, 1234, 0, "cuda")
    kb = ops.keep_mask_to_bool(keep, T)
    assert kb.shape == (B, H, T, T)
    rate = float(kb.float().mean())
    assert abs(rate - (1 - p)) < 5e-4, rate                      # 8.4 M draws: sigma = 1e-4
    # per-row and per-column rates (no structure along either axis), adjacent-key correlation
    assert float((kb.float().mean(-1) - (1 - p)).abs().max()) < 0.06
    assert float((kb.float().mean(-2) - (1 - p)).abs().max()) < 0.06
    a, b_ = kb[..., :-1].float() - (1 - p), kb[..., 1:].float() - (1 - p)
    assert abs(float((a * b_).mean()) / (p * (1 - p))) < 5e-3
    a, b_ = kb[..., :-1, :].float() - (1 - p), kb[..., 1:, :].float() - (1 - p)
    assert abs(float((a * b_).mean()) / (p * (1 - p))) < 5e-3
    assert torch.equal(ops.attn_keep_mask(B, H, T, p, 1234, 0, "cuda"), keep)       # same (seed, offset): replay
    assert not torch.equal(ops.attn_keep_mask(B, H, T, p, 1234, 4, "cuda"), keep)   # different offset
    assert not torch.equal(ops.attn_keep_mask(B, H, T, p, 1235, 0, "cuda"), keep)   # different seed
    for pp in (0.25, 0.5):
        r = float(ops.keep_mask_to_bool(ops.attn_keep_mask(1, 2, 512, pp, 7, 0, "cuda"), 512).float().mean())
        assert abs(r - (1 - pp)) < 3e-3, (pp, r)


@pytest.mark.parametrize("B,T,H", [(2, 256, 2), (1, 200, 1), (2, 1024, 2)])
@pytest.mark.parametrize("mode", ["none", "interval"])
def test_attn_dropout_against_torch_with_the_same_keep_mask(B, T, H, mode):
    """Forward and backward with dropout, elementwise against an fp32 torch restatement that applies the SAME keep
    mask (unpacked from the bit matrix the kernels read)."""
    from omnibiote_b200 import ops
    d = 128
    C = H * d
    scale = 8.0 / C
    p = 0.1
    torch.manual_seed(11 + T)
    qkv = (torch.randn(B * T, 3 * C, device="cuda") * 1.5).to(BF)
    mask4, spec = None, ops.MaskSpec(None, B, H, T)
    if mode == "interval":
        ids = _doc_ids(B, T, T + 5)
        lo, hi = ops.doc_mask_intervals(ids, 3, False)
        j = torch.arange(T, device="cuda").view(1, 1, T)
        dense = torch.where((j >= lo.unsqueeze(-1)) & (j < hi.unsqueeze(-1)), 0.0, -1e9).to(BF)
        mask4 = dense.unsqueeze(1).expand(-1, H, -1, -1)
        spec = ops.MaskSpec(None, B, H, T, lo, hi)
    keep = ops.attn_keep_mask(B, H, T, p, 42, 8, "cuda")
    kb = ops.keep_mask_to_bool(keep, T)
    y, lse = ops.attention_fwd(qkv, B, T, H, d, scale, spec, p, keep, impl="tc")
    qr = qkv.float().requires_grad_(True)
    ref = _ref_drop(qr, B, T, H, d, scale, mask4, kb, p)
    assert rel_err(y, ref) < 8e-3
    dy = torch.randn(B * T, C, device="cuda").to(BF)
    ref.backward(dy.float())
    dqkv = ops.attention_bwd(qkv, y, dy, lse, B, T, H, d, scale, spec, p, keep, impl="tc")
    torch.cuda.synchronize()
    for name, sl in [("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))]:
        err = rel_err(dqkv[:, sl], qr.grad[:, sl])
        assert err < 1.5e-2, (mode, B, T, H, name, err)


def test_attn_dropout_same_mask_in_both_kernels_and_in_backward():
    from omnibiote_b200 import ops
    B, T, H, d = 2, 256, 2, 128
    C = H * d
    scale = 8.0 / C
    torch.manual_seed(3)
    qkv = (torch.randn(B * T, 3 * C, device="cuda")).to(BF)
    spec = ops.MaskSpec(None, B, H, T)
    p = 0.25
    keep = ops.attn_keep_mask(B, H, T, p, 99, 8, "cuda")
    y_tc, lse_tc = ops.attention_fwd(qkv, B, T, H, d, scale, spec, p, keep, impl="tc")
    y_si, lse_si = ops.attention_fwd(qkv, B, T, H, d, scale, spec, p, keep, impl="simt")
    y_nd, _ = ops.attention_fwd(qkv, B, T, H, d, scale, spec, 0.0, None, impl="tc")
    assert rel_err(y_tc, y_si) < 1e-2           # identical keep mask in both kernels
    assert rel_err(y_tc, y_nd) > 5e-2           # and it really drops something
    dy = torch.randn(B * T, C, device="cuda").to(BF)
    g_tc = ops.attention_bwd(qkv, y_tc, dy, lse_tc, B, T, H, d, scale, spec, p, keep, impl="tc")
    g_si = ops.attention_bwd(qkv, y_si, dy, lse_si, B, T, H, d, scale, spec, p, keep, impl="simt")
    assert rel_err(g_tc, g_si) < 2e-2


@pytest.mark.parametrize("T", [1024, 512, 200])   # 1024 / 512: warp per 32 rows x 4 words; 200: linear mapping
def test_keep_mask_skips_words_outside_the_visible_interval(T):
    """With an interval mask the generator stores all-ones for the 32-key words no key of which the row can see, and the
    SAME bits as the full draw everywhere else; fully-masked rows (uniform attention over every key) are drawn in full.
    Attention with either mask is bit-identical."""
    from omnibiote_b200 import ops
    B, H, d, p = 2, 2, 128, 0.1
    C = H * d
    ids = _doc_ids(B, T, T + 9)
    ids[1, T - 40:] = 1
    lo, hi = ops.doc_mask_intervals(ids, 3, True)
    assert bool((lo >= hi).any())
    spec = ops.MaskSpec(None, B, H, T, lo, hi)
    full = ops.attn_keep_mask(B, H, T, p, 5, 16, "cuda")
    part = ops.attn_keep_mask(B, H, T, p, 5, 16, "cuda", spec)
    nw = full.shape[-1]
    w = torch.arange(nw, device="cuda").view(1, 1, nw) * 32
    dead = (lo >= hi).unsqueeze(-1)
    visible = dead | ((w < hi.unsqueeze(-1)) & (w + 32 > lo.unsqueeze(-1)))       # [B,T,nw]
    visible = visible.unsqueeze(1).expand(B, H, T, nw)
    assert torch.equal(part[visible], full[visible])
    assert bool((part[~visible] == -1).all())
    assert 0.2 < float((~visible).float().mean()) < 0.95                            # the test really exercises both
    torch.manual_seed(T)
    qkv = (torch.randn(B * T, 3 * C, device="cuda") * 1.5).to(BF)
    scale = 8.0 / C
    y0, lse0 = ops.attention_fwd(qkv, B, T, H, d, scale, spec, p, full, impl="tc")
    y1, lse1 = ops.attention_fwd(qkv, B, T, H, d, scale, spec, p, part, impl="tc")
    assert torch.equal(y0, y1) and torch.equal(lse0, lse1)
    dy = torch.randn(B * T, C, device="cuda").to(BF)
    g0 = ops.attention_bwd(qkv, y0, dy, lse0, B, T, H, d, scale, spec, p, full, impl="tc")
    g1 = ops.attention_bwd(qkv, y1, dy, lse1, B, T, H, d, scale, spec, p, part, impl="tc")
    assert torch.equal(g0, g1)
    # a dense-bias MaskSpec carries no intervals: full draw
    assert torch.equal(ops.attn_keep_mask(B, H, T, p, 5, 16, "cuda", ops.MaskSpec(None, B, H, T)), full)


def test_tile_metadata_of_interval_masks():
    """obt_attn_tile_meta against a torch restatement: per 128-query tile {min lo, max hi, any fully-masked row}, per
    128-key tile the relevance bits of the 64-query sub-tiles."""
    from omnibiote_b200 import ops
    B, T = 3, 700
    ids = _doc_ids(B, T, 5)
    ids[1, 600:] = 1  # padded tail -> fully-masked rows with padding=True
    lo, hi = ops.doc_mask_intervals(ids, 3, True)
    spec = ops.MaskSpec(None, B, 2, T, lo, hi)
    qmeta, kmeta = spec.tile_meta()
    assert spec.tile_meta()[0] is qmeta                                 # cached: one launch per micro-batch
    nT = (T + 127) // 128
    lo_c, hi_c = lo.cpu(), hi.cpu()
    for b in range(B):
        for t in range(nT):
            rows = slice(t * 128, min(T, (t + 1) * 128))
            l, h = lo_c[b, rows], hi_c[b, rows]
            dead = l >= h
            live = ~dead
            want = [int(l[live].min()) if live.any() else T, int(h[live].max()) if live.any() else 0, int(dead.any()), 0]
            assert qmeta[b, t].tolist() == want, (b, t)
            bits = 0
            for it in range((T + 63) // 64):
                r = slice(it * 64, min(T, (it + 1) * 64))
                ll, hh = lo_c[b, r], hi_c[b, r]
                if bool(((ll >= hh) | ((ll < t * 128 + 128) & (hh > t * 128))).any()):
                    bits |= 1 << it
            got = sum((int(kmeta[b, t, w]) & 0xffffffff) << (32 * w) for w in range(4))
            assert got == bits, (b, t, hex(got), hex(bits))
