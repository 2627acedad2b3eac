"""GPU: BASELINE-size shapes through size-independent properties and on-device fp32 restatements.

* small (8L/1024d/8h, ctx 1024) and large (n_embd 2048, 16 heads) widths: one full block forward + backward against
  the oracle's block math evaluated in fp32 on the GPU (same bf16 weights) — tolerance = the bf16 noise floor.
* ctx 4096 attention (BASELINE config 5): tensor-core kernel vs the generic CUDA-core kernel.
* gradient accumulation: two micro-batches accumulated in place == sum of the separate gradients (bf16 `+=`).
* idempotence / determinism: the same batch twice gives bit-identical logits (no atomics on the forward path).
"""
import copy
import contextlib
import io
import warnings

import pytest
import torch

import omnibiota_oracle as orc
from conftest import rel_err

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def make_model(n_layer, n_embd, n_head, vocab, block_size, seed=0):
    from omnibiote_b200.model import OmniBioTA, OmniBioTAConfig
    from omnibiote_b200.mup import set_base_shapes
    torch.manual_seed(seed)
    cfg = OmniBioTAConfig()
    cfg.vocab_size, cfg.block_size, cfg.n_layer, cfg.n_head, cfg.n_embd, cfg.dropout = vocab, block_size, n_layer, n_head, n_embd, 0.0
    cfg.flash = True
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = OmniBioTA(cfg)
        c2 = copy.copy(cfg); c2.n_embd, c2.n_head = 24, 3
        c3 = copy.copy(cfg); c3.n_embd, c3.n_head = 48, 12
        set_base_shapes(m, OmniBioTA(c2), delta=OmniBioTA(c3))
        m.to(BF)
    return m.cuda()


@pytest.mark.parametrize("n_embd,n_head,T,B", [(1024, 8, 1024, 2), (2048, 16, 512, 1)])
def test_full_width_model_against_fp32_restatement_on_device(n_embd, n_head, T, B):
    vocab, L = 2048, 2
    model = make_model(L, n_embd, n_head, vocab, T).train()
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(20, vocab, (B, T), generator=g)
    ids[:, T // 3] = orc.EOS_TOKEN
    ids[:, (2 * T) // 3] = orc.EOS_TOKEN
    lm = (torch.rand(B, T, generator=g) < 0.15) & (ids != orc.EOS_TOKEN)
    masked = ids.masked_fill(lm, orc.MASK_TOKEN)
    mask3 = orc.create_attention_mask(torch.ones(B, T, T, dtype=BF) * -1e9, ids, padding=False).cuda()
    mask4 = mask3.unsqueeze(1).expand(-1, n_head, -1, -1)
    loss, _ = model.mlm_loss(masked.cuda(), ids.cuda(), lm.cuda(), attn_mask=mask4, n_accum=1)
    loss.backward()
    # fp32 twin on the GPU holding the same bf16-rounded weights and the same real cosine table
    p = {k: v.detach().float().clone().requires_grad_(v.is_floating_point() and "freqs" not in k)
         for k, v in model.state_dict().items()}
    logits = orc.forward(p, L, n_head, masked.cuda(), mask4.float(), readout_width_mult=model.lm_head.width_mult())
    ref = orc.mlm_loss(logits, ids.cuda(), lm.cuda(), 1)
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 2 ** -6 * float(ref), (float(loss), float(ref))
    worst = {}
    for n, q in model.named_parameters():
        worst[n] = rel_err(q.grad, p[n].grad)
    print({k: f"{v:.1e}" for k, v in worst.items()})
    for n, v in worst.items():
        assert v < (6e-2 if "wte" in n else 3e-2), (n, v)


def test_attention_ctx4096_tc_vs_generic():
    from omnibiote_b200 import ops
    B, T, H, d = 1, 4096, 2, 128
    C = H * d
    qkv = torch.randn(B * T, 3 * C, device="cuda").to(BF)
    spec = ops.MaskSpec(None, B, H, T)
    y1, l1 = ops.attention_fwd(qkv, B, T, H, d, 8.0 / 1024, spec, 0.0, None, impl="tc")
    y2, l2 = ops.attention_fwd(qkv, B, T, H, d, 8.0 / 1024, spec, 0.0, None, impl="simt")
    assert rel_err(y1, y2) < 8e-3
    dy = torch.randn(B * T, C, device="cuda").to(BF)
    g1 = ops.attention_bwd(qkv, y1, dy, l1, B, T, H, d, 8.0 / 1024, spec, 0.0, None, impl="tc")
    g2 = ops.attention_bwd(qkv, y2, dy, l2, B, T, H, d, 8.0 / 1024, spec, 0.0, None, impl="simt")
    assert rel_err(g1, g2) < 1.5e-2


def test_gradient_accumulation_in_place_equals_sum():
    from omnibiote_b200 import functional as Fn
    from omnibiote_b200.parallel import FlatGradBuckets, model_buckets
    vocab, T, B = 1024, 256, 2
    model = make_model(2, 256, 2, vocab, T).train()
    g = torch.Generator().manual_seed(11)
    batches = []
    for _ in range(2):
        ids = torch.randint(20, vocab, (B, T), generator=g).cuda()
        lm = (torch.rand(B, T, generator=g) < 0.15).cuda()
        batches.append((ids.masked_fill(lm, 2), ids, lm))
    separate = []
    for x, y, m in batches:
        model.zero_grad(set_to_none=True)
        loss, _ = model.mlm_loss(x, y, m, n_accum=2)
        loss.backward()
        separate.append({n: p.grad.float().clone() for n, p in model.named_parameters()})
    model.zero_grad(set_to_none=True)
    buckets = FlatGradBuckets(model_buckets(model))
    with Fn.direct_grad_accumulation(True):
        for x, y, m in batches:
            loss, _ = model.mlm_loss(x, y, m, n_accum=2)
            loss.backward()
    for n, p in model.named_parameters():
        want = separate[0][n] + separate[1][n]
        assert p.grad.data_ptr() >= buckets.flat.data_ptr()
        assert rel_err(p.grad, want) < 1e-2, n


def test_forward_is_deterministic():
    model = make_model(2, 256, 2, 1024, 256).eval()
    ids = torch.randint(20, 1024, (3, 200), device="cuda")
    with torch.no_grad():
        a = model(ids)
        b = model(ids)
    assert torch.equal(a, b)
