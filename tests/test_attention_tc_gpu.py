"""GPU: tcgen05 attention forward (head_dim 128) against an fp32 torch restatement of
F.scaled_dot_product_attention(q,k,v,attn_mask,scale=8/n_embd) and against the generic CUDA-core kernel."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


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
    y, lse = ops.attention_fwd(qkv, B, T, H, d, scale, spec, 0.0, 0, 0, impl="tc")
    torch.cuda.synchronize()
    ref = _ref(qkv, B, T, H, d, scale, mask4)
    err = rel_err(y, ref)
    assert err < 8e-3, (mode, B, T, H, err)
    y2, lse2 = ops.attention_fwd(qkv, B, T, H, d, scale, spec, 0.0, 0, 0, impl="simt")
    assert rel_err(y, y2) < 8e-3
    # log-sum-exp bookkeeping agrees with the generic kernel (needed by the backward)
    tot, tot2 = lse[..., 0] + lse[..., 1], lse2[..., 0] + lse2[..., 1]
    fin = tot2.abs() < 1e6  # rows with a -1e9 row max keep (max, logsum) apart; compare the logsum part there
    assert float((tot[fin] - tot2[fin]).abs().max()) < 2e-2
