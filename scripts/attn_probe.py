#!/usr/bin/env python
"""Attention kernels alone at the benchmark's shape and mask mix (B=32, H=8, T=1024, d=128, packed-document interval
masks from bench.synth_ids, dropout 0.1): CUDA-event time of the forward and of the whole backward call
(delta + dQ + dK/dV), model TFLOP/s, and an element check of both against the fp32 torch restatement on a slice.
Used for A/B runs of kernel variants (one JSON line per run)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from omnibiote_b200 import ops  # noqa: E402


def main():
    B, H, T, d = int(os.environ.get("PROBE_B", 32)), 8, int(os.environ.get("PROBE_T", 1024)), 128
    p = float(os.environ.get("PROBE_P", 0.1))
    reps = int(os.environ.get("PROBE_REPS", 50))
    C, dev = H * d, torch.device("cuda", 0)
    scale = 8.0 / C
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = (torch.randn(B * T, 3 * C, generator=g, device=dev) * 1.0).to(torch.bfloat16)
    dy = (torch.randn(B * T, C, generator=g, device=dev) * 0.1).to(torch.bfloat16)
    ids = torch.from_numpy(bench.synth_ids(B, T, np.random.RandomState(1234))).to(dev)
    lo, hi = ops.doc_mask_intervals(ids, 3, False)
    spec = ops.MaskSpec(None, B, H, T, lo, hi) if os.environ.get("PROBE_NOMASK") is None else ops.MaskSpec(None, B, H, T)
    keep = ops.attn_keep_mask(B, H, T, p, 1, 0, dev, spec) if p > 0 else None

    def timed(fn):
        for _ in range(min(10, max(3, reps // 4))):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, out

    ms_f, (y, lse) = timed(lambda: ops.attention_fwd(qkv, B, T, H, d, scale, spec, p, keep, impl="tc"))
    ms_b, dqkv = timed(lambda: ops.attention_bwd(qkv, y, dy, lse, B, T, H, d, scale, spec, p, keep, impl="tc"))
    ms_k, _ = timed(lambda: ops.attn_keep_mask(B, H, T, p, 1, 0, dev, spec)) if p > 0 else (0.0, None)

    # element check on batch row 0 against fp32 torch with the same keep mask
    b0 = slice(0, T)
    q, k, v = [t.view(T, H, d).transpose(0, 1).float() for t in qkv[b0].float().split(C, dim=1)]
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    s = (q @ k.transpose(-1, -2)) * scale
    if spec.row_lo is not None:
        j = torch.arange(T, device=dev)
        vis = (j[None, :] >= lo[0][:, None]) & (j[None, :] < hi[0][:, None])
        s = s + torch.where(vis, 0.0, -1e9)[None]
    P = torch.softmax(s, dim=-1)
    if p > 0:
        P = P * ops.keep_mask_to_bool(keep[:1], T)[0].float() / (1 - p)
    ref = (P @ v).transpose(0, 1).reshape(T, C)
    ref.backward(dy[b0].float())
    ref_g = torch.cat([t.grad.transpose(0, 1).reshape(T, C) for t in (q, k, v)], dim=1)
    rel = lambda a, b: float((a.float() - b).norm() / (b.norm() + 1e-30))
    flops_f = 4.0 * B * H * T * T * d
    out = {"B": B, "T": T, "drop_p": p, "fwd_ms": ms_f, "bwd_ms": ms_b, "keep_mask_ms": ms_k,
           "fwd_model_tflops": flops_f / ms_f / 1e9, "bwd_model_tflops": 2.5 * flops_f / ms_b / 1e9,
           "fwd_rel_err": rel(y[b0], ref.detach()), "bwd_rel_err": rel(dqkv[b0], ref_g),
           "variant": os.environ.get("OBT_ATTN_VARIANT", "default")}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
