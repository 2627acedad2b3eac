#!/usr/bin/env python
"""encode() inference throughput (BASELINE configs 1 and 5): omnibiote-small, eval(), variable-length padded
mixed nucleotide/peptide batches, methods all / max / mean at ctx 1024 and 4096 (no attn_mask: that is what
``OmniBioTA.encode`` does, training/model.py:268). Prints one JSON line per case."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def cpu_baseline(B=2, T=1024):
    """BASELINE config 1 / SURVEY §8d: the reference's CPU path (oracle port, bit-exact with training/model.py on CPU),
    omnibiote-small, fp32, encode(method="mean"), batch 2, ctx 1024, all host threads; 2 warm-ups + best of 5."""
    import time
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import omnibiota_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    C, V, L, H = 1024, 65536, 8, 8
    lin = lambda o, i: (torch.rand(o, i) * 2 - 1) / np.sqrt(i)
    p = {"transformer.wte.weight": torch.randn(V, C)}
    for l in range(L):
        pre = f"transformer.h.{l}."
        p[pre + "ln_1.weight"] = torch.ones(C)
        p[pre + "attn.freqs_cis"] = orc.precompute_freqs_cis(C // H, T)
        p[pre + "attn.c_attn.weight"] = lin(3 * C, C)
        p[pre + "attn.c_proj.weight"] = lin(C, C)
        p[pre + "ln_2.weight"] = torch.ones(C)
        p[pre + "mlp.c_fc.weight"] = lin(4 * C, C)
        p[pre + "mlp.c_proj.weight"] = lin(C, 4 * C)
    p["transformer.ln_f.weight"] = torch.ones(C)
    ids = torch.from_numpy(bench.synth_ids(B, T, np.random.RandomState(7), padded=True))
    best = float("inf")
    with torch.no_grad():
        for it in range(7):
            t0 = time.perf_counter()
            emb = orc.forward(p, L, H, ids, None, return_embeddings=True)
            out = emb.mean(dim=1)
            dt = time.perf_counter() - t0
            if it >= 2:
                best = min(best, dt)
    r = {"metric": "encode_sequences_per_s", "impl": "reference (oracle port, CPU fp32)", "method": "mean", "ctx_len": T,
         "batch": B, "value": B / best, "unit": "sequences/s", "ms_per_batch": best * 1e3, "cores": os.cpu_count(),
         "out_shape": list(out.shape)}
    print(json.dumps(r))
    return r


def main():
    if "--cpu-baseline" in sys.argv:
        cpu_baseline()
        if "--cpu-only" in sys.argv:
            return []
    device = torch.device("cuda", 0)
    torch.manual_seed(0)
    results = []
    for T, B in ((1024, 32), (4096, 8)):
        bench.SMALL["block_size"] = T
        model = bench.build_model(device, 0.0).eval()
        rng = np.random.RandomState(7)
        ids = torch.from_numpy(bench.synth_ids(B, T, rng, padded=True)).to(device)
        flops_tok = 24 * 8 * 1024 ** 2 + 4 * 8 * 1024 * T  # 2*12*L*C^2 + 4*L*C*T
        for method in ("all", "max", "mean"):
            with torch.no_grad():
                for _ in range(3):
                    model.encode(ids, method)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 10
                e0.record()
                for _ in range(n):
                    out = model.encode(ids, method)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            r = {"metric": "encode_sequences_per_s", "method": method, "ctx_len": T, "batch": B,
                 "value": B / (ms / 1e3), "unit": "sequences/s", "ms_per_batch": ms,
                 "tflops": flops_tok * B * T / (ms / 1e3) / 1e12, "out_shape": list(out.shape)}
            results.append(r)
            print(json.dumps(r))
        del model
        torch.cuda.empty_cache()
    return results


if __name__ == "__main__":
    main()
