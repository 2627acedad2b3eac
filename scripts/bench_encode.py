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


def main():
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
