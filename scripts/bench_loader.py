"""CPU throughput of the batch loader: omnibiote_b200/loader.py against the oracle restatement of the reference's
training/loader.py on the same synthetic shards (tokens/s of packed ctx-1024 batches, torch int64 output)."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import loader_oracle as orc  # noqa: E402
from omnibiote_b200 import loader as ours  # noqa: E402


def shards(d, n_files=4, tokens_per_file=4_000_000, seed=0):
    rng = np.random.RandomState(seed)
    names = []
    for f in range(n_files):
        lens = np.clip(rng.lognormal(np.log(200.0), 1.0, size=tokens_per_file // 300), 8, 4096).astype(np.int64)
        arr = rng.randint(20, 65533, size=int(lens.sum()) + len(lens)).astype(np.uint16)
        arr[np.cumsum(lens + 1) - 1] = 3
        p = os.path.join(d, f"s{f}.npy")
        np.save(p, arr)
        names.append(p)
    return names


def time_it(mod, files, ctx, batch, n_batches, torch_out):
    np.random.seed(0)
    gens = [mod.get_sequence(mod.line_reader(list(files), banned_tokens=[65533]), ctx, False)]
    bg = mod.get_batch(gens, [batch], return_pt=True) if torch_out else mod.get_batch(gens, [batch])
    next(bg)  # first chunk load
    t0 = time.perf_counter()
    for _ in range(n_batches):
        b = next(bg)
        if not torch_out:
            b = torch.tensor(b.tolist() if False else b, dtype=torch.long)
    return batch * ctx * n_batches / (time.perf_counter() - t0)


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as d:
        files = shards(d)
        ctx, batch, n = 1024, 256, 8
        ref_tps = time_it(orc, files, ctx, batch, n, torch_out=False)
        our_tps = time_it(ours, files, ctx, batch, n, torch_out=True)
    print(json.dumps({"metric": "loader_tokens_per_s", "ctx_len": ctx, "batch": batch,
                      "reference_port": ref_tps, "omnibiote_b200": our_tps, "speedup": our_tps / ref_tps,
                      "note": "one host thread each; reference port yields lists of Python ints (torch.tensor of the "
                              "stacked array timed on top), this loader yields int32 rows and one int64 tensor"}))
