"""Generates tests/golden/loader_golden.npz with the UNMODIFIED reference loader (/root/reference/training/loader.py):
synthetic shards (deterministic), then batches for padding on / off. Run in the build container only."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_shards(d, n_files, seed, lo=20, hi=400, banned=65533):
    rng = np.random.RandomState(seed)
    names = []
    for f in range(n_files):
        toks = []
        for _ in range(rng.randint(30, 60)):
            n = int(np.clip(rng.lognormal(np.log(40.0), 1.0), 1, 400))
            body = rng.randint(lo, hi, size=n)
            body[rng.rand(n) < 0.02] = banned
            toks.append(np.concatenate([[4], body, [3]]))
        if f % 2 == 0:
            toks.append(rng.randint(lo, hi, size=7))  # trailing piece without EOS
        arr = np.concatenate(toks).astype(np.uint16)
        p = os.path.join(d, f"shard_{f:02d}.npy")
        np.save(p, arr)
        names.append(p)
    return names


def run(mod, files_a, files_b, ctx, padding, n_batches, seed):
    np.random.seed(seed)
    readers = [mod.line_reader(list(files_a), banned_tokens=[65533]), mod.line_reader(list(files_b), banned_tokens=[65533, 21])]
    gens = [mod.get_sequence(r, ctx, padding) for r in readers]
    bg = mod.get_batch(gens, [5, 3])
    return np.stack([np.asarray(next(bg)) for _ in range(n_batches)])


if __name__ == "__main__":
    sys.path.insert(0, "/root/reference/training")
    import loader as ref  # the unmodified reference; only needed to (re)generate the golden file
    out = {}
    with tempfile.TemporaryDirectory() as d:
        a = make_shards(d, 12, seed=1)
        out["shard_seed"] = np.array([1])
    # the shards are regenerated from the seed by the tests (same make_shards), only the batches are stored
    with tempfile.TemporaryDirectory() as d:
        files = make_shards(d, 12, seed=1)
        for padding in (False, True):
            out[f"batches_pad{int(padding)}"] = run(ref, files[:7], files[7:], 64, padding, 6, seed=7).astype(np.int32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "loader_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})
