"""CPU: the parts of bench.py that do not need a GPU: the workload generator, the FLOP model and the reference arm's
launch contract (under torchrun rank 0 alone prints ONE JSON line, the other ranks exit 0 silently)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_flop_model_is_the_references_formula():
    # train_encoder.py:360: 6 N + 12 L C T with N = non-embedding parameters
    assert bench.flops_per_token(8, 1024, 1024, bench.N_NONEMB_SMALL) == 1_107_400_704
    n = 8 * (12 * 1024 * 1024 + 2 * 1024) + 1024 + 65536 * 1024  # blocks (4 matrices = 12 C^2, 2 LN) + ln_f + lm_head
    assert n == bench.N_NONEMB_SMALL


def test_synthetic_batches_look_like_the_loaders_output():
    rng = np.random.RandomState(0)
    ids = bench.synth_ids(16, 1024, rng)
    assert ids.shape == (16, 1024) and ids.dtype == np.int64
    assert ids.min() >= 3 and ids.max() < 65536 and not np.any(ids == 65533)   # no PAD when packing without padding
    assert np.all(np.isin(ids[:, 0], [4, 18]))                                  # rows start with a modality tag
    assert (ids == 3).sum() >= 16                                               # EOS-delimited documents
    padded = bench.synth_ids(8, 1024, np.random.RandomState(1), padded=True)
    first_pad = (padded == 1).argmax(axis=1)
    for r in range(8):                                                          # whole documents, then PAD to the end
        if (padded[r] == 1).any():
            assert np.all(padded[r, first_pad[r]:] == 1) and padded[r, first_pad[r] - 1] == 3


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mlm_train_tokens_per_s" and d["n_gpus"] == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
