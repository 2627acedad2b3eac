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


def test_parameter_count_and_flop_model_of_both_configs():
    assert bench.n_nonembedding(bench.SMALL) == bench.N_NONEMB_SMALL
    # SURVEY §8d / BASELINE.md §2: large (32L / 2048 / 16h): N = 1 744 963 584, 11 275 087 872 FLOP per token
    assert bench.n_nonembedding(bench.LARGE) == 1_744_963_584
    assert bench.flops_per_token(32, 2048, 1024, bench.n_nonembedding(bench.LARGE)) == 11_275_087_872
    # encode (forward, no head): 234 881 024 FLOP per token at T = 1024, 335 544 320 at T = 4096
    assert bench.encode_flops_per_token(bench.SMALL, 1024) == 234_881_024
    assert bench.encode_flops_per_token(bench.SMALL, 4096) == 335_544_320


def test_both_arms_report_the_same_config():
    a = bench.workload_config("small", 1024, 32, 0.1, 1)
    assert a["global_batch"] == 1024 and a["mini_batch_size"] == 32 and a["grad_accum_per_rank"] == 32
    assert a["seq_len"] == 1024 and "8L/1024d/8h" in a["workload"]
    assert bench.workload_config("large", 1024, 32, 0.1, 8)["grad_accum_per_rank"] == 4


def test_gemm_dram_traffic_is_read_from_the_committed_capture(tmp_path, monkeypatch):
    prof = tmp_path / "profiles"
    prof.mkdir()
    (prof / "r02_gemm_dram_bytes.csv").write_text(
        "# comment\nmetric,bytes_per_launch\ndram__bytes_read.sum,200000000\ndram__bytes_write.sum,100000000\n")
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.read_gemm_dram_traffic() == 300000000.0
    (prof / "r02_gemm_dram_bytes.csv").unlink()
    assert bench.read_gemm_dram_traffic() is None


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
    assert d["steps"] == 1 and d["warmup"] == 0                       # the arm honours --steps / --warmup
    assert d["config"] == bench.workload_config("small", 1024, 32, 0.1, 2)   # same config as the B200 arm
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
