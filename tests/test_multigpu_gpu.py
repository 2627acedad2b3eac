"""GPU, >= 2 devices: the data-parallel path over NCCL (train_encoder.py:185,284-311,334-336,350-356), launched under
torchrun exactly as the bench is. scripts/ddp_check.py asserts, on every rank:
  * flat-bucket gradients (in-place wgrad accumulation + bucketed all-reduce overlapped with the last backward)
    == single-process loop over all ranks' micro-batches (mean), relative error <= 1.5e-2 (bf16 summation order);
  * the drop-in module wrapped in stock torch.nn.parallel.DistributedDataParallel gives the same mean gradient;
  * MLMTrainer's on-device step bookkeeping: [loss_sum, n_masked, n_tokens] all-reduced once per step equals the
    gathered per-rank values, and every rank holds identical parameters after the step.
Skipped on a single-GPU box (the CPU/gloo tests in test_parallel_cpu.py cover the host logic there)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_data_parallel_paths_across_two_ranks():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(ROOT, "scripts", "ddp_check.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "DDP_OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
