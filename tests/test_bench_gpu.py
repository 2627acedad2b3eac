"""GPU: the auxiliary legs of bench.py on a tiny geometry (they must never cost the headline line an exception): the
drop-in module under the reference's own loop, the eager GPU restatement of the reference, the encode leg's model."""
import math
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

pytestmark = pytest.mark.gpu

TINY = dict(bench.SMALL, n_layer=2, n_embd=256, n_head=2, block_size=256)   # head_dim 128: the tensor-core attention


def test_dropin_module_loop_leg():
    """train_encoder.py:270-318 driven through the nn.Module API only: dense (b, h, t, t) bias, full logits, ATen
    cross-entropy, loss.item(), torch clip_grad_norm_ and MuAdamW.step()."""
    out = bench.dropin_module_loop(torch.device("cuda", 0), TINY, global_batch=4, mbs=2, dropout=0.1, n_micro=1)
    assert out["unit"] == "tokens/s" and out["value"] > 0
    assert math.isfinite(out["loss"]) and 0.0 < out["loss"] < 20.0           # ln(65536) / n_accum = 5.5 at init
    assert out["ms_per_step"] == pytest.approx(2 * out["ms_per_micro_batch"] + out["ms_optimizer"], rel=1e-6)


def test_gpu_eager_reference_leg():
    out = bench.gpu_eager_reference(torch.device("cuda", 0), TINY, global_batch=4, mbs=2, n_micro=1)
    assert out.get("value", 0) > 0 and out["mini_batch_size"] == 2 and out["kind"] == "port"
