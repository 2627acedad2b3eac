"""GPU: the whole optimizer step (train.MLMTrainer.step = train_encoder.py:270-318) as a black box: determinism under
a fixed seed, the optional masked-rows-only head against the dense head, attention dropout through the full step."""
import contextlib
import copy
import io
import warnings

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _model(dropout, seed=0):
    from omnibiote_b200.model import OmniBioTA, OmniBioTAConfig
    from omnibiote_b200.mup import set_base_shapes
    torch.manual_seed(seed)
    cfg = OmniBioTAConfig()
    cfg.vocab_size, cfg.block_size, cfg.n_layer, cfg.n_head, cfg.n_embd, cfg.dropout = 1024, 256, 2, 2, 256, dropout
    cfg.flash = True
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = OmniBioTA(cfg)
        c2 = copy.copy(cfg); c2.n_embd, c2.n_head = 24, 3
        c3 = copy.copy(cfg); c3.n_embd, c3.n_head = 48, 12
        set_base_shapes(m, OmniBioTA(c2), delta=OmniBioTA(c3))
        m.to(BF)
    return m.cuda().train()


def _ids(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(20, 1024, (B, T), generator=g)
    ids[:, T // 3] = 3
    ids[:, (2 * T) // 3] = 3
    return ids.cuda()


def _run(masked_rows_head, dropout, steps=2):
    from omnibiote_b200.train import MLMTrainer
    model = _model(dropout)
    tr = MLMTrainer(model, global_batch=8, mini_batch_size=4, ctx_len=256, lr=1e-3, token_budget=1e9,
                    masked_rows_head=masked_rows_head)
    torch.manual_seed(123)  # MLM mask / dropout streams
    losses = []
    for s in range(steps):
        losses.append(float(tr.step(_ids(8, 256, 50 + s))) / tr.n_accum)
    tr.check_head_overflow()
    return model, losses


def test_trainer_step_is_deterministic_and_finite():
    m1, l1 = _run(False, 0.1)
    m2, l2 = _run(False, 0.1)
    assert l1 == l2 and all(l == l and 0.0 < l < 20.0 for l in l1), (l1, l2)
    for (n, p), (_, q) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.equal(p, q), n


@pytest.mark.parametrize("dropout", [0.0, 0.1])
def test_masked_rows_head_trains_like_the_dense_head(dropout):
    """Same seeds, same batches: the head restricted to the rows inside the MLM mask must give the same losses and
    (up to the summation order of the head's weight gradient) the same parameters after two optimizer steps."""
    md, ld = _run(False, dropout)
    mm, lm = _run(True, dropout)
    for a, b in zip(ld, lm):
        assert abs(a - b) <= 2 ** -7 * abs(a), (ld, lm)
    for (n, p), (_, q) in zip(md.named_parameters(), mm.named_parameters()):
        assert rel_err(q, p) < 2e-3, (n, rel_err(q, p))
