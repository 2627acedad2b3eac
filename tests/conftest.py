import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE = os.path.join(ROOT, "oracle")
if ORACLE not in sys.path:
    sys.path.insert(0, ORACLE)  # test-only: the oracle (and its mup shim) is importable from tests
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu", weights_only=False)
    return load


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """relative Frobenius error of a against reference b"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def make_model(n_layer, n_head, n_embd, vocab_size=512, block_size=256, dropout=0.0, seed=0, checkpoint_freq=0):
    """Random-init OmniBioTA on the B200 kernels with the reference's muP setup (train_encoder.py:145-170):
    target / base (n_embd 24, 3 heads) / delta (n_embd 48, 12 heads) models, set_base_shapes, bf16."""
    import contextlib
    import copy
    import io
    import warnings
    from omnibiote_b200.model import OmniBioTA, OmniBioTAConfig
    from omnibiote_b200.mup import set_base_shapes
    cfg = OmniBioTAConfig()
    cfg.vocab_size, cfg.block_size, cfg.n_layer, cfg.n_head, cfg.n_embd = vocab_size, block_size, n_layer, n_head, n_embd
    cfg.dropout, cfg.checkpoint_freq, cfg.flash = dropout, checkpoint_freq, True
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = OmniBioTA(cfg)
        c2 = copy.copy(cfg); c2.n_embd, c2.n_head = 24, 3
        c3 = copy.copy(cfg); c3.n_embd, c3.n_head = 48, 12
        set_base_shapes(m, OmniBioTA(c2), delta=OmniBioTA(c3))
        m.to(torch.bfloat16)
    return m
