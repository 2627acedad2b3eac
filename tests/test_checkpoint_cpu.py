"""CPU: reading the reference's whole-module checkpoints (torch.save(model), train_encoder.py:413,430) without the
reference's ``model`` module or ``mup`` being importable (omnibiote_b200/checkpoint.py). Fixture: pickles of the
UNMODIFIED reference model written by oracle/gen_ckpt_golden.py."""
import io
import os
import sys

import pytest
import torch

from conftest import GOLDEN


@pytest.fixture(scope="module")
def fixture():
    return torch.load(os.path.join(GOLDEN, "ref_checkpoint.pt"), map_location="cpu", weights_only=False)


@pytest.mark.parametrize("tag", ["fp32", "bf16"])
def test_reference_module_pickle_loads_without_reference_code(fixture, tag):
    from omnibiote_b200 import checkpoint
    from omnibiote_b200.model import OmniBioTA
    from omnibiote_b200.mup import MuReadout
    c = fixture[tag]
    m = checkpoint.load_reference_checkpoint(io.BytesIO(c["pickle"]))
    assert isinstance(m, OmniBioTA) and isinstance(m.lm_head, MuReadout)
    # architecture inferred from the tensors, NOT from the pickled config (which says n_embd = 48, n_head = 12)
    assert c["pickled_config_n_embd"] == 48
    assert (m.config.n_embd, m.config.n_head, m.config.n_layer, m.config.vocab_size, m.config.block_size) == (64, 4, 2, 96, 24)
    assert m.transformer.h[0].attn.n_head == 4 and abs(m.config.dropout - 0.1) < 1e-12
    sd = m.state_dict()
    assert list(sd.keys()) == list(c["state_dict"].keys())
    for k, v in c["state_dict"].items():
        assert sd[k].dtype == v.dtype and torch.equal(sd[k], v), k          # bit-exact, freqs_cis form preserved
    # (module.to(dtype) of train_encoder.py:170 turns the complex table into a real one for fp32 and bf16 alike)
    assert sd["transformer.h.0.attn.freqs_cis"].is_complex() == c["state_dict"]["transformer.h.0.attn.freqs_cis"].is_complex()
    # muP: same readout width multiplier, and the already-rescaled head weight must not be rescaled again
    assert abs(m.lm_head.width_mult() - c["width_mult"]) < 1e-12
    with pytest.raises(Exception):
        m.lm_head._rescale_parameters()


def test_state_dict_round_trip(fixture, tmp_path):
    from omnibiote_b200 import checkpoint
    m = checkpoint.load_reference_checkpoint(io.BytesIO(fixture["bf16"]["pickle"]))
    p = tmp_path / "sd.pt"
    checkpoint.save_state_dict(m, p)
    m2 = checkpoint.load_reference_checkpoint(p)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k


def test_rejects_foreign_state_dict():
    from omnibiote_b200 import checkpoint
    with pytest.raises(Exception):
        checkpoint.model_from_state_dict({"transformer.wte.weight": torch.zeros(8, 4)})
