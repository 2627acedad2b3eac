"""TEST INFRASTRUCTURE ONLY — writes tests/golden/ref_checkpoint.pt: whole-module pickles of the UNMODIFIED reference
model (torch.save(model), as train_encoder.py:413,430 does; oracle/mup stands in for the missing mup package), built
exactly like train_encoder.py:144-170 (including the config object that keeps being mutated), plus the state_dict the
loader must reproduce. Run in the build container only (needs /root/reference)."""
import io
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/training")
from model import OmniBioTA, OmniBioTAConfig  # noqa: E402  (the reference)
from mup import set_base_shapes  # noqa: E402

out = {}
for tag, dtype in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
    torch.manual_seed(3)
    config = OmniBioTAConfig()
    config.vocab_size, config.dropout, config.block_size = 96, 0.1, 24
    config.n_embd, config.n_layer, config.n_head = 64, 2, 4
    config.flash, config.checkpoint_freq = True, 0
    m = OmniBioTA(config)
    config.n_embd, config.n_head = 24, 3
    base = OmniBioTA(config)
    config.n_embd, config.n_head = 48, 12
    delta = OmniBioTA(config)
    set_base_shapes(m, base, delta=delta)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m.to(dtype)
    buf = io.BytesIO()
    torch.save(m, buf)
    out[tag] = {"pickle": buf.getvalue(), "state_dict": {k: v.clone() for k, v in m.state_dict().items()},
                "width_mult": float(m.lm_head.width_mult()), "pickled_config_n_embd": m.config.n_embd}
torch.save(out, os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_checkpoint.pt"))
print({k: (len(v["pickle"]), v["pickled_config_n_embd"], v["width_mult"]) for k, v in out.items()})
