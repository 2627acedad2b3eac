"""TEST INFRASTRUCTURE ONLY — deterministic weights for the full-size golden fixture (tests/golden/small_shape.pt).

The omnibiote-small weights (235 M parameters, 470 MB) are too large to commit, so the fixture stores only inputs and
(sub-sampled) outputs of the UNMODIFIED reference; both the generator (oracle/gen_golden_small.py, run in the build
container against /root/reference) and the GPU test rebuild the SAME weights from this recipe: every tensor is drawn
from its own torch CPU generator seeded by a hash of its state_dict name, with the scale of the reference's default
initialisation (training/model.py has no custom init: Embedding N(0,1), Linear ~ 1/sqrt(fan_in), LayerNorm ~ 1) and
the muP readout rescale sqrt(width_mult) that set_base_shapes applies (train_encoder.py:166). ``freqs_cis`` buffers
are left to each implementation (computed from the config)."""
from __future__ import annotations

import hashlib
import math

import torch


def _seed(name: str, salt: int) -> int:
    return int.from_bytes(hashlib.sha256(f"{salt}:{name}".encode()).digest()[:4], "little")


def recipe_tensor(name: str, shape, n_embd: int, salt: int = 0) -> torch.Tensor:
    """bf16 tensor for the parameter `name` of shape `shape`."""
    g = torch.Generator().manual_seed(_seed(name, salt))
    x = torch.randn(tuple(shape), generator=g, dtype=torch.float32)
    if name.endswith("wte.weight"):
        pass
    elif "ln_" in name:
        x = 1.0 + 0.1 * x
    elif name == "lm_head.weight":
        x = x * (math.sqrt(n_embd / 24.0) / math.sqrt(shape[1]) * 0.5)
    else:  # nn.Linear weights [out, in]
        x = x * (0.5 / math.sqrt(shape[1]))
    return x.to(torch.bfloat16)


def load_recipe_weights(model, n_embd: int, salt: int = 0) -> None:
    """Overwrite every parameter of `model` (the reference or the drop-in; same names) with the recipe's values."""
    with torch.no_grad():
        for name, p in model.named_parameters():
            p.copy_(recipe_tensor(name, p.shape, n_embd, salt).to(p.dtype))


def sample_indices(name: str, numel: int, k: int = 4096) -> torch.Tensor:
    """Fixed pseudo-random element positions used to sub-sample the gradient of `name`."""
    g = torch.Generator().manual_seed(_seed("idx:" + name, 7))
    return torch.randint(0, numel, (min(k, numel),), generator=g)
