"""muP semantics needed by the hot path, restated for the B200 implementation.

The reference depends on ``microsoft/mup==1.0.0`` (README.md:16) for three things only:
``MuReadout`` (training/model.py:19,208), ``set_base_shapes`` (training/train_encoder.py:166) and ``MuAdamW``
(training/train_encoder.py:199). The package is not vendored in the reference tree and not installable here, so
its published behaviour is restated: every parameter gets an ``infshape`` (a dimension is "infinite" iff it differs
between the base and delta models), ``MuReadout`` divides its input by ``width_mult`` (= fan-in / base fan-in) and
has its weight multiplied by ``sqrt(width_mult)`` once at ``set_base_shapes`` time, and ``MuAdamW`` gives
matrix-like parameters (two infinite dims) ``lr / width_mult`` and ``weight_decay * width_mult``.
"""
from __future__ import annotations

from collections import defaultdict

import torch
import torch.nn as nn

from . import functional as Fn


class InfDim:
    """One tensor dimension with its base size; ``base_dim is None`` marks a finite (width-independent) dim."""

    def __init__(self, base_dim, dim):
        self.base_dim = base_dim
        self.dim = dim

    def isinf(self) -> bool:
        return self.base_dim is not None

    def width_mult(self) -> float:
        return self.dim / self.base_dim if self.base_dim is not None else 1

    def __repr__(self):
        return f"InfDim({self.base_dim}, {self.dim})"


class InfShape(tuple):
    """Tuple of InfDim; ``main`` is the last infinite dimension (fan-in for inf x inf matrices)."""

    def __new__(cls, dims):
        return super().__new__(cls, dims)

    def __init__(self, dims):
        self.main = None
        for dim in reversed(self):
            if dim.isinf():
                self.main = dim
                break

    def ninf(self) -> int:
        return sum(1 for d in self if d.isinf())

    def width_mult(self) -> float:
        return self.main.width_mult() if self.main is not None else 1

    def __reduce__(self):
        return (InfShape, (list(self),))


def _zip_infshape(base_shape, shape, delta_shape=None) -> InfShape:
    dims = []
    for i, (b, s) in enumerate(zip(base_shape, shape)):
        if delta_shape is not None:
            inf = b != delta_shape[i]
        else:
            inf = b != s
        dims.append(InfDim(b if inf else None, s))
    return InfShape(dims)


class MuReadout(nn.Linear):
    """Drop-in for ``mup.MuReadout``: ``forward(x) = Linear(output_mult * x / width_mult())`` on the B200 kernels."""

    def __init__(self, *args, readout_zero_init=False, output_mult=1.0, **kwargs):
        self.output_mult = output_mult
        self.readout_zero_init = readout_zero_init
        super().__init__(*args, **kwargs)

    def reset_parameters(self) -> None:
        if self.readout_zero_init:
            self.weight.data.zero_()
            if self.bias is not None:
                self.bias.data.zero_()
        else:
            super().reset_parameters()

    def width_mult(self) -> float:
        if not hasattr(self.weight, "infshape"):
            raise AssertionError(
                "Please call set_base_shapes(...). If using torch.nn.DataParallel, switch to distributed training "
                "with torch.nn.parallel.DistributedDataParallel instead")
        return self.weight.infshape.width_mult()

    def _rescale_parameters(self) -> None:
        if getattr(self, "_has_rescaled_params", False):
            raise RuntimeError("`_rescale_parameters` has been called once before already.")
        if self.bias is not None:
            self.bias.data *= self.width_mult() ** 0.5
        self.weight.data *= self.width_mult() ** 0.5
        self._has_rescaled_params = True

    def readout_div(self) -> float:
        return self.width_mult() / self.output_mult

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.bias is not None:
            raise RuntimeError("omnibiote_b200.MuReadout: bias is not supported (the reference uses bias=False)")
        shape = x.shape
        logits = Fn.ReadoutFunction.apply(x.reshape(-1, shape[-1]), self.weight, float(self.readout_div()))
        return logits.view(*shape[:-1], self.weight.shape[0])


def set_base_shapes(model: nn.Module, base: nn.Module, rescale_params: bool = True, delta: nn.Module | None = None):
    """Attach ``infshape`` to every parameter of ``model`` and rescale MuReadout weights (train_encoder.py:166)."""
    base_shapes = {n: tuple(p.shape) for n, p in base.named_parameters()}
    delta_shapes = {n: tuple(p.shape) for n, p in delta.named_parameters()} if delta is not None else None
    for name, p in model.named_parameters():
        if name not in base_shapes:
            raise KeyError(f"set_base_shapes: parameter {name} missing from the base model")
        p.infshape = _zip_infshape(base_shapes[name], tuple(p.shape), delta_shapes[name] if delta_shapes else None)
    if rescale_params:
        for module in model.modules():
            if isinstance(module, MuReadout):
                module._rescale_parameters()
    return model


def mu_param_groups(params, lr: float, weight_decay: float = 0.0, decoupled_wd: bool = False):
    """Param-group split of mup.MuAdam: [matrix-like groups (one per width_mult) ..., vector-like group]."""
    param_groups = list(params)
    if not param_groups:
        raise ValueError("optimizer got an empty parameter list")
    if not isinstance(param_groups[0], dict):
        param_groups = [{"params": param_groups}]
    out = []
    for group in param_groups:
        group = dict(group)
        group.setdefault("lr", lr)
        group.setdefault("weight_decay", weight_decay)

        def new_group():
            g = {k: v for k, v in group.items() if k != "params"}
            g["params"] = []
            return g

        matrix_like = defaultdict(new_group)
        vector_like = new_group()
        for p in group["params"]:
            if not hasattr(p, "infshape"):
                raise AssertionError(f"A parameter with shape {tuple(p.shape)} does not have `infshape` attribute. "
                                     "Did you forget to call `set_base_shapes` on the model?")
            ninf = p.infshape.ninf()
            if ninf == 2:
                matrix_like[p.infshape.width_mult()]["params"].append(p)
            elif ninf > 2:
                raise NotImplementedError("more than 2 inf dimensions")
            else:
                vector_like["params"].append(p)
        for width_mult, g in matrix_like.items():
            g["lr"] /= width_mult
            if not decoupled_wd:
                g["weight_decay"] *= width_mult
        out.extend(list(matrix_like.values()) + [vector_like])
    return out
