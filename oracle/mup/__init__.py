"""TEST INFRASTRUCTURE ONLY — minimal stand-in for ``microsoft/mup==1.0.0`` (README.md:16 of the reference).

The reference imports ``from mup import MuReadout`` (training/model.py:19) and ``set_base_shapes, MuAdamW``
(training/train_encoder.py:7). mup is a third-party dependency that is neither vendored under /root/reference nor
installable here (no network), so its published algorithm is restated: just enough for the *unmodified* reference
``training/model.py`` to import and run on CPU when ``oracle/`` is put on ``sys.path``. Only tests/,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this.

Restated behaviour (mup/layer.py, mup/shape.py, mup/infshape.py, mup/optim.py of mup 1.0.0):
  * ``MuReadout(nn.Linear)``: ``forward(x) = Linear(output_mult * x / width_mult())``,
    ``width_mult() = weight.infshape.width_mult()``; ``_rescale_parameters`` multiplies weight by sqrt(width_mult).
  * ``set_base_shapes(model, base, delta=...)``: a dim is infinite iff base != delta; attaches ``infshape`` to each
    parameter; rescales every MuReadout once.
  * ``MuAdamW``: params with two infinite dims get lr / width_mult and weight_decay * width_mult; the rest are
    unchanged; groups ordered [matrix-like..., vector-like]; then ``torch.optim.AdamW``.
"""
from collections import defaultdict

import torch
from torch import nn


class _Dim:
    def __init__(self, base, size):
        self.base, self.size = base, size

    def isinf(self):
        return self.base is not None

    def width_mult(self):
        return self.size / self.base if self.base is not None else 1


class _Shape(tuple):
    def ninf(self):
        return sum(d.isinf() for d in self)

    def width_mult(self):
        for d in reversed(self):  # "main" = last infinite dim (fan-in for inf x inf)
            if d.isinf():
                return d.width_mult()
        return 1


class MuReadout(nn.Linear):
    def __init__(self, *args, readout_zero_init=False, output_mult=1.0, **kwargs):
        self.output_mult = output_mult
        self.readout_zero_init = readout_zero_init
        super().__init__(*args, **kwargs)

    def reset_parameters(self):
        if self.readout_zero_init:
            self.weight.data[:] = 0
            if self.bias is not None:
                self.bias.data[:] = 0
        else:
            super().reset_parameters()

    def width_mult(self):
        assert hasattr(self.weight, "infshape"), "call set_base_shapes first"
        return self.weight.infshape.width_mult()

    def _rescale_parameters(self):
        if getattr(self, "_has_rescaled_params", False):
            raise RuntimeError("already rescaled")
        if self.bias is not None:
            self.bias.data *= self.width_mult() ** 0.5
        self.weight.data *= self.width_mult() ** 0.5
        self._has_rescaled_params = True

    def forward(self, x):
        return super().forward(self.output_mult * x / self.width_mult())


def set_base_shapes(model, base, rescale_params=True, delta=None, **_unused):
    bs = {n: p.shape for n, p in base.named_parameters()}
    ds = {n: p.shape for n, p in delta.named_parameters()} if delta is not None else None
    for n, p in model.named_parameters():
        dims = []
        for i, (b, s) in enumerate(zip(bs[n], p.shape)):
            inf = (b != ds[n][i]) if ds is not None else (b != s)
            dims.append(_Dim(b if inf else None, s))
        p.infshape = _Shape(dims)
    if rescale_params:
        for m in model.modules():
            if isinstance(m, MuReadout):
                m._rescale_parameters()
    return model


def _mu_groups(params, decoupled_wd=False, **kwargs):
    groups = list(params)
    if not isinstance(groups[0], dict):
        groups = [{"params": groups}]
    out = []
    for g in groups:
        g.setdefault("lr", kwargs["lr"])
        g.setdefault("weight_decay", kwargs.get("weight_decay", 0.0))

        def new():
            n = {k: v for k, v in g.items() if k != "params"}
            n["params"] = []
            return n

        mat, vec = defaultdict(new), new()
        for p in g["params"]:
            assert hasattr(p, "infshape")
            if p.infshape.ninf() == 2:
                mat[p.infshape.width_mult()]["params"].append(p)
            elif p.infshape.ninf() > 2:
                raise NotImplementedError
            else:
                vec["params"].append(p)
        for wm, ng in mat.items():
            ng["lr"] /= wm
            if not decoupled_wd:
                ng["weight_decay"] *= wm
        out.extend(list(mat.values()) + [vec])
    return out


def MuAdamW(params, **kwargs):
    return torch.optim.AdamW(_mu_groups(params, **kwargs), **kwargs)
