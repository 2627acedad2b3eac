"""Weight interchange with the reference.

The reference checkpoints the WHOLE module object (``torch.save(model.module, ...)``, train_encoder.py:413,430) and the
evals load it back with ``torch.load(path)`` (evals/gue.py:279), which needs the reference's ``model`` module and the
third-party ``mup`` package to be importable. ``load_reference_checkpoint`` reads such a pickle WITHOUT either: class
look-ups for ``model.*`` / ``mup.*`` are redirected while unpickling, the parameters and buffers are taken from the
restored module tree and poured into a freshly constructed ``omnibiote_b200.model.OmniBioTA``:

  * the pickled ``config`` cannot be trusted: train_encoder.py:145-162 keeps mutating the one config object after the
    model was built (a saved checkpoint says n_embd=48, n_head=12), so the architecture is inferred from the tensors;
  * ``freqs_cis`` keeps the form it was saved in (complex64 = rotation, real = cosine scaling);
  * µP shapes are re-derived with the reference's base / delta widths (24 / 48, train_encoder.py:158-164) and the
    readout weight is NOT rescaled again (the checkpoint already holds the rescaled weight).
A plain ``state_dict`` file (what ``save_state_dict`` writes) is accepted as well.
"""
from __future__ import annotations

import copy
import io
import pickle
import warnings

import torch
from torch import nn

from . import model as _model
from . import mup as _mup


class _Opaque:
    """Stands in for pickled classes whose content is not needed (mup's InfShape / InfDim, optimizer hooks ...)."""

    def __new__(cls, *args, **kwargs):
        return object.__new__(cls)

    def __init__(self, *args, **kwargs):
        pass

    def __setstate__(self, state):
        self._state = state

    def append(self, _x):
        pass

    def extend(self, _x):
        pass

    def __setitem__(self, _k, _v):
        pass


class _RemapUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        root = module.split(".")[0]
        if module == "model" or module.endswith(".model") and root != "torch":
            if hasattr(_model, name):
                return getattr(_model, name)
            return _Opaque
        if root == "mup":
            return _mup.MuReadout if name == "MuReadout" else _Opaque
        return super().find_class(module, name)


class _RemapPickle:
    """``pickle_module`` for torch.load."""
    __name__ = "omnibiote_b200_remap_pickle"
    Unpickler = _RemapUnpickler
    load = staticmethod(lambda f, **kw: _RemapUnpickler(f, **kw).load())
    loads = staticmethod(lambda b, **kw: _RemapUnpickler(io.BytesIO(b), **kw).load())
    dump, dumps, Pickler = pickle.dump, pickle.dumps, pickle.Pickler
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL


def _infer_config(sd: dict, dropout: float = 0.1) -> "_model.OmniBioTAConfig":
    cfg = _model.OmniBioTAConfig()
    wte = sd["transformer.wte.weight"]
    cfg.vocab_size, cfg.n_embd = int(wte.shape[0]), int(wte.shape[1])
    cfg.n_layer = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.h."))
    fc = sd["transformer.h.0.attn.freqs_cis"]
    cfg.block_size = int(fc.shape[0])
    cfg.n_head = cfg.n_embd // (2 * int(fc.shape[1]))   # freqs_cis is [block_size, head_dim / 2]
    cfg.dropout = dropout
    cfg.bias = any(k.endswith(".bias") for k in sd)
    cfg.flash = True
    return cfg


def model_from_state_dict(sd: dict, dropout: float = 0.1) -> "_model.OmniBioTA":
    """Fresh OmniBioTA with the architecture, dtype, parameters and buffers of a (reference or own) state_dict."""
    import contextlib
    cfg = _infer_config(sd, dropout)
    with contextlib.redirect_stdout(io.StringIO()):
        m = _model.OmniBioTA(cfg)
        c2 = copy.copy(cfg); c2.n_embd, c2.n_head = 24, 3      # train_encoder.py:158-160
        c3 = copy.copy(cfg); c3.n_embd, c3.n_head = 48, 12     # :162-164
        _mup.set_base_shapes(m, _model.OmniBioTA(c2), delta=_model.OmniBioTA(c3), rescale_params=False)
    for mod in m.modules():
        if isinstance(mod, _mup.MuReadout):
            mod._has_rescaled_params = True  # the stored weight is already multiplied by sqrt(width_mult)
    dtype = sd["transformer.wte.weight"].dtype
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # complex -> real cast warning of freqs_cis; the buffers are replaced below
        m.to(dtype)
    own = dict(m.named_parameters())
    missing = [k for k in own if k not in sd]
    extra = [k for k in sd if k not in own and not k.endswith("freqs_cis")]
    if missing or extra:
        raise RuntimeError(f"omnibiote_b200: state_dict does not match the OmniBioTA layout: missing {missing[:4]}, "
                           f"unexpected {extra[:4]}")
    with torch.no_grad():
        for k, p in own.items():
            if tuple(p.shape) != tuple(sd[k].shape):
                raise RuntimeError(f"omnibiote_b200: shape of {k}: {tuple(sd[k].shape)} vs {tuple(p.shape)}")
            p.copy_(sd[k])
    for i, blk in enumerate(m.transformer.h):      # keep the saved form of the rotary table (complex or real)
        blk.attn.register_buffer("freqs_cis", sd[f"transformer.h.{i}.attn.freqs_cis"].clone())
    return m


def load_reference_checkpoint(path_or_file, map_location="cpu", dropout: float | None = None) -> "_model.OmniBioTA":
    """Reads a reference whole-module pickle (or a state_dict file) and returns an ``omnibiote_b200`` OmniBioTA holding
    the same weights. ``dropout``: the pickled config's value when None and available, else 0.1 (reference default)."""
    obj = torch.load(path_or_file, map_location=map_location, weights_only=False, pickle_module=_RemapPickle)
    if isinstance(obj, nn.Module):
        sd = obj.state_dict()
        if dropout is None:
            dropout = float(getattr(getattr(obj, "config", None), "dropout", 0.1))
    elif isinstance(obj, dict):
        sd = obj.get("state_dict", obj)
    else:
        raise RuntimeError(f"omnibiote_b200: unsupported checkpoint object {type(obj)!r}")
    return model_from_state_dict(sd, 0.1 if dropout is None else dropout)


def save_state_dict(model: nn.Module, path) -> None:
    """Portable checkpoint: the reference's state_dict keys (incl. freqs_cis buffers); loadable by either code base."""
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, path)
