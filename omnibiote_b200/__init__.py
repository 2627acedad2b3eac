"""omnibiote_b200 — B200-native (sm_100a) implementation of the OmniBioTA encoder hot path.

Drop-in for the reference's ``training/model.py``: ``from omnibiote_b200.model import OmniBioTA, OmniBioTAConfig``.
"""
from .model import OmniBioTA, OmniBioTAConfig  # noqa: F401
from .mup import MuReadout, set_base_shapes  # noqa: F401

__all__ = ["OmniBioTA", "OmniBioTAConfig", "MuReadout", "set_base_shapes"]
