"""aline_b200 -- B200-native (sm_100a) rollout + sPCE hot path of ALINE behind the reference's Python API.

Sub-packages mirror the reference's module paths (``model``, ``loss``, ``tasks``, ``utils``); see
INTEGRATION.md for how ``train_aline.py`` / the hydra ``_target_`` strings bind to them.
"""
from ._lib import AlineError, LIB_PATH, kernel_launches, lib  # noqa: F401

__all__ = ["AlineError", "LIB_PATH", "kernel_launches", "lib"]
