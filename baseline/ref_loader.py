"""Import the unmodified reference from ``baseline/_ref`` (installed by ``baseline/install_ref.py``).

Only bench.py's reference legs and tests that pin against the reference use this.  The reference needs two things
this image lacks: the ``attrdictionary`` package (an attribute dict; 8-line stand-in below) and, for
``utils/__init__.py``, hydra / omegaconf / termcolor -- so ``utils/eval.py`` and ``utils/target_mask.py`` are loaded by
file path, which leaves their source untouched.
"""
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


class AttrDict(dict):
    """Stand-in for ``attrdictionary.AttrDict``."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]


def available():
    return os.path.exists(os.path.join(REF, "model", "base.py"))


_cache = None


def load():
    """Returns a namespace with the reference's classes / functions: Aline, Embedder, Encoder, OutputHead,
    HiddenLocation, CESTask, PsychometricTask, GPTask, EIGStepLoss, PCELoss, NMCLoss, eval (module), target_mask (module)."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise FileNotFoundError(f"{REF} is empty: run `python baseline/install_ref.py` where /root/reference exists")
    if "attrdictionary" not in sys.modules:
        m = types.ModuleType("attrdictionary")
        m.AttrDict = AttrDict
        sys.modules["attrdictionary"] = m
    # the reference's top-level package names (model, loss, tasks, utils, distributions) must resolve to baseline/_ref
    for name in ("model", "loss", "tasks", "distributions", "utils"):
        for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
            mod = sys.modules[k]
            f = getattr(mod, "__file__", "") or ""
            if not f.startswith(REF):
                del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        ns = types.SimpleNamespace()
        from model.base import Aline
        from model.embedder import Embedder
        from model.encoder import Encoder
        from model.head import OutputHead
        from tasks.location_finding import HiddenLocation
        from tasks.ces import CESTask
        from tasks.psychometric import PsychometricTask
        from tasks.gaussian_process import GPTask
        from loss.eig import EIGStepLoss, PCELoss, NMCLoss
        ns.Aline, ns.Embedder, ns.Encoder, ns.OutputHead = Aline, Embedder, Encoder, OutputHead
        ns.HiddenLocation, ns.CESTask, ns.PsychometricTask, ns.GPTask = HiddenLocation, CESTask, PsychometricTask, GPTask
        ns.EIGStepLoss, ns.PCELoss, ns.NMCLoss = EIGStepLoss, PCELoss, NMCLoss

        def _by_path(name, rel):
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod

        ns.eval = _by_path("_aline_ref_eval", "utils/eval.py")
        ns.target_mask = _by_path("_aline_ref_target_mask", "utils/target_mask.py")
        ns.AttrDict = AttrDict
    finally:
        sys.path.remove(REF)
    _cache = ns
    return ns
