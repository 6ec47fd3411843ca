"""Helpers of the reference ``utils/misc.py`` that sit on the forward hot path.

``calculate_gmm_variance`` (reference 244-279) is the acquisition score of the uncertainty-sampling baseline
(notebooks/eval_al.ipynb cell 1); here it runs as one kernel (``aline_gmm_variance``), and
``Aline.rollout(..., acquisition="uncertainty_sampling")`` uses the variant fused with the GMM head so that
``posterior_out_query`` is never written to memory.
"""
from __future__ import annotations

from .. import rollout as _ro


def calculate_gmm_variance(mixture_means, mixture_stds, mixture_weights):
    """mixture_means / mixture_stds [B, n_query, C]; mixture_weights [B, n_query, C] or [B, C] -> variance [B, n_query]:
    ``sum_c w_c (sigma_c^2 + (mu_c - sum_c w_c mu_c)^2)``."""
    return _ro.gmm_variance(mixture_means, mixture_stds, mixture_weights)


# ---- checkpoint / run-directory compatibility (reference utils/misc.py:28-135, 174-241) ----
import importlib  # noqa: E402
import os  # noqa: E402
import re  # noqa: E402

import torch  # noqa: E402

from ..attrdict import AttrDict  # noqa: E402


def save_state_dict(model, dir, name="aline.pth"):
    """``<dir>/model/<name>`` <- model.state_dict() (reference 28-42); the keys are the reference's own."""
    file_path = os.path.join(dir, "model")
    os.makedirs(file_path, exist_ok=True)
    file_path = os.path.join(file_path, name)
    torch.save(model.state_dict(), file_path)
    return file_path


def load_state_dict(model, dir, name="aline.pth"):
    """Load ``<dir>/model/<name>`` (a reference ``.pth``) into the model (reference 45-56)."""
    file_path = os.path.join(dir, "model", name)
    dev = next(model.parameters()).device
    model.load_state_dict(torch.load(file_path, map_location=dev, weights_only=True))
    return model


def load_checkpoint_weights(model, ckpt_path):
    """Model part of a reference training checkpoint ``ckpt.tar`` (save_checkpoint, reference 59-88): the ``"model"``
    entry, or a bare state dict.  Optimizer / scheduler / RNG state belong to the training loop.
    Returns the stored epoch (None for a bare state dict)."""
    if not os.path.exists(ckpt_path):
        raise FileNotFoundError(f"Checkpoint file not found: {ckpt_path}")
    dev = next(model.parameters()).device
    ckpt = torch.load(ckpt_path, map_location=dev, weights_only=False)
    if isinstance(ckpt, dict) and "model" in ckpt:
        model.load_state_dict(ckpt["model"])
        return ckpt.get("epoch")
    model.load_state_dict(ckpt)
    return None


_TARGETS = {            # hydra ``_target_`` paths of the reference (config/{embedder,encoder,head}/*.yaml:1) -> classes here
    "model.embedder.Embedder": ("aline_b200.model.embedder", "Embedder"),
    "model.encoder.Encoder": ("aline_b200.model.encoder", "Encoder"),
    "model.head.OutputHead": ("aline_b200.model.head", "OutputHead"),
    "model.base.Aline": ("aline_b200.model.base", "Aline"),
}
_INTERP = re.compile(r"\$\{([^${}:]+)\}")


_EXP_FLOAT = re.compile(r"[-+]?\d+(\.\d*)?[eE][-+]?\d+")


def _coerce(v):
    # PyYAML (YAML 1.1) reads "1e-4" (no dot) as a string; OmegaConf's loader, which hydra uses, reads a float
    if isinstance(v, str) and _EXP_FLOAT.fullmatch(v):
        return float(v)
    return v


def _lookup(root, dotted):
    node = root
    for part in dotted.split("."):
        node = node[part]
    return node


def _resolve(node, root, depth=0):
    """``${a.b}`` interpolations of a composed hydra config (OmegaConf semantics for plain key references: a value that
    is exactly one interpolation keeps the referenced type); resolver calls such as ``${now:...}`` stay as they are."""
    if depth > 20:
        raise ValueError("interpolation cycle in the config")
    if isinstance(node, dict):
        return {k: _resolve(v, root, depth) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root, depth) for v in node]
    if isinstance(node, str):
        m = _INTERP.fullmatch(node)
        if m:
            return _resolve(_lookup(root, m.group(1).strip()), root, depth + 1)
        if _INTERP.search(node):
            return _INTERP.sub(lambda mm: str(_resolve(_lookup(root, mm.group(1).strip()), root, depth + 1)), node)
        return _coerce(node)
    return node


def _to_attr(node):
    if isinstance(node, dict):
        return AttrDict({k: _to_attr(v) for k, v in node.items()})
    if isinstance(node, list):
        return [_to_attr(v) for v in node]
    return node


def instantiate(node):
    """Minimal ``hydra.utils.instantiate`` for the three model components: ``_target_`` is mapped from the reference's
    module path to the class of the same name here, the remaining keys become constructor kwargs."""
    kwargs = {k: _coerce(v) for k, v in dict(node).items() if k != "_target_"}
    target = node["_target_"]
    if target not in _TARGETS:
        raise ValueError(f"_target_ {target!r} has no aline_b200 counterpart (known: {sorted(_TARGETS)})")
    mod, cls = _TARGETS[target]
    return getattr(importlib.import_module(mod), cls)(**kwargs)


def load_config(path, config_name="config.yaml"):
    """Resolved ``<path>/.hydra/<config_name>`` of a reference run directory as an AttrDict."""
    import yaml
    config_dir = os.path.join(path, ".hydra")
    if not os.path.isdir(config_dir):
        raise FileNotFoundError(f"Config path not found: {config_dir}")
    with open(os.path.join(config_dir, config_name)) as f:
        raw = yaml.safe_load(f)
    return _to_attr(_resolve(raw, raw))


def load_config_and_model(path, config_name="config.yaml", file_name="aline.pth", load_type="ckpt", device=None):
    """Reference ``load_config_and_model`` (174-241) without hydra / omegaconf: reads the run directory's
    ``.hydra/config.yaml``, builds Embedder / Encoder / OutputHead / Aline from its ``_target_`` nodes and loads
    ``<path>/<file_name>`` (``"ckpt"``: a ``.tar`` with a ``"model"`` entry or a bare state dict; ``"pth"``: a state
    dict).  Returns (resolved_cfg, model) with the model on ``device`` (default: the current CUDA device)."""
    full_dir = os.path.abspath(path)
    cfg = load_config(full_dir, config_name)
    model_cls = getattr(importlib.import_module(_TARGETS["model.base.Aline"][0]), "Aline")
    model = model_cls(instantiate(cfg.embedder), instantiate(cfg.encoder), instantiate(cfg.head))
    file_path = os.path.join(full_dir, file_name)
    if not os.path.exists(file_path):
        raise FileNotFoundError(f"Model file not found: {file_path}")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    try:
        if load_type.lower() == "ckpt":
            ckpt = torch.load(file_path, map_location="cpu", weights_only=False)
            model.load_state_dict(ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt else ckpt)
        elif load_type.lower() == "pth":
            model.load_state_dict(torch.load(file_path, map_location="cpu", weights_only=True))
        else:
            raise ValueError(f"Invalid load_type: {load_type}. Must be 'ckpt' or 'pth'")
    except Exception as e:  # noqa: BLE001  (same wrapping as the reference, 237-238)
        raise RuntimeError(f"Failed to load model from {file_path}: {str(e)}")
    return cfg, model.to(device)
