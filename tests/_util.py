"""Shared helpers for the test-suite (fixture loading; oracle import)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def state_dict_of(g, device="cpu"):
    return {k[3:]: torch.from_numpy(v).to(device) for k, v in g.items() if k.startswith("sd/")}


def step_batch(g, t, device="cpu"):
    """The batch dict the reference's model.forward saw at teacher-forced step t."""
    b = {k: torch.from_numpy(g[f"step{t}/{k}"]).to(device) for k in ("context_x", "context_y", "query_x", "query_y")}
    b["target_all"] = torch.from_numpy(g["target_all"]).to(device)
    if "target_x" in g:
        b["target_x"] = torch.from_numpy(g["target_x"]).to(device)
    if "target_mask" in g:
        b["target_mask"] = torch.from_numpy(g["target_mask"]).to(device)
    if f"step{t}/t" in g:
        b["t"] = torch.from_numpy(g[f"step{t}/t"])          # time token (utils/eval.py:25-26), a host scalar
    return b


def mode_of(g):
    if "target_x" in g:
        return "mix" if "sd/embedder.theta_tokens" in g else "data"
    return "theta"


def n_head_of(sd):
    return sd["encoder.encoder.layers.0.linear2.weight"].shape[0] // 8


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).abs() / b.abs().clamp_min(1e-12)).max().item()


def abs_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return (a - b).abs().max().item()
