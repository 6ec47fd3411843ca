"""Task base: design (un)normalisation and the query->context move of a design step.

Mirror of the reference ``tasks/base_task.py`` (``Task`` 10-154): same method names and return
conventions.  ``update_batch`` (reference 133-154: gather the chosen (x, y), compact the query set
order-preservingly, append to the context) runs as one CUDA kernel for CUDA batches; a resident
T-step rollout never calls it (``aline_b200.rollout`` appends in place on the device).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib


def _shape_list(size):
    if isinstance(size, int):
        return [size]
    return list(size)


class Task(nn.Module):
    def __init__(self, dim_x: int = 2, dim_y: int = 1, dim_theta: int = 0, mode: str = "data",
                 design_scale: float = 1.0, outcome_scale: float = 1.0, device=None, **kwargs) -> None:
        super().__init__()
        self.dim_x = dim_x
        self.dim_y = dim_y
        self.dim_theta = dim_theta
        self.mode = mode
        self.design_scale = design_scale
        self.outcome_scale = outcome_scale
        self.device = device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu")
        if mode in ("theta", "mix") and dim_theta <= 0:
            raise ValueError(f"dim_theta must be positive for mode '{mode}'")

    # ---- abstract simulator interface (tasks/base_task.py:44-101) ----
    @torch.no_grad()
    def sample_theta(self, size, **kwargs):
        raise NotImplementedError("Child classes must implement sample_theta")

    @torch.no_grad()
    def forward(self, xi, theta):
        raise NotImplementedError("Child classes must implement forward")

    def log_likelihood(self, y, xi, theta):
        raise NotImplementedError("Child classes must implement log_likelihood")

    # ---- design / outcome scaling (tasks/base_task.py:58-73) ----
    def to_design_space(self, xi):
        return xi * self.design_scale

    def normalise_design(self, x):
        return x / self.design_scale

    def unnormalise_design(self, x):
        return x * self.design_scale

    def normalise_outcomes(self, y):
        return y / self.outcome_scale

    # ---- design step bookkeeping (tasks/base_task.py:103-154) ----
    def update_batch_query(self, query, idx):
        """Remove row ``idx[b]`` of ``query [B, N, D]`` for every b, keeping the order of the rest."""
        from ..rollout import remove_rows
        return remove_rows(query, idx)

    def update_batch_context(self, context, new):
        from ..rollout import append_rows
        return append_rows(context, new)

    def update_batch(self, batch, idx):
        """Move the selected candidate from the query set to the end of the context (in the batch dict)."""
        from ..rollout import move_selected
        (batch.query_x, batch.context_x), (batch.query_y, batch.context_y) = move_selected(
            [(batch.query_x, batch.context_x), (batch.query_y, batch.context_y)], idx)
        return batch

    # ---- native likelihood hook ----
    def aline_lik(self) -> "_lib.AlineLik":
        raise _lib.AlineError(f"{type(self).__name__} has no sm_100a likelihood kernel")

    def _native_log_likelihood(self, y, xi, theta, n_param_dims):
        """Route ``log_likelihood(y, xi, theta)`` calls of the shapes the EIG estimators use
        (y [1,B,1] / [B,1], xi [1,B,dx] / [B,dx], theta [n_rows,B,...] or [B,...]) to the kernel."""
        from .. import spce
        squeeze_rows = theta.dim() == 1 + n_param_dims          # [B, ...] -> one row
        th = theta.unsqueeze(0) if squeeze_rows else theta
        if th.dim() != 2 + n_param_dims:
            raise _lib.AlineError(f"log_likelihood: unsupported theta shape {tuple(theta.shape)}")
        out = spce.log_likelihood(self.aline_lik(), y, xi, th)
        return out[0] if squeeze_rows else out
