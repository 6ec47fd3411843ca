"""``Embedder`` -- same constructor, parameters and state-dict keys as the reference ``model/embedder.py``
(x_embedder / y_embedder = Linear -> ReLU -> Linear, learnable ``theta_tokens``; reference 17-65).

On the hot path (``model.base.Aline.forward`` / ``rollout``) the module only owns the parameters: the arithmetic runs
inside the sm_100a kernels (embedding phase of ``aline_embed_queries`` / ``aline_ctx_stack``).  Calling the sub-module
directly (reference 67-95; no shipped caller does) composes the same embedding from torch ops on the module's device
(``model/grad_path.py``), differentiable."""
from __future__ import annotations

from typing import Any

import torch
import torch.nn as nn


class Embedder(nn.Module):
    def __init__(self, dim_x: int, dim_y: int, dim_embedding: int, dim_feedforward: int, n_target_theta: int = 0,
                 embedding_type: str = "data", **kwargs: Any) -> None:
        super().__init__()
        self.dim_x = dim_x
        self.dim_y = dim_y
        self.dim_embedding = dim_embedding
        self.dim_feedforward = dim_feedforward
        self.n_target_theta = n_target_theta
        self.embedding_type = embedding_type
        if embedding_type not in ("data", "theta", "mix"):
            raise ValueError(f"Unknown embedding type: {embedding_type}")
        self.x_embedder = nn.Sequential(nn.Linear(dim_x, dim_feedforward), nn.ReLU(),
                                        nn.Linear(dim_feedforward, dim_embedding))
        self.y_embedder = nn.Sequential(nn.Linear(dim_y, dim_feedforward), nn.ReLU(),
                                        nn.Linear(dim_feedforward, dim_embedding))
        if embedding_type in ("theta", "mix"):
            if self.n_target_theta <= 0:
                raise ValueError("dim_theta must be positive for theta or mix embedding type")
            self.theta_tokens = nn.Parameter(torch.randn(self.n_target_theta, dim_embedding))

    def forward(self, batch):
        """[B, n_context + n_query + n_target, dim_embedding], token order [context | query | target] (reference 67-95)."""
        from . import grad_path
        return torch.cat(grad_path.embed(self, batch), dim=1)
