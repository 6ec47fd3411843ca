"""Heads -- same classes, constructor arguments and state-dict keys as the reference ``model/head.py``:
``AcquisitionHead`` (9-44: ``predictor.{0,2}``), ``GMMTargetHead`` (115-266: ``heads.{c}.{0,2}``), ``OutputHead``
(270-393).  Parameters only; the arithmetic runs in ``aline_query_stream`` / ``aline_select`` / ``aline_gmm_head``.
``ValueHead`` (84-111: ``predictor.{0,2}``, ``empty_value``) runs in ``aline_value_head`` on the context tokens' final
encodings.  Continuous heads / ``single_head`` are outside the hot path and refused."""
from __future__ import annotations

from typing import Any

import torch
import torch.nn as nn

from ..rollout import gmm_log_likelihood


class AcquisitionHead(nn.Module):
    def __init__(self, dim_embedding: int, dim_feedforward: int, time_token: bool, **kwargs: Any) -> None:
        super().__init__()
        self.time_token = bool(time_token)
        d_in = dim_embedding + (1 if time_token else 0)
        self.predictor = nn.Sequential(nn.Linear(d_in, dim_feedforward), nn.ReLU(), nn.Linear(dim_feedforward, 1),
                                       nn.Flatten(start_dim=-2), nn.Softmax(dim=-1))


class ValueHead(nn.Module):
    def __init__(self, dim_embedding: int, dim_feedforward: int, **kwargs: Any) -> None:
        super().__init__()
        self.predictor = nn.Sequential(nn.Linear(dim_embedding, dim_feedforward), nn.ReLU(), nn.Linear(dim_feedforward, 1))
        self.empty_value = nn.Parameter(torch.zeros(1))      # value for zero context (never reached: n_context >= 1)


class GMMTargetHead(nn.Module):
    def __init__(self, dim_y: int, dim_embedding: int, dim_feedforward: int, num_components: int,
                 single_head: bool = False, std_min: float = 1e-4, **kwargs: Any) -> None:
        super().__init__()
        if single_head:
            raise NotImplementedError("single_head GMM heads are not part of the B200 hot path (config: False)")
        if dim_y != 1:
            raise NotImplementedError("dim_y must be 1 (as in every reference task)")
        self.dim_embedding = dim_embedding
        self.dim_feedforward = dim_feedforward
        self.dim_y = dim_y
        self.single_head = single_head
        self.num_components = num_components
        self.std_min = std_min
        self.heads = nn.ModuleList([
            nn.Sequential(nn.Linear(dim_embedding, dim_feedforward), nn.ReLU(), nn.Linear(dim_feedforward, dim_y * 3))
            for _ in range(num_components)])

    @staticmethod
    def compute_ll(value, means, stds, weights):
        """GMM log-likelihood (reference model/head.py:233-249)."""
        return gmm_log_likelihood(value, means, stds, weights)


class OutputHead(nn.Module):
    def __init__(self, dim_x: int, dim_y: int, dim_embedding: int, dim_feedforward: int, num_components: int = 10,
                 single_head: bool = False, std_min: float = 1e-4, value_head: bool = False, time_token: bool = False,
                 **kwargs: Any) -> None:
        super().__init__()
        self.dim_x = dim_x
        self.dim_y = dim_y
        self.time_token = bool(time_token)
        self.acquisition_head = AcquisitionHead(dim_embedding=dim_embedding, dim_feedforward=dim_feedforward,
                                                time_token=time_token)
        self.target_head = GMMTargetHead(dim_y=dim_y, dim_embedding=dim_embedding, dim_feedforward=dim_feedforward,
                                         num_components=num_components, single_head=single_head, std_min=std_min)
        self.value_head = value_head
        if value_head:
            self.value_head = ValueHead(dim_embedding=dim_embedding, dim_feedforward=dim_feedforward)

    def forward(self, batch, z):
        """Stand-alone call (reference 319-393; the hot path never takes it): heads on a given encoding z [B, N, d],
        composed from torch ops on the module's device (``model/grad_path.py``)."""
        from . import grad_path
        n_c, n_q = batch.context_x.shape[1], batch.query_x.shape[1]
        return grad_path.heads(self, batch, z[:, :n_c], z[:, n_c:n_c + n_q], z[:, n_c + n_q:], self.training, None, True)
