"""Psychometric-function task -- mirror of the reference ``tasks/psychometric.py``.

theta = (alpha ~ U(-3,3), beta ~ U(0.1,2), gamma ~ U(0.1,0.9), lambda ~ U(0,0.5)) (reference 43-88);
``p = lambda*gamma + (1-lambda)*(1 - exp(-10^((x-alpha)/beta)))`` (107-134); Bernoulli outcomes
(158-176); Bernoulli log-likelihood with the reference's 1e-10 guards (178-195) on the sm_100a kernel.

Unlike the reference (whose ``sample_theta`` only accepts an int and whose likelihood indexes
``theta[:, k, :]``, so that sPCE cannot run -- SURVEY.md section 7), ``sample_theta`` also takes a
tuple ``(L, B)`` and ``log_likelihood`` accepts ``[L, B, 4(, 1)]``, which is what the EIG estimators need.
"""
from __future__ import annotations

import torch

from .. import _lib
from ..attrdict import AttrDict
from .base_task import Task, _shape_list

_PRIOR = ((-3.0, 3.0), (0.1, 2.0), (0.1, 0.9), (0.0, 0.5))     # alpha, beta, gamma, lambda


class PsychometricTask(Task):
    def __init__(self, name: str = "Psychometric", dim_x: int = 1, dim_y: int = 1, embedding_type="theta",
                 n_target_theta: int = 4, n_context_init: int = 5, n_query_init: int = 300, design_scale: int = 5,
                 **kwargs) -> None:
        super().__init__(dim_x=dim_x, dim_y=dim_y)
        self.name = name
        self.n_target_theta = n_target_theta
        self.n_context_init = n_context_init
        self.n_query_init = n_query_init
        (self.alpha_lower, self.alpha_upper), (self.beta_lower, self.beta_upper), \
            (self.gamma_lower, self.gamma_upper), (self.lambda_lower, self.lambda_upper) = _PRIOR
        self.design_scale = design_scale

    def aline_lik(self):
        return _lib.AlineLik(_lib.TASK_PSYCHOMETRIC, 1, 1, 4, 0.0, 0.0, 0.0, 0.0)

    @torch.no_grad()
    def sample_theta(self, batch_size):
        """[B, 4, 1] for an int (as the reference, four successive uniform draws); [*shape, 4] for a tuple."""
        shape = _shape_list(batch_size)
        cols = [lo + torch.rand(shape) * (hi - lo) for lo, hi in _PRIOR]
        if isinstance(batch_size, int):
            return torch.stack(cols, dim=1).reshape(batch_size, 4, 1)
        return torch.stack(cols, dim=-1)

    @torch.no_grad()
    def sample_data(self, batch_size, n_data):
        return torch.rand(batch_size, n_data, self.dim_x) * 2 * self.design_scale - self.design_scale

    def psychometric_function(self, x, theta):
        """x [B,1] (or [B,T,1]); theta [B,4,1] / [B,4] -> response probability."""
        if theta.dim() == 2:
            theta = theta.unsqueeze(-1)
        alpha, beta, gamma, lmbda = theta[:, 0, :], theta[:, 1, :], theta[:, 2, :], theta[:, 3, :]
        z = (x - alpha) / beta
        F = 1 - torch.exp(-10 ** z)
        return lmbda * gamma + (1 - lmbda) * F

    def to_design_space(self, xi):
        return xi

    def normalise_outcomes(self, y):
        return y

    def forward(self, xi, theta):
        return torch.bernoulli(self.psychometric_function(self.to_design_space(xi), theta))

    def log_likelihood(self, y, xi, theta):
        """theta [B,4,1] / [B,4] (reference shapes) or [n_rows,B,4(,1)] with y [1,B,1], xi [1,B,1]."""
        if theta.shape[-1] == 1 and theta.dim() >= 3 and theta.shape[-2] == 4:
            theta = theta.squeeze(-1)
        return self._native_log_likelihood(y, xi, theta, 1)

    def sample_batch(self, batch_size):
        """Reference 197-222; outcomes are simulated column by column so the RNG order is the reference's."""
        theta = self.sample_theta(batch_size)
        n = self.n_context_init + self.n_query_init
        x = self.sample_data(batch_size, n)
        y = torch.empty(batch_size, n, self.dim_y)
        for i in range(n):
            y[:, i, :] = self.forward(x[:, i, :], theta)
        batch = AttrDict()
        batch.context_x = x[:, :self.n_context_init]
        batch.context_y = y[:, :self.n_context_init]
        batch.query_x = x[:, self.n_context_init:]
        batch.query_y = y[:, self.n_context_init:]
        batch.target_all = batch.target_theta = theta
        batch.n_target_theta = self.n_target_theta
        return batch

    def __str__(self) -> str:
        info = {k: v for k, v in self.__dict__.items() if not k.startswith("_")}
        return f"PsychometricTask({', '.join(f'{k}={v}' for k, v in info.items())})"
