"""Constant-elasticity-of-substitution preference task -- mirror of the reference ``tasks/ces.py``.

theta = [rho, alpha_1..3, log u] with rho = 0.01 + 0.99 Beta(1,1), alpha ~ Dirichlet(1,1,1),
log u ~ N(1, 3) (reference 36-81); designs are two baskets U(0, design_scale)^3 (83-94);
``U(x) = (sum_i alpha_i x_i^rho)^(1/rho)``, response = censored sigmoid-normal with mean (U1-U2) u and
std (1+|x1-x2|) noise u (129-167); ``log_likelihood`` (169-210) runs on the sm_100a kernel.
Designs are used un-normalised (normalise/unnormalise are identities, 118-127).
"""
from __future__ import annotations

import torch
import torch.distributions as dist

from .. import _lib
from ..attrdict import AttrDict
from ..distributions import CensoredSigmoidNormal
from .base_task import Task, _shape_list


class CESTask(Task):
    def __init__(self, name: str = "CES", dim_x: int = 6, dim_y: int = 1, embedding_type="theta", n_theta: int = 5,
                 n_context_init: int = 5, n_query_init: int = 300, design_scale: int = 100,
                 noise_scale: float = 0.005, epsilon: float = 2 ** (-22), **kwargs) -> None:
        super().__init__(dim_x=dim_x, dim_y=dim_y)
        self.name = name
        self.basket_dim = 3
        self.n_theta = n_theta
        self.n_target_theta = n_theta
        self.n_context_init = n_context_init
        self.n_query_init = n_query_init
        self.design_scale = design_scale
        self.noise_scale = noise_scale
        self.epsilon = epsilon
        self.rho_a, self.rho_b = 1.0, 1.0
        self.alpha_concentration = torch.ones(self.basket_dim)
        self.u_mu, self.u_sigma = 1.0, 3.0

    def aline_lik(self):
        return _lib.AlineLik(_lib.TASK_CES, 6, 1, 5, float(self.noise_scale), float(self.epsilon), 0.0, 0.0)

    @torch.no_grad()
    def sample_theta(self, batch_size):
        """[*batch_size, 5]; the three priors are drawn in the reference's order (52-81)."""
        shape = _shape_list(batch_size)
        rho = 0.01 + 0.99 * dist.Beta(self.rho_a, self.rho_b).sample(shape)
        alpha = dist.Dirichlet(self.alpha_concentration).sample(shape)
        log_u = dist.Normal(self.u_mu, self.u_sigma).sample(shape)
        return torch.cat([rho.unsqueeze(-1), alpha, log_u.unsqueeze(-1)], dim=-1)

    @torch.no_grad()
    def sample_data(self, batch_size, n_data):
        b1 = torch.rand(batch_size, n_data, self.basket_dim) * self.design_scale
        b2 = torch.rand(batch_size, n_data, self.basket_dim) * self.design_scale
        return torch.cat([b1, b2], dim=-1)

    def utility(self, x, rho, alpha):
        return torch.sum(alpha * x ** rho, dim=-1, keepdim=True) ** (1.0 / rho)

    def normalise_design(self, x):
        return x

    def unnormalise_design(self, x):
        return x

    def normalise_outcomes(self, y):
        return y

    def response_params(self, xi, theta):
        """Mean and std of the latent normal (reference 143-165); xi [..., 6], theta [..., 5]."""
        rho, alpha, u = theta[..., 0:1], theta[..., 1:4], torch.exp(theta[..., 4:5])
        xi = torch.clamp(xi, min=0.01, max=100.0)
        b1, b2 = xi[..., :self.basket_dim], xi[..., self.basket_dim:]
        mu = (self.utility(b1, rho, alpha) - self.utility(b2, rho, alpha)) * u
        sigma = (1 + torch.norm(b1 - b2, dim=-1, p=2, keepdim=True)) * self.noise_scale * u
        return mu, sigma

    def forward(self, xi, theta):
        mu, sigma = self.response_params(xi, theta)
        return CensoredSigmoidNormal(mu, sigma, self.epsilon, 1 - self.epsilon).rsample()

    def log_likelihood(self, y, xi, theta):
        """theta [n_rows, B, 5] (or [B, 5]) with y [1,B,1], xi [1,B,6]: fused sm_100a kernel."""
        return self._native_log_likelihood(y, xi, theta, 1)

    @torch.no_grad()
    def sample_batch(self, batch_size):
        theta = self.sample_theta(batch_size).reshape(batch_size, self.n_theta, 1)
        x = self.sample_data(batch_size, self.n_context_init + self.n_query_init)
        y = self.forward(x, theta.squeeze(-1).unsqueeze(-2))
        x = self.normalise_design(x)
        batch = AttrDict()
        batch.context_x = x[:, :self.n_context_init]
        batch.context_y = y[:, :self.n_context_init]
        batch.query_x = x[:, self.n_context_init:]
        batch.query_y = y[:, self.n_context_init:]
        batch.target_all = batch.target_theta = theta
        batch.n_theta = self.n_theta
        return batch

    def __str__(self) -> str:
        info = {k: v for k, v in self.__dict__.items() if not k.startswith("_") and not callable(v)}
        return f"CESTask({', '.join(f'{k}={v}' for k, v in info.items())})"
