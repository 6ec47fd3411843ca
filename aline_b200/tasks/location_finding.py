"""Location-finding (hidden sources) task -- mirror of the reference ``tasks/location_finding.py``.

Prior U(0,1)^{K x D} (or Normal), designs U(0,1)^D, signal ``log(base + sum_k 1/(max + |xi-theta_k|^2))``
(reference 110-130), outcome N(signal, noise_scale) (132-147), ``log_likelihood`` on the sm_100a
kernel (149-164 -> aline_log_likelihood).  The draws consume the global torch RNG in the same order
as the reference, so a seed reproduces the reference's batches.
"""
from __future__ import annotations

import torch

from .. import _lib
from ..attrdict import AttrDict
from .base_task import Task, _shape_list


class HiddenLocation(Task):
    def __init__(self, name: str = "Location", dim_x: int = 2, dim_y: int = 1, embedding_type="theta",
                 n_target_theta: int = 2, n_context_init: int = 1, n_query_init: int = 200, K: int = 1,
                 theta_loc=None, theta_cov=None, theta_dist="uniform", design_scale=None, outcome_scale=10,
                 noise_scale=0.5, base_signal: float = 0.1, max_signal: float = 1e-4, **kwargs) -> None:
        super().__init__(dim_x=dim_x, dim_y=dim_y)
        self.name = name
        self.theta_dist = theta_dist
        if theta_dist == "normal":
            self.theta_loc = theta_loc if theta_loc is not None else torch.zeros((K, dim_x))
            self.theta_cov = theta_cov if theta_cov is not None else torch.eye(dim_x)
            self._data_low, self._data_high = -4.0, 4.0
        elif theta_dist == "uniform":
            self.theta_loc = theta_loc if theta_loc is not None else torch.zeros((K, dim_x))   # low
            self.theta_cov = theta_cov if theta_cov is not None else torch.ones((K, dim_x))    # high
            self._data_low, self._data_high = 0.0, 1.0
        else:
            raise ValueError(f"Prior distribution type {theta_dist} is not supported!")
        # U(0,1) prior (the shipped config): checked once here, not per draw (a device tensor would force a sync)
        self._unit_prior = theta_dist == "uniform" and bool((torch.as_tensor(self.theta_loc) == 0).all()) \
            and bool((torch.as_tensor(self.theta_cov) == 1).all())
        self.design_scale = design_scale if design_scale is not None else torch.max(self.theta_cov)
        self.outcome_scale = outcome_scale
        self.register_buffer("noise_scale", noise_scale * torch.tensor(1.0, dtype=torch.float32))
        self.base_signal = base_signal
        self.max_signal = max_signal
        self.n_target_theta = n_target_theta
        self.n_context_init = n_context_init
        self.n_query_init = n_query_init
        self.K = K
        assert self.n_target_theta == self.K * self.dim_x, "n_theta must be equal to K * dim_x"

    def aline_lik(self):
        return _lib.AlineLik(_lib.TASK_LOCATION, int(self.dim_x), int(self.K), int(self.K * self.dim_x),
                             float(self.noise_scale), float(self.base_signal), float(self.max_signal), 0.0)

    @torch.no_grad()
    def sample_theta(self, batch_size):
        """Prior draw [*batch_size, K, D] (reference 85-98)."""
        shape = _shape_list(batch_size) + [self.K, self.dim_x]
        if self.theta_dist == "uniform":
            low, high = torch.as_tensor(self.theta_loc), torch.as_tensor(self.theta_cov)
            u = torch.rand(shape)
            if self._unit_prior:
                return u       # 0 + u * (1 - 0) == u bit for bit: skip two more passes over a [L+1, B, K, D] tensor
            return u.mul_((high - low).to(u.device)).add_(low.to(u.device))
        if self.dim_x == 1:
            return torch.as_tensor(self.theta_loc) + torch.as_tensor(self.theta_cov) * torch.randn(shape)
        chol = torch.linalg.cholesky(torch.as_tensor(self.theta_cov, dtype=torch.float32))
        eps = torch.randn(shape)
        return torch.as_tensor(self.theta_loc).to(eps.device) + eps @ chol.T.to(eps.device)

    @torch.no_grad()
    def sample_data(self, batch_size, n_data):
        """Designs U(low, high)^D, [B, N, D].  The reference draws [B, N, K, D] and keeps source 0
        (100-106); the same amount of randomness is consumed here."""
        u = torch.rand([batch_size, n_data, self.K, self.dim_x])
        return (self._data_low + u * (self._data_high - self._data_low))[..., 0, :]

    def total_density(self, xi, theta):
        """log(base + sum_k 1 / (max + |xi - theta_k|^2)); xi [..., D], theta [..., K, D] -> [..., 1]."""
        sq = (xi.unsqueeze(-2) - theta).pow(2).sum(-1)
        return torch.log(self.base_signal + (self.max_signal + sq).pow(-1).sum(-1, keepdim=True))

    def forward(self, xi, theta):
        """Simulated outcome N(signal, noise_scale); xi is the real (unnormalised) design."""
        signal = self.total_density(xi, theta)
        return signal + self.noise_scale.to(signal.device) * torch.randn(signal.shape)

    def log_likelihood(self, y, xi, theta):
        """log N(y; signal(xi, theta), noise_scale) on the sm_100a kernel; theta [n_rows, B, K, D] or [B, K, D]."""
        return self._native_log_likelihood(y, xi, theta, 2)

    @torch.no_grad()
    def sample_batch(self, batch_size, with_query=True):
        """One context point + n_query_init candidates with pre-simulated outcomes (reference 167-192)."""
        theta = self.sample_theta(batch_size)
        if not with_query:
            self.n_query_init = 1
        n = self.n_context_init + self.n_query_init
        x = self.sample_data(batch_size, n)
        y = self.forward(self.unnormalise_design(x), theta.unsqueeze(1).expand(batch_size, n, self.K, self.dim_x))
        theta = theta.reshape(batch_size, self.n_target_theta, 1)
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
        return f"HiddenLocation({', '.join(f'{k}={v}' for k, v in info.items())})"
