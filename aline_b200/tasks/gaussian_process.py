"""Gaussian-process active-learning task -- mirror of the reference ``tasks/gaussian_process.py`` (GPTask).

Hyper-prior (reference 84-105): per-dim lengthscales U(lower, upper) * sqrt(dx), isotropic with prob p_iso
(dim 0 copied), output scale U(0.1, 1).  Kernel family ~ multinomial(kernel_weights) over rbf / matern12 /
matern32 / matern52 (319-343).  ``generate_gp_data`` (366-417) -- kernel matrix + jitter, Cholesky,
``f = L z``, ``y = f + noise eps`` per batch element in a Python loop upstream -- is one batched sm_100a
kernel here (``aline_gp_sample``: one thread block per matrix, lower-packed in shared memory).
"""
from __future__ import annotations

import torch

from .. import _lib
from ..attrdict import AttrDict
from .base_task import Task


class GPTask(Task):
    def __init__(self, name: str = "AL_mix", dim_x: int = 1, dim_y: int = 1, embedding_type="mix",
                 n_context_init: int = 5, n_query_init: int = 10, n_target_theta: int = 2, n_target_data: int = 5,
                 design_scale=None, noise_scale: float = 0.01, p_iso: float = 0.5, kernel_weights=None,
                 lengthscale_lower: float = 0.1, lengthscale_upper: float = 2.0, reference_rng: bool = False,
                 **kwargs) -> None:
        super().__init__(dim_x=dim_x, dim_y=dim_y)
        self.name = name
        self.n_context_init = n_context_init
        self.n_query_init = n_query_init
        self.n_target_theta = n_target_theta
        self.n_target_data = n_target_data
        self.embedding_type = embedding_type
        self.jitter = 1e-5
        self.p_iso = p_iso
        self.kernel_weights = kernel_weights if kernel_weights is not None else [1 / 3, 0, 1 / 3, 1 / 3]
        self.kernel_types = list(_lib.GP_KERNELS)
        if self.embedding_type in ("mix", "theta"):
            if self.n_target_theta != dim_x + 1:
                raise ValueError("n_target_theta must be equal to dim_x + 1 for theta or mix embedding type")
        else:
            self.n_target_theta = 0
        root = torch.sqrt(torch.tensor(dim_x, dtype=torch.float))
        self.lengthscale_lower = lengthscale_lower * root
        self.lengthscale_upper = lengthscale_upper * root
        self.scale_lower, self.scale_upper = 0.1, 1.0
        self.noise_scale = noise_scale
        self.design_scale = torch.tensor(design_scale) if design_scale is not None else torch.tensor(5.0)
        # reproduce the reference's per-element randn call order (two calls per batch element) instead of one
        # batched draw; only matters when replaying a reference seed
        self.reference_rng = reference_rng

    @torch.no_grad()
    def sample_theta(self, batch_size):
        """[B, dx + 1, 1] = per-dim lengthscales then output scale."""
        span = self.lengthscale_upper - self.lengthscale_lower
        ls = self.lengthscale_lower + span * torch.rand(batch_size, self.dim_x)
        iso = torch.bernoulli(torch.ones(batch_size) * self.p_iso).bool()
        ls = torch.where(iso.unsqueeze(1), ls[:, :1], ls)     # (the reference's boolean-index assignment syncs the host)
        scale = self.scale_lower + (self.scale_upper - self.scale_lower) * torch.rand(batch_size)
        return torch.cat([ls, scale.unsqueeze(1)], dim=1).unsqueeze(2)

    @torch.no_grad()
    def sample_data(self, batch_size, n_data):
        return torch.rand(batch_size, n_data, self.dim_x) * 2 * self.design_scale - self.design_scale

    def to_design_space(self, xi):
        return xi * self.design_scale

    def normalise_outcomes(self, y):
        return y

    def _sample_kernel_index(self, batch_size):
        """Kernel family per batch element as an index tensor (stays on the device: no per-element host reads)."""
        w = torch.tensor(self.kernel_weights, dtype=torch.float)
        return torch.multinomial(w / w.sum(), batch_size, replacement=True)

    def sample_kernel_type(self, batch_size):
        return [self.kernel_types[i] for i in self._sample_kernel_index(batch_size).tolist()]

    def compute_kernel_matrix(self, x1, x2, lengthscales, scale, kernel_type):
        """Kernel matrix [N, M] of one batch element (no jitter) on the device kernel-matrix entry point."""
        from ..gp import kernel_matrix
        if kernel_type not in self.kernel_types:
            raise ValueError(f"Unknown kernel type: {kernel_type}")
        return kernel_matrix(x1, x2, lengthscales, scale, self.kernel_types.index(kernel_type))

    def generate_gp_data(self, x, theta, kernel_types=None, z=None, eps=None):
        """GP prior draws with observation noise, [B, N, 1].  ``kernel_types`` / ``z`` / ``eps`` may be given
        explicitly (value-exact tests); otherwise drawn here in the reference's order: kernel types for the
        whole batch, then the normal variates."""
        from ..gp import gp_sample
        B, N, _ = x.shape
        if kernel_types is None:
            kt = self._sample_kernel_index(B).to(device=x.device, dtype=torch.int32)
        elif torch.is_tensor(kernel_types):
            kt = kernel_types.to(device=x.device, dtype=torch.int32)
        else:
            kt = torch.tensor([self.kernel_types.index(k) if isinstance(k, str) else int(k) for k in kernel_types],
                              dtype=torch.int32, device=x.device)
        if z is None or eps is None:
            if self.reference_rng:
                pairs = [(torch.randn(N), torch.randn(N)) for _ in range(B)]
                z = torch.stack([p[0] for p in pairs])
                eps = torch.stack([p[1] for p in pairs])
            else:
                ze = torch.randn(B, 2, N)
                z, eps = ze[:, 0], ze[:, 1]
        ls = theta[:, :self.dim_x, 0]
        scale = theta[:, self.dim_x, 0]
        y = gp_sample(x, ls, scale, kt, z.to(x.device), eps.to(x.device), self.jitter, self.noise_scale)
        return y.unsqueeze(-1)

    def forward(self, xi, theta):
        x = self.to_design_space(xi)
        if x.dim() == 2:
            return self.generate_gp_data(x.unsqueeze(1), theta).squeeze(1)
        return self.generate_gp_data(x, theta)

    def sample_batch(self, batch_size):
        """Reference 450-530.  Note the reference passes the raw (already design-scale) x straight to
        ``generate_gp_data``."""
        batch = AttrDict()
        theta = self.sample_theta(batch_size)
        nc, nq = self.n_context_init, self.n_query_init
        n_total = nc + nq + (0 if self.embedding_type == "theta" else self.n_target_data)
        x = self.sample_data(batch_size, n_total)
        y = self.generate_gp_data(x, theta)
        batch.context_x, batch.context_y = x[:, :nc], y[:, :nc]
        batch.query_x, batch.query_y = x[:, nc:nc + nq], y[:, nc:nc + nq]
        if self.embedding_type == "theta":
            batch.target_all = batch.target_theta = theta
            batch.target_x = None
            batch.target_y = None
        else:
            batch.target_x, batch.target_y = x[:, nc + nq:], y[:, nc + nq:]
            if self.embedding_type == "data":
                batch.target_all = batch.target_y
                batch.target_theta = None
            else:
                batch.target_theta = theta
                batch.target_all = torch.cat([batch.target_y, batch.target_theta], dim=1)
        batch.n_target_theta = self.n_target_theta
        return batch

    def __str__(self) -> str:
        info = {k: v for k, v in self.__dict__.items() if not k.startswith("_")}
        return f"Active learning task with GP prior data({', '.join(f'{k}={v}' for k, v in info.items())})"
