"""EIG bound estimators on the sm_100a sPCE kernels -- same classes and call signatures as the
reference's ``loss/eig.py`` (EIGBounds 8-51, PCELoss 55-86, NMCLoss 120-151, EIGStepLoss 154-209).

``log_prob`` is still the task's bound ``log_likelihood`` method, as in the reference; it is not
called -- the native path dispatches on the task behind it (``aline_b200.spce.lik_of``).  A callable
without a kernel raises; there is no PyTorch fallback.  Forward-only (the reference evaluates these
under ``torch.no_grad``, utils/eval.py:42).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import spce as _spce


class EIGBounds(nn.Module):
    """Whole-history bounds, evaluated at the end of the trajectories (loss/eig.py:8-51)."""

    def __init__(self, L: int, T: int, log_prob, reduction=None) -> None:
        super().__init__()
        self.L = L
        self.T = T
        self.log_prob = log_prob
        self.reduction = reduction
        self._lik = _spce.lik_of(log_prob)

    def _partials(self, y_outcomes, xi_designs, thetas):
        """One fused pass over thetas [L, B, (K,) D] for the whole history; row 0 is theta_0."""
        B, T = xi_designs.shape[:2]
        n_rows = thetas.shape[0]
        m, s, lp0 = _spce.spce_history(self._lik, y_outcomes, xi_designs, thetas, seq=None, skip_rows=1)
        return m[:, -1], s[:, -1], lp0[:, -1], None

    @torch.no_grad()
    def compute_seq_logprobs(self, y_outcomes, xi_designs, thetas):
        """Sequential joint log-likelihood [L, B] (loss/eig.py:22-48)."""
        B = xi_designs.shape[0]
        seq = torch.zeros((thetas.shape[0], B), dtype=torch.float32, device=thetas.device)
        _spce.spce_history(self._lik, y_outcomes, xi_designs, thetas, seq=seq, skip_rows=0)
        return seq

    def forward(self, y_outcomes, xi_designs, thetas):
        return self.compute_seq_logprobs(y_outcomes, xi_designs, thetas)

    def _reduce(self, loss):
        return torch.mean(loss) if self.reduction == "mean" else loss


class PCELoss(EIGBounds):
    """sPCE loss ``logsumexp_{l=0..L} - lp[0]`` (loss/eig.py:55-86)."""

    def __init__(self, L: int, T: int, log_prob, reduction="mean") -> None:
        super().__init__(L, T, log_prob, reduction)

    @torch.no_grad()
    def forward(self, y_outcomes, xi_designs, thetas):
        m, s, lp0, _ = self._partials(y_outcomes, xi_designs, thetas)
        pce, _ = _spce.lse_combine(m, s, lp0)
        return self._reduce(pce)


class NMCLoss(EIGBounds):
    """sNMC loss ``logsumexp_{l=1..L} - lp[0]`` (loss/eig.py:120-151)."""

    def __init__(self, L: int, T: int, log_prob, reduction="mean") -> None:
        super().__init__(L, T, log_prob, reduction)

    @torch.no_grad()
    def forward(self, y_outcomes, xi_designs, thetas):
        m, s, lp0, _ = self._partials(y_outcomes, xi_designs, thetas)
        _, nmc = _spce.lse_combine(m, s, lp0)
        return self._reduce(nmc)


class EIGStepLoss(nn.Module):
    """Step-wise sPCE + sNMC with the running ``seq_logprobs [L+1, M]`` state (loss/eig.py:154-209)."""

    def __init__(self, L: int, M: int, log_prob, reduction=None) -> None:
        super().__init__()
        self.L = L
        self.M = M
        self.log_prob = log_prob
        self.reduction = reduction
        self._lik = _spce.lik_of(log_prob)
        self.seq_logprobs = None
        self.reset()

    def reset(self):
        """Reset the sequential log-likelihood (allocated on the default device, like the reference's
        ``torch.zeros((L + 1, M))``, loss/eig.py:168-172)."""
        self.seq_logprobs = torch.zeros((self.L + 1, self.M), dtype=torch.float32)

    def _ensure_device(self, thetas):
        if self.seq_logprobs.device != thetas.device:
            self.seq_logprobs = self.seq_logprobs.to(thetas.device)

    @torch.no_grad()
    def step(self, y_outcomes, xi_designs, thetas):
        self._ensure_device(thetas)
        self._last = _spce.spce_step(self._lik, y_outcomes, xi_designs, thetas, self.seq_logprobs, skip_rows=1)
        return self.seq_logprobs

    @torch.no_grad()
    def forward(self, y_outcomes, xi_designs, thetas):
        self.step(y_outcomes, xi_designs, thetas)
        m, s, lp0 = self._last
        pce_loss, nmc_loss = _spce.lse_combine(m, s, lp0)
        if self.reduction == "mean":
            pce_loss, nmc_loss = torch.mean(pce_loss), torch.mean(nmc_loss)
        return pce_loss, nmc_loss
