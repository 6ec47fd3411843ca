"""Censored logit-normal observation model of the CES task.

Mirror of the reference ``distributions/censored_sigmoid_normal.py`` (``CensoredSigmoidNormal`` 8-111):
``y = clamp(sigmoid(N(loc, scale)), lower, upper)``; the density is the logit-normal one inside the
limits and the censored probability mass at a limit, with the reference's asymptotic switch when the
fp32 cdf underflows (47-86).  ``log_prob`` runs on the sm_100a element-wise kernel and raises
``ArithmeticError`` on NaN / inf exactly where the reference does (83-84).
"""
from __future__ import annotations

import ctypes

import torch

from .. import _lib


class CensoredSigmoidNormal:
    has_rsample = True

    def __init__(self, loc, scale, lower_lim, upper_lim, validate_args=None):
        loc, scale = torch.as_tensor(loc, dtype=torch.float32), torch.as_tensor(scale, dtype=torch.float32)
        self.loc, self.scale = torch.broadcast_tensors(loc, scale)
        self.lower_lim, self.upper_lim = float(lower_lim), float(upper_lim)

    @property
    def batch_shape(self):
        return self.loc.shape

    def z(self, value):
        fi = torch.finfo(torch.float32)
        v = torch.clamp(value, fi.tiny, 1.0 - fi.eps)
        return ((v.log() - (-v).log1p()) - self.loc) / self.scale

    def rsample(self, sample_shape=torch.Size()):
        shape = torch.Size(sample_shape) + self.loc.shape
        eps = torch.randn(shape, dtype=self.loc.dtype, device=self.loc.device)
        x = torch.sigmoid(self.loc + eps * self.scale)
        # sigmoid saturates in fp32; torch's SigmoidTransform clamps to [tiny, 1 - eps] before the limits apply
        fi = torch.finfo(x.dtype)
        x = torch.clamp(x, fi.tiny, 1.0 - fi.eps)
        return torch.clamp(x, min=self.lower_lim, max=self.upper_lim)

    @torch.no_grad()
    def sample(self, sample_shape=torch.Size()):
        return self.rsample(sample_shape)

    def log_prob(self, value):
        value = torch.as_tensor(value, dtype=torch.float32, device=self.loc.device)
        loc, scale, value = torch.broadcast_tensors(self.loc, self.scale, value)
        loc, scale, value = _lib.f32c(loc), _lib.f32c(scale), _lib.f32c(value)
        out = torch.empty_like(loc)
        bad = torch.zeros((1,), dtype=torch.int32, device=loc.device)
        with torch.cuda.device(loc.device):
            _lib.check(_lib.lib().aline_censored_sigmoid_normal_log_prob(
                _lib.dptr(loc), _lib.dptr(scale), _lib.dptr(value), ctypes.c_float(self.lower_lim),
                ctypes.c_float(self.upper_lim), loc.numel(), _lib.dptr(out), _lib.dptr(bad, torch.int32),
                _lib.stream_ptr(loc.device)))
        if int(bad.item()):
            raise ArithmeticError("NaN in log_prob")
        return out
