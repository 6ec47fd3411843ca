"""Host wrappers of the GP prior-draw kernels (C ABI: include/aline_b200.h).
reference: tasks/gaussian_process.py:194-317 (kernels), 366-417 (generate_gp_data)."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import dptr

F32, I32 = torch.float32, torch.int32


def gp_sample(x, lengthscales, scale, kernel_type, z, eps, jitter=1e-5, noise_scale=0.01, return_factors=False,
              check=True):
    """x [B,N,dx], lengthscales [B,dx], scale [B], kernel_type [B] int32, z / eps [B,N] -> y [B,N]
    (and L, K [B,N,N] if return_factors).  Raises if a matrix is not positive definite (the reference's fallback,
    MultivariateNormal, raises on the same input)."""
    x = _lib.f32c(x)
    B, N, dx = x.shape
    dev = x.device
    ls, sc = _lib.f32c(lengthscales.to(dev)), _lib.f32c(scale.to(dev)).reshape(B)
    kt = kernel_type.to(dev, I32).contiguous()
    z, eps = _lib.f32c(z.to(dev)).reshape(B, N), _lib.f32c(eps.to(dev)).reshape(B, N)
    y = torch.empty((B, N), dtype=F32, device=dev)
    Lm = torch.empty((B, N, N), dtype=F32, device=dev) if return_factors else None
    Km = torch.empty((B, N, N), dtype=F32, device=dev) if return_factors else None
    info = torch.zeros((B,), dtype=I32, device=dev)
    lib = _lib.lib()
    nbytes = lib.aline_gp_scratch_bytes(B, N)
    sc_buf = _lib.scratch(nbytes, dev) if nbytes else None
    with torch.cuda.device(dev):
        _lib.check(lib.aline_gp_sample(dptr(x), B, N, dx, dptr(ls), dptr(sc), dptr(kt, I32), dptr(z), dptr(eps),
                                       ctypes.c_float(jitter), ctypes.c_float(noise_scale), dptr(y), dptr(Lm), dptr(Km),
                                       dptr(info, I32), None if sc_buf is None else ctypes.c_void_p(sc_buf.data_ptr()),
                                       nbytes, _lib.stream_ptr(dev)))
    if check and bool(info.any()):
        raise RuntimeError("GP kernel matrix is not positive definite (Cholesky failed)")
    return (y, Lm, Km) if return_factors else y


def kernel_matrix(x1, x2, lengthscales, scale, kernel_type: int):
    """One batch element: x1 [N,dx], x2 [M,dx], lengthscales [dx], scale scalar tensor -> K [N,M] (no jitter)."""
    x1, x2 = _lib.f32c(x1), _lib.f32c(x2)
    dev = x1.device
    N, dx = x1.shape
    M = x2.shape[0]
    ls = _lib.f32c(lengthscales.to(dev)).reshape(dx)
    sc = _lib.f32c(torch.as_tensor(scale, dtype=F32).to(dev)).reshape(1)
    K = torch.empty((N, M), dtype=F32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().aline_gp_kernel_matrix(dptr(x1), dptr(x2), N, M, dx, dptr(ls), dptr(sc), int(kernel_type),
                                                     dptr(K), _lib.stream_ptr(dev)))
    return K
