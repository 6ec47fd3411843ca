"""Host wrappers of the sPCE / sNMC kernels (C ABI: include/aline_b200.h).

reference: loss/eig.py:154-209 (EIGStepLoss), loss/eig.py:22-151 (PCELoss / NMCLoss),
utils/eval.py:43-80 (compute_EIG_from_history).
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import _lib
from ._lib import AlineError, AlineLik, TASK_CES, TASK_LOCATION, TASK_PSYCHOMETRIC, dptr


def lik_of(obj) -> AlineLik:
    """Likelihood descriptor of a task (or of its bound ``log_likelihood`` method).

    The reference passes ``experiment.log_likelihood`` -- a Python bound method -- into the EIG
    criteria (utils/eval.py:56); the native path dispatches on the task behind it and reads its
    constants.  Anything else raises: there is no generic / CPU fallback.
    """
    if isinstance(obj, AlineLik):
        return obj
    task = getattr(obj, "__self__", obj)
    name = type(task).__name__
    if hasattr(task, "aline_lik"):
        return task.aline_lik()
    if name == "HiddenLocation":          # tasks/location_finding.py:8-83
        K, D = int(task.K), int(task.dim_x)
        return AlineLik(TASK_LOCATION, D, K, K * D, float(task.noise_scale), float(task.base_signal),
                        float(task.max_signal), 0.0)
    if name == "CESTask":                 # tasks/ces.py
        return AlineLik(TASK_CES, 6, 1, 5, float(task.noise_scale), float(task.epsilon), 0.0, 0.0)
    if name == "PsychometricTask":        # tasks/psychometric.py
        return AlineLik(TASK_PSYCHOMETRIC, 1, 1, 4, 0.0, 0.0, 0.0, 0.0)
    raise AlineError(f"log_prob callable of type {name!r} has no sm_100a likelihood kernel "
                     "(supported: HiddenLocation, CESTask, PsychometricTask); no CPU fallback exists")


def _flat_thetas(thetas: torch.Tensor, lik: AlineLik) -> torch.Tensor:
    if thetas.dim() < 3:
        raise AlineError(f"thetas must be [L, B, (K,) D], got {tuple(thetas.shape)}")
    th = _lib.f32c(thetas)
    n_rows, B = th.shape[:2]
    th = th.reshape(n_rows, B, -1)
    if th.shape[-1] != lik.dim_theta:
        raise AlineError(f"thetas trailing size {th.shape[-1]} != dim_theta {lik.dim_theta}")
    return th


def raise_if_bad(bad: torch.Tensor):
    """The reference raises ArithmeticError when the CES log-prob has NaN / inf
    (distributions/censored_sigmoid_normal.py:83-84); the kernels set a device flag instead."""
    if int(bad.item()) != 0:
        raise ArithmeticError("NaN in log_prob")


def spce_history(lik, y, xi, thetas, seq=None, skip_rows=1, check=True, last_only=False):
    """Partial log-sum-exp terms of the step-wise bounds over this caller's rows.

    y [B,T] or [B,T,1]; xi [B,T,dx] (unnormalised designs); thetas [n_rows,B,(K,)D];
    seq [n_rows,B] in/out accumulator or None.  Returns (m, s, lp0), each [B,T]:
    running max / sum exp(. - m) over rows >= skip_rows and the row-0 value (zeros if skip_rows = 0).
    ``last_only``: the caller only uses [:, T-1] (stepwise=False); the fused pass then skips the per-step exponentials.
    """
    lik = lik_of(lik)
    th = _flat_thetas(thetas, lik)
    n_rows, B = th.shape[:2]
    xi = _lib.f32c(xi)
    if xi.dim() != 3 or xi.shape[0] != B or xi.shape[2] != lik.dim_x:
        raise AlineError(f"xi must be [B={B}, T, dim_x={lik.dim_x}], got {tuple(xi.shape)}")
    T = xi.shape[1]
    y = _lib.f32c(y).reshape(B, T)
    dev = th.device
    m = torch.empty((B, T), dtype=torch.float32, device=dev)
    s = torch.empty((B, T), dtype=torch.float32, device=dev)
    lp0 = (torch.empty if skip_rows else torch.zeros)((B, T), dtype=torch.float32, device=dev)
    bad = torch.zeros((1,), dtype=torch.int32, device=dev) if lik.task == TASK_CES else None
    if seq is not None and tuple(seq.shape) != (n_rows, B):
        raise AlineError(f"seq must be [{n_rows}, {B}], got {tuple(seq.shape)}")
    L = _lib.lib()
    flags = 0
    if seq is None and T > L.aline_spce_pass_len(ctypes.byref(lik), B):
        # multi-pass: the accumulated log-likelihood is carried between passes in a scratch buffer the kernels own
        # (never read before it is written: no zero fill needed), which also enables the shifted fast pass
        seq = torch.empty((n_rows, B), dtype=torch.float32, device=dev)
        flags = 1 | (2 if last_only else 0)      # ALINE_SPCE_SEQ_SCRATCH | ALINE_SPCE_LAST_ONLY
    nbytes = L.aline_spce_scratch_bytes(B, T)
    sc = _lib.scratch(nbytes, dev)
    with torch.cuda.device(dev):
        _lib.check(L.aline_spce_history_ex(ctypes.byref(lik), dptr(y), dptr(xi), dptr(th), dptr(seq, name="seq"),
                                           n_rows, B, T, int(skip_rows), dptr(m), dptr(s), dptr(lp0),
                                           dptr(bad, torch.int32), ctypes.c_void_p(sc.data_ptr()), nbytes, flags,
                                           _lib.stream_ptr(dev)))
    if check and lik.task == TASK_CES:
        raise_if_bad(bad)
    return m, s, lp0


def spce_step(lik, y, xi, thetas, seq, skip_rows=1, check=True):
    """EIGStepLoss.step + the reductions of .forward for one history point (loss/eig.py:174-209).
    y [B] or [B,1]; xi [B,dx]; seq [n_rows,B] updated in place.  Returns (m, s, lp0) each [B]."""
    lik = lik_of(lik)
    th = _flat_thetas(thetas, lik)
    n_rows, B = th.shape[:2]
    xi = _lib.f32c(xi).reshape(B, lik.dim_x)
    y = _lib.f32c(y).reshape(B)
    dev = th.device
    if seq is None or tuple(seq.shape) != (n_rows, B):
        raise AlineError(f"seq must be [{n_rows}, {B}]")
    m = torch.empty((B,), dtype=torch.float32, device=dev)
    s = torch.empty((B,), dtype=torch.float32, device=dev)
    lp0 = (torch.empty if skip_rows else torch.zeros)((B,), dtype=torch.float32, device=dev)
    bad = torch.zeros((1,), dtype=torch.int32, device=dev) if lik.task == TASK_CES else None
    L = _lib.lib()
    nbytes = L.aline_spce_scratch_bytes(B, 1)
    sc = _lib.scratch(nbytes, dev)
    with torch.cuda.device(dev):
        _lib.check(L.aline_spce_step(ctypes.byref(lik), dptr(y), dptr(xi), dptr(th), dptr(seq, name="seq"), n_rows, B,
                                     int(skip_rows), dptr(m), dptr(s), dptr(lp0), dptr(bad, torch.int32),
                                     ctypes.c_void_p(sc.data_ptr()), nbytes, _lib.stream_ptr(dev)))
    if check and lik.task == TASK_CES:
        raise_if_bad(bad)
    return m, s, lp0


def lse_combine(m, s, lp0):
    """(m, s) [R, ...] partials of R shards + lp0 [...] -> (pce_loss, nmc_loss) [...] (loss/eig.py:200-202)."""
    if m.dim() == lp0.dim():
        m, s = m.unsqueeze(0), s.unsqueeze(0)
    m, s, lp0 = _lib.f32c(m), _lib.f32c(s), _lib.f32c(lp0)
    R, n = m.shape[0], lp0.numel()
    pce = torch.empty_like(lp0)
    nmc = torch.empty_like(lp0)
    with torch.cuda.device(lp0.device):
        _lib.check(_lib.lib().aline_lse_combine(dptr(m), dptr(s), dptr(lp0), R, n, dptr(pce), dptr(nmc),
                                                _lib.stream_ptr(lp0.device)))
    return pce, nmc


def log_likelihood(lik, y, xi, thetas, check=True):
    """Task.log_likelihood with y [1,B,1] / [B,1] / [B], xi [1,B,dx] / [B,dx], thetas [n_rows,B,(K,)D] -> [n_rows,B,1]."""
    lik = lik_of(lik)
    th = _flat_thetas(thetas, lik)
    n_rows, B = th.shape[:2]
    xi = _lib.f32c(xi).reshape(B, lik.dim_x)
    y = _lib.f32c(y).reshape(B)
    dev = th.device
    out = torch.empty((n_rows, B), dtype=torch.float32, device=dev)
    bad = torch.zeros((1,), dtype=torch.int32, device=dev)
    L = _lib.lib()
    nbytes = L.aline_spce_scratch_bytes(B, 1)
    sc = _lib.scratch(nbytes, dev)
    with torch.cuda.device(dev):
        _lib.check(L.aline_log_likelihood(ctypes.byref(lik), dptr(y), dptr(xi), dptr(th), dptr(out), n_rows, B,
                                          dptr(bad, torch.int32), ctypes.c_void_p(sc.data_ptr()), nbytes,
                                          _lib.stream_ptr(dev)))
    if check and lik.task == TASK_CES:
        raise_if_bad(bad)
    return out.unsqueeze(-1)


# ------------------------------------------------------------------ multi-GPU ----
def shard_rows(L: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of the L contrastive rows (1-based rows 1..L of thetas) owned by `rank`."""
    per, rem = divmod(L, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def all_gather_partials(m, s, group=None):
    """One all-gather of the per-(b,t) (max, sum-exp) pairs: the only collective of the sharded bound
    (SURVEY.md section 8e).  Returns m, s of shape [world, ...]."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    ms = torch.stack([m, s], 0).contiguous()
    parts = [torch.empty_like(ms) for _ in range(world)]
    dist.all_gather(parts, ms, group=group)
    out = torch.stack(parts, 0)
    return out[:, 0].contiguous(), out[:, 1].contiguous()


def bounds_from_losses(pce_loss, nmc_loss, L):
    """utils/eval.py:77-78."""
    return math.log(L + 1) - pce_loss, math.log(L) - nmc_loss
