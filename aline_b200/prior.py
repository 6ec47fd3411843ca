"""Device-side prior draws: Philox4x32-10 streams keyed by (seed, global row, column)
instead of torch's global generator.  Statistical -- not value -- parity with ``Task.sample_theta``; the draws do not
depend on how the contrastive rows are sharded over ranks.

    sample_theta_device(task, n_rows, B, seed)            -> thetas [n_rows, B, (K,) D]      (aline_prior_sample)
    spce_history_device_prior(task, y, xi, theta_0, L, seed) -> (m, s, lp0) with the L contrastive rows drawn inside
        the fused sPCE pass (location K=1, D=2: they never touch HBM); other tasks materialise the draws first.
    sample_batch_device(task, B, seed, batch_offset)       -> the AttrDict of ``task.sample_batch(B)``: theta_0, designs
        and pre-simulated outcomes of B rollouts in one kernel (aline_sample_batch)

reference: tasks/location_finding.py:85-98, tasks/psychometric.py:70-89, tasks/ces.py:52-81 (the priors);
utils/eval.py:61-62 (where the contrastive draws are made); tasks/location_finding.py:167-192, tasks/ces.py:213-234,
tasks/psychometric.py:197-222 (sample_batch).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, spce as _spce
from ._lib import AlineError, dptr

PRIOR_BOX, PRIOR_CES = 0, 1


class AlinePrior(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("dim_theta", ctypes.c_int32), ("lo", ctypes.c_float * 16),
                ("hi", ctypes.c_float * 16)]


def _box(lo, hi):
    lo, hi = [float(v) for v in lo], [float(v) for v in hi]
    if len(lo) != len(hi) or not 1 <= len(lo) <= 16:
        raise AlineError(f"box prior with {len(lo)} / {len(hi)} bounds (1..16 supported)")
    p = AlinePrior(PRIOR_BOX, len(lo))
    for i, (a, b) in enumerate(zip(lo, hi)):
        p.lo[i], p.hi[i] = a, b
    return p


def prior_of(task) -> AlinePrior:
    """Prior descriptor of a task; raises for priors without a device generator (no silent torch fallback)."""
    name = type(task).__name__
    if name == "HiddenLocation":
        if task.theta_dist != "uniform":
            raise AlineError("device-side prior draws: HiddenLocation supports theta_dist='uniform'")
        lo = torch.as_tensor(task.theta_loc, dtype=torch.float32).cpu().reshape(-1).tolist()
        hi = torch.as_tensor(task.theta_cov, dtype=torch.float32).cpu().reshape(-1).tolist()
        return _box(lo, hi)
    if name == "PsychometricTask":
        from .tasks.psychometric import _PRIOR
        return _box([a for a, _ in _PRIOR], [b for _, b in _PRIOR])
    if name == "CESTask":
        p = AlinePrior(PRIOR_CES, 5)
        p.lo[4], p.hi[4] = float(task.u_mu), float(task.u_sigma)
        return p
    raise AlineError(f"{name} has no device-side prior generator")


def _theta_tail(task):
    name = type(task).__name__
    if name == "HiddenLocation":
        return [int(task.K), int(task.dim_x)]
    return [4] if name == "PsychometricTask" else [5]


def sample_theta_device(task, n_rows, B, seed, row_offset=0, device=None):
    """Prior draws of the global rows row_offset .. row_offset+n_rows-1: [n_rows, B, (K,) D] like
    ``task.sample_theta((n_rows, B))``."""
    pr = prior_of(task)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty([int(n_rows), int(B)] + _theta_tail(task), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().aline_prior_sample(ctypes.byref(pr), ctypes.c_uint64(int(seed) & (2 ** 64 - 1)),
                                                 int(row_offset), int(n_rows), int(B), dptr(out), _lib.stream_ptr(dev)))
    return out


def theta_shape(task, B):
    """Shape of ``task.sample_theta(B)`` (what utils/eval.py:18 reads to reshape ``target_theta``)."""
    tail = _theta_tail(task)
    return [int(B)] + tail + ([1] if type(task).__name__ == "PsychometricTask" else [])


def sample_batch_device(task, B, seed, batch_offset=0, device=None):
    """``task.sample_batch(B)`` drawn on the device for the rollouts batch_offset .. batch_offset + B - 1: same fields
    and shapes, Philox streams keyed by (seed, global rollout index, point) -- a rollout's draw does not depend on the
    batch size or on which rank simulates it.  Statistical parity with the torch-generator path."""
    from .attrdict import AttrDict
    name = type(task).__name__
    pr = prior_of(task)
    lik = task.aline_lik()
    n_c, n_q = int(task.n_context_init), int(task.n_query_init)
    n = n_c + n_q
    if name == "HiddenLocation":
        x_lo, x_hi, scale = float(task._data_low), float(task._data_high), float(task.design_scale)
    elif name == "CESTask":
        x_lo, x_hi, scale = 0.0, float(task.design_scale), 1.0
    elif name == "PsychometricTask":
        x_lo, x_hi, scale = -float(task.design_scale), float(task.design_scale), 1.0
    else:
        raise AlineError(f"{name} has no device-side sample_batch")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    B = int(B)
    theta = torch.empty((B, lik.dim_theta), dtype=torch.float32, device=dev)
    x = torch.empty((B, n, lik.dim_x), dtype=torch.float32, device=dev)
    y = torch.empty((B, n, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().aline_sample_batch(ctypes.byref(lik), ctypes.byref(pr),
                                                 ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), int(batch_offset), B, n,
                                                 ctypes.c_float(x_lo), ctypes.c_float(x_hi), ctypes.c_float(scale),
                                                 dptr(theta), dptr(x), dptr(y), _lib.stream_ptr(dev)))
    batch = AttrDict()
    batch.context_x, batch.context_y = x[:, :n_c], y[:, :n_c]
    batch.query_x, batch.query_y = x[:, n_c:], y[:, n_c:]
    batch.target_all = batch.target_theta = theta.reshape(B, lik.dim_theta, 1)
    if name == "CESTask":
        batch.n_theta = task.n_theta
    else:
        batch.n_target_theta = task.n_target_theta
    return batch


def spce_history_device_prior(task, y, xi, theta_0, L, seed, row_offset=0, check=True):
    """(m, s, lp0), each [B, T], of the step-wise bounds with contrastive rows 1..L drawn on the device from
    (seed, row_offset + row, b).  Location K=1, D=2: drawn inside the fused pass.  If a shifted sum was invalid (device
    flag; `check=True` reads it, one synchronisation) or the task has no in-kernel generator, the same draws are
    materialised with ``sample_theta_device`` and the explicit kernels run on them."""
    lik = _spce.lik_of(task)
    xi = _lib.f32c(xi)
    B, T = xi.shape[:2]
    y = _lib.f32c(y).reshape(B, T)
    dev = xi.device
    th0 = _lib.f32c(theta_0).reshape(1, B, -1)
    if th0.shape[-1] != lik.dim_theta:
        raise AlineError(f"theta_0 trailing size {th0.shape[-1]} != dim_theta {lik.dim_theta}")
    pr = prior_of(task)

    def explicit():
        rows = sample_theta_device(task, L + 1, B, seed, row_offset=row_offset, device=dev).reshape(L + 1, B, -1)
        rows[0] = th0[0]              # row 0 of the stream is unused: theta_0 takes its place (utils/eval.py:61-62)
        return _spce.spce_history(lik, y, xi, rows, seq=None, skip_rows=1, check=check)

    in_kernel = (lik.task == _lib.TASK_LOCATION and lik.K == 1 and lik.dim_x == 2 and pr.kind == PRIOR_BOX)
    if not in_kernel:
        return explicit()
    m = torch.empty((B, T), dtype=torch.float32, device=dev)
    s = torch.empty((B, T), dtype=torch.float32, device=dev)
    lp0 = torch.empty((B, T), dtype=torch.float32, device=dev)
    redo = torch.zeros((1,), dtype=torch.int32, device=dev)
    lib = _lib.lib()
    # scratch for the accumulated log-likelihoods: one row (theta_0's) when the history fits one pass, else [L + 1, B]
    seq_rows = int(lib.aline_spce_device_prior_seq_rows(ctypes.byref(lik), L + 1, T))
    seq = torch.empty((seq_rows, B), dtype=torch.float32, device=dev)
    nbytes = lib.aline_spce_scratch_bytes(B, T)
    sc = _lib.scratch(nbytes, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.aline_spce_history_device_prior(
            ctypes.byref(lik), ctypes.byref(pr), ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), int(row_offset), dptr(y),
            dptr(xi), dptr(th0), dptr(seq), L + 1, B, T, dptr(m), dptr(s), dptr(lp0), dptr(redo, torch.int32),
            ctypes.c_void_p(sc.data_ptr()), nbytes, _lib.stream_ptr(dev)))
    if check and int(redo.item()) != 0:
        return explicit()
    return m, s, lp0
