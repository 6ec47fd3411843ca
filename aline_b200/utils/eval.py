"""Evaluation drivers -- same functions and signatures as the reference ``utils/eval.py``
(get_traces 9-39, compute_EIG_from_history 43-80, eval_EIG_from_history 84-140, eval_boed 143-198,
compute_ll 200-207, compute_rmse 210-232), running on the resident rollout and the fused sPCE kernel.

Multi-GPU: when ``torch.distributed`` is initialised, ``compute_EIG_from_history``
draws only this rank's slice of the L contrastive thetas and combines the per-(b,t) partial
log-sum-exp terms with one all-gather; rollouts (``eval_boed``'s outer loop) are split over ranks with
no collective until the final gather of the bounds.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import spce as _spce
from ..attrdict import AttrDict


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


@torch.no_grad()
def get_traces(model, experiment, T=30, batch_size=40, time_token=False, sampler="torch", seed=None, batch_offset=0):
    """T greedy design steps of ``batch_size`` rollouts, resident on the device (reference 9-39).

    Returns theta_0 [B, (K,) D], x = unnormalised designs [B, n_ctx0 + T, Dx], y [B, n_ctx0 + T, Dy].
    ``sampler="device"`` simulates the batch with one kernel from Philox streams keyed by (``seed``, global rollout
    index ``batch_offset`` + b) instead of ``experiment.sample_batch`` (torch's generator): nothing is drawn on, or
    copied from, the host.
    """
    model.eval()
    if sampler == "device":
        from .. import prior as _prior
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,), device="cpu").item())
        dev = next(model.parameters()).device
        theta_shape = _prior.theta_shape(experiment, batch_size)
        batch = _prior.sample_batch_device(experiment, batch_size, seed, batch_offset=batch_offset, device=dev)
    elif sampler == "torch":
        theta_shape = experiment.sample_theta((batch_size)).shape
        batch = experiment.sample_batch(batch_size)
    else:
        raise ValueError(f"unknown sampler {sampler!r} ('torch' or 'device')")
    batch = model.rollout(batch, T, time_token=time_token)
    theta_0 = batch.target_theta.reshape(*theta_shape)
    x = experiment.unnormalise_design(batch.context_x)
    y = batch.context_y
    return theta_0, x, y


@torch.no_grad()
def compute_EIG_from_history(experiment, theta_0, x, y, L=int(1e6), batch_size=40, stepwise=False, thetas=None,
                             shard=False, prior="torch", seed=None):
    """sPCE (lower) and sNMC (upper) EIG bounds from a minibatch of histories (reference 43-80).

    theta_0 [B, (K,) D]; x [B, T, Dx]; y [B, T, Dy].  Returns (pce, nmc), each [B, T] if stepwise else [B].
    ``thetas`` optionally supplies the L contrastive draws [L, B, (K,) D] (for value-exact comparisons);
    by default they are drawn from the prior exactly like the reference does (61-62).
    ``prior="device"`` draws the contrastive thetas from Philox streams keyed by (``seed``, global row, column) on the
    device (inside the fused pass for location K=1, D=2: they never touch HBM) instead of ``experiment.sample_theta``;
    the bounds are then statistically, not value-, identical to a torch-seeded run, and independent of the sharding.
    ``shard=True`` under an initialised ``torch.distributed`` group splits the L draws over the ranks: every rank
    must hold the SAME histories and draw DIFFERENT thetas (distinct RNG streams); the per-(b,t) partial
    (max, sum-exp) pairs are combined with one all-gather and every rank returns the full bounds.
    """
    dist = _dist() if shard else None
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    lo, hi = _spce.shard_rows(L, rank, world)
    n_local = hi - lo
    dev = x.device
    if prior not in ("torch", "device"):
        raise ValueError(f"unknown prior source {prior!r} ('torch' or 'device')")
    if thetas is None and prior == "device":
        from .. import prior as _prior
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,), device="cpu").item())     # one draw of the global generator
            if dist:                                                             # every rank must use the same key
                t = torch.tensor([seed], dtype=torch.int64, device=dev)
                dist.broadcast(t, 0)
                seed = int(t.item())
        m, s, lp0 = _prior.spce_history_device_prior(experiment, y, x, theta_0.to(dev), n_local, seed, row_offset=lo)
        if dist:
            m, s = _spce.all_gather_partials(m, s)
        pce_loss, nmc_loss = _spce.lse_combine(m, s, lp0)
        if not stepwise:
            pce_loss, nmc_loss = pce_loss[:, -1], nmc_loss[:, -1]
        return math.log(L + 1) - pce_loss, math.log(L) - nmc_loss
    if thetas is None:
        # L contrastive prior draws with theta_0 as row 0 (utils/eval.py:61-62).  One extra row is drawn and
        # overwritten instead of concatenating, which would copy the whole [L, B, .] tensor once more.
        rows = experiment.sample_theta((n_local + 1, batch_size)).to(dev)
        rows[0] = theta_0.to(dev)
    else:
        if world > 1:
            thetas = thetas[lo:hi]
        # row 0 = theta_0 on every rank (both bounds need its likelihood; the contrastive sum skips it)
        rows = torch.cat([theta_0.unsqueeze(0).to(dev), thetas.to(dev)], dim=0)
    m, s, lp0 = _spce.spce_history(experiment.log_likelihood, y, x, rows, seq=None, skip_rows=1)
    if dist:
        m, s = _spce.all_gather_partials(m, s)
    pce_loss, nmc_loss = _spce.lse_combine(m, s, lp0)
    if not stepwise:
        pce_loss, nmc_loss = pce_loss[:, -1], nmc_loss[:, -1]
    return math.log(L + 1) - pce_loss, math.log(L) - nmc_loss


def _summarise(pce, nmc, err_type):
    M = pce.shape[0]
    pce_mean, nmc_mean = torch.mean(pce, dim=0), torch.mean(nmc, dim=0)
    pce_err, nmc_err = torch.std(pce, dim=0), torch.std(nmc, dim=0)
    if err_type == "se":
        pce_err, nmc_err = pce_err / np.sqrt(M), nmc_err / np.sqrt(M)
    elif err_type == "ci":
        pce_err, nmc_err = 1.96 * pce_err / np.sqrt(M), 1.96 * nmc_err / np.sqrt(M)
    elif err_type != "std":
        raise ValueError(f"Unknown err_type: {err_type}")
    return AttrDict(pce_mean=pce_mean.cpu(), pce_err=pce_err.cpu(), nmc_mean=nmc_mean.cpu(), nmc_err=nmc_err.cpu())


@torch.no_grad()
def eval_EIG_from_history(experiment, theta_0, x, y, L=int(1e6), M=2000, batch_size=40, stepwise=False,
                          err_type="se"):
    """Bounds for M stored histories in minibatches (reference 84-140)."""
    pce_list, nmc_list = [], []
    for start in range(0, M, batch_size):
        end = min(start + batch_size, M)
        p, n = compute_EIG_from_history(experiment, theta_0[start:end], x[start:end], y[start:end], L, end - start,
                                        stepwise)
        pce_list.append(p)
        nmc_list.append(n)
    return _summarise(torch.cat(pce_list, 0), torch.cat(nmc_list, 0), err_type)


@torch.no_grad()
def eval_boed(model, experiment, T=30, L=int(1e6), M=2000, batch_size=40, time_token=False, stepwise=False,
              err_type="se", verbose=True, prior="torch", seed=None):
    """Final evaluation of the EIG bounds (reference 143-198): ceil(M / batch_size) x (rollout, bounds).

    Under ``torch.distributed`` the outer batches are dealt round-robin to the ranks (independent rollouts, no
    collective on the data path) and the per-rollout bounds are all-gathered once at the end, so every rank
    returns the statistics over all M outer samples.

    ``prior="device"`` (with an integer ``seed``, the same on every rank) runs the whole M-loop resident: batches are
    simulated on the device (``get_traces(sampler="device")``, rollout g keyed by its global index, so the result does
    not depend on the number of ranks) and the contrastive thetas are drawn on the device with key ``seed + 1 + step``."""
    model.eval()
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    if prior not in ("torch", "device"):
        raise ValueError(f"unknown prior source {prior!r} ('torch' or 'device')")
    if prior == "device" and seed is None:
        raise ValueError("eval_boed(prior='device') needs an integer seed (the same on every rank)")
    pce_list, nmc_list = [], []
    n_steps = (M + batch_size - 1) // batch_size
    for step in range(rank, n_steps, world):
        if prior == "device":
            theta_0, x, y = get_traces(model, experiment, T, batch_size, time_token, sampler="device", seed=seed,
                                       batch_offset=step * batch_size)
            pce, nmc = compute_EIG_from_history(experiment, theta_0, x, y, L, batch_size, stepwise, shard=False,
                                                prior="device", seed=int(seed) + 1 + step)
        else:
            theta_0, x, y = get_traces(model, experiment, T, batch_size, time_token)
            pce, nmc = compute_EIG_from_history(experiment, theta_0, x, y, L, batch_size, stepwise, shard=False)
        pce_list.append(pce)
        nmc_list.append(nmc)
        if verbose:
            print(f"Step {step}: PCE {pce.mean(dim=0)}, NMC {nmc.mean(dim=0)}")
    if pce_list:
        pce, nmc = torch.cat(pce_list, 0), torch.cat(nmc_list, 0)
    else:
        # more ranks than outer batches: this rank simulated nothing and only takes part in the gather
        dev = next(model.parameters()).device
        tail = (int(experiment.n_context_init) + int(T),) if stepwise else ()
        pce = torch.empty((0,) + tail, dtype=torch.float32, device=dev)
        nmc = torch.empty((0,) + tail, dtype=torch.float32, device=dev)
    if dist:
        pce, nmc = _gather_rows(dist, pce, n_steps, batch_size), _gather_rows(dist, nmc, n_steps, batch_size)
    return _summarise(pce, nmc, err_type)


def _gather_rows(dist, t, n_steps, batch_size):
    """All-gather per-rank result rows (ranks may own different numbers of outer batches)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    most = ((n_steps + world - 1) // world) * batch_size
    pad = torch.zeros((most,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    keep = [len(range(r, n_steps, world)) * batch_size for r in range(world)]
    return torch.cat([o[:k] for o, k in zip(out, keep)], 0)


def compute_ll(value, means, stds, weights):
    """GMM log-likelihood ``logsumexp_c(log N(value; mu_c, sigma_c) + log w_c)`` (reference 200-207)."""
    from ..rollout import gmm_log_likelihood
    return gmm_log_likelihood(value, means, stds, weights)


def compute_rmse(target_values, mixture_means, mixture_stds, mixture_weights):
    """RMSE of the mixture mean against the targets (reference 210-232)."""
    weighted = torch.sum(mixture_weights * mixture_means, dim=-1)
    return torch.sqrt(torch.mean((target_values.squeeze(-1) - weighted) ** 2, dim=-1))
