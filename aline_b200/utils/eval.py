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
def get_traces(model, experiment, T=30, batch_size=40, time_token=False, sampler="torch", seed=None, batch_offset=0,
               batch=None):
    """T greedy design steps of ``batch_size`` rollouts, resident on the device (reference 9-39).

    Returns theta_0 [B, (K,) D], x = unnormalised designs [B, n_ctx0 + T, Dx], y [B, n_ctx0 + T, Dy].
    ``sampler="device"`` simulates the batch with one kernel from Philox streams keyed by (``seed``, global rollout
    index ``batch_offset`` + b) instead of ``experiment.sample_batch`` (torch's generator): nothing is drawn on, or
    copied from, the host.  ``batch`` supplies an already simulated batch (the fields of ``experiment.sample_batch``;
    host tensors -- pinned for an asynchronous copy -- or device tensors) instead of sampling one.
    """
    model.eval()
    if batch is not None:
        dev = next(model.parameters()).device
        src = batch
        batch = AttrDict({k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in dict(src).items()})
        if "target_theta" not in batch:
            batch.target_theta = batch.target_all
        if batch.context_x.shape[0] != batch_size:
            raise ValueError(f"the supplied batch has {batch.context_x.shape[0]} rollouts, expected {batch_size}")
        try:                                          # shape of sample_theta(batch_size) without consuming the generator
            from .. import prior as _prior
            theta_shape = _prior.theta_shape(experiment, batch_size)
        except Exception:   # noqa: BLE001 -- a task without a prior descriptor (GP): ask it
            theta_shape = experiment.sample_theta((batch_size)).shape
    elif sampler == "device":
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
    m, s, lp0 = _spce.spce_history(experiment.log_likelihood, y, x, rows, seq=None, skip_rows=1, last_only=not stepwise)
    if dist:
        m, s = _spce.all_gather_partials(m, s)
    pce_loss, nmc_loss = _spce.lse_combine(m, s, lp0)
    if not stepwise:
        pce_loss, nmc_loss = pce_loss[:, -1], nmc_loss[:, -1]
    return math.log(L + 1) - pce_loss, math.log(L) - nmc_loss


def _summarise(pce, nmc, err_type, nmc_pre_scaled=False):
    """Mean and error of the bounds over the M outer samples (reference 168-196).  ``nmc_pre_scaled`` reproduces a quirk of
    the reference's ``eval_EIG_from_history`` only: there the sNMC standard deviation is divided by sqrt(M) BEFORE the
    error-type switch (utils/eval.py:119), so its 'se' / 'ci' carry 1/M and its 'std' is really a standard error."""
    M = pce.shape[0]
    pce_mean, nmc_mean = torch.mean(pce, dim=0), torch.mean(nmc, dim=0)
    pce_err, nmc_err = torch.std(pce, dim=0), torch.std(nmc, dim=0)
    if nmc_pre_scaled:
        nmc_err = nmc_err / np.sqrt(M)
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
    return _summarise(torch.cat(pce_list, 0), torch.cat(nmc_list, 0), err_type, nmc_pre_scaled=True)


def rank_chunks(n_rollouts, batch_size, rank=0, world=1):
    """This rank's share of ``n_rollouts`` outer samples as a list of (global offset, size) mini-batches.

    Ranks own contiguous, balanced slices (sizes differ by at most one); a slice is cut into the fewest balanced
    mini-batches.  One rank: mini-batches of at most ``batch_size`` -- ceil(n / batch_size) batches of ``batch_size`` when
    it divides n, the reference's loop (utils/eval.py:155-158).  Several ranks: ``batch_size`` is a memory knob (the
    contrastive draws are L x B x dim_theta floats), and a rollout of B trajectories is latency-bound (34 dependent
    steps), so a slice that exceeds ``batch_size`` by at most 25 % is NOT split: M = 2000, batch 200 on 8 ranks is one
    mini-batch of 250 per rank (2 x 125 measured 16.8 ms per evaluation against 115.6 ms on one GPU = 0.86 efficiency;
    the round-1 round-robin deal of whole batches gave two ranks 400 rollouts and six ranks 200)."""
    lo, hi = _spce.shard_rows(n_rollouts, rank, world)
    n = hi - lo
    if n <= 0:
        return []
    cap = batch_size if world == 1 else max(batch_size, (batch_size * 5) // 4)
    k = (n + cap - 1) // cap
    base, rem = divmod(n, k)
    out, off = [], lo
    for i in range(k):
        sz = base + (1 if i < rem else 0)
        out.append((off, sz))
        off += sz
    return out


class _TwoStage:
    """Software pipeline of the evaluation's outer loop on two CUDA streams: mini-batch i + 1's rollout (34 dependent
    steps of small kernels, latency-bound, a third of the SMs idle in its context kernels) runs on a high-priority
    stream while mini-batch i's bound (one long MUFU-bound pass over L x B draws) runs on a second stream.  The HOST
    issues the work in the serial order -- sample, rollout, draw thetas, bound -- so torch's generator is consumed in
    exactly the serial order and the bounds are bit-identical to the unpipelined loop."""

    _streams = {}       # device index -> (rollout stream, bound stream): persistent, so that CUDA graphs and scratch
                        # buffers keyed by stream are reused from one evaluation to the next

    def __init__(self, device, enabled):
        self.enabled = bool(enabled)
        self.device = device
        if self.enabled:
            self.main = torch.cuda.current_stream(device)
            key = torch.device(device).index
            if key not in _TwoStage._streams:
                _TwoStage._streams[key] = (torch.cuda.Stream(device, priority=-1), torch.cuda.Stream(device))
            self.s_roll, self.s_bound = _TwoStage._streams[key]
            self.s_roll.wait_stream(self.main)
            self.s_bound.wait_stream(self.main)

    def rollout(self, fn):
        if not self.enabled:
            return fn()
        with torch.cuda.stream(self.s_roll):
            out = fn()
        self.s_bound.wait_stream(self.s_roll)
        for t in out:
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(self.s_bound)
        return out

    def bound(self, fn):
        if not self.enabled:
            return fn()
        with torch.cuda.stream(self.s_bound):
            out = fn()
        for t in out:
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(self.main)
        return out

    def join(self):
        if self.enabled:
            self.main.wait_stream(self.s_roll)
            self.main.wait_stream(self.s_bound)


def _decorrelate_rank_generators(dist, device):
    """Ranks that were seeded identically (the reference's ``set_seed(cfg.seed)`` runs on every rank) would simulate
    the same rollouts and draw the same contrastive thetas: M / world distinct samples counted ``world`` times, and an
    understated standard error.  Detect it (all-gather of the generators' seeds) and re-seed rank r's CUDA generator
    with seed + 7919 r (rank 0 keeps its stream)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    seed = int(torch.cuda.initial_seed()) if device.type == "cuda" else int(torch.initial_seed())
    mine = torch.tensor([seed & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=device)
    seeds = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(seeds, mine)
    seeds = [int(t.item()) for t in seeds]
    if rank > 0 and seeds[rank] in seeds[:rank]:
        import warnings
        warnings.warn(f"eval_boed: rank {rank} shares its RNG seed with a lower rank; re-seeding with seed + 7919 * rank "
                      "so that the ranks simulate different rollouts")
        if device.type == "cuda":
            torch.cuda.manual_seed(seed + 7919 * rank)
        else:
            torch.manual_seed(seed + 7919 * rank)


@torch.no_grad()
def eval_boed(model, experiment, T=30, L=int(1e6), M=2000, batch_size=40, time_token=False, stepwise=False,
              err_type="se", verbose=True, prior="torch", seed=None, overlap=True, batches=None):
    """Final evaluation of the EIG bounds (reference 143-198): ceil(M / batch_size) x (rollout, bounds).

    Under ``torch.distributed`` the ceil(M / batch_size) * batch_size outer samples are dealt to the ranks as balanced
    contiguous slices (``rank_chunks``; independent rollouts, no collective on the data path) and the per-rollout bounds
    are all-gathered once at the end, so every rank returns the statistics over all outer samples.  With the default
    ``prior="torch"`` every rank draws from its own torch generator: ranks must be seeded differently, and ranks found
    sharing a seed are re-seeded (``_decorrelate_rank_generators``).

    ``overlap=True`` pipelines the loop on two streams (``_TwoStage``): same bounds, bit for bit, as the serial loop.
    The per-step ``print`` of the reference (utils/eval.py:165) would synchronise the host with every mini-batch, so
    with ``verbose`` the same lines are printed after the loop.

    ``prior="device"`` (with an integer ``seed``, the same on every rank) runs the whole M-loop resident: batches are
    simulated on the device (``get_traces(sampler="device")``, rollout g keyed by its global index, so the result does
    not depend on the number of ranks) and the contrastive thetas are drawn on the device with key
    ``seed + 1 + (global offset of the mini-batch)``.

    ``batches`` (a sequence with one already simulated batch per mini-batch of THIS rank, sizes as in ``rank_chunks``;
    host -- ideally pinned -- or device tensors) replaces ``experiment.sample_batch``: the copies to the device are part
    of the pipeline."""
    model.eval()
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    if prior not in ("torch", "device"):
        raise ValueError(f"unknown prior source {prior!r} ('torch' or 'device')")
    if prior == "device" and seed is None:
        raise ValueError("eval_boed(prior='device') needs an integer seed (the same on every rank)")
    dev = next(model.parameters()).device
    if dist and prior == "torch":
        _decorrelate_rank_generators(dist, dev)
    n_steps = (M + batch_size - 1) // batch_size
    n_total = n_steps * batch_size                    # the reference evaluates whole mini-batches (>= M rollouts)
    chunks = rank_chunks(n_total, batch_size, rank, world)
    pipe = _TwoStage(dev, overlap and dev.type == "cuda" and len(chunks) > 1)
    pce_list, nmc_list = [], []
    if batches is not None and len(batches) != len(chunks):
        raise ValueError(f"batches has {len(batches)} entries, this rank evaluates {len(chunks)} mini-batches")
    for ci, (off, bs) in enumerate(chunks):
        if batches is not None:
            theta_0, x, y = pipe.rollout(lambda: get_traces(model, experiment, T, bs, time_token, batch=batches[ci]))
            pce, nmc = pipe.bound(lambda: compute_EIG_from_history(
                experiment, theta_0, x, y, L, bs, stepwise, shard=False, prior=prior,
                seed=None if seed is None else int(seed) + 1 + off))
        elif prior == "device":
            theta_0, x, y = pipe.rollout(lambda: get_traces(model, experiment, T, bs, time_token, sampler="device",
                                                            seed=seed, batch_offset=off))
            pce, nmc = pipe.bound(lambda: compute_EIG_from_history(experiment, theta_0, x, y, L, bs, stepwise,
                                                                   shard=False, prior="device",
                                                                   seed=int(seed) + 1 + off))
        else:
            theta_0, x, y = pipe.rollout(lambda: get_traces(model, experiment, T, bs, time_token))
            pce, nmc = pipe.bound(lambda: compute_EIG_from_history(experiment, theta_0, x, y, L, bs, stepwise,
                                                                   shard=False))
        pce_list.append(pce)
        nmc_list.append(nmc)
    pipe.join()
    if verbose:
        for (off, bs), pce, nmc in zip(chunks, pce_list, nmc_list):
            print(f"Step {off // batch_size}: PCE {pce.mean(dim=0)}, NMC {nmc.mean(dim=0)}")
    if pce_list:
        pce, nmc = torch.cat(pce_list, 0), torch.cat(nmc_list, 0)
    else:
        # more ranks than outer samples: this rank simulated nothing and only takes part in the gather
        tail = (int(experiment.n_context_init) + int(T),) if stepwise else ()
        pce = torch.empty((0,) + tail, dtype=torch.float32, device=dev)
        nmc = torch.empty((0,) + tail, dtype=torch.float32, device=dev)
    if dist:
        pce, nmc = gather_rows(dist, pce, n_total), gather_rows(dist, nmc, n_total)
    return _summarise(pce, nmc, err_type)


def gather_rows(dist, t, n_total):
    """All-gather the ranks' result rows ([n_r, ...], n_r = size of rank r's slice of ``n_total``) in global order:
    the one collective of the rollout-sharded evaluation (NCCL all-gather of at most ceil(n_total / world) rows)."""
    world = dist.get_world_size()
    sizes = [hi - lo for lo, hi in (_spce.shard_rows(n_total, r, world) for r in range(world))]
    most = max(sizes)
    pad = torch.zeros((most,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = torch.empty((world * most,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad)
    out = out.reshape((world, most) + tuple(t.shape[1:]))
    return torch.cat([out[r, :k] for r, k in enumerate(sizes)], 0)


def compute_ll(value, means, stds, weights):
    """GMM log-likelihood ``logsumexp_c(log N(value; mu_c, sigma_c) + log w_c)`` (reference 200-207)."""
    from ..rollout import gmm_log_likelihood
    return gmm_log_likelihood(value, means, stds, weights)


def compute_rmse(target_values, mixture_means, mixture_stds, mixture_weights):
    """RMSE of the mixture mean against the targets (reference 210-232)."""
    weighted = torch.sum(mixture_weights * mixture_means, dim=-1)
    return torch.sqrt(torch.mean((target_values.squeeze(-1) - weighted) ** 2, dim=-1))
