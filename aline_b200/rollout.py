"""Host wrappers of the forward / rollout kernels (C ABI: include/aline_b200.h).

reference: model/base.py:32-50 (Aline.forward), tasks/base_task.py:103-154 (Task.update_batch),
utils/eval.py:9-39 (get_traces), utils/eval.py:200-207 (compute_ll).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import AlineError, dptr

I32, I64, U8, F32 = torch.int32, torch.int64, torch.uint8, torch.float32


def _st(dev):
    return _lib.stream_ptr(dev)


_slot_cache = {}


def target_slots(n_t, target_mask, device):
    """int32 [n_t]: rank of target i among the targets the candidate queries attend to, -1 if not attended.
    ``target_mask`` None = attend to all (model/encoder.py:108-124).  The device copy is cached per (mask, device): a
    pageable host->device copy per rollout would make the host wait for everything queued on the stream."""
    if target_mask is None:
        key, n_sel = (n_t, None, str(device)), n_t
    else:
        tm = torch.as_tensor(target_mask).to("cpu", torch.bool).reshape(-1)
        if tm.numel() != n_t:
            raise AlineError(f"target_mask has {tm.numel()} entries, the batch has {n_t} targets")
        key, n_sel = (n_t, bytes(tm.to(torch.uint8).tolist()), str(device)), int(tm.sum())
    hit = _slot_cache.get(key)
    if hit is not None:
        return hit, n_sel
    # explicit host tensors: callers may run under torch.set_default_device("cuda") like train_aline.py:189
    if target_mask is None:
        slots = torch.arange(n_t, dtype=I32, device="cpu")
    else:
        slots = torch.where(tm, torch.cumsum(tm.to(I32), 0, dtype=I32) - 1, torch.full((n_t,), -1, dtype=I32, device="cpu"))
    if len(_slot_cache) > 256:
        _slot_cache.clear()
    dev_slots = _slot_cache[key] = slots.to(device)
    return dev_slots, n_sel


def embed_queries(pm, query_x, row_major=False):
    """query_x [B, nq, dx] -> eq [B, d, nq]; with ``row_major`` also eq_rm [B, nq, d] (the input layout of the
    two-threads-per-row tensor-core query stream): returns (eq, eq_rm)."""
    qx = _lib.f32c(query_x)
    B, nq, _ = qx.shape
    eq = torch.empty((B, pm.dims["d"], nq), dtype=F32, device=qx.device)
    eq_rm = torch.empty((B, nq, pm.dims["d"]), dtype=F32, device=qx.device) if row_major else None
    with torch.cuda.device(qx.device):
        _lib.check(_lib.lib().aline_embed_queries_ex(pm.ref, dptr(qx), B, nq, dptr(eq), dptr(eq_rm), _st(qx.device)))
    return (eq, eq_rm) if row_major else eq


def alloc_tc_kv(pm, B, n_keys, device):
    """bf16 key / value operand blocks of the fast tensor-core query stream for up to `n_keys` keys (uint8 buffer,
    fully rewritten by every ctx_stack call), or None when the model / key count has no fast kernel."""
    n_keys = min(n_keys, pm.tc_fast_max_keys if pm.tc_blob is not None else 0)
    if n_keys < 1:
        return None
    nbytes = int(_lib.lib().aline_tc_kv_bytes(pm.ref, B, n_keys))
    return torch.empty((nbytes,), dtype=torch.uint8, device=device)


def _vp(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def ctx_stack(pm, cx, cy, n_c, target_x, slots, n_sel, kv=None, kv_slots=None, want_z=True, tc_kv=None, z_ctx=None):
    """cx [B, cap, dx], cy [B, cap(,1)] with n_c valid points -> (kv [n_layer, B, kv_slots, 2, d], z_tgt or None).
    z_ctx: optional [B, n_c, d] output buffer for the context tokens' final encodings (value head)."""
    B, cap = cx.shape[:2]
    d, nl = pm.dims["d"], pm.dims["n_layer"]
    n_td = 0 if target_x is None else target_x.shape[1]
    n_t = n_td + pm.dims["n_theta_tok"]
    if kv_slots is None:
        kv_slots = n_c + n_sel
    if kv is None:
        kv = torch.empty((nl, B, kv_slots, 2, d), dtype=F32, device=cx.device)
    z = torch.empty((B, n_t, d), dtype=F32, device=cx.device) if want_z else None
    n_keys = n_c + n_sel
    if tc_kv is not None and n_keys > pm.tc_fast_max_keys:
        tc_kv = None
    with torch.cuda.device(cx.device):
        _lib.check(_lib.lib().aline_ctx_stack_ex(pm.ref, dptr(cx), dptr(cy), B, n_c, cap, dptr(target_x), n_td,
                                                 dptr(slots, I32), dptr(kv), kv_slots, dptr(z), dptr(z_ctx), _vp(tc_kv),
                                                 n_keys if tc_kv is not None else 0, _st(cx.device)))
    return kv, z


def use_tensor_cores(pm, precision, n_keys, have_tc_kv=True):
    """precision 'bf16' -> the tcgen05 query stream when the model shape and key count have one (d = 32; the fast
    kernels up to `tc_fast_max_keys` = 160 keys need the bf16 operand blocks `tc_kv`, the general kernel holds up to
    `tc_max_keys` fp32 keys in shared memory); 'fp32' -> the FFMA kernels.  The choice is explicit, never a silent
    downgrade of 'fp32'."""
    if precision == "fp32":
        return False
    if precision != "bf16":
        raise AlineError(f"unknown precision {precision!r} (use 'fp32' or 'bf16')")
    if pm.tc_blob is None:
        return False
    return n_keys <= pm.tc_max_keys or (have_tc_kv and n_keys <= pm.tc_fast_max_keys)


def query_stream(pm, eq, alive, kv, n_keys, t_value=0.0, want_z=False, precision="fp32", tc_kv=None, eq_rm=None):
    B, d, nq = eq.shape
    logits = torch.empty((B, nq), dtype=F32, device=eq.device)
    zq = torch.empty((B, nq, d), dtype=F32, device=eq.device) if want_z else None
    with torch.cuda.device(eq.device):
        if tc_kv is not None and n_keys > pm.tc_fast_max_keys:
            tc_kv = None
        if use_tensor_cores(pm, precision, n_keys, tc_kv is not None):
            _lib.check(_lib.lib().aline_query_stream_tc_ex(pm.ref, ctypes.c_void_p(pm.tc_blob.data_ptr()), dptr(eq),
                                                           dptr(eq_rm), dptr(alive, U8), B, nq, dptr(kv), n_keys,
                                                           kv.shape[2], ctypes.c_float(t_value), dptr(logits), dptr(zq),
                                                           _vp(tc_kv), _st(eq.device)))
        else:
            _lib.check(_lib.lib().aline_query_stream(pm.ref, dptr(eq), dptr(alive, U8), B, nq, dptr(kv), n_keys,
                                                     kv.shape[2], ctypes.c_float(t_value), dptr(logits), dptr(zq),
                                                     _st(eq.device)))
    return logits, zq


def select(logits, want_zt=True):
    """Softmax + first-argmax + log-prob over a fully live candidate set: (idx [B,1] int64, log_prob [B], zt [B,nq])."""
    B, nq = logits.shape
    dev = logits.device
    idx = torch.empty((B, 1), dtype=I64, device=dev)
    lp = torch.empty((B,), dtype=F32, device=dev)
    zt = torch.empty((B, nq), dtype=F32, device=dev) if want_zt else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().aline_select(dptr(logits), None, B, nq, None, None, 0, 0, None, None, 0, 0,
                                           dptr(idx, I64), 1, dptr(lp), 1, None, dptr(zt), _st(dev)))
    return idx, lp, zt


def select_sample(logits, seed, step, want_zt=True):
    """Train-mode design choice (model/head.py:350-354) fused with the softmax: idx ~ Categorical(softmax(logits)) from
    a Philox stream keyed by (seed, rollout, step): (idx [B,1] int64, log_prob [B], zt [B,nq])."""
    B, nq = logits.shape
    dev = logits.device
    idx = torch.empty((B, 1), dtype=I64, device=dev)
    lp = torch.empty((B,), dtype=F32, device=dev)
    zt = torch.empty((B, nq), dtype=F32, device=dev) if want_zt else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().aline_select_sample(dptr(_lib.f32c(logits)), None, B, nq, None, None, 0, 0, None, None, 0, 0,
                                                  dptr(idx, I64), 1, dptr(lp), 1, None, dptr(zt),
                                                  ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), int(step) & 0x7FFFFFFF,
                                                  _st(dev)))
    return idx, lp, zt


def value_head(vh, z_ctx):
    """ValueHead.forward (model/head.py:97-111): z_ctx [B, n_c, d] -> value [B]."""
    B, n_c, d = z_ctx.shape
    w1, b1 = _lib.f32c(vh.predictor[0].weight.detach()), _lib.f32c(vh.predictor[0].bias.detach())
    w2, b2 = _lib.f32c(vh.predictor[2].weight.detach()).reshape(-1), _lib.f32c(vh.predictor[2].bias.detach())
    out = torch.empty((B,), dtype=F32, device=z_ctx.device)
    with torch.cuda.device(z_ctx.device):
        _lib.check(_lib.lib().aline_value_head(dptr(z_ctx), B, n_c, d, w1.shape[0], dptr(w1), dptr(b1), dptr(w2),
                                               dptr(b2), dptr(out), _st(z_ctx.device)))
    return out


def gmm_head(pm, z):
    """z [..., d] -> (means, stds, weights) each [..., n_comp]   (model/head.py:152-186)."""
    lead = z.shape[:-1]
    zf = _lib.f32c(z).reshape(-1, z.shape[-1])
    n, C = zf.shape[0], pm.dims["n_comp"]
    out = [torch.empty((n, C), dtype=F32, device=z.device) for _ in range(3)]
    with torch.cuda.device(z.device):
        _lib.check(_lib.lib().aline_gmm_head(pm.ref, dptr(zf), n, dptr(out[0]), dptr(out[1]), dptr(out[2]),
                                             _st(z.device)))
    return tuple(o.reshape(*lead, C) for o in out)


def gmm_head_variance(pm, z):
    """z [..., d] -> predictive variance of the GMM head's mixture [...] (head + utils/misc.py:244-279, fused)."""
    lead = z.shape[:-1]
    zf = _lib.f32c(z).reshape(-1, z.shape[-1])
    out = torch.empty((zf.shape[0],), dtype=F32, device=z.device)
    with torch.cuda.device(z.device):
        _lib.check(_lib.lib().aline_gmm_head_variance(pm.ref, dptr(zf), zf.shape[0], dptr(out), _st(z.device)))
    return out.reshape(lead)


def gmm_variance(means, stds, weights):
    """calculate_gmm_variance (utils/misc.py:244-279): means/stds [B,nq,C], weights [B,nq,C] or [B,C] -> [B,nq]."""
    means, stds, weights = _lib.f32c(means), _lib.f32c(stds), _lib.f32c(weights)
    if means.dim() != 3 or stds.shape != means.shape:
        raise AlineError(f"mixture_means / mixture_stds must both be [B, n_query, C], got {tuple(means.shape)} / "
                         f"{tuple(stds.shape)}")
    B, nq, C = means.shape
    if weights.dim() == 2 and tuple(weights.shape) == (B, C):
        tok_per_w = nq
    elif tuple(weights.shape) == (B, nq, C):
        tok_per_w = 1
    else:
        raise AlineError(f"mixture_weights must be [B, n_query, C] or [B, C], got {tuple(weights.shape)}")
    out = torch.empty((B, nq), dtype=F32, device=means.device)
    with torch.cuda.device(means.device):
        _lib.check(_lib.lib().aline_gmm_variance(dptr(means), dptr(stds), dptr(weights), B * nq, C, tok_per_w,
                                                 dptr(out), _st(means.device)))
    return out


def gmm_log_likelihood(value, means, stds, weights):
    """compute_ll (utils/eval.py:200-207): value [..., 1] (or [...]), params [..., C] -> [...].
    When autograd is recording through the mixture parameters (the training loss, train_aline.py:92-95) the same
    formula is composed from torch ops on the tensors' device, so that it carries gradient (row f2 interim)."""
    if torch.is_grad_enabled() and (means.requires_grad or stds.requires_grad or weights.requires_grad):
        import math
        v = value if value.dim() == means.dim() else value.unsqueeze(-1)
        lp = -((v - means) ** 2) / (2 * stds ** 2) - stds.log() - math.log(math.sqrt(2 * math.pi))
        return torch.logsumexp(lp + torch.log(weights), dim=-1)
    means, stds, weights = _lib.f32c(means), _lib.f32c(stds), _lib.f32c(weights)
    lead, C = means.shape[:-1], means.shape[-1]
    v = _lib.f32c(value)
    if v.dim() == means.dim():
        v = v.squeeze(-1)
    v = v.expand(lead).contiguous()
    out = torch.empty(lead, dtype=F32, device=means.device)
    n = out.numel()
    with torch.cuda.device(means.device):
        _lib.check(_lib.lib().aline_gmm_log_likelihood(dptr(v), dptr(means), dptr(stds), dptr(weights), n, C,
                                                       dptr(out), _st(means.device)))
    return out


# ---- Task.update_batch pieces (tasks/base_task.py:103-154) ----
def move_selected(pairs, idx):
    """For each (query [B,N,D], context [B,M,D]) pair: drop row idx[b] from the query (order preserving) and append
    it to the context.  One kernel per pair."""
    idx = idx.reshape(-1).to(I64).contiguous()
    out = []
    for q, c in pairs:
        q, c = _lib.f32c(q), _lib.f32c(c)
        B, N, D = q.shape
        M = c.shape[1]
        nq = torch.empty((B, N - 1, D), dtype=F32, device=q.device)
        nc = torch.empty((B, M + 1, D), dtype=F32, device=q.device)
        with torch.cuda.device(q.device):
            _lib.check(_lib.lib().aline_move_selected(dptr(q), dptr(c), dptr(idx, I64), B, N, M, D, dptr(nq), dptr(nc),
                                                      _st(q.device)))
        out.append((nq, nc))
    return out


def remove_rows(query, idx):
    q = _lib.f32c(query)
    dummy = torch.empty((q.shape[0], 0, q.shape[2]), dtype=F32, device=q.device)
    return move_selected([(q, dummy)], idx)[0][0]


def append_rows(context, new):
    return torch.cat([context, new], dim=1)


def select_append(scores, alive, qx, qy, cx, cy, n_c, idx_out, lp_out, t):
    """argmax of `scores` over the live candidates (first maximum), in-place Task.update_batch: the chosen (x, y) is
    appended at context position n_c and the candidate retired.  idx_out / lp_out [B, T]: column t is written."""
    B, nq = scores.shape
    T = idx_out.shape[1]
    with torch.cuda.device(scores.device):
        _lib.check(_lib.lib().aline_select(
            dptr(scores), dptr(alive, U8), B, nq, dptr(qx), dptr(qy), qx.shape[2], qy.shape[2] if qy.dim() == 3 else 1,
            dptr(cx), dptr(cy), n_c, cx.shape[1], ctypes.c_void_p(idx_out.data_ptr() + 8 * t), T,
            ctypes.c_void_p(lp_out.data_ptr() + 4 * t), T, None, None, _st(scores.device)))


def rollout_uncertainty(pm, context_x, context_y, query_x, query_y, target_x, T, precision="fp32"):
    """T steps of the uncertainty-sampling baseline (notebooks/eval_al.ipynb cell 1: acquisition
    "uncertainty_sampling", target_mask None): the next design is the live candidate whose GMM posterior predictive has
    the largest variance.  Same resident state as `rollout`; per step ctx_stack, query_stream (encodings only),
    fused GMM-head variance, select + append -- enqueued back to back, no host synchronisation."""
    cx0, cy0 = _lib.f32c(context_x), _lib.f32c(context_y)
    qx, qy = _lib.f32c(query_x), _lib.f32c(query_y)
    dev = qx.device
    B, n_c0, dx = cx0.shape
    nq = qx.shape[1]
    dy = cy0.shape[2] if cy0.dim() == 3 else 1
    if T > nq:
        raise AlineError(f"rollout of T={T} steps needs at least T candidates (n_query={nq})")
    cap = n_c0 + T
    cx = torch.empty((B, cap, dx), dtype=F32, device=dev)
    cy = torch.empty((B, cap, dy), dtype=F32, device=dev)
    cx[:, :n_c0] = cx0
    cy[:, :n_c0] = cy0.reshape(B, n_c0, dy)
    qy = qy.reshape(B, nq, dy)
    tx = None if target_x is None else _lib.f32c(target_x)
    n_t = (0 if tx is None else tx.shape[1]) + pm.dims["n_theta_tok"]
    slots, n_sel = target_slots(n_t, None, dev)
    kv_slots = cap + n_sel
    kv = torch.empty((pm.dims["n_layer"], B, kv_slots, 2, pm.dims["d"]), dtype=F32, device=dev)
    alive = torch.ones((B, nq), dtype=U8, device=dev)
    idx = torch.empty((B, T), dtype=I64, device=dev)
    lp = torch.empty((B, T), dtype=F32, device=dev)
    eq = embed_queries(pm, qx)
    for t in range(T):
        n_c = n_c0 + t
        ctx_stack(pm, cx, cy, n_c, tx, slots, n_sel, kv=kv, kv_slots=kv_slots, want_z=False)
        _, zq = query_stream(pm, eq, alive, kv, n_c + n_sel, want_z=True, precision=precision)
        var = gmm_head_variance(pm, zq)                 # retired candidates: rows never written, masked by `alive`
        select_append(var, alive, qx, qy, cx, cy, n_c, idx, lp, t)
    return dict(context_x=cx, context_y=cy, alive=alive, idx=idx, log_prob=None)


# ---- resident rollout (utils/eval.py:21-30) ----
def _graphs_enabled():
    import os
    return os.environ.get("ALINE_ROLLOUT_GRAPH", "1") != "0"


class _RolloutPlan:
    """Static device buffers of one rollout shape and, from the second call on, the CUDA graph of its T-step chain.

    A rollout is T x (context stack [+ fused select] -> candidate stream -> conditional robust candidate stream) + the
    last select: ~3 T dependent launches of 20-150 us kernels, i.e. host-launch-bound when enqueued one by one
    (round-1 driver run: 9.2 ms stand-alone against 6.7 ms inside the fused step).  The launches depend only on the
    shape key, so they are captured once into a CUDA graph whose nodes read / write the plan's static buffers
    (programmatic-dependent-launch edges included); a call copies its inputs in, replays the graph and returns small
    copies of the results, so the caller never aliases the plan's state.
    """

    def __init__(self, pm, B, nq, n_c0, T, dx, dy, n_td, target_mask, t_values, precision, dev):
        self.pm, self.B, self.nq, self.n_c0, self.T = pm, B, nq, n_c0, T
        self.dev = dev
        cap = n_c0 + T
        n_t = n_td + pm.dims["n_theta_tok"]
        self.slots, self.n_sel = target_slots(n_t, target_mask, dev)
        self.kv_slots = cap + self.n_sel
        d, nl = pm.dims["d"], pm.dims["n_layer"]
        e = lambda shape, dt=F32: torch.empty(shape, dtype=dt, device=dev)   # noqa: E731
        self.qx, self.qy = e((B, nq, dx)), e((B, nq, dy))
        self.cx0, self.cy0 = e((B, n_c0, dx)), e((B, n_c0, dy))
        self.tx = e((B, n_td, dx)) if n_td else None
        self.cx, self.cy = e((B, cap, dx)), e((B, cap, dy))
        self.kv = e((nl, B, self.kv_slots, 2, d))
        self.alive = e((B, nq), U8)
        self.logits = e((B, nq))
        self.idx, self.lp = e((B, T), I64), e((B, T))
        self.eq = e((B, d, nq))
        self.eq_rm = None
        self.tv = (ctypes.c_float * T)(*[float(v) for v in t_values]) if t_values is not None else None
        self.tcw, self.tckv = None, None
        if use_tensor_cores(pm, precision, cap - 1 + self.n_sel):
            self.tcw = ctypes.c_void_p(pm.tc_blob.data_ptr())
            self.tckv = alloc_tc_kv(pm, B, cap - 1 + self.n_sel, dev)
            if self.tckv is not None:
                self.eq_rm = e((B, nq, d))                 # row-major embeddings: two-threads-per-row candidate stream
        self.graph = None
        self.graph_failed = False
        self.n_kernels = 0
        self.calls = 0

    def enqueue(self):
        """All launches of the rollout on the current stream, reading / writing only the plan's buffers."""
        pm, B, nq, n_c0, T, dev = self.pm, self.B, self.nq, self.n_c0, self.T, self.dev
        self.cx[:, :n_c0] = self.cx0
        self.cy[:, :n_c0] = self.cy0
        self.alive.fill_(1)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().aline_embed_queries_ex(pm.ref, dptr(self.qx), B, nq, dptr(self.eq), dptr(self.eq_rm),
                                                         _st(dev)))
            _lib.check(_lib.lib().aline_rollout_ex(
                pm.ref, dptr(self.qx), dptr(self.qy), dptr(self.alive, U8), dptr(self.eq), dptr(self.eq_rm), dptr(self.cx),
                dptr(self.cy),
                B, nq, n_c0, n_c0 + T, dptr(self.tx), 0 if self.tx is None else self.tx.shape[1], dptr(self.slots, I32),
                self.n_sel, dptr(self.kv), self.kv_slots, dptr(self.logits), T, self.tv, dptr(self.idx, I64),
                dptr(self.lp), self.tcw, _vp(self.tckv), _st(dev)))

    def run(self, cx0, cy0, qx, qy, tx):
        self.cx0.copy_(cx0.reshape(self.cx0.shape), non_blocking=True)
        self.cy0.copy_(cy0.reshape(self.cy0.shape), non_blocking=True)
        self.qx.copy_(qx, non_blocking=True)
        self.qy.copy_(qy.reshape(self.qy.shape), non_blocking=True)
        if self.tx is not None:
            self.tx.copy_(tx, non_blocking=True)
        self.calls += 1
        if self.graph is None and self.calls >= 2 and not self.graph_failed and _graphs_enabled():
            # second call with this shape: everything lazily initialised on the C side (shared-memory opt-ins, the flag
            # ring) exists after the first, eager, call -- capture the chain
            g = torch.cuda.CUDAGraph()
            n0 = int(_lib.lib().aline_kernel_launches())
            try:
                with torch.cuda.graph(g):
                    self.enqueue()
                self.graph = g
                self.n_kernels = int(_lib.lib().aline_kernel_launches()) - n0
                _lib.graph_captured(self.n_kernels)
            except Exception as exc:   # noqa: BLE001  -- the eager chain is the same CUDA path, only launched one by one
                import warnings
                warnings.warn(f"aline_b200: CUDA-graph capture of the rollout failed ({exc}); launching eagerly")
                self.graph_failed = True
                torch.cuda.synchronize(self.dev)
        if self.graph is not None:
            self.graph.replay()
            _lib.graph_replayed(self.n_kernels)
        else:
            self.enqueue()
        return dict(context_x=self.cx.clone(), context_y=self.cy.clone(), alive=self.alive.clone(),
                    idx=self.idx.clone(), log_prob=self.lp.clone())


def rollout(pm, context_x, context_y, query_x, query_y, target_x, target_mask, T, t_values=None, precision="fp32"):
    """T greedy design steps on the device, no host synchronisation.

    Returns dict(context_x [B, n_c0+T, dx], context_y [B, n_c0+T, 1], alive [B, nq] uint8,
    idx [B, T] int64 (index within the live set at each step), log_prob [B, T]).
    """
    cx0, cy0 = _lib.f32c(context_x), _lib.f32c(context_y)
    qx, qy = _lib.f32c(query_x), _lib.f32c(query_y)
    dev = qx.device
    B, n_c0, dx = cx0.shape
    nq = qx.shape[1]
    dy = cy0.shape[2] if cy0.dim() == 3 else 1
    if T < 1 or T > nq:
        raise AlineError(f"rollout of T={T} steps needs 1 <= T <= n_query candidates (n_query={nq})")
    tx = None if target_x is None else _lib.f32c(target_x)
    n_td = 0 if tx is None else tx.shape[1]
    if target_mask is None:
        mkey = None
    else:
        mkey = bytes(torch.as_tensor(target_mask).to("cpu", torch.uint8).reshape(-1).tolist())
    key = (B, nq, n_c0, T, dx, dy, n_td, mkey, None if t_values is None else tuple(float(v) for v in t_values),
           precision, dev.index, torch.cuda.current_stream(dev).cuda_stream)
    plans = pm.__dict__.setdefault("_rollout_plans", {})
    plan = plans.pop(key, None)
    if plan is None:
        while len(plans) >= 6:                         # least recently used shape goes first
            plans.pop(next(iter(plans)))
        plan = _RolloutPlan(pm, B, nq, n_c0, T, dx, dy, n_td, target_mask, t_values, precision, dev)
    plans[key] = plan                                  # (re-)insert as most recently used
    return plan.run(cx0, cy0, qx, qy, tx)
