"""``Aline(embedder, encoder, head)`` -- same constructor, ``forward(batch)`` signature, returned ``AttrDict``
layout and state-dict keys as the reference ``model/base.py`` (13-50); the forward runs on the sm_100a kernels.

``forward(batch)`` (eval / no_grad, which is how utils/eval.py and the notebooks call it) returns

    design_out          idx [B,1] int64, log_prob [B], zt [B,n_query]        (model/head.py:355-358, 383-387)
    posterior_out       mixture_means / mixture_stds / mixture_weights [B,n_target,C]   (model/head.py:365)
    posterior_out_query same over the candidate queries [B,n_query,C]         (model/head.py:366)

``posterior_out_query`` costs as much as the rest of the forward and is only read by the uncertainty-sampling
baseline, so by default it is computed on first access (``query_posterior = "lazy"``; ``"eager"`` / ``"off"``).

``rollout(batch, T)`` is the resident replacement for the ``for t in range(T): forward; update_batch`` loop of
``get_traces`` (utils/eval.py:21-30): T design steps on the device without host synchronisation.

Grad-enabled calls (``train_aline.py:80-110``: ``model.train()`` with autograd recording) take the differentiable
torch-op composition of ``model/grad_path.py`` on the same parameters and CUDA device -- an explicit interim for row f2
(no backward kernels yet); the train-mode design choice is drawn by the fused Philox kernel (``aline_select_sample``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib, rollout as _ro
from ..attrdict import AttrDict
from . import grad_path
from .packing import PackedModel


class _Outputs(AttrDict):
    """AttrDict whose ``posterior_out_query`` entry is materialised on first access."""

    def __missing__(self, key):
        if key == "posterior_out_query":
            fn = object.__getattribute__(self, "_lazy_fn")
            if fn is not None:
                val = fn()
                self[key] = val
                return val
        raise KeyError(key)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def get(self, k, default=None):
        try:
            return self[k]
        except KeyError:
            return default


def _mixture(means, stds, weights):
    return AttrDict(mixture_means=means, mixture_stds=stds, mixture_weights=weights)


class Aline(nn.Module):
    def __init__(self, embedder, encoder, head) -> None:
        super().__init__()
        self.embedder = embedder
        self.encoder = encoder
        self.head = head
        self.query_posterior = "lazy"
        # "bf16": candidate-query stream on the tcgen05 tensor cores (bf16 operands, fp32 accumulation; log-probs
        # within 1e-3 of the fp32 reference); "fp32": everything on the FFMA pipe (log-probs within 1e-5)
        self.precision = "bf16"
        self._packed = None
        self._packed_key = None
        self.differentiable = None          # None: autograd forward in train() mode only (see _needs_grad)
        # train-mode sampling (model/head.py:350-354): Philox key drawn once from torch's CPU generator (so
        # torch.manual_seed controls it) + a per-call counter; `design_sampler` (zt -> idx [B]) overrides it (tests)
        self._sample_seed = None
        self._sample_calls = 0
        self.design_sampler = None

    # ---- parameter blob, re-packed whenever a parameter changes (in-place updates bump ._version) ----
    def packed(self) -> PackedModel:
        params = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._packed is None or key != self._packed_key:
            dev = params[0].device
            if dev.type != "cuda":
                raise _lib.AlineError("Aline parameters are on %s: move the model to a CUDA device "
                                      "(the B200 path has no CPU fallback)" % dev)
            sd = {k: v for k, v in self.state_dict().items()}
            self._packed = PackedModel(sd, self.encoder.n_head, self.head.target_head.std_min, dev)
            self._packed_key = key
        return self._packed

    def _needs_grad(self):
        """Differentiable torch-op forward (row f2) or the kernels?  The kernels have no backward, so the composition is
        used where a backward pass can follow: autograd recording AND `train()` mode with trainable parameters
        (train_aline.py:55,80-132).  In `eval()` mode the kernels run even without `torch.no_grad()` -- that is how the
        reference's notebooks call the model (eval_al.ipynb, eval_psychometric.ipynb: `model.eval()`, then plain
        `model(batch)`), and their outputs are only read, never differentiated.  `model.differentiable = True / False`
        overrides the rule (True: gradients in eval mode, e.g. sensitivity analyses)."""
        if not torch.is_grad_enabled():
            return False
        if self.differentiable is not None:
            return bool(self.differentiable)
        return self.training and any(p.requires_grad for p in self.parameters())

    def _next_sample_key(self):
        if self._sample_seed is None:
            self._sample_seed = int(torch.randint(0, 2 ** 62, (1,), device="cpu").item())
        self._sample_calls += 1
        return self._sample_seed, self._sample_calls

    def _sample_design(self, zt):
        """idx [B] ~ Categorical(zt) on the device (fused Philox inverse-CDF kernel), or the user's `design_sampler`."""
        if self.design_sampler is not None:
            return self.design_sampler(zt)
        seed, step = self._next_sample_key()
        idx, _, _ = _ro.select_sample(torch.log(zt.detach().float()), seed, step, want_zt=False)
        return idx[:, 0]

    def _forward_with_grad(self, batch):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.AlineError(f"Aline parameters are on {dev}: move the model to a CUDA device (the B200 path has no "
                                  "CPU fallback; model.grad_path.forward_torch is the device-agnostic composition)")
        return grad_path.forward_torch(self, batch, sampler=self._sample_design if self.training else None,
                                       with_query_posterior=self.query_posterior != "off")

    @staticmethod
    def _field(batch, name):
        if isinstance(batch, dict):
            return batch.get(name, None)
        return getattr(batch, name, None)

    def forward(self, batch):
        if self._needs_grad():
            return self._forward_with_grad(batch)
        with torch.no_grad():
            pm = self.packed()
            cx, cy = _lib.f32c(batch.context_x), _lib.f32c(batch.context_y)
            qx = _lib.f32c(batch.query_x)
            B, n_c = cx.shape[:2]
            mode = self.embedder.embedding_type
            tx = self._field(batch, "target_x") if mode in ("data", "mix") else None
            if mode in ("data", "mix") and tx is None:
                raise ValueError(f"embedding_type '{mode}' needs batch.target_x")
            tx = None if tx is None else _lib.f32c(tx)
            n_t = (0 if tx is None else tx.shape[1]) + pm.dims["n_theta_tok"]
            if batch.target_all.shape[1] != n_t:
                raise ValueError(f"batch.target_all has {batch.target_all.shape[1]} targets, the embedder produces {n_t}")
            slots, n_sel = _ro.target_slots(n_t, self._field(batch, "target_mask"), cx.device)
            t_value = 0.0
            if self.head.time_token:
                t_value = float(torch.as_tensor(batch.t).reshape(-1)[0])
            tc_kv = None
            if _ro.use_tensor_cores(pm, self.precision, n_c + n_sel) and n_c + n_sel <= pm.tc_fast_max_keys:
                tc_kv = _ro.alloc_tc_kv(pm, B, n_c + n_sel, cx.device)
            eq, eq_rm = _ro.embed_queries(pm, qx, row_major=True) if tc_kv is not None else (_ro.embed_queries(pm, qx), None)
            z_ctx = None
            if isinstance(self.head.value_head, nn.Module):
                z_ctx = torch.empty((B, n_c, pm.dims["d"]), dtype=torch.float32, device=cx.device)
            kv, z_t = _ro.ctx_stack(pm, cx, cy, n_c, tx, slots, n_sel, tc_kv=tc_kv, z_ctx=z_ctx)
            want_zq = self.query_posterior in ("lazy", "eager")
            logits, zq = _ro.query_stream(pm, eq, None, kv, n_c + n_sel, t_value, want_z=want_zq,
                                          precision=self.precision, tc_kv=tc_kv, eq_rm=eq_rm)
            if self.training:       # no_grad + train(): Categorical sample (model/head.py:350-354), fused Philox kernel
                if self.design_sampler is not None:
                    zt = torch.softmax(logits, -1)
                    idx = self.design_sampler(zt).reshape(-1, 1).to(torch.int64)
                    log_prob = torch.log((zt / zt.sum(-1, keepdim=True)).clamp(1.1920929e-07, 1 - 1.1920929e-07)
                                         .gather(1, idx))[:, 0]
                else:
                    seed, step = self._next_sample_key()
                    idx, log_prob, zt = _ro.select_sample(logits, seed, step)
            else:
                idx, log_prob, zt = _ro.select(logits)
            out = _Outputs(posterior_out=_mixture(*_ro.gmm_head(pm, z_t)),
                           design_out=AttrDict(idx=idx, log_prob=log_prob, zt=zt))
            if z_ctx is not None:                   # model/head.py:367-381
                out["value"] = _ro.value_head(self.head.value_head, z_ctx)
            lazy = None
            if self.query_posterior == "eager":
                out["posterior_out_query"] = _mixture(*_ro.gmm_head(pm, zq))
            elif self.query_posterior == "lazy":
                lazy = lambda: _mixture(*_ro.gmm_head(pm, zq))  # noqa: E731
            object.__setattr__(out, "_lazy_fn", lazy)
            return out

    @torch.no_grad()
    def rollout(self, batch, T, time_token=False, acquisition="aae"):
        """T greedy design steps (eval semantics: argmax), resident on the device.  Returns the batch with
        ``context_x / context_y`` extended by the T chosen (design, outcome) pairs, plus ``design_idx [B,T]``
        (index within the live candidate set, the reference's ``design_out.idx`` at each step),
        ``design_log_prob [B,T]`` and ``query_alive [B,n_query]``; ``query_x / query_y`` keep their original
        storage (retired candidates are flagged, not compacted)."""
        pm = self.packed()
        mode = self.embedder.embedding_type
        tx = self._field(batch, "target_x") if mode in ("data", "mix") else None
        if acquisition == "uncertainty_sampling":     # baseline of notebooks/eval_al.ipynb cell 1 (target_mask = None)
            r = _ro.rollout_uncertainty(pm, batch.context_x, batch.context_y, batch.query_x, batch.query_y, tx, T,
                                        precision=self.precision)
            batch.context_x, batch.context_y = r["context_x"], r["context_y"]
            batch.query_alive, batch.design_idx, batch.design_log_prob = r["alive"], r["idx"], None
            return batch
        if acquisition != "aae":
            raise ValueError(f"unknown acquisition {acquisition!r} ('aae' or 'uncertainty_sampling')")
        tv = [(T - t) / T for t in range(T)] if (time_token and self.head.time_token) else None   # utils/eval.py:26
        r = _ro.rollout(pm, batch.context_x, batch.context_y, batch.query_x, batch.query_y, tx,
                        self._field(batch, "target_mask"), T, tv, precision=self.precision)
        batch.context_x, batch.context_y = r["context_x"], r["context_y"]
        batch.query_alive = r["alive"]
        batch.design_idx, batch.design_log_prob = r["idx"], r["log_prob"]
        return batch
