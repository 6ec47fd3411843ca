"""CPU oracle for the ALINE rollout + sPCE hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement (fp32, torch-CPU used purely as an
array library: matmul / exp / log / erf / softmax written out explicitly, no
``torch.nn`` modules, no reference code) of the algorithm the reference runs on
its hot path.  Every function cites the reference ``file:line`` it follows.

Who may import this: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs -- as the *checker* or the timed
CPU baseline, never as the product.  Nothing under ``aline_b200/`` imports it.

Parity pinning: the reference ships no golden vectors or known-answer tests
for this path.  The oracle is therefore pinned against
outputs of the reference itself, produced in the build container by
``tests/golden/make_golden.py`` (which imports /root/reference unchanged) and
committed as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
every function below against those fixtures.

Parameter containers are plain dicts keyed exactly like the reference
``state_dict`` (model/embedder.py, model/encoder.py, model/head.py).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch

Tensor = torch.Tensor
F32 = torch.float32


# --------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------
def _linear(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """y = x W^T + b (torch.nn.Linear semantics)."""
    return x @ w.t() + b


def _mlp2(x: Tensor, p: Dict[str, Tensor], prefix: str) -> Tensor:
    """Linear -> ReLU -> Linear, parameters ``prefix.0.*`` and ``prefix.2.*``.

    reference: model/embedder.py:47-57 (x/y embedders), model/head.py:27-33,
    model/head.py:214-224 (GMM heads).
    """
    h = torch.relu(_linear(x, p[prefix + ".0.weight"], p[prefix + ".0.bias"]))
    return _linear(h, p[prefix + ".2.weight"], p[prefix + ".2.bias"])


def _layer_norm(x: Tensor, g: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """LayerNorm over the last dim, biased variance (torch native_layer_norm)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


# --------------------------------------------------------------------------
# embedder  (reference: model/embedder.py:67-214)
# --------------------------------------------------------------------------
def embed(p: Dict[str, Tensor], batch: dict, mode: str) -> Tensor:
    """Token embeddings in the order [context | query | target].

    theta mode: model/embedder.py:128-168; mix mode: 170-214; data mode: 97-126.
    """
    cx, cy, qx = batch["context_x"], batch["context_y"], batch["query_x"]
    B = cx.shape[0]
    e_ctx = _mlp2(cx, p, "embedder.x_embedder") + _mlp2(cy, p, "embedder.y_embedder")
    e_q = _mlp2(qx, p, "embedder.x_embedder")
    parts = [e_ctx, e_q]
    if mode in ("data", "mix"):
        parts.append(_mlp2(batch["target_x"], p, "embedder.x_embedder"))
    if mode in ("theta", "mix"):
        tok = p["embedder.theta_tokens"]
        parts.append(tok.unsqueeze(0).expand(B, -1, -1))
    if mode not in ("data", "theta", "mix"):
        raise ValueError(f"Unknown embedding type: {mode}")
    return torch.cat(parts, dim=1)


# --------------------------------------------------------------------------
# encoder  (reference: model/encoder.py:48-141 + torch post-norm layer)
# --------------------------------------------------------------------------
def attention_mask(n_ctx: int, n_q: int, n_tgt: int, target_mask: Optional[Tensor]) -> Tensor:
    """Additive [N, N] mask of {0, -inf}.  reference: model/encoder.py:83-126."""
    N = n_ctx + n_q + n_tgt
    m = torch.full((N, N), float("-inf"), dtype=F32)
    m[:, :n_ctx] = 0.0
    q0, q1 = n_ctx, n_ctx + n_q
    if target_mask is not None:
        sel = torch.where(target_mask)[0] + q1
        m[q0:q1, sel] = 0.0
    else:
        m[q0:q1, q1:] = 0.0
    return m


def _num_layers(p: Dict[str, Tensor]) -> int:
    n = 0
    while f"encoder.encoder.layers.{n}.linear1.weight" in p:
        n += 1
    return n


def _layer_tail(x: Tensor, attn_out: Tensor, p: Dict[str, Tensor], pre: str) -> Tensor:
    """out-proj, residual + LN1, FF (ReLU), residual + LN2 (norm_first=False).

    reference: model/encoder.py:76-79 builds nn.TransformerEncoderLayer with
    defaults (post-norm, ReLU, eps 1e-5); SURVEY.md appendix A.1.
    """
    y = _linear(attn_out, p[pre + "self_attn.out_proj.weight"], p[pre + "self_attn.out_proj.bias"])
    h = _layer_norm(x + y, p[pre + "norm1.weight"], p[pre + "norm1.bias"])
    f = _linear(torch.relu(_linear(h, p[pre + "linear1.weight"], p[pre + "linear1.bias"])),
                p[pre + "linear2.weight"], p[pre + "linear2.bias"])
    return _layer_norm(h + f, p[pre + "norm2.weight"], p[pre + "norm2.bias"])


def _mha(xq: Tensor, xkv: Tensor, p: Dict[str, Tensor], pre: str, n_head: int,
         mask: Optional[Tensor]) -> Tensor:
    """Multi-head attention of rows xq over keys/values xkv (pre-out-proj).

    Follows torch._native_multi_head_attention: q is scaled by 1/sqrt(head_dim)
    before QK^T; a float mask is added to the scores.
    """
    W, b = p[pre + "self_attn.in_proj_weight"], p[pre + "self_attn.in_proj_bias"]
    d = W.shape[1]
    dh = d // n_head
    q = _linear(xq, W[:d], b[:d]) / math.sqrt(dh)
    k = _linear(xkv, W[d:2 * d], b[d:2 * d])
    v = _linear(xkv, W[2 * d:], b[2 * d:])
    B, Nq, _ = q.shape
    Nk = k.shape[1]
    q = q.view(B, Nq, n_head, dh).transpose(1, 2)
    k = k.view(B, Nk, n_head, dh).transpose(1, 2)
    v = v.view(B, Nk, n_head, dh).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if mask is not None:
        s = s + mask
    a = torch.softmax(s, dim=-1)
    return (a @ v).transpose(1, 2).reshape(B, Nq, d)


def encode_dense(p: Dict[str, Tensor], emb: Tensor, n_ctx: int, n_q: int, n_tgt: int,
                 n_head: int, target_mask: Optional[Tensor]) -> Tensor:
    """Dense masked encoder: exactly what the reference executes in eval mode
    (the N x N additive mask, every row scored against every column).

    reference: model/encoder.py:128-141.  Cost O(N^2); use for small cases and
    for the timed CPU baseline (it is the reference's cost model).
    """
    mask = attention_mask(n_ctx, n_q, n_tgt, target_mask)
    x = emb
    for l in range(_num_layers(p)):
        pre = f"encoder.encoder.layers.{l}."
        x = _layer_tail(x, _mha(x, x, p, pre, n_head, mask), p, pre)
    return x


def encode_structured(p: Dict[str, Tensor], emb: Tensor, n_ctx: int, n_q: int, n_tgt: int,
                      n_head: int, target_mask: Optional[Tensor]) -> Tensor:
    """Structured equivalent of :func:`encode_dense`:
    context rows attend to context; target rows attend to context; query rows
    attend to context + selected targets.  Same result up to fp32 summation
    order; O(N * n_keys).  Used for oracle runs at sizes where dense is too slow.
    """
    xc, xq, xt = emb[:, :n_ctx], emb[:, n_ctx:n_ctx + n_q], emb[:, n_ctx + n_q:]
    if target_mask is None:
        sel = torch.arange(n_tgt)
    else:
        sel = torch.where(target_mask)[0]
    for l in range(_num_layers(p)):
        pre = f"encoder.encoder.layers.{l}."
        kq = torch.cat([xc, xt[:, sel]], dim=1)
        xc_n = _layer_tail(xc, _mha(xc, xc, p, pre, n_head, None), p, pre)
        xt_n = _layer_tail(xt, _mha(xt, xc, p, pre, n_head, None), p, pre)
        xq_n = _layer_tail(xq, _mha(xq, kq, p, pre, n_head, None), p, pre)
        xc, xq, xt = xc_n, xq_n, xt_n
    return torch.cat([xc, xq, xt], dim=1)


# --------------------------------------------------------------------------
# heads  (reference: model/head.py:9-44, 115-266, 319-393)
# --------------------------------------------------------------------------
def acquisition_logits(p: Dict[str, Tensor], z_q: Tensor, t: Optional[Tensor] = None) -> Tensor:
    """Pre-softmax acquisition scores [B, n_q].  reference: model/head.py:27-31; with ``time_token`` the scalar
    ``batch.t`` is appended to every query encoding (model/head.py:342-345)."""
    if p["head.acquisition_head.predictor.0.weight"].shape[1] == z_q.shape[-1] + 1:
        if t is None:
            raise ValueError("time_token model: batch['t'] is required")
        B, n_q = z_q.shape[:2]
        time_info = torch.as_tensor(t, dtype=F32).reshape(-1)[:1].expand(B).unsqueeze(1).unsqueeze(1).expand(B, n_q, 1)
        z_q = torch.cat([z_q, time_info], dim=-1)
    return _mlp2(z_q, p, "head.acquisition_head.predictor").squeeze(-1)


def gmm_head(p: Dict[str, Tensor], z: Tensor, std_min: float = 1e-4) -> Dict[str, Tensor]:
    """GMM head: C MLPs 32->128->3; component c yields (mean, raw_std, raw_w).

    reference: model/head.py:152-186 (forward), 252-266 (_map_raw_output):
    stack -> movedim(0,-1) -> flatten(-2,-1) -> chunk(3) means output j of head
    c lands in chunk j at position c.
    """
    C = 0
    while f"head.target_head.heads.{C}.0.weight" in p:
        C += 1
    outs = torch.stack([_mlp2(z, p, f"head.target_head.heads.{c}") for c in range(C)], dim=-1)  # [B,T,3,C]
    mean, raw_std, raw_w = outs[..., 0, :], outs[..., 1, :], outs[..., 2, :]
    std = torch.nn.functional.softplus(raw_std) + std_min
    w = torch.softmax(raw_w, dim=-1)
    return {"mixture_means": mean, "mixture_stds": std, "mixture_weights": w}


def compute_ll(value: Tensor, means: Tensor, stds: Tensor, weights: Tensor) -> Tensor:
    """GMM log-likelihood.  reference: utils/eval.py:200-207 (= model/head.py:233-249)."""
    lp = -((value - means) ** 2) / (2 * stds ** 2) - stds.log() - math.log(math.sqrt(2 * math.pi))
    return torch.logsumexp(lp + torch.log(weights), dim=-1)


def gmm_variance(means: Tensor, stds: Tensor, weights: Tensor) -> Tensor:
    """Predictive variance of a Gaussian mixture per query point.  reference: utils/misc.py:244-279
    (calculate_gmm_variance): weights [B,n_q,C] or [B,C]."""
    B, n_q, C = means.shape
    w = weights.unsqueeze(1).expand(B, n_q, C) if weights.dim() == 2 else weights
    wm = torch.sum(w * means, dim=2)
    return torch.sum(w * (stds ** 2 + (means - wm.unsqueeze(2)) ** 2), dim=2)


def rollout_uncertainty(p: Dict[str, Tensor], batch: dict, T: int, mode: str, n_head: int = 4) -> dict:
    """T steps of the uncertainty-sampling baseline.  reference: notebooks/eval_al.ipynb cell 1 (acquisition
    "uncertainty_sampling": target_mask None, index = argmax of calculate_gmm_variance(posterior_out_query))."""
    batch = dict(batch)
    batch["target_mask"] = None
    idxs, gaps = [], []
    for _ in range(T):
        o = forward(p, batch, mode, n_head, dense=True, with_query_posterior=True)
        pq = o["posterior_out_query"]
        var = gmm_variance(pq["mixture_means"], pq["mixture_stds"], pq["mixture_weights"])
        idx = torch.argmax(var, dim=1, keepdim=True)
        top2 = torch.topk(var, 2, dim=1).values
        gaps.append((top2[:, 0] - top2[:, 1]) / top2[:, 0].abs().clamp_min(1e-30))
        idxs.append(idx[:, 0])
        batch = update_batch(batch, idx)
    return {"batch": batch, "idx": torch.stack(idxs, 1), "rel_gap": torch.stack(gaps, 1)}


def forward(p: Dict[str, Tensor], batch: dict, mode: str, n_head: int = 4,
            dense: bool = True, with_query_posterior: bool = True) -> dict:
    """Aline.forward in eval mode.  reference: model/base.py:32-50, model/head.py:319-393.

    Returns dict(idx [B,1] int64, log_prob [B], zt [B,n_q], logits [B,n_q],
    posterior_out{...}, posterior_out_query{...}, encoding [B,N,d]).
    """
    n_ctx = batch["context_x"].shape[1]
    n_q = batch["query_x"].shape[1]
    n_tgt = batch["target_all"].shape[1]
    tm = batch.get("target_mask", None)
    emb = embed(p, batch, mode)
    enc = (encode_dense if dense else encode_structured)(p, emb, n_ctx, n_q, n_tgt, n_head, tm)
    z_q, z_t = enc[:, n_ctx:n_ctx + n_q], enc[:, n_ctx + n_q:]
    logits = acquisition_logits(p, z_q, batch.get("t", None))
    zt = torch.softmax(logits, dim=-1)
    pmax, idx = torch.max(zt, -1)              # first maximal index on ties (head.py:355-358)
    out = {
        "idx": idx.unsqueeze(1), "log_prob": torch.log(pmax), "zt": zt, "logits": logits,
        "posterior_out": gmm_head(p, z_t), "encoding": enc,
    }
    if with_query_posterior:
        out["posterior_out_query"] = gmm_head(p, z_q)
    if "head.value_head.predictor.0.weight" in p:
        # ValueHead: mean over the context tokens of the 2-layer MLP.  reference: model/head.py:97-111, 367-370
        out["value"] = _mlp2(enc[:, :n_ctx], p, "head.value_head.predictor").squeeze(-1).mean(1)
    return out


def update_batch(batch: dict, idx: Tensor) -> dict:
    """Move the chosen (x, y) from the query set to the end of the context,
    removing it from the query set order-preservingly.

    reference: tasks/base_task.py:103-154.
    """
    B = idx.shape[0]
    out = dict(batch)
    ar = torch.arange(B)
    for kq, kc in (("query_x", "context_x"), ("query_y", "context_y")):
        q = batch[kq]
        D = q.shape[-1]
        nxt = q[ar, idx[:, 0]].unsqueeze(1)
        keep = torch.ones(q.shape[:2], dtype=torch.bool)
        keep[ar, idx[:, 0]] = False
        out[kq] = q[keep].view(B, -1, D)
        out[kc] = torch.cat([batch[kc], nxt], dim=1)
    return out


def rollout(p: Dict[str, Tensor], batch: dict, T: int, mode: str, n_head: int = 4,
            dense: bool = True) -> dict:
    """T greedy design steps.  reference: utils/eval.py:9-39 (get_traces),
    without the task's ``sample_batch`` (the batch is an input here)."""
    idxs, lps = [], []
    for _ in range(T):
        o = forward(p, batch, mode, n_head, dense=dense, with_query_posterior=False)
        idxs.append(o["idx"][:, 0])
        lps.append(o["log_prob"])
        batch = update_batch(batch, o["idx"])
    return {"batch": batch, "idx": torch.stack(idxs, 1), "log_prob": torch.stack(lps, 1)}


# --------------------------------------------------------------------------
# simulator likelihoods
# --------------------------------------------------------------------------
_LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


def normal_log_prob(v: Tensor, loc: Tensor, scale: Tensor) -> Tensor:
    """torch.distributions.Normal.log_prob written out (torch 2.x source)."""
    return -((v - loc) ** 2) / (2 * scale ** 2) - scale.log() - _LOG_SQRT_2PI


def location_log_likelihood(y: Tensor, xi: Tensor, theta: Tensor, noise_scale: float = 0.5,
                            base_signal: float = 0.1, max_signal: float = 1e-4) -> Tensor:
    """reference: tasks/location_finding.py:110-130 (total_density), 149-164.

    y [..,1], xi [..,D], theta [..,K,D] -> [..,1].
    """
    sq = ((xi.unsqueeze(-2) - theta) ** 2).sum(-1)                     # [.., K]
    inv = (max_signal + sq).pow(-1)
    signal = torch.log(base_signal + inv.sum(-1, keepdim=True))        # [.., 1]
    return normal_log_prob(y, signal, torch.tensor(noise_scale, dtype=F32))


def _sigmoid_inv(y: Tensor) -> Tensor:
    """torch SigmoidTransform._inverse."""
    fi = torch.finfo(y.dtype)
    y = y.clamp(min=fi.tiny, max=1.0 - fi.eps)
    return y.log() - (-y).log1p()


def ces_log_likelihood(y: Tensor, xi: Tensor, theta: Tensor, noise_scale: float = 0.005,
                       epsilon: float = 2.0 ** -22, check: bool = True) -> Tensor:
    """reference: tasks/ces.py:169-210 + distributions/censored_sigmoid_normal.py:47-86.

    y [1,B,1], xi [1,B,6], theta [L,B,5] = [rho, a1, a2, a3, log u] -> [L,B,1].
    Raises ArithmeticError on NaN/inf exactly like the reference (line 83-84).
    """
    softplus = torch.nn.functional.softplus
    rho, alpha, log_u = theta[..., 0:1], theta[..., 1:4], theta[..., 4:5]
    u = torch.exp(log_u)
    x = torch.clamp(xi, min=0.01, max=100.0)
    b1, b2 = x[..., :3], x[..., 3:]
    U1 = torch.sum(alpha * b1 ** rho, dim=-1, keepdim=True) ** (1.0 / rho)
    U2 = torch.sum(alpha * b2 ** rho, dim=-1, keepdim=True) ** (1.0 / rho)
    mu = (U1 - U2) * u
    sigma = (1 + torch.norm(b1 - b2, dim=-1, p=2, keepdim=True)) * noise_scale * u
    lo = torch.tensor(epsilon, dtype=F32)
    hi = torch.tensor(1 - epsilon, dtype=F32)
    mu, sigma = torch.broadcast_tensors(mu, sigma)
    value = y.expand_as(mu)

    def base_log_prob(v):
        t = _sigmoid_inv(v)
        return normal_log_prob(t, mu, sigma) - (-softplus(-t) - softplus(t))

    def base_cdf(v):
        t = _sigmoid_inv(v)
        return 0.5 * (1 + torch.erf((t - mu) * sigma.reciprocal() / math.sqrt(2)))

    lp = base_log_prob(value)
    upper_cdf = 1.0 - base_cdf(hi)
    lower_cdf = base_cdf(lo)
    crit = 2 * torch.finfo(F32).tiny
    z_up = (_sigmoid_inv(hi) - mu) / sigma
    z_lo = (_sigmoid_inv(lo) - mu) / sigma
    asym_up = base_log_prob(hi) - (crit + z_up.abs()).log()
    asym_lo = base_log_prob(lo) - (crit + z_lo.abs()).log()
    up_log = torch.where(upper_cdf < crit, asym_up, upper_cdf.log())
    lo_log = torch.where(lower_cdf < crit, asym_lo, lower_cdf.log())
    lp = torch.where(value == hi, up_log, lp)
    lp = torch.where(value == lo, lo_log, lp)
    lp = torch.where(value > hi, torch.tensor(float("-inf")), lp)
    lp = torch.where(value < lo, torch.tensor(float("-inf")), lp)
    if check and (torch.isnan(lp).any() or torch.isinf(lp).any()):
        raise ArithmeticError("NaN in log_prob")
    return lp


def psychometric_log_likelihood(y: Tensor, x: Tensor, theta: Tensor) -> Tensor:
    """reference: tasks/psychometric.py:107-134, 178-195.

    theta [.., 4] = [alpha, beta, gamma, lambda]; x, y [.., 1] -> [.., 1].
    (The reference indexes theta[:, k, :] on [B,4,1]; this restatement takes the
    parameter axis last so that it also covers [L,B,4] -- see SURVEY.md section 7:
    the reference cannot run sPCE on this task, parity is at the formula level.)
    """
    a, b, g, lam = theta[..., 0:1], theta[..., 1:2], theta[..., 2:3], theta[..., 3:4]
    z = (x - a) / b
    Fz = 1 - torch.exp(-10 ** z)
    pr = lam * g + (1 - lam) * Fz
    return y * torch.log(pr + 1e-10) + (1 - y) * torch.log(1 - pr + 1e-10)


# --------------------------------------------------------------------------
# sPCE / sNMC  (reference: loss/eig.py:154-209, utils/eval.py:43-80)
# --------------------------------------------------------------------------
def spce_history(log_lik: Callable[[Tensor, Tensor, Tensor], Tensor], y: Tensor, x: Tensor,
                 thetas: Tensor, stepwise: bool = True) -> dict:
    """Step-wise sPCE / sNMC bounds over a history.

    y [B,T,Dy], x [B,T,Dx] (unnormalised designs), thetas [L+1,B,(K,)D] with
    row 0 = theta_0.  Returns pce, nmc [B,T] (or [B] if not stepwise) and the
    final accumulated seq_logprobs [L+1,B].

    reference: EIGStepLoss.step/forward loss/eig.py:174-209; constants
    log(L+1), log(L) utils/eval.py:77-78.
    """
    L = thetas.shape[0] - 1
    B, T = y.shape[:2]
    seq = torch.zeros((L + 1, B), dtype=F32)
    pces, nmcs = [], []
    for t in range(T):
        lp = log_lik(y[:, t].unsqueeze(0), x[:, t].unsqueeze(0), thetas).squeeze(-1)
        seq = seq + lp
        pces.append(seq.logsumexp(0) - seq[0])
        nmcs.append(seq[1:].logsumexp(0) - seq[0])
    if stepwise:
        pce, nmc = torch.stack(pces, -1), torch.stack(nmcs, -1)
    else:
        pce, nmc = pces[-1], nmcs[-1]
    pce = torch.log(torch.tensor(L + 1)) - pce
    nmc = torch.log(torch.tensor(L)) - nmc
    return {"pce": pce, "nmc": nmc, "seq_logprobs": seq}


def pce_loss_whole_history(log_lik, y: Tensor, x: Tensor, thetas: Tensor, nmc: bool = False) -> Tensor:
    """PCELoss / NMCLoss with reduction=None.  reference: loss/eig.py:22-48, 68-86, 133-151."""
    T = x.shape[1]
    th = thetas.unsqueeze(2)
    th = th.expand(-1, -1, T, *thetas.shape[2:])
    lp = log_lik(y.unsqueeze(0), x.unsqueeze(0), th).sum(dim=(-2, -1))
    return (lp[1:] if nmc else lp).logsumexp(0) - lp[0]


def combine_partials(m: Tensor, s: Tensor, lp0: Tensor, L: int) -> dict:
    """Combine R per-rank partial logsumexp terms.

    m, s: [R, ...] running max and sum of exp(lp - m) over each rank's slice of
    the contrastive rows 1..L;  lp0: [...] the theta_0 row.
    """
    M = m.max(0).values
    S = (s * torch.exp(m - M)).sum(0)
    nmc = math.log(L) - (M + torch.log(S) - lp0)
    M2 = torch.maximum(M, lp0)
    pce = math.log(L + 1) - (M2 + torch.log(S * torch.exp(M - M2) + torch.exp(lp0 - M2)) - lp0)
    return {"pce": pce, "nmc": nmc}


# --------------------------------------------------------------------------
# GP prior draw  (reference: tasks/gaussian_process.py:194-317, 366-417)
# --------------------------------------------------------------------------
KERNEL_TYPES = ("rbf", "matern12", "matern32", "matern52")


def gp_kernel_matrix(x: Tensor, lengthscales: Tensor, scale: Tensor, kernel_type: str) -> Tensor:
    """x [N,dx], lengthscales [dx], scale scalar -> K [N,N] (no jitter)."""
    sq = ((x.unsqueeze(1) - x.unsqueeze(0)) ** 2 / (lengthscales ** 2).view(1, 1, -1)).sum(-1)
    if kernel_type == "rbf":
        return scale * torch.exp(-0.5 * sq)
    d = torch.sqrt(sq)
    if kernel_type == "matern12":
        return scale * torch.exp(-d)
    if kernel_type == "matern32":
        s3 = torch.sqrt(torch.tensor(3.0))
        return scale * (1 + s3 * d) * torch.exp(-s3 * d)
    if kernel_type == "matern52":
        s5 = torch.sqrt(torch.tensor(5.0))
        return scale * (1 + s5 * d + (5.0 / 3.0) * d ** 2) * torch.exp(-s5 * d)
    raise ValueError(f"Unknown kernel type: {kernel_type}")


def cholesky_lower(K: Tensor) -> Tensor:
    """Unblocked left-looking (Cholesky-Crout, column by column) factorisation in
    fp32.  The reference calls torch.linalg.cholesky (LAPACK potrf; un-vendored
    dependency torch==2.6.0) at gaussian_process.py:403; this restates the
    published algorithm.  Summation order differs from potrf, so L matches to
    ~1e-4 relative on these cond ~1e6 matrices, not bitwise.
    """
    n = K.shape[0]
    Lm = torch.zeros_like(K)
    for j in range(n):
        c = K[j:, j] - Lm[j:, :j] @ Lm[j, :j]
        d = torch.sqrt(c[0])
        Lm[j, j] = d
        Lm[j + 1:, j] = c[1:] / d
    return Lm


def gp_draw(x: Tensor, lengthscales: Tensor, scale: Tensor, kernel_type: str, z: Tensor,
            eps: Tensor, jitter: float = 1e-5, noise_scale: float = 0.01,
            lapack: bool = True) -> dict:
    """One GP prior draw with the normal variates given explicitly.

    reference: gaussian_process.py:391-415: K + jitter I -> L -> f = L z ->
    y = f + noise_scale * eps.
    """
    n = x.shape[0]
    K = gp_kernel_matrix(x, lengthscales, scale, kernel_type) + jitter * torch.eye(n)
    Lm = torch.linalg.cholesky(K) if lapack else cholesky_lower(K)
    f = Lm @ z
    return {"K": K, "L": Lm, "f": f, "y": f + noise_scale * eps}


# --------------------------------------------------------------------------
# target mask helpers  (reference: utils/target_mask.py:5-125)
# --------------------------------------------------------------------------
def create_target_mask_deterministic(mask_type: str, embedding_type: str, n_target_data: int,
                                     n_target_theta: int, predefined_masks=None, mask_index=None,
                                     attend_to=None) -> Tensor:
    """The branches of create_target_mask that do not draw random numbers."""
    n = n_target_data + n_target_theta
    m = torch.zeros(n, dtype=torch.bool)
    if mask_type == "all":
        m[:] = True
    elif mask_type == "none":
        pass
    elif mask_type == "predefined":
        for i, v in enumerate(predefined_masks[mask_index]):
            if i < n and v:
                m[i] = True
    elif mask_type == "split" and embedding_type == "mix":
        if attend_to == "data":
            m[:n_target_data] = True
        else:
            m[n_target_data:] = True
    return m


# ---------------------------------------------------------------------------------------
# Device-side prior draws (row f1): Philox4x32-10 restated from the published algorithm
# (Salmon et al., SC'11), and the maps from its 32-bit outputs to the tasks' priors.
# ---------------------------------------------------------------------------------------
def philox4x32_10(counter, key):
    """counter: uint32 array [..., 4]; key: (k0, k1).  Returns uint32 [..., 4]."""
    import numpy as np
    c = [np.asarray(counter[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), \
        np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        n0 = ((p1 >> np.uint64(32)) ^ c[1] ^ k0) & MASK
        n2 = ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & MASK
        c = [n0, p1 & MASK, n2, p0 & MASK]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return np.stack(c, axis=-1).astype(np.uint32)


def prior_uniforms(seed: int, row_offset: int, n_rows: int, B: int, n_calls: int):
    """u [n_rows, B, 4 * n_calls] float32 in [0, 1): the stream layout of csrc/philox.cuh
    (key = seed, counter = (row_lo, row_hi, b, call))."""
    import numpy as np
    rows = (np.arange(n_rows, dtype=np.uint64) + np.uint64(row_offset))[:, None, None]
    b = np.arange(B, dtype=np.uint64)[None, :, None]
    call = np.arange(n_calls, dtype=np.uint64)[None, None, :]
    ctr = np.stack(np.broadcast_arrays(rows & np.uint64(0xFFFFFFFF), rows >> np.uint64(32), b, call), axis=-1)
    x = philox4x32_10(ctr.astype(np.uint32), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    u = (x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    return u.reshape(n_rows, B, 4 * n_calls)


def prior_box(seed: int, row_offset: int, n_rows: int, B: int, lo, hi) -> Tensor:
    """theta_i = lo_i + u_i (hi_i - lo_i).  reference priors: tasks/location_finding.py:85-98, tasks/psychometric.py:70-89."""
    import numpy as np
    lo, hi = np.asarray(lo, dtype=np.float32), np.asarray(hi, dtype=np.float32)
    d = lo.shape[0]
    u = prior_uniforms(seed, row_offset, n_rows, B, (d + 3) // 4)[..., :d]
    return torch.from_numpy((lo + u.astype(np.float64) * (hi - lo).astype(np.float64)).astype(np.float32))


# ---- Task.sample_batch from Philox streams (restates csrc/prior.cu: sample_batch_kernel) ----
BATCH_CALL = 0x80000000


def batch_uniforms(seed: int, g, col, call: int):
    """u [..., 4] float32 of the counters (g lo, g hi, col, 0x80000000 | call); g, col broadcastable uint64 arrays."""
    import numpy as np
    g, col = np.asarray(g, dtype=np.uint64), np.asarray(col, dtype=np.uint64)
    cc = np.uint64(BATCH_CALL | call)
    ctr = np.stack(np.broadcast_arrays(g & np.uint64(0xFFFFFFFF), g >> np.uint64(32), col, cc), axis=-1)
    x = philox4x32_10(ctr.astype(np.uint32), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    return (x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def _box_muller(u0, u1):
    import numpy as np
    return np.sqrt(-2.0 * np.log(1.0 - u0.astype(np.float64))) * np.cos(2.0 * np.pi * u1.astype(np.float64))


def sample_batch_philox(task: str, seed: int, batch_offset: int, B: int, n_pts: int, dim_x: int, x_lo: float,
                        x_hi: float, design_scale: float, lo=None, hi=None, K: int = 1, noise_scale: float = 0.5,
                        base_signal: float = 0.1, max_signal: float = 1e-4, epsilon: float = 2.0 ** -22,
                        u_mu: float = 1.0, u_sigma: float = 3.0, theta_override=None) -> dict:
    """theta [B, dim_theta], x [B, n_pts, dim_x] (float32, bit-level restatement of the draws), y [B, n_pts, 1] and
    the Bernoulli margin |u - p| (psychometric), with the simulators evaluated in float64.  `theta_override`
    (float32 [B, dim_theta]) simulates from given thetas instead of the restated draw; CES also returns `y_tol`, the
    fp32 conditioning of each outcome (pow round-off amplified by u / rho).
    reference simulators: tasks/location_finding.py:110-147,167-192; tasks/ces.py:129-167,213-234 +
    distributions/censored_sigmoid_normal.py:43-45; tasks/psychometric.py:107-176,197-222."""
    import numpy as np
    g = np.arange(B, dtype=np.uint64) + np.uint64(batch_offset)
    if task == "ces":
        u0, u1 = batch_uniforms(seed, g, 0, 0).astype(np.float64), batch_uniforms(seed, g, 0, 1)
        e = -np.log(1.0 - u0[:, 1:4])
        theta = np.concatenate([0.01 + 0.99 * u0[:, :1], e / e.sum(1, keepdims=True),
                                (u_mu + u_sigma * _box_muller(u1[:, 0], u1[:, 1]))[:, None]], axis=1)
    else:
        lo32, hi32 = np.asarray(lo, dtype=np.float32), np.asarray(hi, dtype=np.float32)
        d = lo32.shape[0]
        u = np.concatenate([batch_uniforms(seed, g, 0, c) for c in range((d + 3) // 4)], axis=-1)[:, :d]
        theta = lo32 + u.astype(np.float64) * (hi32 - lo32).astype(np.float64)
    theta32 = theta.astype(np.float32)
    if theta_override is not None:
        theta32 = np.asarray(theta_override, dtype=np.float32).reshape(theta32.shape)
    th = theta32.astype(np.float64)                        # the kernel simulates from the float32 draw
    col = (np.arange(n_pts, dtype=np.uint64) + np.uint64(1))[None, :]
    ux = np.concatenate([batch_uniforms(seed, g[:, None], col, c) for c in range((dim_x + 3) // 4)], axis=-1)[..., :dim_x]
    span = np.float32(np.float32(x_hi) - np.float32(x_lo))
    x32 = (np.float64(np.float32(x_lo)) + ux.astype(np.float64) * np.float64(span)).astype(np.float32)
    xi = (x32 * np.float32(design_scale)).astype(np.float64)
    un = batch_uniforms(seed, g[:, None], col, 2)
    z = _box_muller(un[..., 0], un[..., 1])
    margin = y_tol = None
    if task == "location":
        thk = th.reshape(B, 1, K, dim_x)
        sq = ((xi[:, :, None, :] - thk) ** 2).sum(-1)
        y = np.log(base_signal + (1.0 / (max_signal + sq)).sum(-1)) + noise_scale * z
    elif task == "ces":
        v = np.clip(xi, 0.01, 100.0)
        rho, alpha, uu = th[:, None, 0:1], th[:, None, 1:4], np.exp(th[:, None, 4])
        U1 = (alpha * v[..., :3] ** rho).sum(-1) ** (1.0 / rho[..., 0])
        U2 = (alpha * v[..., 3:] ** rho).sum(-1) ** (1.0 / rho[..., 0])
        mu = (U1 - U2) * uu
        sigma = (1.0 + np.sqrt(((v[..., :3] - v[..., 3:]) ** 2).sum(-1))) * noise_scale * uu
        with np.errstate(over="ignore"):
            s = 1.0 / (1.0 + np.exp(-(mu + sigma * z)))
        # |dy| <= s(1-s) |d(mu + sigma z)|, with a few fp32 ulps on every pow and the 1/rho power of their sum
        y_tol = s * (1.0 - s) * (2e-6 * (U1 + U2) * uu * (1.0 + 1.0 / rho[..., 0]) + 1e-5 * (1.0 + np.abs(mu + sigma * z))) + 3e-7
        s = np.clip(s, np.finfo(np.float32).tiny, 1.0 - np.finfo(np.float32).eps)
        y = np.clip(s, epsilon, 1.0 - epsilon)
    elif task == "psychometric":
        zz = (xi[..., 0] - th[:, None, 0]) / th[:, None, 1]
        with np.errstate(over="ignore"):
            F = 1.0 - np.exp(-(10.0 ** zz))
        p = th[:, None, 3] * th[:, None, 2] + (1.0 - th[:, None, 3]) * F
        y = (un[..., 2].astype(np.float64) < p).astype(np.float64)
        margin = np.abs(un[..., 2].astype(np.float64) - p)
    else:
        raise ValueError(task)
    out = dict(theta=torch.from_numpy(theta32), x=torch.from_numpy(x32),
               y=torch.from_numpy(y.astype(np.float32)).unsqueeze(-1))
    if margin is not None:
        out["margin"] = torch.from_numpy(margin)
    if y_tol is not None:
        out["y_tol"] = torch.from_numpy(y_tol).unsqueeze(-1)
    return out
