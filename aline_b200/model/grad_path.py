"""Grad-enabled ``Aline.forward`` (row f2 of the scope table, minimum slice): the forward of ``train_aline.py:80-110``.

The sm_100a kernels implement the inference path and have no backward.  When autograd is recording, the forward is
composed from torch ops ON THE MODEL'S OWN PARAMETERS AND DEVICE (a CUDA device in ``Aline.forward``; there is no CPU
product path) -- explicitly an interim until backward kernels exist ("hard parts").  It is still not
the reference's arithmetic: the ``[N, N]`` mask is never built and nothing is scored against masked columns -- context
rows attend to context, target rows to context, candidate rows to context + the selected targets, which
is what the reference's ``_sa_block`` computes with a dense masked softmax (model/encoder.py:7-46).  The train-mode design
choice (``Categorical(zt).sample()``, model/head.py:350-354) is drawn by the fused Philox kernel
(``aline_select_sample``); its log-prob is recomputed with torch ops so that it carries gradient.

reference: model/embedder.py:97-214, model/encoder.py:7-46,128-141, model/head.py:152-186,252-266,319-393.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from ..attrdict import AttrDict

_EPS = torch.finfo(torch.float32).eps


def _field(batch, name):
    if isinstance(batch, dict):
        return batch.get(name, None)
    return getattr(batch, name, None)


def embed(embedder, batch):
    """Token embeddings (context, query, target) -- model/embedder.py:97-214."""
    mode = embedder.embedding_type
    e_ctx = embedder.x_embedder(batch.context_x) + embedder.y_embedder(batch.context_y)
    e_q = embedder.x_embedder(batch.query_x)
    parts = []
    if mode in ("data", "mix"):
        tx = _field(batch, "target_x")
        if tx is None:
            raise ValueError(f"embedding_type '{mode}' needs batch.target_x")
        parts.append(embedder.x_embedder(tx))
    if mode in ("theta", "mix"):
        parts.append(embedder.theta_tokens.unsqueeze(0).expand(e_ctx.shape[0], -1, -1))
    return e_ctx, e_q, torch.cat(parts, dim=1)


def _attend(layer, n_head, xq, xkv):
    """Multi-head attention of rows xq over keys / values xkv, before the output projection (torch's
    _native_multi_head_attention: q scaled by 1/sqrt(head_dim))."""
    W, b = layer.self_attn.in_proj_weight, layer.self_attn.in_proj_bias
    d = W.shape[1]
    dh = d // n_head
    q = F.linear(xq, W[:d], b[:d]) * (1.0 / math.sqrt(dh))
    k = F.linear(xkv, W[d:2 * d], b[d:2 * d])
    v = F.linear(xkv, W[2 * d:], b[2 * d:])
    B, nq_, nk = q.shape[0], q.shape[1], k.shape[1]
    q = q.view(B, nq_, n_head, dh).transpose(1, 2)
    k = k.view(B, nk, n_head, dh).transpose(1, 2)
    v = v.view(B, nk, n_head, dh).transpose(1, 2)
    a = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    return (a @ v).transpose(1, 2).reshape(B, nq_, d)


def _tail(layer, x, attn):
    """out-proj, residual + LN1, feed-forward (ReLU), residual + LN2 (post-norm layer, dropout 0)."""
    h = layer.norm1(x + layer.self_attn.out_proj(attn))
    return layer.norm2(h + layer.linear2(F.relu(layer.linear1(h))))


def encode(encoder, e_ctx, e_q, e_t, selected):
    """Structured encoder: per layer, with the layer-INPUT activations, context <- context, target <- context,
    query <- context + selected targets.  ``selected``: index tensor of the targets the queries attend to."""
    xc, xq, xt = e_ctx, e_q, e_t
    for layer in encoder.encoder.layers:
        kq = torch.cat([xc, xt.index_select(1, selected)], dim=1) if selected.numel() else xc
        xc_n = _tail(layer, xc, _attend(layer, encoder.n_head, xc, xc))
        xt_n = _tail(layer, xt, _attend(layer, encoder.n_head, xt, xc))
        xq_n = _tail(layer, xq, _attend(layer, encoder.n_head, xq, kq))
        xc, xq, xt = xc_n, xq_n, xt_n
    return xc, xq, xt


def gmm_head(target_head, z):
    """GMMTargetHead.forward + _map_raw_output (model/head.py:152-186, 252-266): output j of component head c lands in
    chunk j at position c."""
    outs = torch.stack([h(z) for h in target_head.heads], dim=-1)            # [B, T, 3, C]
    mean, raw_std, raw_w = outs[..., 0, :], outs[..., 1, :], outs[..., 2, :]
    return AttrDict(mixture_means=mean, mixture_stds=F.softplus(raw_std) + target_head.std_min,
                    mixture_weights=torch.softmax(raw_w, dim=-1))


def forward_torch(model, batch, sampler=None, with_query_posterior=True):
    """``Aline.forward`` composed from torch ops (differentiable).  ``sampler(zt) -> idx [B]`` supplies the train-mode
    design choice (default: ``torch.multinomial``; ``Aline.forward`` passes the fused Philox kernel); in eval mode the
    argmax is taken (model/head.py:355-358)."""
    emb, enc, head = model.embedder, model.encoder, model.head
    e_ctx, e_q, e_t = embed(emb, batch)
    n_t = e_t.shape[1]
    if batch.target_all.shape[1] != n_t:
        raise ValueError(f"batch.target_all has {batch.target_all.shape[1]} targets, the embedder produces {n_t}")
    tm = _field(batch, "target_mask")
    dev = e_ctx.device
    if tm is None:
        selected = torch.arange(n_t, device=dev)
    else:
        selected = torch.where(torch.as_tensor(tm).to(dev).reshape(-1))[0]
    z_c, z_q, z_t = encode(enc, e_ctx, e_q, e_t, selected)
    return heads(head, batch, z_c, z_q, z_t, model.training, sampler, with_query_posterior)


def heads(head, batch, z_c, z_q, z_t, training, sampler=None, with_query_posterior=True):
    """OutputHead.forward (model/head.py:319-393) on the context / query / target encodings."""
    dev = z_q.device
    zq_in = z_q
    if head.time_token:
        t = torch.as_tensor(batch.t, dtype=z_q.dtype, device=dev).reshape(-1)[:1]
        zq_in = torch.cat([z_q, t.expand(z_q.shape[0]).unsqueeze(1).unsqueeze(1).expand(-1, z_q.shape[1], 1)], dim=-1)
    zt = head.acquisition_head.predictor(zq_in)                               # probabilities [B, n_q]
    if training:
        with torch.no_grad():
            idx = (sampler(zt) if sampler is not None else torch.multinomial(zt, 1)[:, 0]).reshape(-1).to(torch.int64)
        probs = zt / zt.sum(-1, keepdim=True)                                 # Categorical(probs=zt) normalises ...
        log_prob = torch.log(probs.clamp(min=_EPS, max=1 - _EPS)).gather(1, idx.unsqueeze(1))[:, 0]   # ... and clamps
    else:
        p, idx = torch.max(zt, -1)
        log_prob = torch.log(p)
    out = AttrDict(posterior_out=gmm_head(head.target_head, z_t),
                   design_out=AttrDict(idx=idx.unsqueeze(1), log_prob=log_prob, zt=zt))
    if with_query_posterior:
        out.posterior_out_query = gmm_head(head.target_head, z_q)
    if isinstance(head.value_head, torch.nn.Module):
        out.value = head.value_head.predictor(z_c).squeeze(-1).mean(1)        # model/head.py:97-111
    return out
