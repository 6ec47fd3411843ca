"""``Encoder`` -- same constructor and state-dict keys as the reference ``model/encoder.py`` (56-81):
``encoder.layers.{l}.self_attn.{in_proj_weight,in_proj_bias,out_proj.*}``, ``linear1/2``, ``norm1/2``.

The parameters are held by a stock ``nn.TransformerEncoder`` (so reference checkpoints load unchanged); the
forward never runs it.  The reference's ``[N, N]`` additive mask (``create_mask``, 83-126) is not materialised on
the hot path -- its structure is compiled into the kernels -- but ``create_mask`` is kept for callers / tests."""
from __future__ import annotations

import torch
import torch.nn as nn


class Encoder(nn.Module):
    def __init__(self, dim_embedding, dim_feedforward, n_head, dropout, num_layers):
        super().__init__()
        if dropout not in (0, 0.0):
            raise ValueError("aline_b200 implements the inference path: dropout must be 0.0 (reference config)")
        self.dim_embedding = dim_embedding
        self.dim_feedforward = dim_feedforward
        self.n_head = n_head
        self.num_layers = num_layers
        layer = nn.TransformerEncoderLayer(dim_embedding, n_head, dim_feedforward, dropout, batch_first=True)
        self.encoder = nn.TransformerEncoder(layer, num_layers, enable_nested_tensor=False)

    def create_mask(self, batch):
        """The additive {0, -inf} mask the reference builds (model/encoder.py:83-126)."""
        n_c, n_q, n_t = batch.context_x.shape[1], batch.query_x.shape[1], batch.target_all.shape[1]
        n = n_c + n_q + n_t
        mask = torch.full((n, n), float("-inf"), device=batch.context_x.device)
        mask[:, :n_c] = 0.0
        tm = batch.get("target_mask", None) if hasattr(batch, "get") else getattr(batch, "target_mask", None)
        if tm is not None:
            sel = torch.where(tm)[0].to(mask.device) + n_c + n_q
            mask[n_c:n_c + n_q, sel] = 0.0
        else:
            mask[n_c:n_c + n_q, n_c + n_q:] = 0.0
        return mask

    def forward(self, batch, embeddings):
        """Stand-alone call (reference 128-141; the hot path never takes it): the structured equivalent of the masked
        encoder -- context <- context, target <- context, query <- context + selected targets -- from torch ops on the
        module's device (``model/grad_path.py``); same [B, N, d] layout as the input."""
        from . import grad_path
        n_c, n_q = batch.context_x.shape[1], batch.query_x.shape[1]
        n_t = embeddings.shape[1] - n_c - n_q
        tm = batch.get("target_mask", None) if hasattr(batch, "get") else getattr(batch, "target_mask", None)
        dev = embeddings.device
        sel = torch.arange(n_t, device=dev) if tm is None else torch.where(torch.as_tensor(tm).to(dev).reshape(-1))[0]
        z = grad_path.encode(self, embeddings[:, :n_c], embeddings[:, n_c:n_c + n_q], embeddings[:, n_c + n_q:], sel)
        return torch.cat(z, dim=1)
