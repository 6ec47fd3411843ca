"""The algebra behind the folded operands of the tensor-core candidate stream (csrc/query_tc3.cu FOLD, csrc/query_fast.cuh,
DESIGN.md 4f), checked in float64 against the attention block the reference computes (model/encoder.py:128-141 =
nn.TransformerEncoderLayer self-attention restricted to a candidate row: q from the candidate, k / v from the context):

    S_h = c (x Wq_h^T + bq_h) (K_h - K_0h)^T  =  x K'_h^T + bias_h,   K'_h = c (K_h - K_0h) Wq_h,  bias_h = c (K_h - K_0h) bq_h
    y   = concat_h(softmax_h V_h) Wo^T + bo   =  sum_h Pn_h V'_h + bo, V'_h = V_h Wo[:, head h]^T,  Pn_h = 2^S_h / rowsum

with c = log2(e) / sqrt(head_dim): scores relative to key 0 (the softmax is shift-invariant), base-2 exponentials, masked /
unused key slots pushed to -200 (probability 2^-200 = 0 in fp32)."""
import math

import torch


def _reference(x, K, V, Wq, bq, Wo, bo, H):
    n, D = x.shape
    hd = D // H
    q = (x @ Wq.T + bq) / math.sqrt(hd)
    out = []
    for h in range(H):
        sl = slice(hd * h, hd * (h + 1))
        p = torch.softmax(q[:, sl] @ K[:, sl].T, dim=-1)
        out.append(p @ V[:, sl])
    return torch.cat(out, dim=-1) @ Wo.T + bo


def _folded(x, K, V, Wq, bq, Wo, bo, H, n_pad):
    n, D = x.shape
    hd = D // H
    c = math.log2(math.e) / math.sqrt(hd)
    nk = K.shape[0]
    y = bo.expand(n, D).clone()
    for h in range(H):
        sl = slice(hd * h, hd * (h + 1))
        Kd = K[:, sl] - K[0, sl]
        Kp = torch.zeros(n_pad, D, dtype=x.dtype)
        bias = torch.full((n_pad,), -200.0, dtype=x.dtype)         # slots without a key
        Kp[:nk] = c * Kd @ Wq[sl, :]
        bias[:nk] = c * Kd @ bq[sl]
        Vp = torch.zeros(n_pad, D, dtype=x.dtype)
        Vp[:nk] = V[:, sl] @ Wo[:, sl].T
        P = torch.exp2(x @ Kp.T + bias)
        P = torch.where(P < 2.0 ** -149, torch.zeros_like(P), P)   # what fp32 keeps of 2^-200
        y = y + (P / P.sum(-1, keepdim=True)) @ Vp
    return y


def test_fold_identity():
    g = torch.Generator().manual_seed(3)
    D, H = 32, 4
    for nk, n_pad in ((1, 8), (3, 8), (16, 16), (17, 24), (32, 32), (37, 48)):
        x = torch.randn(50, D, generator=g, dtype=torch.float64)
        K, V = torch.randn(nk, D, generator=g, dtype=torch.float64), torch.randn(nk, D, generator=g, dtype=torch.float64)
        Wq, Wo = torch.randn(D, D, generator=g, dtype=torch.float64) / 4, torch.randn(D, D, generator=g, dtype=torch.float64) / 4
        bq, bo = torch.randn(D, generator=g, dtype=torch.float64), torch.randn(D, generator=g, dtype=torch.float64)
        ref = _reference(x, K, V, Wq, bq, Wo, bo, H)
        got = _folded(x, K, V, Wq, bq, Wo, bo, H, n_pad)
        assert (ref - got).abs().max().item() < 1e-11 * max(1.0, ref.abs().max().item())
