"""Shared helpers for the test-suite (fixture loading; oracle import)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def state_dict_of(g, device="cpu"):
    return {k[3:]: torch.from_numpy(v).to(device) for k, v in g.items() if k.startswith("sd/")}


def step_batch(g, t, device="cpu"):
    """The batch dict the reference's model.forward saw at teacher-forced step t."""
    b = {k: torch.from_numpy(g[f"step{t}/{k}"]).to(device) for k in ("context_x", "context_y", "query_x", "query_y")}
    b["target_all"] = torch.from_numpy(g["target_all"]).to(device)
    if "target_x" in g:
        b["target_x"] = torch.from_numpy(g["target_x"]).to(device)
    if "target_mask" in g:
        b["target_mask"] = torch.from_numpy(g["target_mask"]).to(device)
    if f"step{t}/t" in g:
        b["t"] = torch.from_numpy(g[f"step{t}/t"])          # time token (utils/eval.py:25-26), a host scalar
    return b


def mode_of(g):
    if "target_x" in g:
        return "mix" if "sd/embedder.theta_tokens" in g else "data"
    return "theta"


def n_head_of(sd):
    return sd["encoder.encoder.layers.0.linear2.weight"].shape[0] // 8


def reinforce_loss(log_probs, nlls_for_query, nlls_for_prediction, gamma=1.0, alpha=1.0):
    """The training objective of the reference's caller, restated (train_aline.py:112-132): rewards
    R_t = gamma^t * clamp(nll_q[t-1] - nll_q[t], 0) (detached), normalised over the batch per step;
    design_loss = -mean(log_probs[:, :-1] * R); predict_loss = mean over steps and rollouts of the target NLL;
    loss = alpha * design_loss + predict_loss.  log_probs [B, T]; the two lists hold T tensors [B]."""
    T = log_probs.shape[1]
    R = torch.stack([(gamma ** t) * torch.clamp(nlls_for_query[t - 1] - nlls_for_query[t], min=0.0).detach()
                     for t in range(1, T)], 1)
    R = (R - R.mean(dim=0, keepdim=True)) / (R.std(dim=0, keepdim=True) + 1e-9)
    design_loss = -torch.mean(log_probs[:, :-1] * R)
    predict_loss = torch.mean(torch.stack(nlls_for_prediction))
    return design_loss * alpha + predict_loss, design_loss, predict_loss


def train_inner_loop(model, update_batch, compute_ll, select_targets_by_mask, batch, T, mix_n_theta=0, gamma=0.99,
                     alpha=1.0, time_token=False, tensor=torch.tensor):
    """The T-step experiment of train_aline.py:80-110 on any implementation of the boundary (the reference's classes
    when the fixture is generated, aline_b200's in the tests): forward (train mode: sampled design), update_batch,
    compute_ll on the targets, masked mean as the reward signal.  Returns (loss, design_loss, predict_loss, idx [T,B])."""
    B = batch.context_x.shape[0]
    log_probs, nq_, np_, idxs = [], [], [], []
    for t in range(T):
        if time_token:
            batch.t = tensor([t / T])
        pred = model.forward(batch)
        d, post = pred.design_out, pred.posterior_out
        batch = update_batch(batch, d.idx)
        log_probs.append(d.log_prob)
        idxs.append(d.idx.reshape(B).detach().clone())
        ll = compute_ll(batch.target_all, post.mixture_means, post.mixture_stds, post.mixture_weights)
        tm = batch.get("target_mask", None) if hasattr(batch, "get") else getattr(batch, "target_mask", None)
        mll = select_targets_by_mask(ll, tm) if tm is not None else ll
        if mix_n_theta:
            np_.append(-(ll[:, :-mix_n_theta].mean(-1) + ll[:, -mix_n_theta:].mean(-1)))
        else:
            np_.append(-ll.mean(-1))
        nq_.append(-mll.mean(-1))
    loss, dl, pl = reinforce_loss(torch.stack(log_probs, 1), nq_, np_, gamma, alpha)
    return loss, dl, pl, torch.stack(idxs, 0)


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).abs() / b.abs().clamp_min(1e-12)).max().item()


def abs_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return (a - b).abs().max().item()
