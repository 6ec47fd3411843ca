"""Row f1: device-side prior draws (Philox streams) -- bit-level agreement of the sampler kernel with the oracle's
restatement of Philox4x32-10, exact agreement of the in-kernel generation with the materialised draws, independence of
the sharding, and statistical agreement with the torch-generator path."""
import math

import numpy as np
import pytest
import torch

from oracle import aline_oracle as O

pytestmark = pytest.mark.gpu


def test_box_prior_matches_philox_oracle():
    from aline_b200.prior import sample_theta_device
    from aline_b200.tasks import HiddenLocation, PsychometricTask
    task = HiddenLocation(design_scale=1)
    th = sample_theta_device(task, 50, 7, seed=0x1234_5678_9ABC, row_offset=3)
    assert th.shape == (50, 7, 1, 2)
    ref = O.prior_box(0x1234_5678_9ABC, 3, 50, 7, [0.0, 0.0], [1.0, 1.0])
    assert torch.equal(th.reshape(50, 7, 2).cpu(), ref)                         # unit box: bit exact
    th2 = sample_theta_device(HiddenLocation(K=2, n_target_theta=4, design_scale=1), 9, 5, seed=77)
    assert th2.shape == (9, 5, 2, 2)
    assert torch.equal(th2.reshape(9, 5, 4).cpu(), O.prior_box(77, 0, 9, 5, [0.0] * 4, [1.0] * 4))
    psy = sample_theta_device(PsychometricTask(), 1000, 3, seed=5, row_offset=2 ** 33)     # 64-bit row counter
    refp = O.prior_box(5, 2 ** 33, 1000, 3, [-3, 0.1, 0.1, 0.0], [3, 2, 0.9, 0.5])
    assert psy.shape == (1000, 3, 4) and (psy.cpu() - refp).abs().max().item() < 1e-6
    # rows of a shard = the same rows of the unsharded stream
    a = sample_theta_device(task, 100, 4, seed=9)
    b = sample_theta_device(task, 40, 4, seed=9, row_offset=60)
    assert torch.equal(a[60:], b)


def test_ces_prior_moments():
    from aline_b200.prior import sample_theta_device
    from aline_b200.tasks import CESTask
    th = sample_theta_device(CESTask(), 200_000, 2, seed=11).reshape(-1, 5)
    rho, alpha, logu = th[:, 0], th[:, 1:4], th[:, 4]
    assert rho.min().item() >= 0.01 and rho.max().item() <= 1.0 and abs(rho.mean().item() - 0.505) < 3e-3
    assert (alpha >= 0).all() and (alpha.sum(1) - 1).abs().max().item() < 1e-5
    assert (alpha.mean(0) - 1 / 3).abs().max().item() < 3e-3
    assert abs(alpha[:, 0].var().item() - 1 / 18) < 2e-3                        # Dirichlet(1,1,1): var = 2/(9*4)
    assert abs(logu.mean().item() - 1.0) < 0.03 and abs(logu.std().item() - 3.0) < 0.03
    ref = CESTask().sample_theta((200_000, 2)).reshape(-1, 5)                    # torch-generator draws of the same prior
    for j in range(5):
        qs = torch.tensor([0.1, 0.5, 0.9])
        assert (torch.quantile(th[:, j].cpu(), qs) - torch.quantile(ref[:, j], qs)).abs().max().item() < 0.06


@pytest.mark.parametrize("B,T,L", [(8, 35, 20001), (200, 13, 3000), (3, 2, 50), (16, 40, 5000)])    # T = 40: multi-pass kernel
def test_in_kernel_draws_equal_materialised_draws(B, T, L):
    from aline_b200 import spce
    from aline_b200.prior import sample_theta_device, spce_history_device_prior
    from aline_b200.tasks import HiddenLocation
    torch.manual_seed(B + T)
    task = HiddenLocation(design_scale=1)
    theta0 = torch.rand(B, 1, 2, device="cuda")
    x = torch.rand(B, T, 2, device="cuda")
    d2 = ((x - theta0) ** 2).sum(-1, keepdim=True)
    y = torch.log(0.1 + 1.0 / (1e-4 + d2)) + 0.5 * torch.randn(B, T, 1, device="cuda")
    seed = 424242
    m, s, lp0 = spce_history_device_prior(task, y, x, theta0, L, seed)
    pl, nl = spce.lse_combine(m, s, lp0)
    rows = sample_theta_device(task, L + 1, B, seed)
    rows[0] = theta0
    seq = torch.zeros(L + 1, B, device="cuda")
    m2, s2, lp02 = spce.spce_history(task.log_likelihood, y, x, rows, seq=seq)
    pl2, nl2 = spce.lse_combine(m2, s2, lp02)
    assert torch.allclose(pl, pl2, rtol=1e-5, atol=1e-5) and torch.allclose(nl, nl2, rtol=1e-5, atol=1e-5)
    # two shards with global row offsets combine to the unsharded bound
    ms, ss = [], []
    for r in range(2):
        lo, hi = spce.shard_rows(L, r, 2)
        mr, sr, lp0r = spce_history_device_prior(task, y, x, theta0, hi - lo, seed, row_offset=lo)
        assert torch.allclose(lp0r, lp0)
        ms.append(mr)
        ss.append(sr)
    pl3, nl3 = spce.lse_combine(torch.stack(ms), torch.stack(ss), lp0)
    assert torch.allclose(pl3, pl, rtol=1e-5, atol=1e-5) and torch.allclose(nl3, nl, rtol=1e-5, atol=1e-5)


def test_device_prior_is_statistically_the_torch_prior():
    from aline_b200.tasks import HiddenLocation, PsychometricTask
    from aline_b200.utils.eval import compute_EIG_from_history
    torch.manual_seed(21)
    task = HiddenLocation(design_scale=1)
    B, T, L = 64, 20, 200_000
    theta0 = torch.rand(B, 1, 2, device="cuda")
    x = torch.rand(B, T, 2, device="cuda")
    d2 = ((x - theta0) ** 2).sum(-1, keepdim=True)
    y = torch.log(0.1 + 1.0 / (1e-4 + d2)) + 0.5 * torch.randn(B, T, 1, device="cuda")
    p_t, n_t = compute_EIG_from_history(task, theta0, x, y, L=L, batch_size=B, stepwise=True)
    p_d, n_d = compute_EIG_from_history(task, theta0, x, y, L=L, batch_size=B, stepwise=True, prior="device", seed=3)
    p_e, _ = compute_EIG_from_history(task, theta0, x, y, L=L, batch_size=B, stepwise=True, prior="device", seed=4)
    mc = (p_d - p_e).abs().mean().item()                      # Monte-Carlo spread between two device seeds
    assert (p_d - p_t).abs().mean().item() < 3 * mc + 0.02
    assert abs(p_d.mean().item() - p_t.mean().item()) < 0.05 and abs(n_d.mean().item() - n_t.mean().item()) < 0.05
    # a task without an in-kernel generator: draws are materialised on the device, same API
    psy = PsychometricTask()
    th0 = psy.sample_theta((6,)).cuda()
    xs = (torch.rand(6, 5, 1) * 10 - 5).cuda()
    ys = torch.bernoulli(torch.full((6, 5, 1), 0.5)).cuda()
    pp, nn = compute_EIG_from_history(psy, th0, xs, ys, L=5000, batch_size=6, stepwise=True, prior="device", seed=1)
    assert pp.shape == (6, 5) and torch.isfinite(pp).all() and torch.isfinite(nn).all()
    with pytest.raises(ValueError):
        compute_EIG_from_history(task, theta0, x, y, L=10, batch_size=B, prior="cpu")


def _oracle_batch(task, seed, off, B, theta=None):
    name = type(task).__name__
    n = task.n_context_init + task.n_query_init
    if name == "HiddenLocation":
        return O.sample_batch_philox("location", seed, off, B, n, task.dim_x, 0.0, 1.0, float(task.design_scale),
                                     lo=[0.0] * (task.K * task.dim_x), hi=[1.0] * (task.K * task.dim_x), K=task.K,
                                     noise_scale=float(task.noise_scale), base_signal=task.base_signal,
                                     max_signal=task.max_signal)
    if name == "CESTask":
        return O.sample_batch_philox("ces", seed, off, B, n, 6, 0.0, float(task.design_scale), 1.0, theta_override=theta,
                                     noise_scale=float(task.noise_scale), epsilon=float(task.epsilon))
    return O.sample_batch_philox("psychometric", seed, off, B, n, 1, -float(task.design_scale), float(task.design_scale),
                                 1.0, lo=[-3, 0.1, 0.1, 0.0], hi=[3, 2, 0.9, 0.5])


def test_sample_batch_kernel_matches_philox_oracle():
    """aline_sample_batch against the oracle's restatement (same Philox counters, simulators in float64): designs and
    box-prior thetas bit for bit, outcomes within fp32 round-off, Bernoulli outcomes equal away from u == p."""
    from aline_b200.prior import sample_batch_device
    from aline_b200.tasks import CESTask, HiddenLocation, PsychometricTask
    for task in (HiddenLocation(n_query_init=300, design_scale=1),
                 HiddenLocation(K=2, n_target_theta=4, n_query_init=50, design_scale=1)):
        got = sample_batch_device(task, 33, seed=0xABCDEF0123, batch_offset=2 ** 32 - 5)
        ref = _oracle_batch(task, 0xABCDEF0123, 2 ** 32 - 5, 33)
        x = torch.cat([got.context_x, got.query_x], 1).cpu()
        y = torch.cat([got.context_y, got.query_y], 1).cpu()
        assert got.context_x.shape == (33, 1, 2) and got.query_y.shape == (33, task.n_query_init, 1)
        assert got.target_all.shape == (33, task.n_target_theta, 1) and got.target_theta is got.target_all
        assert torch.equal(x, ref["x"]) and torch.equal(got.target_all.reshape(33, -1).cpu(), ref["theta"])
        assert (y - ref["y"]).abs().max().item() < 2e-5
    ces = CESTask(n_context_init=1, n_query_init=400)
    got = sample_batch_device(ces, 64, seed=99, batch_offset=7)
    ref = _oracle_batch(ces, 99, 7, 64)
    x = torch.cat([got.context_x, got.query_x], 1).cpu()
    y = torch.cat([got.context_y, got.query_y], 1).cpu()
    assert (x - ref["x"]).abs().max().item() < 1e-5 and got.n_theta == 5
    assert (got.target_all.reshape(64, 5).cpu() - ref["theta"]).abs().max().item() < 2e-5
    eps = 2.0 ** -22
    assert y.min().item() >= eps and y.max().item() <= 1 - eps
    # outcomes: simulate the oracle from the kernel's own float32 thetas; fp32 pow round-off is amplified by u / rho
    # where the response is not saturated, so the gate is the per-outcome conditioning the oracle reports
    ref = _oracle_batch(ces, 99, 7, 64, theta=got.target_all.reshape(64, 5).cpu().numpy())
    excess = ((y - ref["y"]).abs().double() - ref["y_tol"]).reshape(-1)
    assert (excess <= 0).float().mean().item() > 0.999 and excess.max().item() < 1e-3, excess.max().item()
    assert 0.02 < ((y > eps) & (y < 1 - eps)).float().mean().item() < 0.9       # censored and interior outcomes both occur
    psy = PsychometricTask(n_context_init=1, n_query_init=500)
    got = sample_batch_device(psy, 40, seed=5)
    ref = _oracle_batch(psy, 5, 0, 40)
    x = torch.cat([got.context_x, got.query_x], 1).cpu()
    y = torch.cat([got.context_y, got.query_y], 1).cpu()
    assert (x - ref["x"]).abs().max().item() < 1e-6 and got.target_all.shape == (40, 4, 1)
    clear = ref["margin"] > 1e-5
    assert torch.equal(y.squeeze(-1)[clear], ref["y"].squeeze(-1)[clear]) and clear.float().mean().item() > 0.999
    assert set(y.unique().tolist()) <= {0.0, 1.0}
    # global rollout index: a later slice of a larger batch is the same draw
    a = sample_batch_device(psy, 16, seed=8)
    b = sample_batch_device(psy, 6, seed=8, batch_offset=10)
    assert torch.equal(a.query_y[10:], b.query_y) and torch.equal(a.target_all[10:], b.target_all)


def test_resident_eval_boed_matches_the_torch_sampled_one():
    """eval_boed(prior='device'): batches simulated and contrastive thetas drawn on the device; the bounds agree with
    the torch-generator run within Monte-Carlo error and repeat exactly for the same seed."""
    from aline_b200.model import Aline, Embedder, Encoder, OutputHead
    from aline_b200.tasks import HiddenLocation
    from aline_b200.utils.eval import eval_boed, get_traces
    torch.manual_seed(123)
    model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128)).cuda().eval()
    task = HiddenLocation(n_query_init=60, design_scale=1)
    th0, x, y = get_traces(model, task, T=6, batch_size=9, sampler="device", seed=4, batch_offset=18)
    assert th0.shape == (9, 1, 2) and x.shape == (9, 7, 2) and y.shape == (9, 7, 1) and x.is_cuda
    kw = dict(T=8, L=4000, M=256, batch_size=64, stepwise=True, verbose=False)
    with torch.device("cuda"):                    # the torch-generator path samples on the default device (train_aline.py:189)
        r_t = eval_boed(model, task, **kw)
    r_d = eval_boed(model, task, prior="device", seed=11, **kw)
    r_d2 = eval_boed(model, task, prior="device", seed=11, **kw)
    assert torch.equal(r_d.pce_mean, r_d2.pce_mean) and torch.equal(r_d.nmc_mean, r_d2.nmc_mean)
    se = torch.sqrt(r_t.pce_err ** 2 + r_d.pce_err ** 2)
    assert ((r_t.pce_mean - r_d.pce_mean).abs() < 5 * se + 0.02).all(), (r_t.pce_mean, r_d.pce_mean, se)
    with pytest.raises(ValueError):
        eval_boed(model, task, prior="device", **kw)
