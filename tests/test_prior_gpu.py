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


@pytest.mark.parametrize("B,T,L", [(8, 35, 20001), (200, 13, 3000), (3, 2, 50)])
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
