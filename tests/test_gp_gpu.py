"""GPU parity of the batched GP prior-draw kernel against the golden fixture (reference GPTask.generate_gp_data
replayed with explicit normal variates) and the oracle."""
import numpy as np
import pytest
import torch

from oracle import aline_oracle as O
from _util import load_golden

pytestmark = pytest.mark.gpu


def test_gp_draws_golden():
    from aline_b200.gp import gp_sample, kernel_matrix
    g = {k: torch.from_numpy(v) for k, v in load_golden("gp_draws").items()}
    x, theta = g["x"].cuda(), g["theta"].cuda()
    y, L, K = gp_sample(x, theta[:, :2, 0], theta[:, 2, 0], g["ktype"].cuda(), g["z"].cuda(), g["eps"].cuda(),
                        jitter=1e-5, noise_scale=0.01, return_factors=True)
    B = x.shape[0]
    for b in range(B):
        Kr, Lr = g["K"][b].double(), g["L"][b].double()
        assert (K[b].cpu().double() - Kr).abs().max().item() < 2e-6                       # kernel matrix entries
        L2 = L[b].cpu().double()
        assert ((L2 @ L2.T - Kr).norm() / Kr.norm()).item() < 1e-6                         # reconstruction
        f_ref = Lr @ g["z"][b].double()
        f = L2 @ g["z"][b].double()
        assert ((f - f_ref).norm() / f_ref.norm()).item() < 2e-3                           # SURVEY.md section 7 gate
        y_ref = g["y"][b, :, 0].double()
        assert ((y[b].cpu().double() - y_ref).norm() / y_ref.norm()).item() < 2e-3
    for i, kt in enumerate(O.KERNEL_TYPES):
        Km = kernel_matrix(x[0], x[0], theta[0, :2, 0], theta[0, 2, 0], i)
        assert (Km.cpu() - g["K_" + kt]).abs().max().item() < 2e-6, kt


def test_gp_task_sample_batch_cfg4():
    """cfg4 shapes: 2-D, mix embedding, 1 + 200 + 100 points -> N = 301 (lower triangle packed in shared memory)."""
    from aline_b200.tasks import GPTask
    torch.manual_seed(4)
    task = GPTask(dim_x=2, embedding_type="mix", n_context_init=1, n_query_init=200, n_target_theta=3,
                  n_target_data=100, design_scale=5)
    torch.set_default_device("cuda")
    try:
        batch = task.sample_batch(200)
    finally:
        torch.set_default_device("cpu")
    assert batch.context_x.shape == (200, 1, 2) and batch.query_y.shape == (200, 200, 1)
    assert batch.target_all.shape == (200, 103, 1) and batch.target_x.shape == (200, 100, 2)
    ys = torch.cat([batch.context_y, batch.query_y, batch.target_y], 1)
    assert torch.isfinite(ys).all()
    # prior variance of a draw = output scale (+ jitter + noise^2): check the batch-level moment
    scale = batch.target_theta[:, 2, 0]
    ratio = (ys.squeeze(-1).var(dim=1) / scale).mean().item()
    assert 0.3 < ratio < 1.3


def test_gp_large_matrix_uses_global_scratch():
    from aline_b200.gp import gp_sample
    torch.manual_seed(5)
    B, N = 3, 400
    x = (torch.rand(B, N, 2) * 10 - 5).cuda()
    ls = torch.full((B, 2), 1.5).cuda()
    sc = torch.tensor([0.5, 1.0, 0.2]).cuda()
    kt = torch.tensor([0, 2, 3], dtype=torch.int32).cuda()
    z, eps = torch.randn(B, N).cuda(), torch.randn(B, N).cuda()
    y, L, K = gp_sample(x, ls, sc, kt, z, eps, return_factors=True)
    for b in range(B):
        Kd, Ld = K[b].double(), L[b].double()
        assert ((Ld @ Ld.T - Kd).norm() / Kd.norm()).item() < 1e-6
        assert torch.allclose(y[b].double(), Ld @ z[b].double() + 0.01 * eps[b].double(), atol=1e-4)


def test_gp_not_positive_definite_raises():
    from aline_b200.gp import gp_sample
    x = torch.zeros(1, 8, 1).cuda()           # identical points, no jitter -> singular
    with pytest.raises(RuntimeError):
        gp_sample(x, torch.ones(1, 1).cuda(), torch.ones(1).cuda(), torch.zeros(1, dtype=torch.int32).cuda(),
                  torch.randn(1, 8).cuda(), torch.randn(1, 8).cuda(), jitter=-1e-3)
