"""CPU-only host logic of the device-side prior draws (row f1): prior descriptors of the tasks, the oracle's Philox
stream layout, and the refusal of priors without a generator."""
import numpy as np
import pytest
import torch

from oracle import aline_oracle as O


def test_prior_descriptors():
    from aline_b200.prior import PRIOR_BOX, PRIOR_CES, prior_of
    from aline_b200.tasks import CESTask, GPTask, HiddenLocation, PsychometricTask
    from aline_b200 import AlineError
    p = prior_of(HiddenLocation(K=2, n_target_theta=4, design_scale=1))
    assert p.kind == PRIOR_BOX and p.dim_theta == 4 and list(p.lo)[:4] == [0.0] * 4 and list(p.hi)[:4] == [1.0] * 4
    p = prior_of(HiddenLocation(theta_loc=torch.tensor([[-1.0, 2.0]]), theta_cov=torch.tensor([[3.0, 5.0]]), design_scale=5))
    assert (p.lo[0], p.lo[1], p.hi[0], p.hi[1]) == (-1.0, 2.0, 3.0, 5.0)
    p = prior_of(PsychometricTask())
    assert p.kind == PRIOR_BOX and p.dim_theta == 4
    assert [round(v, 6) for v in list(p.lo)[:4]] == [-3.0, 0.1, 0.1, 0.0]
    assert [round(v, 6) for v in list(p.hi)[:4]] == [3.0, 2.0, 0.9, 0.5]
    p = prior_of(CESTask())
    assert p.kind == PRIOR_CES and p.dim_theta == 5 and (p.lo[4], p.hi[4]) == (1.0, 3.0)
    with pytest.raises(AlineError):
        prior_of(HiddenLocation(theta_dist="normal", design_scale=1))
    with pytest.raises(AlineError):
        prior_of(GPTask(dim_x=1))


def test_stream_layout_is_sharding_invariant():
    """Row r of a shard that starts at global row `off` == row off + r of the unsharded stream; columns and seeds give
    different streams; uniforms have the 24-bit resolution of torch.rand."""
    full = O.prior_uniforms(99, 0, 64, 5, 2)
    part = O.prior_uniforms(99, 40, 24, 5, 2)
    assert np.array_equal(full[40:], part)
    assert not np.array_equal(full[:, 0], full[:, 1]) and not np.array_equal(full, O.prior_uniforms(100, 0, 64, 5, 2))
    assert np.array_equal(full * 2 ** 24, np.round(full * 2 ** 24))
    u = O.prior_uniforms(3, 2 ** 32 - 2, 4, 2, 1)            # the row counter carries into the high word
    assert len({tuple(r) for r in u.reshape(-1, 4).tolist()}) == 8
    big = O.prior_uniforms(7, 0, 20000, 2, 1)
    assert abs(big.mean() - 0.5) < 5e-3 and abs(big.var() - 1 / 12) < 2e-3
    th = O.prior_box(7, 0, 100, 3, [-3, 0.1], [3, 2.0])
    assert th.shape == (100, 3, 2) and (th[..., 0] >= -3).all() and (th[..., 0] < 3).all() and (th[..., 1] >= 0.1).all()


def _oracle_batch(task, seed, off, B):
    name = type(task).__name__
    n = task.n_context_init + task.n_query_init
    if name == "HiddenLocation":
        return O.sample_batch_philox("location", seed, off, B, n, task.dim_x, 0.0, 1.0, float(task.design_scale),
                                     lo=[0.0] * (task.K * task.dim_x), hi=[1.0] * (task.K * task.dim_x), K=task.K,
                                     noise_scale=float(task.noise_scale), base_signal=task.base_signal,
                                     max_signal=task.max_signal)
    if name == "CESTask":
        return O.sample_batch_philox("ces", seed, off, B, n, 6, 0.0, float(task.design_scale), 1.0,
                                     noise_scale=float(task.noise_scale), epsilon=float(task.epsilon))
    return O.sample_batch_philox("psychometric", seed, off, B, n, 1, -float(task.design_scale), float(task.design_scale),
                                 1.0, lo=[-3, 0.1, 0.1, 0.0], hi=[3, 2, 0.9, 0.5])


def test_philox_sample_batch_is_the_task_distribution():
    """The oracle's Philox restatement of Task.sample_batch (the function csrc/prior.cu evaluates) against the
    torch-generator mirrors of the reference simulators: same marginals of theta, designs and outcomes."""
    from aline_b200.tasks import CESTask, HiddenLocation, PsychometricTask
    torch.manual_seed(5)
    qs = torch.tensor([0.1, 0.3, 0.5, 0.7, 0.9])
    for task, tol in [(HiddenLocation(n_query_init=40, design_scale=1), 0.08),
                      (HiddenLocation(K=2, n_target_theta=4, n_query_init=40, design_scale=1), 0.08),
                      (CESTask(n_context_init=1, n_query_init=40), 0.03),
                      (PsychometricTask(n_context_init=1, n_query_init=40), 0.03)]:
        B = 1500
        ref = task.sample_batch(B)
        got = _oracle_batch(task, 17, 1000, B)
        ry = torch.cat([ref.context_y, ref.query_y], 1).reshape(-1).float()
        rx = torch.cat([ref.context_x, ref.query_x], 1).float()
        assert got["x"].shape == rx.shape and got["y"].shape[:2] == rx.shape[:2]
        assert got["theta"].shape == ref.target_all.shape[:2]
        assert (torch.quantile(got["y"].reshape(-1), qs) - torch.quantile(ry, qs)).abs().max().item() < tol, type(task)
        sx = float(rx.max() - rx.min())
        assert (torch.quantile(got["x"].reshape(-1), qs) - torch.quantile(rx.reshape(-1), qs)).abs().max().item() < 0.02 * sx
        rt = ref.target_all.reshape(B, -1).float()
        for j in range(rt.shape[1]):
            st = float(rt[:, j].max() - rt[:, j].min())
            assert (torch.quantile(got["theta"][:, j], qs) - torch.quantile(rt[:, j], qs)).abs().max().item() < 0.06 * st
    # a rollout's draw depends on its global index only
    a = _oracle_batch(HiddenLocation(n_query_init=10, design_scale=1), 3, 0, 8)
    b = _oracle_batch(HiddenLocation(n_query_init=10, design_scale=1), 3, 4, 4)
    assert all(torch.equal(a[k][4:], b[k]) for k in ("theta", "x", "y"))
