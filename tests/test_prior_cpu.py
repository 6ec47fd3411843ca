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
