"""The BASELINE.json configurations other than the headline one, at FULL size, through size-independent properties
(the oracle needs minutes to hours at these sizes): cfg1 (location B=1000), cfg3 (CES, L=1e7 sharded draws),
cfg4 (GP mix, T=50), cfg5 (psychometric, predefined target masks)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(dx, n_theta, mode, precision):
    from aline_b200.model import Aline, Embedder, Encoder, OutputHead
    torch.manual_seed(123)
    m = Aline(Embedder(dx, 1, 32, 128, n_theta, mode), Encoder(32, 128, 4, 0.0, 3), OutputHead(dx, 1, 32, 128))
    m = m.cuda().eval()
    m.precision = precision
    return m


def _rollout_and_check(task, model, B, steps, target_mask=None):
    """Every step retires exactly one live candidate of each rollout; the appended (design, outcome) pairs are exactly
    the retired candidates with their pre-simulated outcomes, in selection order; log-probs are finite and <= 0."""
    from aline_b200.attrdict import AttrDict
    hb = task.sample_batch(B)
    b = AttrDict({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in hb.items()})
    if target_mask is not None:
        b.target_mask = target_mask
    qx, qy = b.query_x.clone(), b.query_y.clone()
    n0 = b.context_x.shape[1]
    nq = qx.shape[1]
    out = model.rollout(b, steps)
    assert out.context_x.shape[1] == n0 + steps and out.context_y.shape[1] == n0 + steps
    alive = out.query_alive.bool()
    assert (alive.sum(1) == nq - steps).all()
    assert torch.isfinite(out.design_log_prob).all() and (out.design_log_prob <= 0).all()
    # design_idx is the index within the compacted live set (the reference's design_out.idx): replay it
    idx = out.design_idx.cpu()
    for r in (0, B // 2, B - 1):
        live = list(range(nq))
        for t in range(steps):
            j = live.pop(int(idx[r, t]))
            assert torch.equal(out.context_x[r, n0 + t].cpu(), qx[r, j].cpu())
            assert torch.equal(out.context_y[r, n0 + t].cpu(), qy[r, j].cpu())
        assert sorted(live) == torch.where(alive[r])[0].cpu().tolist()
    return out, hb


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cfg1_location_b1000(precision):
    from aline_b200.tasks import HiddenLocation
    from aline_b200.utils.eval import compute_EIG_from_history
    torch.manual_seed(1)
    task = HiddenLocation(n_query_init=200, design_scale=1)
    out, hb = _rollout_and_check(task, _model(2, 2, "theta", precision), 1000, 29)
    L = 10_000
    theta_0 = hb["target_all"].reshape(1000, 1, 2).cuda()
    thetas = torch.rand(L, 1000, 1, 2, device="cuda")
    pce, nmc = compute_EIG_from_history(task, theta_0, out.context_x, out.context_y, L=L, batch_size=1000, stepwise=True,
                                        thetas=thetas)
    assert pce.shape == (1000, 30) and torch.isfinite(pce).all() and torch.isfinite(nmc).all()
    assert (pce <= nmc + 1e-4).all() and (pce <= math.log(L + 1) + 1e-4).all()
    perm = torch.randperm(L, device="cuda")                      # the bound does not depend on the order of the draws
    pce2, _ = compute_EIG_from_history(task, theta_0, out.context_x, out.context_y, L=L, batch_size=1000, stepwise=True,
                                       thetas=thetas[perm])
    assert torch.allclose(pce2, pce, rtol=1e-4, atol=5e-5)


def test_cfg3_ces_l1e7_sharded():
    """CES eval-final: B=20, 2000 candidates, 14 steps, L=1e7 contrastive draws.  The draws are evaluated as 8 row
    shards (what 8 ranks would hold, 1.25e6 each) and combined with the all-gather formula; two different shardings
    of the same draws must agree, and sPCE <= sNMC, sPCE <= log(L+1)."""
    from aline_b200 import spce
    from aline_b200.tasks import CESTask
    torch.manual_seed(3)
    task = CESTask(n_context_init=1, n_query_init=2000)
    out, hb = _rollout_and_check(task, _model(6, 5, "theta", "bf16"), 20, 14)
    B, L = 20, 10_000_000
    x, y = task.unnormalise_design(out.context_x), out.context_y
    theta_0 = hb["target_all"].reshape(1, B, 5).cuda()
    rows = task.sample_theta((L, B)).cuda()

    def bounds(n_shards):
        ms, ss, lp0 = [], [], None
        for r in range(n_shards):
            lo, hi = spce.shard_rows(L, r, n_shards)
            part = torch.cat([theta_0, rows[lo:hi]], 0)
            m, s, lp0 = spce.spce_history(task.log_likelihood, y, x, part, seq=None, skip_rows=1)
            ms.append(m)
            ss.append(s)
            del part
        pl, nl = spce.lse_combine(torch.stack(ms), torch.stack(ss), lp0)
        return math.log(L + 1) - pl, math.log(L) - nl

    pce8, nmc8 = bounds(8)
    assert pce8.shape == (B, 15) and torch.isfinite(pce8).all() and torch.isfinite(nmc8).all()
    assert (pce8 <= nmc8 + 1e-4).all() and (pce8 <= math.log(L + 1) + 1e-4).all()
    pce3, nmc3 = bounds(3)
    assert torch.allclose(pce3, pce8, rtol=1e-4, atol=5e-5) and torch.allclose(nmc3, nmc8, rtol=1e-4, atol=5e-5)


@pytest.mark.parametrize("attend_to", ["theta", "data"])
def test_cfg4_gp_mix_t50(attend_to):
    from aline_b200.tasks import GPTask
    from aline_b200.utils.target_mask import create_target_mask
    torch.manual_seed(4)
    task = GPTask(dim_x=2, embedding_type="mix", n_context_init=1, n_query_init=200, n_target_theta=3, n_target_data=100,
                  design_scale=5)
    tm = create_target_mask("split", "mix", 100, 3, None, None, None, None, attend_to)
    assert tm.shape == (103,) and int(tm.sum()) == (3 if attend_to == "theta" else 100)
    torch.set_default_device("cuda")
    try:
        _rollout_and_check(task, _model(2, 3, "mix", "bf16"), 200, 50, target_mask=tm)
    finally:
        torch.set_default_device("cpu")


@pytest.mark.parametrize("mask", [[False, False, True, True], [True, True, False, False]])
def test_cfg5_psychometric_masks(mask):
    from aline_b200.tasks import PsychometricTask
    torch.manual_seed(5)
    task = PsychometricTask(n_context_init=1, n_query_init=200, design_scale=5)
    out, hb = _rollout_and_check(task, _model(1, 4, "theta", "bf16"), 200, 30, target_mask=torch.tensor(mask))
    assert set(torch.unique(out.context_y).cpu().tolist()) <= {0.0, 1.0}        # Bernoulli outcomes
