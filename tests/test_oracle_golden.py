"""The CPU oracle against fixtures produced by the reference itself (tests/golden/make_golden.py).

This is what pins the oracle: the reference has no golden vectors of its own for
this path.  CPU-only; runs in the build container and on the GPU box.
"""
import numpy as np
import pytest
import torch

from oracle import aline_oracle as O
from _util import load_golden, state_dict_of, step_batch, mode_of, n_head_of, abs_err, rel_err

ROLLOUTS = ["rollout_location", "rollout_location_sharp", "rollout_ces", "rollout_psychometric_a",
            "rollout_psychometric_b", "rollout_psychometric_d64", "rollout_gpmix_data", "rollout_gpmix_theta",
            "rollout_gpmix_all", "rollout_gpmix_none", "rollout_gpmix_absent", "rollout_location_tt", "rollout_location_value", "rollout_gp_data", "rollout_gp_theta"]


@pytest.mark.parametrize("name", ROLLOUTS)
@pytest.mark.parametrize("dense", [True, False])
def test_forward_teacher_forced(name, dense):
    g = load_golden(name)
    sd = state_dict_of(g)
    for t in range(int(g["n_steps"])):
        b = step_batch(g, t)
        o = O.forward(sd, b, mode_of(g), n_head_of(sd), dense=dense)
        pre = f"step{t}/"
        scale = max(1.0, float(np.abs(g[pre + "logits"]).max()))
        assert abs_err(o["logits"], g[pre + "logits"]) < 2e-5 * scale
        assert rel_err(o["zt"], g[pre + "zt"]) < 5e-5 * scale
        assert rel_err(o["log_prob"], g[pre + "log_prob"]) < 1e-5
        for k in ("mixture_means", "mixture_stds", "mixture_weights"):
            assert abs_err(o["posterior_out"][k], g[pre + "post/" + k]) < 2e-5, k
            assert abs_err(o["posterior_out_query"][k], g[pre + "postq/" + k]) < 2e-5, k
        if pre + "value" in g:
            assert rel_err(o["value"], g[pre + "value"]) < 1e-5
        po = o["posterior_out"]
        ll = O.compute_ll(b["target_all"], po["mixture_means"], po["mixture_stds"], po["mixture_weights"])
        assert abs_err(ll, g[pre + "target_ll"]) < 5e-5
        # index parity: exact unless the reference's own top-2 logit gap is below the fp32 error bound
        ref_idx = torch.from_numpy(g[pre + "idx"])
        lg = torch.from_numpy(g[pre + "logits"])
        top2 = lg.topk(2, dim=-1).values
        gap = top2[:, 0] - top2[:, 1]
        differs = (o["idx"] != ref_idx)[:, 0]
        assert not (differs & (gap > 1e-5 * scale)).any()
        # update_batch
        nb = O.update_batch(b, ref_idx)
        nxt = f"step{t + 1}/" if t + 1 < int(g["n_steps"]) else "final/"
        for k in ("context_x", "context_y", "query_x", "query_y"):
            assert np.array_equal(nb[k].numpy(), g[nxt + k]), k


def test_sharpened_indices_exact():
    g = load_golden("rollout_location_sharp")
    sd = state_dict_of(g)
    for t in range(int(g["n_steps"])):
        o = O.forward(sd, step_batch(g, t), "theta", 4, dense=True)
        assert np.array_equal(o["idx"].numpy(), g[f"step{t}/idx"])


def test_free_running_traces():
    g = load_golden("traces_location")
    sd = state_dict_of(g)
    b = {k: torch.from_numpy(g["batch0/" + k]) for k in ("context_x", "context_y", "query_x", "query_y", "target_all")}
    r = O.rollout(sd, b, 6, "theta", 4)
    assert abs_err(r["batch"]["context_x"], g["x"]) == 0.0        # design_scale = 1
    assert abs_err(r["batch"]["context_y"], g["y"]) == 0.0


def test_mask_truth_table():
    g = load_golden("mask_truth")
    for k in ("absent", "none_attr", "all_true", "all_false", "predef", "mixed"):
        tm = None if k in ("absent", "none_attr") else torch.from_numpy(g["tm_" + k])
        m = O.attention_mask(2, 3, 4, tm)
        assert np.array_equal(m.numpy(), g["mask_" + k]), k


@pytest.mark.parametrize("name,K", [("spce_location_k1", 1), ("spce_location_k2", 2)])
def test_spce_location(name, K):
    g = {k: torch.from_numpy(v) for k, v in load_golden(name).items()}
    r = O.spce_history(O.location_log_likelihood, g["y"], g["x"], g["thetas"], stepwise=True)
    assert rel_err(r["pce"], g["pce"]) < 1e-5 and rel_err(r["nmc"], g["nmc"]) < 1e-5
    r2 = O.spce_history(O.location_log_likelihood, g["y"], g["x"], g["thetas"], stepwise=False)
    assert rel_err(r2["pce"], g["pce_last"]) < 1e-5 and rel_err(r2["nmc"], g["nmc_last"]) < 1e-5
    for t in range(2):
        ll = O.location_log_likelihood(g["y"][:, t].unsqueeze(0), g["x"][:, t].unsqueeze(0), g["thetas"]).squeeze(-1)
        assert rel_err(ll, g["ll01"][t]) < 1e-6
    assert rel_err(O.pce_loss_whole_history(O.location_log_likelihood, g["y"], g["x"], g["thetas"]), g["pce_loss"]) < 1e-5
    assert rel_err(O.pce_loss_whole_history(O.location_log_likelihood, g["y"], g["x"], g["thetas"], nmc=True), g["nmc_loss"]) < 1e-5


def test_spce_ces():
    g = {k: torch.from_numpy(v) for k, v in load_golden("spce_ces").items()}
    for t in range(2):
        ll = O.ces_log_likelihood(g["y"][:, t].unsqueeze(0), g["x"][:, t].unsqueeze(0), g["thetas"]).squeeze(-1)
        assert torch.allclose(ll, g["ll01"][t], rtol=2e-6, atol=2e-6)   # same libm; 1-ulp op-order differences
    r = O.spce_history(O.ces_log_likelihood, g["y"], g["x"], g["thetas"], stepwise=True)
    assert rel_err(r["pce"], g["pce"]) < 1e-6 and rel_err(r["nmc"], g["nmc"]) < 1e-6


def test_ces_raises_on_inf():
    y = torch.tensor([[[2.0]]])            # outside (lo, hi): -inf -> ArithmeticError (censored_sigmoid_normal.py:81-84)
    xi = torch.rand(1, 1, 6) * 100
    th = torch.tensor([[[0.5, 0.3, 0.3, 0.4, 1.0]]])
    with pytest.raises(ArithmeticError):
        O.ces_log_likelihood(y, xi, th)


def test_psychometric_loglik():
    g = {k: torch.from_numpy(v) for k, v in load_golden("loglik_psychometric").items()}
    ll = O.psychometric_log_likelihood(g["y"], g["x"], g["theta"].squeeze(-1))
    assert torch.equal(ll, g["ll"])


def test_combine_partials_matches_logsumexp():
    torch.manual_seed(0)
    L, B, R = 999, 5, 3
    seq = torch.randn(L + 1, B) * 30
    ref_pce = seq.logsumexp(0) - seq[0]
    ref_nmc = seq[1:].logsumexp(0) - seq[0]
    chunks = torch.tensor_split(seq[1:], R, dim=0)
    m = torch.stack([c.max(0).values for c in chunks])
    s = torch.stack([torch.exp(c - c.max(0).values).sum(0) for c in chunks])
    out = O.combine_partials(m, s, seq[0], L)
    assert rel_err(np.log(L + 1) - out["pce"], ref_pce) < 1e-5
    assert rel_err(np.log(L) - out["nmc"], ref_nmc) < 1e-5


def test_gp_draws():
    g = {k: torch.from_numpy(v) for k, v in load_golden("gp_draws").items()}
    B = g["x"].shape[0]
    for b in range(B):
        kt = O.KERNEL_TYPES[int(g["ktype"][b])]
        r = O.gp_draw(g["x"][b], g["theta"][b, :2, 0], g["theta"][b, 2, 0], kt, g["z"][b], g["eps"][b])
        assert torch.equal(r["K"], g["K"][b])
        assert torch.equal(r["y"], g["y"][b, :, 0])       # LAPACK path: bitwise
        # restated Cholesky: reconstruction and draw within the SURVEY section-7 gates
        r2 = O.gp_draw(g["x"][b], g["theta"][b, :2, 0], g["theta"][b, 2, 0], kt, g["z"][b], g["eps"][b], lapack=False)
        K = g["K"][b].double()
        L2 = r2["L"].double()
        assert ((L2 @ L2.T - K).norm() / K.norm()).item() < 1e-6
        f_ref = g["L"][b] @ g["z"][b]
        assert ((r2["f"] - f_ref).norm() / f_ref.norm()).item() < 2e-3
    for i, kt in enumerate(O.KERNEL_TYPES):
        K = O.gp_kernel_matrix(g["x"][0], g["theta"][0, :2, 0], g["theta"][0, 2, 0], kt)
        assert torch.equal(K, g["K_" + kt])


def _uncertainty_inputs(g):
    batch = {k: torch.from_numpy(g["in/" + k]) for k in ("context_x", "context_y", "query_x", "query_y", "target_all",
                                                          "target_x")}
    return batch


def test_uncertainty_sampling_baseline():
    """calculate_gmm_variance (utils/misc.py:244-279) and the free-running uncertainty-sampling loop
    (notebooks/eval_al.ipynb cell 1) against the reference's own outputs."""
    g = load_golden("uncertainty_gpmix")
    sd = state_dict_of(g)
    for t in range(int(g["n_steps"])):
        pq = {k: torch.from_numpy(g[f"step{t}/postq/{k}"]) for k in ("mixture_means", "mixture_stds", "mixture_weights")}
        var = O.gmm_variance(pq["mixture_means"], pq["mixture_stds"], pq["mixture_weights"])
        assert rel_err(var, g[f"step{t}/var"]) < 1e-6
        var2 = O.gmm_variance(pq["mixture_means"], pq["mixture_stds"], pq["mixture_weights"][:, 0])
        assert rel_err(var2, g[f"step{t}/var_shared_w"]) < 1e-6
    r = O.rollout_uncertainty(sd, _uncertainty_inputs(g), int(g["n_steps"]), "mix", 4)
    for t in range(int(g["n_steps"])):
        assert torch.equal(r["idx"][:, t], torch.from_numpy(g[f"step{t}/idx"])[:, 0])
    assert abs_err(r["batch"]["context_x"], g["final/context_x"]) == 0.0


def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = O.philox4x32_10(np.array(ctr, dtype=np.uint32), key)
        assert tuple(int(v) for v in got) == want
    u = O.prior_uniforms(7, 0, 4, 3, 1)
    assert u.shape == (4, 3, 4) and u.dtype == np.float32 and (u >= 0).all() and (u < 1).all()


def test_summary_statistics_match_reference():
    """_summarise / eval_EIG_from_history / eval_boed aggregation (mean, unbiased std, 'se' / 'ci' / 'std') against
    what the unmodified reference's loops return for the same per-rollout bounds (fixture summary_stats.npz), including
    the reference's pre-scaled sNMC error in eval_EIG_from_history (utils/eval.py:119)."""
    import torch
    from aline_b200.utils import eval as ev
    g = load_golden("summary_stats")
    pce, nmc = torch.from_numpy(g["pce_all"]), torch.from_numpy(g["nmc_all"])
    pb, nb = torch.from_numpy(g["pce_boed"]), torch.from_numpy(g["nmc_boed"])
    for et in ("se", "ci", "std"):
        h = ev._summarise(pce, nmc, et, nmc_pre_scaled=True)
        b = ev._summarise(pb, nb, et)
        for k in ("pce_mean", "pce_err", "nmc_mean", "nmc_err"):
            assert torch.allclose(h[k], torch.from_numpy(g[f"history/{et}/{k}"]), rtol=1e-6, atol=1e-7), (et, k)
            assert torch.allclose(b[k], torch.from_numpy(g[f"boed/{et}/{k}"]), rtol=1e-6, atol=1e-7), (et, k)
    # eval_EIG_from_history's own loop (mini-batches of stored histories) with the bound computation stubbed out
    M, bs = pce.shape[0], int(g["batch_size"])
    orig = ev.compute_EIG_from_history
    try:
        ev.compute_EIG_from_history = lambda experiment, th, x, y, L, n, stepwise: (
            pce[int(x[0, 0, 0]):int(x[0, 0, 0]) + n], nmc[int(x[0, 0, 0]):int(x[0, 0, 0]) + n])
        x = torch.arange(M, dtype=torch.float32).reshape(M, 1, 1).expand(M, pce.shape[1], 1).contiguous()
        r = ev.eval_EIG_from_history(None, torch.zeros(M, 2), x, x, L=10, M=M, batch_size=bs, stepwise=True, err_type="ci")
    finally:
        ev.compute_EIG_from_history = orig
    for k in ("pce_mean", "pce_err", "nmc_mean", "nmc_err"):
        assert torch.allclose(r[k], torch.from_numpy(g[f"history/ci/{k}"]), rtol=1e-6, atol=1e-7), k


def test_spce_ces_large():
    """The oracle against the reference's CES bound at L = 1e5, B = 20, T = 15 (fixture spce_ces_large.npz; the
    contrastive draws are redrawn from the fixture's seed through the mirror task -- checksum in the fixture)."""
    from aline_b200.tasks import CESTask
    g = load_golden("spce_ces_large")
    L, seed = int(g["L"]), int(g["seed"])
    B = g["x"].shape[0]
    torch.manual_seed(seed)
    thetas = CESTask(n_context_init=1, n_query_init=1).sample_theta((L, B))
    assert abs(float(thetas.double().sum()) - float(g["thetas_checksum"])) < 1e-6 * abs(float(g["thetas_checksum"]))
    th0 = torch.from_numpy(g["theta_0"])
    r = O.spce_history(O.ces_log_likelihood, torch.from_numpy(g["y"]), torch.from_numpy(g["x"]),
                       torch.cat([th0.unsqueeze(0), thetas], 0), stepwise=True)
    assert rel_err(r["pce"], g["pce"]) < 1e-5 and rel_err(r["nmc"], g["nmc"]) < 1e-5
