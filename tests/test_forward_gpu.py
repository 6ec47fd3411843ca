"""GPU parity of Aline.forward / Task.update_batch / Aline.rollout (through the C ABI) against the golden fixtures
(teacher-forced records of the reference) and the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import aline_oracle as O
from _util import load_golden, state_dict_of, step_batch, mode_of, n_head_of, abs_err, rel_err

pytestmark = pytest.mark.gpu

ROLLOUTS = ["rollout_location", "rollout_location_sharp", "rollout_ces", "rollout_psychometric_a",
            "rollout_psychometric_b", "rollout_psychometric_d64", "rollout_gpmix_data", "rollout_gpmix_theta",
            "rollout_gpmix_all", "rollout_gpmix_none", "rollout_gpmix_absent", "rollout_location_tt", "rollout_location_value", "rollout_gp_data", "rollout_gp_theta"]

# BASELINE.json: "encoder/head log-probs match ... to 1e-5 in a full-fp32 mode"
LOGP_RTOL_FP32 = 1e-5


def build_model(sd, mode, precision="fp32"):
    from aline_b200.model import Aline, Embedder, Encoder, OutputHead
    d = sd["embedder.x_embedder.2.weight"].shape[0]
    ff, dx = sd["embedder.x_embedder.0.weight"].shape
    ntok = sd["embedder.theta_tokens"].shape[0] if "embedder.theta_tokens" in sd else 0
    nl = 0
    while f"encoder.encoder.layers.{nl}.linear1.weight" in sd:
        nl += 1
    tt = sd["head.acquisition_head.predictor.0.weight"].shape[1] == d + 1
    model = Aline(Embedder(dx, 1, d, ff, ntok, mode), Encoder(d, ff, d // 8, 0.0, nl),
                  OutputHead(dx, 1, d, ff, time_token=tt, value_head="head.value_head.empty_value" in sd))
    missing, unexpected = model.load_state_dict(sd, strict=True)      # same key names / shapes as the reference
    assert not missing and not unexpected
    model.precision = precision
    return model.cuda().eval()


def attr_batch(b):
    from aline_b200.attrdict import AttrDict
    return AttrDict({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()})


@pytest.mark.parametrize("name", ROLLOUTS)
def test_forward_teacher_forced(name):
    from aline_b200.tasks import Task
    from aline_b200.utils.eval import compute_ll
    g = load_golden(name)
    sd = state_dict_of(g)
    model = build_model(sd, mode_of(g))
    model.query_posterior = "eager" if name.endswith("sharp") else "lazy"
    task = Task(dim_x=g["step0/query_x"].shape[-1], dim_y=1)
    for t in range(int(g["n_steps"])):
        b = attr_batch(step_batch(g, t))
        if "target_mask" in b:
            b.target_mask = b.target_mask.cpu()
        pred = model.forward(b)
        pre = f"step{t}/"
        scale = max(1.0, float(np.abs(g[pre + "logits"]).max()))
        assert rel_err(pred.design_out.zt.cpu(), g[pre + "zt"]) < 5e-5 * scale
        assert rel_err(pred.design_out.log_prob.cpu(), g[pre + "log_prob"]) < LOGP_RTOL_FP32
        if pre + "value" in g:
            assert rel_err(pred.value.cpu(), g[pre + "value"]) < 1e-5
        for k in ("mixture_means", "mixture_stds", "mixture_weights"):
            assert abs_err(pred.posterior_out[k].cpu(), g[pre + "post/" + k]) < 2e-5, k
            assert abs_err(pred.posterior_out_query[k].cpu(), g[pre + "postq/" + k]) < 2e-5, k
        po = pred.posterior_out
        ll = compute_ll(b.target_all, po.mixture_means, po.mixture_stds, po.mixture_weights)
        assert abs_err(ll.cpu(), g[pre + "target_ll"]) < 5e-5
        # index parity: bit-exact except where the reference's own top-2 logit gap is below the fp32 error bound
        ref_idx = torch.from_numpy(g[pre + "idx"])
        lg = torch.from_numpy(g[pre + "logits"])
        top2 = lg.topk(2, dim=-1).values
        gap = top2[:, 0] - top2[:, 1]
        differs = (pred.design_out.idx.cpu() != ref_idx)[:, 0]
        assert pred.design_out.idx.dtype == torch.int64 and tuple(pred.design_out.idx.shape) == tuple(ref_idx.shape)
        assert not (differs & (gap > 1e-5 * scale)).any()
        # Task.update_batch with the reference's index (teacher forcing)
        nb = task.update_batch(b, ref_idx.cuda())
        nxt = f"step{t + 1}/" if t + 1 < int(g["n_steps"]) else "final/"
        for k in ("context_x", "context_y", "query_x", "query_y"):
            assert np.array_equal(nb[k].cpu().numpy(), g[nxt + k]), k


def test_sharpened_indices_exact():
    g = load_golden("rollout_location_sharp")
    model = build_model(state_dict_of(g), "theta")
    for t in range(int(g["n_steps"])):
        pred = model.forward(attr_batch(step_batch(g, t)))
        assert np.array_equal(pred.design_out.idx.cpu().numpy(), g[f"step{t}/idx"])


def test_resident_rollout_matches_reference_traces():
    """Free-running: utils/eval.py get_traces of the reference (fixture) vs the resident rollout."""
    g = load_golden("traces_location")
    model = build_model(state_dict_of(g), "theta")
    b = attr_batch({k: torch.from_numpy(g["batch0/" + k]) for k in
                    ("context_x", "context_y", "query_x", "query_y", "target_all")})
    out = model.rollout(b, 6)
    assert abs_err(out.context_x.cpu(), g["x"]) == 0.0        # design_scale = 1
    assert abs_err(out.context_y.cpu(), g["y"]) == 0.0
    assert int(out.query_alive.sum()) == out.query_alive.numel() - 6 * out.query_alive.shape[0]


@pytest.mark.parametrize("name,T", [("rollout_location_sharp", 8), ("rollout_gpmix_data", 5), ("rollout_ces", 4),
                                    ("rollout_psychometric_d64", 4), ("rollout_gpmix_theta", 3), ("rollout_gpmix_none", 3),
                                    ("rollout_psychometric_a", 3), ("rollout_gp_data", 3), ("rollout_gp_theta", 3)])
def test_resident_rollout_vs_oracle(name, T):
    """Resident rollout (retired-candidate bitmap, in-place append) vs the oracle's forward + update_batch loop."""
    g = load_golden(name)
    sd = state_dict_of(g)
    mode = mode_of(g)
    model = build_model(sd, mode)
    b0 = step_batch(g, 0)
    ref = O.rollout(sd, dict(b0), T, mode, n_head_of(sd), dense=False)
    out = model.rollout(attr_batch(b0), T)
    same = (out.design_idx.cpu() == ref["idx"])
    # trajectories may only part ways at a near-tie; with the sharpened / trained-like weights they must not
    if name.endswith("sharp"):
        assert same.all()
    first_div = torch.where(~same.all(0))[0]
    upto = int(first_div[0]) if len(first_div) else T
    assert upto >= 1
    assert rel_err(out.design_log_prob.cpu()[:, :upto], ref["log_prob"][:, :upto]) < 1e-4
    if same.all():
        assert abs_err(out.context_x.cpu(), ref["batch"]["context_x"]) == 0.0
        assert abs_err(out.context_y.cpu(), ref["batch"]["context_y"]) == 0.0


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_large_rollout_properties(precision):
    """cfg2-sized rollout (B=200, 2000 candidates, 34 steps): every step retires exactly one live candidate,
    appended designs are members of the candidate set with their pre-simulated outcome, no repeats."""
    from aline_b200.model import Aline, Embedder, Encoder, OutputHead
    from aline_b200.tasks import HiddenLocation
    torch.manual_seed(123)
    model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128))
    model = model.cuda().eval()
    model.precision = precision
    task = HiddenLocation(n_query_init=2000, design_scale=1)
    batch = task.sample_batch(200)
    for k in ("context_x", "context_y", "query_x", "query_y", "target_all"):
        batch[k] = batch[k].cuda()
    qx, qy = batch.query_x.clone(), batch.query_y.clone()
    out = model.rollout(batch, 34)
    assert out.context_x.shape == (200, 35, 2)
    alive = out.query_alive.bool()
    assert (alive.sum(1) == 2000 - 34).all()
    chosen_x = out.context_x[:, 1:]
    for b in (0, 57, 199):
        dead = torch.where(~alive[b])[0]
        got = {tuple(r.tolist()) for r in torch.cat([chosen_x[b], out.context_y[b, 1:]], -1).cpu()}
        want = {tuple(r.tolist()) for r in torch.cat([qx[b, dead], qy[b, dead]], -1).cpu()}
        assert got == want
    assert torch.isfinite(out.design_log_prob).all() and (out.design_log_prob <= 0).all()


# ---------------------------------------------------------------- bf16 tensor-core query stream ----
# BASELINE.json: "encoder/head log-probs match to 1e-3 relative with bf16 operands and fp32 accumulation"
LOGP_RTOL_BF16 = 1e-3
LOGIT_ABS_BF16 = 4e-3        # fixed logit bound of the bf16 mode (tests/test_query_parity_gpu.py)
POSTQ_ABS_BF16 = 1e-2        # GMM means / stds on candidate rows in bf16 mode
POSTQ_W_ABS_BF16 = 2e-3      # GMM mixture weights on candidate rows in bf16 mode
TC_FIXTURES = list(ROLLOUTS)          # d = 32: csrc/query_tc3.cu / query_tc4.cu / query_tc.cu; d = 64: csrc/query_tc5.cu


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_under_no_grad_runs_the_kernels(precision):
    """`Aline.forward` without autograd must go through the C ABI (the differentiable torch-op composition of row f2 is
    only for grad mode): the library's launch counter moves, and in bf16 mode the log-probs carry bf16 rounding -- they
    are NOT the fp32 composition's values (which agree with the reference to ~1e-7)."""
    from aline_b200 import _lib
    g = load_golden("rollout_location")
    model = build_model(state_dict_of(g), mode_of(g), precision=precision)
    assert not torch.is_grad_enabled()
    n0 = _lib.kernel_launches()
    pred = model.forward(attr_batch(step_batch(g, 0)))
    torch.cuda.synchronize()
    assert _lib.kernel_launches() - n0 >= 3, "forward did not launch the library's kernels"
    err = rel_err(pred.design_out.log_prob.cpu(), g["step0/log_prob"])
    if precision == "bf16":
        assert 1e-6 < err < LOGP_RTOL_BF16, err
    else:
        assert err < 1e-5, err


@pytest.mark.needs_grad
def test_eval_mode_forward_runs_the_kernels_even_with_autograd_enabled():
    """notebooks/eval_al.ipynb and eval_psychometric.ipynb call `model.eval()` and then `model(batch)` WITHOUT
    torch.no_grad(): that call must run the kernels (nothing is differentiated there).  The autograd composition is for
    train() mode, or on request (`model.differentiable = True`)."""
    from aline_b200 import _lib
    g = load_golden("rollout_location")
    model = build_model(state_dict_of(g), mode_of(g), precision="fp32")          # .eval()
    assert torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters())
    n0 = _lib.kernel_launches()
    pred = model(attr_batch(step_batch(g, 0)))
    torch.cuda.synchronize()
    assert _lib.kernel_launches() - n0 >= 3 and not pred.design_out.log_prob.requires_grad
    assert rel_err(pred.design_out.log_prob.cpu(), g["step0/log_prob"]) < 1e-5
    model.differentiable = True
    pred = model(attr_batch(step_batch(g, 0)))
    assert pred.posterior_out.mixture_means.requires_grad
    assert rel_err(pred.design_out.log_prob.detach().cpu(), g["step0/log_prob"]) < 1e-5


@pytest.mark.parametrize("name", TC_FIXTURES)
def test_forward_teacher_forced_bf16(name):
    g = load_golden(name)
    sd = state_dict_of(g)
    model = build_model(sd, mode_of(g), precision="bf16")
    from aline_b200 import rollout as ro
    pm = model.packed()
    assert pm.tc_blob is not None and (pm.tc_fast_max_keys >= 48 if "d64" in name else pm.tc_max_keys >= 60)
    for t in range(int(g["n_steps"])):
        b = attr_batch(step_batch(g, t))
        pred = model.forward(b)
        pre = f"step{t}/"
        # the sharpened fixture multiplies the last acquisition weight by 100, and with it the bf16 operand
        # rounding error of the logits: its tolerance scales accordingly (it exists to exercise index parity)
        amp = 100.0 if name.endswith("sharp") else 1.0
        assert rel_err(pred.design_out.log_prob.cpu(), g[pre + "log_prob"]) < LOGP_RTOL_BF16 * amp
        # targets run on the fp32 path: same tolerance as in fp32 mode
        for k in ("mixture_means", "mixture_stds", "mixture_weights"):
            assert abs_err(pred.posterior_out[k].cpu(), g[pre + "post/" + k]) < 2e-5, k
        # candidate rows carry the bf16 operand rounding of the query stream (final states off by ~3e-3) through the
        # GMM head's two-layer MLP: FIXED budgets, ~2x the largest error measured over all fixtures
        # (tools/measure_bf16_errors.py: means / stds <= 4.5e-3, weights <= 6.5e-4)
        for k, tol in (("mixture_means", POSTQ_ABS_BF16), ("mixture_stds", POSTQ_ABS_BF16), ("mixture_weights", POSTQ_W_ABS_BF16)):
            assert abs_err(pred.posterior_out_query[k].cpu(), g[pre + "postq/" + k]) < tol, k
        # log-softmax error per row: a FIXED bound of twice the logit bound (the normaliser moves by at most the logit
        # error); measured 1.0e-3 .. 3.0e-3.  Index parity: exact unless the reference's own top-2 gap is below it.
        zt, zr = pred.design_out.zt.cpu().double(), torch.from_numpy(g[pre + "zt"]).double()
        row_err = (zt.log() - zr.log()).abs().max(-1).values
        assert (row_err < 2 * LOGIT_ABS_BF16 * amp).all(), float(row_err.max())
        lg = torch.from_numpy(g[pre + "logits"]).double()
        top2 = lg.topk(2, dim=-1).values
        gap = top2[:, 0] - top2[:, 1]
        differs = (pred.design_out.idx.cpu() != torch.from_numpy(g[pre + "idx"]))[:, 0]
        assert not (differs & (gap > 2 * LOGIT_ABS_BF16 * amp)).any()


def test_bf16_and_fp32_rollouts_agree_when_not_near_tie():
    """Sharpened acquisition head: the bf16 and fp32 resident rollouts pick the same designs except at near-ties."""
    g = load_golden("rollout_location_sharp")
    sd = state_dict_of(g)
    b0 = step_batch(g, 0)
    out32 = build_model(sd, "theta", "fp32").rollout(attr_batch(b0), 8)
    out16 = build_model(sd, "theta", "bf16").rollout(attr_batch(b0), 8)
    same = (out32.design_idx == out16.design_idx)
    assert same[:, 0].float().mean() >= 0.5
    assert torch.isfinite(out16.design_log_prob).all()
    # log-probs of the first step (same inputs in both modes); the logits are sharpened x100, and so is their error
    first = same[:, 0]
    assert rel_err(out16.design_log_prob[first, 0].cpu(), out32.design_log_prob[first, 0].cpu()) < 0.1


def _query_logits(model, b, fast):
    """Candidate logits of one forward through the tensor-core query stream: fast kernel (+ conditional fallback)
    or the general kernel alone."""
    from aline_b200 import rollout as ro
    pm = model.packed()
    n_c = b.context_x.shape[1]
    slots, n_sel = ro.target_slots(pm.dims["n_theta_tok"], None, "cuda")
    eq, eq_rm = ro.embed_queries(pm, b.query_x, row_major=True)
    tc_kv = ro.alloc_tc_kv(pm, b.context_x.shape[0], n_c + n_sel, "cuda") if fast else None
    assert (tc_kv is not None) == fast
    kv, _ = ro.ctx_stack(pm, b.context_x, b.context_y, n_c, None, slots, n_sel, tc_kv=tc_kv)
    logits, _ = ro.query_stream(pm, eq, None, kv, n_c + n_sel, precision="bf16", tc_kv=tc_kv, eq_rm=eq_rm if fast else None)
    return logits


def test_fast_tc_kernel_is_used_and_overflow_falls_back():
    """The fast tcgen05 kernel takes the softmax relative to key 0; when a score exceeds key 0's by more than 127
    log2-units it flags the launch and the general kernel recomputes it (bit-identical to running it alone)."""
    g = load_golden("rollout_location")
    sd = state_dict_of(g)
    b = attr_batch(step_batch(g, 3))
    model = build_model(sd, "theta", "bf16")
    fast, general = _query_logits(model, b, True), _query_logits(model, b, False)
    assert not torch.equal(fast, general)                      # two different kernels ...
    assert abs_err(fast.cpu(), general.cpu()) < 0.05           # ... that agree to bf16 operand rounding
    sd_hot = {k: v.clone() for k, v in sd.items()}
    for l in range(3):                                         # blow up q and k: |scores| in the thousands
        sd_hot[f"encoder.encoder.layers.{l}.self_attn.in_proj_weight"][:64] *= 60.0
    hot = build_model(sd_hot, "theta", "bf16")
    fast, general = _query_logits(hot, b, True), _query_logits(hot, b, False)
    assert torch.isfinite(fast).all()
    assert torch.equal(fast, general)


def test_uncertainty_sampling_baseline():
    """Row f3: calculate_gmm_variance as a kernel, its fusion with the GMM head, and the resident
    uncertainty-sampling rollout, against the reference's own free-running loop (fixture uncertainty_gpmix)."""
    from aline_b200.utils import calculate_gmm_variance
    from aline_b200 import rollout as ro
    g = load_golden("uncertainty_gpmix")
    sd = state_dict_of(g)
    model = build_model(sd, "mix")
    T = int(g["n_steps"])
    for t in range(T):
        pq = {k: torch.from_numpy(g[f"step{t}/postq/{k}"]).cuda() for k in ("mixture_means", "mixture_stds",
                                                                             "mixture_weights")}
        var = calculate_gmm_variance(pq["mixture_means"], pq["mixture_stds"], pq["mixture_weights"])
        assert rel_err(var.cpu(), g[f"step{t}/var"]) < 1e-5
        var2 = calculate_gmm_variance(pq["mixture_means"], pq["mixture_stds"], pq["mixture_weights"][:, 0].contiguous())
        assert rel_err(var2.cpu(), g[f"step{t}/var_shared_w"]) < 1e-5
    with pytest.raises(Exception):
        calculate_gmm_variance(pq["mixture_means"], pq["mixture_stds"], pq["mixture_weights"][:, :3].contiguous())
    b = attr_batch({k: torch.from_numpy(g["in/" + k]) for k in ("context_x", "context_y", "query_x", "query_y",
                                                                 "target_all", "target_x")})
    # forward + lazily materialised posterior_out_query + variance == the fused head-variance kernel == the reference
    out = model.forward(b)
    pq = out.posterior_out_query
    var = calculate_gmm_variance(pq.mixture_means, pq.mixture_stds, pq.mixture_weights)
    assert rel_err(var.cpu(), g["step0/var"]) < 2e-4
    b = attr_batch({k: torch.from_numpy(g["in/" + k]) for k in ("context_x", "context_y", "query_x", "query_y",
                                                                 "target_all", "target_x")})
    r = model.rollout(b, T, acquisition="uncertainty_sampling")
    for t in range(T):
        assert torch.equal(r.design_idx[:, t].cpu(), torch.from_numpy(g[f"step{t}/idx"])[:, 0]), t
    assert abs_err(r.context_x.cpu(), g["final/context_x"]) == 0.0
    assert abs_err(r.context_y.cpu(), g["final/context_y"]) == 0.0
    with pytest.raises(ValueError):
        model.rollout(b, 1, acquisition="random")
