"""Value parity of the candidate-query stream at BENCHMARK size (through the C ABI) against the CPU oracle.

The teacher-forced fixtures have <= 40 candidates = one 128-token tile; the benchmarked configuration has 2000
candidates = 16 tiles per rollout, tile groups > 0, left-over half-units and the <4> -> <2> -> general kernel
hand-over as the context grows.  These tests compare the candidate LOGITS of the CUDA kernels with the oracle's
structured forward (`O.forward(..., dense=False)`, reference: model/encoder.py:128-141, model/head.py:27-31,355-358)
on the same seeded inputs, at 300 and 2000 candidates, for key counts that cross every kernel boundary
(16 / 32 / 48 padded keys; 49+ keys = general tensor-core kernel), in both precision modes, with the masks of
utils/target_mask.py.

Tolerances (BASELINE.json north_star):
  fp32 mode   log_prob 1e-5 relative; logits 2e-5 absolute; design index exact unless the oracle's own top-2 logit
              gap is < 1e-5 (near-tie)
  bf16 mode   log_prob 1e-3 relative; logits LOGIT_ABS_BF16 = 4e-3 absolute -- a FIXED bound (bf16 operand rounding
              of a d=32 three-layer stack with random-init weights, logit spread ~0.03; SURVEY.md section 7 measured
              2.4e-3 for an emulation of the same arithmetic); design index exact whenever the oracle's top-2 gap
              exceeds 2 x that bound
"""
import pytest
import torch

from oracle import aline_oracle as O
from _util import load_golden, state_dict_of, abs_err, rel_err
from test_forward_gpu import build_model, attr_batch

pytestmark = pytest.mark.gpu

LOGIT_ABS_FP32 = 2e-5
LOGIT_ABS_BF16 = 4e-3
LOGP_RTOL = {"fp32": 1e-5, "bf16": 1e-3}


def _location_sd():
    return state_dict_of(load_golden("rollout_location"))          # random-init reference weights, theta mode, 2 tokens


def _batch(B, n_c, nq, dx=2, n_t=2, seed=0, target_x=None):
    g = torch.Generator().manual_seed(1000 * seed + 17 * n_c + nq + B)
    b = dict(context_x=torch.rand(B, n_c, dx, generator=g), context_y=torch.randn(B, n_c, 1, generator=g),
             query_x=torch.rand(B, nq, dx, generator=g), query_y=torch.randn(B, nq, 1, generator=g),
             target_all=torch.rand(B, n_t, 1, generator=g))
    if target_x is not None:
        b["target_x"] = target_x
    return b


def _gpu_logits(model, b, target_mask=None, general=False, one_thread_per_row=False):
    """Logits of one forward through the lower-level wrappers (same calls as Aline.forward); general=True withholds the
    fast kernel's operand blocks so the general tensor-core kernel is the primary path; one_thread_per_row=True withholds
    the row-major embeddings so the fast kernel is query_tc3 (one thread per row) instead of query_tc4 (two)."""
    from aline_b200 import rollout as ro
    pm = model.packed()
    cx, cy, qx = b["context_x"].cuda(), b["context_y"].cuda(), b["query_x"].cuda()
    tx = b["target_x"].cuda() if "target_x" in b else None
    n_c = cx.shape[1]
    n_t = (0 if tx is None else tx.shape[1]) + pm.dims["n_theta_tok"]
    slots, n_sel = ro.target_slots(n_t, target_mask, cx.device)
    eq, eq_rm = ro.embed_queries(pm, qx, row_major=True)
    tc_kv = None
    if (not general and ro.use_tensor_cores(pm, model.precision, n_c + n_sel)
            and n_c + n_sel <= pm.tc_fast_max_keys):
        tc_kv = ro.alloc_tc_kv(pm, cx.shape[0], n_c + n_sel, cx.device)
    kv, _ = ro.ctx_stack(pm, cx, cy, n_c, tx, slots, n_sel, tc_kv=tc_kv)
    logits, _ = ro.query_stream(pm, eq, None, kv, n_c + n_sel, precision=model.precision, tc_kv=tc_kv,
                                eq_rm=None if one_thread_per_row else eq_rm)
    return logits.cpu()


def _check(model, sd, b, mode, precision, target_mask=None, general=False):
    ob = dict(b)
    if target_mask is not None:
        ob["target_mask"] = target_mask
    n_head = sd["embedder.x_embedder.2.weight"].shape[0] // 8
    ref = O.forward(sd, ob, mode, n_head, dense=False, with_query_posterior=False)
    tol = LOGIT_ABS_FP32 if precision == "fp32" else LOGIT_ABS_BF16
    from aline_b200 import _lib
    # both fast kernels; the one-thread-per-row kernel with and without the folded operands (K' = Wq^T K, V' = V Wo^T)
    variants = ((False, -1), (True, 1), (True, 0)) if (precision == "bf16" and not general) else ((False, -1),)
    for tc3, fold in variants:
        _lib.set_option("query_tc4", 0 if tc3 else 1)
        _lib.set_option("query_fold", fold)
        try:
            lg = _gpu_logits(model, b, target_mask, general, one_thread_per_row=tc3)
        finally:
            _lib.set_option("query_tc4", -1)
            _lib.set_option("query_fold", -1)
        err = (lg.double() - ref["logits"].double()).abs().max().item()
        assert err < tol, (f"logits differ from the oracle by {err:.3e} (bound {tol:.1e}, one_thread_per_row={tc3}, "
                           f"fold={fold})")
    if general:
        return
    ab = attr_batch(b)
    if target_mask is not None:
        ab.target_mask = target_mask
    pred = model.forward(ab)
    assert rel_err(pred.design_out.log_prob.cpu(), ref["log_prob"]) < LOGP_RTOL[precision]
    top2 = ref["logits"].topk(2, dim=-1).values
    gap = top2[:, 0] - top2[:, 1]
    differs = (pred.design_out.idx.cpu() != ref["idx"])[:, 0]
    near = 1e-5 if precision == "fp32" else 2 * LOGIT_ABS_BF16
    assert not (differs & (gap > near)).any(), "design index differs outside a near-tie"


# context lengths -> key counts n_c + 2: 3, 16, 17, 32, 33, 37, 49, 51, 82 (kernel <4> up to 32 keys, <2> up to 48,
# general tcgen05 kernel up to its shared-memory limit, FFMA kernel beyond)
CTX = [1, 14, 15, 30, 31, 35, 47, 49, 80]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("B,nq", [(3, 300), (8, 2000), (3, 2000), (8, 300), (5, 200), (6, 100)])
def test_logits_vs_oracle_multi_tile(precision, B, nq):
    sd = _location_sd()
    model = build_model(sd, "theta", precision)
    model.query_posterior = "off"
    for n_c in CTX:
        if precision == "fp32" and nq == 2000 and n_c not in (1, 31, 80):
            continue                                    # the FFMA kernel has one code path for every key count
        _check(model, sd, _batch(B, n_c, nq, seed=B), "theta", precision)


def test_d64_fast_kernel_is_selected():
    """dim_embedding 64 / 8 heads (config/model/aline_psychometric.yaml): csrc/query_tc5.cu covers up to 48 keys."""
    sd = state_dict_of(load_golden("rollout_psychometric_d64"))
    pm = build_model(sd, "theta", "bf16").packed()
    assert pm.dims["d"] == 64 and pm.tc_blob is not None and pm.tc_max_keys == 0 and pm.tc_fast_max_keys == 48


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("B,nq", [(3, 200), (5, 300), (200, 200), (151, 130)])
def test_d64_logits_vs_oracle(precision, B, nq):
    """The d = 64 / 8-head model (psychometric configuration, 4 theta tokens -> key counts n_c + 4): weight-streaming
    tcgen05 kernel up to 48 keys (1, 2 and 3 blocks of 16 keys), FFMA kernel beyond and in fp32 mode.  Launch shapes:
    one and two tiles per rollout, more units than SMs (left-over units split into single-tile sub-units), a
    left-over that is not split (151 units on 148 SMs)."""
    sd = state_dict_of(load_golden("rollout_psychometric_d64"))
    model = build_model(sd, "theta", precision)
    model.query_posterior = "off"
    for n_c in ((1, 12, 13, 28, 29, 44, 45, 60) if B <= 5 else (1, 27, 40)):
        if precision == "fp32" and n_c not in (1, 28, 60):
            continue
        _check(model, sd, _batch(B, n_c, nq, dx=1, n_t=4, seed=B), "theta", precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_logits_vs_oracle_cfg2_shape(precision):
    """B = 200 rollouts x 2000 candidates (the benchmarked launch shape: 3200 tiles over 148 CTAs, left-over
    half-units), mid-rollout context (18 points -> 20 keys), and the first (3 keys) / last (37 keys) steps."""
    sd = _location_sd()
    model = build_model(sd, "theta", precision)
    model.query_posterior = "off"
    for n_c in ((1, 18, 35) if precision == "bf16" else (18,)):
        _check(model, sd, _batch(200, n_c, 2000, seed=2), "theta", precision)


@pytest.mark.parametrize("n_c", [1, 14, 30, 46, 47, 60, 78])
def test_general_tc_kernel_vs_oracle(n_c):
    """query_stream_tc_kernel (max-subtracted softmax; the overflow fallback of the fast kernel and the primary path
    above 48 keys) against the oracle directly -- not against the fast kernel."""
    sd = _location_sd()
    model = build_model(sd, "theta", "bf16")
    pm = model.packed()
    if n_c + 2 > pm.tc_max_keys:
        pytest.skip("beyond the general kernel's key limit")
    _check(model, sd, _batch(5, n_c, 700, seed=3), "theta", "bf16", general=True)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("mask", ["absent", "theta", "data", "all", "none"])
def test_masks_vs_oracle_multi_tile(precision, mask):
    """GP-mix model (100 data targets + 3 theta tokens): the candidate rows attend to context + the SELECTED targets
    (utils/target_mask.py 'split' -> theta (3 keys more) or data (100 keys more); 'all' 103; 'none' 0)."""
    g = load_golden("rollout_gpmix_data")
    sd = state_dict_of(g)
    n_td = 100                       # cfg4: 100 data targets (the fixture's weights do not depend on the count)
    n_tok = sd["embedder.theta_tokens"].shape[0]
    n_t = n_td + n_tok
    tm = {"absent": None,
          "theta": torch.cat([torch.zeros(n_td, dtype=torch.bool), torch.ones(n_tok, dtype=torch.bool)]),
          "data": torch.cat([torch.ones(n_td, dtype=torch.bool), torch.zeros(n_tok, dtype=torch.bool)]),
          "all": torch.ones(n_t, dtype=torch.bool), "none": torch.zeros(n_t, dtype=torch.bool)}[mask]
    dx = g["step0/query_x"].shape[-1]
    model = build_model(sd, "mix", precision)
    model.query_posterior = "off"
    gen = torch.Generator().manual_seed(5)
    for B, nq, n_c in ((3, 300, 1), (4, 700, 12), (2, 300, 40)):
        tx = torch.rand(B, n_td, dx, generator=gen) * 2 - 1
        b = _batch(B, n_c, nq, dx=dx, n_t=n_t, seed=7, target_x=tx)
        b["context_x"], b["query_x"] = b["context_x"] * 2 - 1, b["query_x"] * 2 - 1
        _check(model, sd, b, "mix", precision, target_mask=tm)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_free_running_cfg2_rollout_teacher_checked(precision):
    """34 free-running design steps over 2000 candidates (the cfg2 shape) on the resident rollout, re-scored step by step
    by the oracle on the GPU's own trajectory (teacher forcing on the GPU's choices).  With 2000 random-init candidates
    the oracle's own top-2 logit gap is below the arithmetic error bound at most steps (median gap 6e-5 of a 0.03
    spread), so index EQUALITY with a free-running oracle is not a meaningful gate here; what must hold at every step
    and for every rollout is: (i) the chosen design's oracle logit is within 2 x the mode's logit bound of the oracle's
    best (the choice is the argmax up to a documented near-tie), (ii) the step log-prob equals the oracle's log-softmax
    at the chosen index (1e-5 / 1e-3 relative + the logit bound), (iii) the appended (design, outcome) pairs are exactly
    the candidates the indices point at (bit-exact history)."""
    sd = _location_sd()
    model = build_model(sd, "theta", precision)
    B, T = 4, 34
    bound = LOGIT_ABS_FP32 if precision == "fp32" else LOGIT_ABS_BF16
    b = _batch(B, 1, 2000, seed=13)
    out = model.rollout(attr_batch(b), T)
    gi, glp = out.design_idx.cpu(), out.design_log_prob.cpu()
    batch, exact = dict(b), 0
    for t in range(T):
        o = O.forward(sd, batch, "theta", 4, dense=False, with_query_posterior=False)
        lg = o["logits"]
        chosen = lg.gather(1, gi[:, t:t + 1])[:, 0]
        assert ((lg.max(-1).values - chosen) <= 2 * bound).all(), f"step {t}: chosen design is not a near-best"
        exact += int((gi[:, t] == o["idx"][:, 0]).sum())
        ref_lp = torch.log_softmax(lg.double(), -1).gather(1, gi[:, t:t + 1])[:, 0]
        assert (glp[:, t].double() - ref_lp).abs().max().item() < 2 * bound + LOGP_RTOL[precision] * ref_lp.abs().max().item()
        batch = O.update_batch(batch, gi[:, t:t + 1])
    assert abs_err(out.context_x.cpu(), batch["context_x"]) == 0.0
    assert abs_err(out.context_y.cpu(), batch["context_y"]) == 0.0
    if precision == "fp32":
        assert exact >= 0.8 * B * T, f"fp32 mode agrees with the oracle's argmax at only {exact} of {B * T} steps"


def test_select_nan_and_inf_logits_follow_torch_max():
    """A NaN or +inf logit makes every softmax probability NaN; torch.max then returns the first candidate with a NaN
    value (model/head.py:355-358).  The kernel must do the same and must not index out of bounds."""
    from aline_b200 import rollout as ro
    for bad in (float("nan"), float("inf")):
        lg = torch.randn(4, 300, device="cuda")
        lg[1, 17] = bad
        lg[3, 299] = bad
        idx, lp, zt = ro.select(lg)
        ref_zt = torch.softmax(lg.cpu(), -1)
        ref_p, ref_i = torch.max(ref_zt, -1)
        assert torch.equal(idx.cpu()[:, 0], ref_i)
        assert torch.equal(torch.isnan(lp.cpu()), torch.isnan(ref_p.log()))
        ok = ~torch.isnan(ref_p)
        assert rel_err(lp.cpu()[ok], ref_p.log()[ok]) < 1e-5
    # in-place append with a NaN row: the first LIVE candidate is retired, nothing is written out of bounds
    B, nq = 2, 130
    lg = torch.randn(B, nq, device="cuda")
    lg[0, 5] = float("nan")
    alive = torch.ones(B, nq, dtype=torch.uint8, device="cuda")
    alive[0, 0] = 0
    qx, qy = torch.rand(B, nq, 2, device="cuda"), torch.randn(B, nq, 1, device="cuda")
    cx, cy = torch.zeros(B, 3, 2, device="cuda"), torch.zeros(B, 3, 1, device="cuda")
    idx = torch.zeros(B, 2, dtype=torch.int64, device="cuda")
    lp = torch.zeros(B, 2, device="cuda")
    ro.select_append(lg, alive, qx, qy, cx, cy, 1, idx, lp, 0)
    torch.cuda.synchronize()
    assert int(idx[0, 0]) == 0 and int(alive[0, 1]) == 0 and int(alive[0].sum()) == nq - 2
    assert torch.equal(cx[0, 1], qx[0, 1]) and bool(torch.isnan(lp[0, 0]))
    assert int(alive[1].sum()) == nq - 1


_GENERIC_CTX_CHILD = """
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import test_query_parity_gpu as T
sd = T._location_sd()
model = T.build_model(sd, "theta", "bf16")
model.query_posterior = "off"
for n_c in (1, 14, 15, 22, 30, 35):
    T._check(model, sd, T._batch(3, n_c, 300, seed=4), "theta", "bf16")
    T._check(model, sd, T._batch(4, n_c, 2000, seed=5), "theta", "bf16")
print("generic-ctx-ok")
"""


def test_generic_context_kernel_emits_the_folded_operands():
    """ALINE_CTX_KERNEL=head selects the lane-per-head context kernel (rollout.cu ctx_stack_kernel<32>: the path of
    d = 32 models whose feed-forward / embedder widths the warp-per-token kernel does not cover).  It folds the plain
    bf16 operand blocks (query_fast.cuh fold_kv_emit) instead of the fp32 rows in shared memory; the switch is read once
    per process, hence the child process.  Key counts 3 / 16 / 17 / 24 / 32 / 37 = every key padding of the folded
    candidate stream plus the two-threads-per-row kernel."""
    import os
    import subprocess
    import sys
    tests = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(tests)
    env = dict(os.environ, ALINE_CTX_KERNEL="head")
    out = subprocess.run([sys.executable, "-c", _GENERIC_CTX_CHILD.format(root=root, tests=tests)], capture_output=True,
                         text=True, env=env, timeout=600)
    assert out.returncode == 0 and "generic-ctx-ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
