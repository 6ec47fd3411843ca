"""Generate golden fixtures by running the UNMODIFIED reference in this container.

Run from the repo root (build container only; /root/reference is not on the GPU box):

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own, so the
fixtures below -- outputs of the reference's own classes on seeded inputs -- are
what pins the oracle (`oracle/aline_oracle.py`) and, through it, the CUDA path.
Only this script reads /root/reference; the resulting `*.npz` files are committed.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("ALINE_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


class AttrDict(dict):
    """Stand-in for the `attrdictionary` package (missing in this image)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]


_m = types.ModuleType("attrdictionary")
_m.AttrDict = AttrDict
sys.modules["attrdictionary"] = _m
sys.path.insert(0, REF)


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


from model.base import Aline  # noqa: E402
from model.embedder import Embedder  # noqa: E402
from model.encoder import Encoder  # noqa: E402
from model.head import OutputHead  # noqa: E402
from tasks.location_finding import HiddenLocation  # noqa: E402
from tasks.ces import CESTask  # noqa: E402
from tasks.psychometric import PsychometricTask  # noqa: E402
from tasks.gaussian_process import GPTask  # noqa: E402
from loss.eig import PCELoss, NMCLoss  # noqa: E402

ref_eval = _load("ref_eval", "utils/eval.py")          # utils/__init__ imports hydra (missing)
ref_tmask = _load("ref_tmask", "utils/target_mask.py")


def npy(t):
    return t.detach().cpu().numpy()


def build_model(dx, n_theta, mode, d=32, ff=128, h=4, layers=3, time_token=False, value_head=False):
    return Aline(Embedder(dx, 1, d, ff, n_theta, mode), Encoder(d, ff, h, 0.0, layers),
                 OutputHead(dx, 1, d, ff, time_token=time_token, value_head=value_head)).eval()


def record_rollout(name, task, model, B, steps, target_mask=None, sharpen=1.0, extra=None, time_T=None):
    """Teacher-forced per-step records of Aline.forward + Task.update_batch."""
    out = {}
    if sharpen != 1.0:
        with torch.no_grad():
            model.head.acquisition_head.predictor[2].weight.mul_(sharpen)
    for k, v in model.state_dict().items():
        out["sd/" + k] = npy(v)
    logits_box = {}
    hook = model.head.acquisition_head.predictor[2].register_forward_hook(
        lambda m, i, o: logits_box.__setitem__("v", o.squeeze(-1).detach().clone()))
    batch = task.sample_batch(B)
    if target_mask is not None:
        batch.target_mask = target_mask
        out["target_mask"] = npy(target_mask)
    out["target_all"] = npy(batch.target_all)
    if batch.get("target_x", None) is not None:
        out["target_x"] = npy(batch.target_x)
    with torch.no_grad():
        for t in range(steps):
            pre = f"step{t}/"
            for k in ("context_x", "context_y", "query_x", "query_y"):
                out[pre + k] = npy(batch[k])
            if time_T is not None:
                batch.t = torch.tensor([(time_T - t) / time_T])        # utils/eval.py:25-26
                out[pre + "t"] = npy(batch.t)
            pred = model.forward(batch)
            out[pre + "zt"] = npy(pred.design_out.zt)
            out[pre + "idx"] = npy(pred.design_out.idx)
            out[pre + "log_prob"] = npy(pred.design_out.log_prob)
            out[pre + "logits"] = npy(logits_box["v"])
            if "value" in pred:
                out[pre + "value"] = npy(pred.value)
            for k in ("mixture_means", "mixture_stds", "mixture_weights"):
                out[pre + "post/" + k] = npy(pred.posterior_out[k])
                out[pre + "postq/" + k] = npy(pred.posterior_out_query[k])
            out[pre + "target_ll"] = npy(ref_eval.compute_ll(
                batch.target_all, pred.posterior_out.mixture_means,
                pred.posterior_out.mixture_stds, pred.posterior_out.mixture_weights))
            batch = task.update_batch(batch, pred.design_out.idx)
        for k in ("context_x", "context_y", "query_x", "query_y"):
            out["final/" + k] = npy(batch[k])
    hook.remove()
    out["n_steps"] = np.int64(steps)
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, sum(v.nbytes for v in out.values()) // 1024, "KiB")


def gen_models():
    torch.manual_seed(123)
    task = HiddenLocation(n_query_init=24, design_scale=1)
    record_rollout("rollout_location", task, build_model(2, 2, "theta"), B=4, steps=4)

    torch.manual_seed(124)
    task = HiddenLocation(n_query_init=40, design_scale=1)
    record_rollout("rollout_location_sharp", task, build_model(2, 2, "theta"), B=6, steps=5, sharpen=100.0)

    torch.manual_seed(134)
    task = HiddenLocation(n_query_init=24, design_scale=1)
    if "tt" in os.environ.get("ALINE_GOLDEN_ONLY", "tt"):
        record_rollout("rollout_location_tt", task, build_model(2, 2, "theta", time_token=True), B=4, steps=4, time_T=4)
    torch.manual_seed(135)
    task = HiddenLocation(n_query_init=24, design_scale=1)
    if "value" in os.environ.get("ALINE_GOLDEN_ONLY", "value"):
        record_rollout("rollout_location_value", task, build_model(2, 2, "theta", value_head=True), B=4, steps=4)

    # the two other shipped active-learning task configs (config/task/al_data.yaml, al_theta.yaml), 1-D
    torch.manual_seed(136)
    task = GPTask(dim_x=1, embedding_type="data", n_context_init=1, n_query_init=20, n_target_theta=0, n_target_data=10,
                  design_scale=5)
    if "al" in os.environ.get("ALINE_GOLDEN_ONLY", "al"):
        record_rollout("rollout_gp_data", task, build_model(1, 0, "data"), B=3, steps=3)
    torch.manual_seed(137)
    task = GPTask(dim_x=1, embedding_type="theta", n_context_init=1, n_query_init=20, n_target_theta=2, n_target_data=0,
                  design_scale=5)
    if "al" in os.environ.get("ALINE_GOLDEN_ONLY", "al"):
        record_rollout("rollout_gp_theta", task, build_model(1, 2, "theta"), B=3, steps=3)
    if os.environ.get("ALINE_GOLDEN_ONLY"):
        return

    torch.manual_seed(125)
    task = CESTask(n_context_init=1, n_query_init=24)
    record_rollout("rollout_ces", task, build_model(6, 5, "theta"), B=4, steps=3)

    for tag, tm in (("a", [False, False, True, True]), ("b", [True, True, False, False])):
        torch.manual_seed(126)
        task = PsychometricTask(n_context_init=1, n_query_init=24)
        record_rollout("rollout_psychometric_" + tag, task, build_model(1, 4, "theta"), B=4, steps=3,
                       target_mask=torch.tensor(tm))
    torch.manual_seed(127)
    task = PsychometricTask(n_context_init=2, n_query_init=16)
    record_rollout("rollout_psychometric_d64", task, build_model(1, 4, "theta", d=64, h=8), B=3, steps=3,
                   target_mask=torch.tensor([True, True, False, False]))

    for attend in ("data", "theta", "all", "none", None):
        torch.manual_seed(128)
        task = GPTask(dim_x=2, embedding_type="mix", n_context_init=1, n_query_init=20,
                      n_target_theta=3, n_target_data=10, design_scale=5)
        if attend in ("data", "theta"):
            tm = ref_tmask.create_target_mask("split", "mix", 10, 3, None, None, None, None, attend)
        elif attend in ("all", "none"):
            tm = ref_tmask.create_target_mask(attend, "mix", 10, 3, None, None, None, None, None)
        else:
            tm = None        # attribute absent -> queries attend to all targets (encoder.py:122-124)
        record_rollout(f"rollout_gpmix_{attend or 'absent'}", task, build_model(2, 3, "mix"), B=3, steps=3,
                       target_mask=tm)


def gen_traces():
    """Free-running get_traces + compute_EIG_from_history (location)."""
    torch.manual_seed(129)
    task = HiddenLocation(n_query_init=30, design_scale=1)
    model = build_model(2, 2, "theta")
    out = {"sd/" + k: npy(v) for k, v in model.state_dict().items()}
    # replicate get_traces but keep the initial batch (utils/eval.py:9-39)
    torch.manual_seed(130)
    task.sample_theta((5))          # get_traces draws this first (utils/eval.py:19)
    batch0 = task.sample_batch(5)
    for k in ("context_x", "context_y", "query_x", "query_y", "target_all"):
        out["batch0/" + k] = npy(batch0[k])
    torch.manual_seed(130)
    theta_0, x, y = ref_eval.get_traces(model, task, T=6, batch_size=5)
    out.update(theta_0=npy(theta_0), x=npy(x), y=npy(y))
    np.savez_compressed(os.path.join(OUT, "traces_location.npz"), **out)
    print("traces_location")


def gen_spce():
    # location K=1 and K=2, CES: compute_EIG_from_history with known thetas.
    def run(name, task, theta_0, x, y, L, seed):
        B = x.shape[0]
        torch.manual_seed(seed)
        thetas = torch.concat([theta_0.unsqueeze(0), task.sample_theta((L, B))], dim=0)
        torch.manual_seed(seed)   # compute_EIG_from_history redraws the same thetas (eval.py:61-62)
        pce, nmc = ref_eval.compute_EIG_from_history(task, theta_0, x, y, L=L, batch_size=B, stepwise=True)
        torch.manual_seed(seed)
        pce_last, nmc_last = ref_eval.compute_EIG_from_history(task, theta_0, x, y, L=L, batch_size=B, stepwise=False)
        # whole-history losses (loss/eig.py:55-151)
        T = x.shape[1]
        pl = PCELoss(L, T, task.log_likelihood, reduction=None)(y, x, thetas)
        nl = NMCLoss(L, T, task.log_likelihood, reduction=None)(y, x, thetas)
        # per-term log-likelihoods for the first two history points
        ll01 = torch.stack([task.log_likelihood(y[:, t].unsqueeze(0), x[:, t].unsqueeze(0), thetas).squeeze(-1)
                            for t in range(min(T, 2))], 0)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), theta_0=npy(theta_0), x=npy(x), y=npy(y),
                            thetas=npy(thetas), pce=npy(pce), nmc=npy(nmc), pce_last=npy(pce_last),
                            nmc_last=npy(nmc_last), pce_loss=npy(pl), nmc_loss=npy(nl), ll01=npy(ll01))
        print(name, "pce", pce[0, -1].item(), "nmc", nmc[0, -1].item())

    for K in (1, 2):
        torch.manual_seed(140 + K)
        task = HiddenLocation(K=K, n_target_theta=2 * K, n_query_init=1, design_scale=1)
        B, T, L = 8, 7, 511
        theta_0 = task.sample_theta(B)
        x = task.sample_data(B, T)
        y = task.forward(x, theta_0.unsqueeze(1).expand(B, T, K, 2))
        run(f"spce_location_k{K}", task, theta_0, x, y, L, 150 + K)

    torch.manual_seed(160)
    task = CESTask(n_context_init=1, n_query_init=1)
    B, T, L = 8, 12, 1023
    theta_0 = task.sample_theta(B)
    x = task.sample_data(B, T)
    y = task.forward(x, theta_0.unsqueeze(1))
    lo, hi = task.epsilon, 1 - task.epsilon
    yy = y.squeeze(-1)
    print("ces censor census: lo", (yy == yy.new_tensor(lo)).sum().item(), "hi", (yy == yy.new_tensor(hi)).sum().item(),
          "interior", ((yy > lo) & (yy < hi)).sum().item())
    run("spce_ces", task, theta_0, x, y, L, 161)

    # psychometric: elementwise log_likelihood on [B,4,1] inputs only (sPCE unsupported upstream)
    torch.manual_seed(170)
    task = PsychometricTask()
    B = 64
    theta = task.sample_theta(B)
    xs = task.sample_data(B, 1)[:, 0]
    ys = task.forward(xs, theta)
    ll = task.log_likelihood(ys, xs, theta)
    np.savez_compressed(os.path.join(OUT, "loglik_psychometric.npz"), theta=npy(theta), x=npy(xs), y=npy(ys), ll=npy(ll))
    print("loglik_psychometric")


def gen_spce_large():
    """CES bound at L = 1e5 (B = 20, T = 15, the cfg3 shape at 1 % of its L): the L = 1023 fixture above is decided by
    fp32 round-off of single terms (pow amplified by 1 / rho <= 100; half of its sPCE entries are saturated at
    log(L + 1)), so it cannot referee a change of the power arithmetic.  The 1e5 x 20 x 5 contrastive draws are not
    stored (40 MB): the test redraws them on the CPU with the same seed through the mirror task, whose sample_theta
    consumes torch's generator exactly like the reference's (checked here, checksum stored)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from aline_b200.tasks import CESTask as MirrorCES
    torch.manual_seed(260)
    task = CESTask(n_context_init=1, n_query_init=1)
    B, T, L, seed = 20, 15, 100_000, 261
    theta_0 = task.sample_theta(B)
    x = task.sample_data(B, T)
    y = task.forward(x, theta_0.unsqueeze(1))
    torch.manual_seed(seed)
    th_ref = task.sample_theta((L, B))
    torch.manual_seed(seed)
    th_mirror = MirrorCES(n_context_init=1, n_query_init=1).sample_theta((L, B))
    assert torch.equal(th_ref, th_mirror), "the mirror task does not reproduce the reference's prior draws"
    torch.manual_seed(seed)
    pce, nmc = ref_eval.compute_EIG_from_history(task, theta_0, x, y, L=L, batch_size=B, stepwise=True)
    yy = y.squeeze(-1)
    lo, hi = task.epsilon, 1 - task.epsilon
    print("ces large censor census: lo", (yy == yy.new_tensor(lo)).sum().item(), "hi", (yy == yy.new_tensor(hi)).sum().item(),
          "interior", ((yy > lo) & (yy < hi)).sum().item(), "saturated pce entries",
          int((pce > np.log(L + 1) - 1e-3).sum()), "of", pce.numel())
    # the same bound with the reference's own code in float64 (same fp32 draws, promoted): the yardstick for "how far is an
    # fp32 evaluation from the exact value of the reference's formula" -- the reference's fp32 run itself is ~1e-2 away
    torch.manual_seed(seed)
    pce64, nmc64 = ref_eval.compute_EIG_from_history(task, theta_0.double(), x.double(), y.double(), L=L, batch_size=B,
                                                     stepwise=True)
    print("reference fp32 vs its own fp64: max abs", float((pce.double() - pce64).abs().max()), "mean signed",
          float((pce.double() - pce64).mean()))
    np.savez_compressed(os.path.join(OUT, "spce_ces_large.npz"), theta_0=npy(theta_0), x=npy(x), y=npy(y), pce=npy(pce),
                        nmc=npy(nmc), pce64=npy(pce64), nmc64=npy(nmc64), seed=np.int64(seed), L=np.int64(L),
                        thetas_checksum=np.float64(th_ref.double().sum().item()))
    print("spce_ces_large pce", pce[0, -1].item(), "nmc", nmc[0, -1].item())


def gen_gp():
    out = {}
    torch.manual_seed(180)
    task = GPTask(dim_x=2, embedding_type="mix", n_context_init=1, n_query_init=40, n_target_theta=3,
                  n_target_data=20, design_scale=5)
    B, N = 9, 61
    theta = task.sample_theta(B)
    x = task.sample_data(B, N)
    torch.manual_seed(181)
    y = task.generate_gp_data(x, theta)
    torch.manual_seed(181)          # replay the RNG order of generate_gp_data (gaussian_process.py:384-415)
    ktypes = task.sample_kernel_type(B)
    zs, es, Ks, Ls = [], [], [], []
    for b in range(B):
        K = task.compute_kernel_matrix(x[b], x[b], theta[b, :2, 0], theta[b, 2, 0], ktypes[b]) + task.jitter * torch.eye(N)
        Ks.append(K)
        Ls.append(torch.linalg.cholesky(K))
        zs.append(torch.randn(N))
        es.append(torch.randn(N))
    z, eps = torch.stack(zs), torch.stack(es)
    y2 = torch.stack([Ls[b] @ z[b] + task.noise_scale * eps[b] for b in range(B)]).unsqueeze(-1)
    assert torch.equal(y, y2), "RNG replay of generate_gp_data failed"
    out.update(x=npy(x), theta=npy(theta), ktype=np.array([task.kernel_types.index(k) for k in ktypes], dtype=np.int32),
               z=npy(z), eps=npy(eps), K=npy(torch.stack(Ks)), L=npy(torch.stack(Ls)), y=npy(y))
    # also one of each kernel type incl. matern12 (weight 0 in the task, still a supported kernel)
    for kt in task.kernel_types:
        out["K_" + kt] = npy(task.compute_kernel_matrix(x[0], x[0], theta[0, :2, 0], theta[0, 2, 0], kt))
    np.savez_compressed(os.path.join(OUT, "gp_draws.npz"), **out)
    print("gp_draws", ktypes)


def gen_masks():
    """Encoder.create_mask truth table (model/encoder.py:83-126)."""
    enc = Encoder(32, 128, 4, 0.0, 1)
    out = {}
    base = AttrDict(context_x=torch.zeros(1, 2, 1), query_x=torch.zeros(1, 3, 1), target_all=torch.zeros(1, 4, 1))
    cases = {"absent": "absent", "none_attr": None, "all_true": [True] * 4, "all_false": [False] * 4,
             "predef": [False, False, True, True], "mixed": [True, False, True, False]}
    for k, tm in cases.items():
        b = AttrDict(base)
        if tm != "absent":
            b.target_mask = None if tm is None else torch.tensor(tm)
        out["mask_" + k] = npy(enc.create_mask(b))
        out["tm_" + k] = np.array([] if tm in ("absent", None) else tm, dtype=bool)
    np.savez_compressed(os.path.join(OUT, "mask_truth.npz"), **out)
    print("mask_truth")


def gen_train():
    """Row f2: the reference's grad-enabled forward (model.train(): Categorical-sampled designs, the _sa_block slow
    path) driven through the T-step inner loop of train_aline.py:80-132; records the sampled indices (so that the test
    can teacher-force them), the loss and the gradient of every parameter."""
    sys.path.insert(0, os.path.dirname(OUT))
    from _util import train_inner_loop

    def record(name, task, model, B, T, target_mask=None, mix_n_theta=0):
        torch.manual_seed(7)
        model.train()
        batch = task.sample_batch(B)
        out = {"sd/" + k: npy(v) for k, v in model.state_dict().items()}
        for k in ("context_x", "context_y", "query_x", "query_y", "target_all"):
            out["batch0/" + k] = npy(batch[k])
        if batch.get("target_x", None) is not None:
            out["batch0/target_x"] = npy(batch.target_x)
        if target_mask is not None:
            batch.target_mask = target_mask
            out["target_mask"] = npy(target_mask)
        model.zero_grad()
        loss, dl, pl, idx = train_inner_loop(model, task.update_batch, ref_eval.compute_ll, ref_tmask.select_targets_by_mask,
                                             batch, T, mix_n_theta=mix_n_theta)
        loss.backward()
        out["idx"] = npy(idx)
        out["loss"], out["design_loss"], out["predict_loss"] = npy(loss), npy(dl), npy(pl)
        for k, p in model.named_parameters():
            out["grad/" + k] = npy(p.grad if p.grad is not None else torch.zeros_like(p))
        out["T"] = np.int64(T)
        out["mix_n_theta"] = np.int64(mix_n_theta)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, sum(v.nbytes for v in out.values()) // 1024, "KiB", float(loss))

    torch.manual_seed(321)
    record("train_location", HiddenLocation(n_query_init=40, design_scale=1), build_model(2, 2, "theta"), B=6, T=4)
    torch.manual_seed(322)
    gp = GPTask(dim_x=2, embedding_type="mix", n_context_init=1, n_query_init=30, n_target_theta=3, n_target_data=10,
                design_scale=5)
    tm = torch.cat([torch.zeros(10, dtype=torch.bool), torch.ones(3, dtype=torch.bool)])
    record("train_gpmix_theta", gp, build_model(2, 3, "mix"), B=5, T=3, target_mask=tm, mix_n_theta=3)


def gen_summary():
    """The aggregation of eval_boed / eval_EIG_from_history (utils/eval.py:84-198: mean, unbiased std, 'se' / 'ci' /
    'std') on known per-rollout bounds: the reference's compute_EIG_from_history / get_traces are replaced by functions
    that hand back slices of stored arrays, so only the loop + statistics of the unmodified reference run."""
    torch.manual_seed(99)
    M, T, bs = 50, 7, 16
    pce_all, nmc_all = torch.randn(M, T) + 3.0, torch.randn(M, T) * 1.5 + 3.5
    out = {"pce_all": npy(pce_all), "nmc_all": npy(nmc_all), "batch_size": np.int64(bs)}
    orig = ref_eval.compute_EIG_from_history
    try:
        def fake(experiment, theta_0, x, y, L, batch_size, stepwise):
            i0 = int(x[0, 0, 0])
            return pce_all[i0:i0 + batch_size], nmc_all[i0:i0 + batch_size]
        ref_eval.compute_EIG_from_history = fake
        x = torch.arange(M, dtype=torch.float32).reshape(M, 1, 1).expand(M, T, 1).contiguous()    # row index rides in x
        for et in ("se", "ci", "std"):
            b = ref_eval.eval_EIG_from_history(None, torch.zeros(M, 2), x, x, L=10, M=M, batch_size=bs, stepwise=True,
                                               err_type=et)
            for k in ("pce_mean", "pce_err", "nmc_mean", "nmc_err"):
                out[f"history/{et}/{k}"] = npy(b[k])
        # eval_boed: whole mini-batches (ceil(M / bs) * bs rollouts)
        state = {"i": 0}
        orig_traces = ref_eval.get_traces
        n_steps = (M + bs - 1) // bs
        pce_b, nmc_b = torch.randn(n_steps * bs, T) + 2.0, torch.randn(n_steps * bs, T) + 2.5
        out["pce_boed"], out["nmc_boed"] = npy(pce_b), npy(nmc_b)

        def fake_traces(model, experiment, T_, batch_size, time_token):
            i0 = state["i"]
            state["i"] += batch_size
            xx = torch.full((batch_size, 1, 1), float(i0))
            return torch.zeros(batch_size, 2), xx, xx

        def fake2(experiment, theta_0, x, y, L, batch_size, stepwise):
            i0 = int(x[0, 0, 0])
            return pce_b[i0:i0 + batch_size], nmc_b[i0:i0 + batch_size]

        class _M:
            def eval(self):
                return self
        ref_eval.get_traces, ref_eval.compute_EIG_from_history = fake_traces, fake2
        for et in ("se", "ci", "std"):
            state["i"] = 0
            b = ref_eval.eval_boed(_M(), None, T=T, L=10, M=M, batch_size=bs, stepwise=True, err_type=et)
            for k in ("pce_mean", "pce_err", "nmc_mean", "nmc_err"):
                out[f"boed/{et}/{k}"] = npy(b[k])
        ref_eval.get_traces = orig_traces
    finally:
        ref_eval.compute_EIG_from_history = orig
    np.savez_compressed(os.path.join(OUT, "summary_stats.npz"), **out)
    print("summary_stats", sum(v.nbytes for v in out.values()) // 1024, "KiB")


def _ref_function(rel, name):
    """One function of a reference module whose top-level imports are not satisfiable here (hydra / omegaconf):
    its source segment is read from the reference file and executed unmodified."""
    import ast
    src = open(os.path.join(REF, rel)).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch, "np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), os.path.join(REF, rel), "exec"), ns)
    return ns[name]


def gen_uncertainty():
    """Uncertainty-sampling baseline (notebooks/eval_al.ipynb cell 1, acquisition "uncertainty_sampling"): free-running
    loop of the reference's model.forward + utils/misc.py:calculate_gmm_variance + argmax + Task.update_batch."""
    calc_var = _ref_function("utils/misc.py", "calculate_gmm_variance")
    torch.manual_seed(133)
    task = GPTask(dim_x=2, embedding_type="mix", n_context_init=1, n_query_init=40, n_target_theta=3,
                  n_target_data=10, design_scale=5)
    model = build_model(2, 3, "mix")
    with torch.no_grad():      # spread the GMM heads so that the variance ranking is not a near-tie
        for h in model.head.target_head.heads:
            h[2].weight.mul_(8.0)
    out = {"sd/" + k: npy(v) for k, v in model.state_dict().items()}
    batch = task.sample_batch(4)
    batch.target_mask = None
    for k in ("context_x", "context_y", "query_x", "query_y", "target_all", "target_x"):
        out["in/" + k] = npy(batch[k])
    steps = 5
    with torch.no_grad():
        for t in range(steps):
            pred = model.forward(batch)
            pq = pred.posterior_out_query
            var = calc_var(pq.mixture_means, pq.mixture_stds, pq.mixture_weights)
            out[f"step{t}/var"] = npy(var)
            out[f"step{t}/var_shared_w"] = npy(calc_var(pq.mixture_means, pq.mixture_stds, pq.mixture_weights[:, 0]))
            for k in ("mixture_means", "mixture_stds", "mixture_weights"):
                out[f"step{t}/postq/" + k] = npy(pq[k])
            idx = torch.argmax(var, dim=1, keepdim=True)
            out[f"step{t}/idx"] = npy(idx)
            batch = task.update_batch(batch, idx)
        for k in ("context_x", "context_y"):
            out["final/" + k] = npy(batch[k])
    out["n_steps"] = np.int64(steps)
    np.savez_compressed(os.path.join(OUT, "uncertainty_gpmix.npz"), **out)
    print("uncertainty_gpmix", sum(v.nbytes for v in out.values()) // 1024, "KiB")


if __name__ == "__main__":
    torch.set_default_dtype(torch.float32)
    only = sys.argv[1:]
    if only:
        for fn in only:
            globals()["gen_" + fn]()
        sys.exit(0)
    gen_models()
    gen_traces()
    gen_spce()
    gen_spce_large()
    gen_gp()
    gen_masks()
    gen_uncertainty()
    gen_train()
    gen_summary()
