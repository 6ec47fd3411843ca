"""design-steps/s and prior-samples/s for the five configurations BASELINE.json names (cfg1..cfg5),
one GPU, CUDA events, synthetic task draws, random-init weights.  cfg2 is the headline (bench.py); the others are the
parity-test cases, timed here for the record:  python tools/bench_configs.py > profiles/r2_configs.json
Round 2 adds: cfg4 with every target-mask variant ('split' -> theta / data, 'all', 'none': 3 / 100 / 103 / 0 target keys),
cfg5 with the d=64 / h=8 trained variant of notebooks/eval_psychometric.ipynb, and the GP prior sampler against batched
torch.linalg.cholesky (cuSOLVER) on the same GPU."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from aline_b200 import spce  # noqa: E402
from aline_b200.attrdict import AttrDict  # noqa: E402
from aline_b200.model import Aline, Embedder, Encoder, OutputHead  # noqa: E402
from aline_b200.tasks import CESTask, GPTask, HiddenLocation, PsychometricTask  # noqa: E402
from aline_b200.utils.eval import compute_EIG_from_history  # noqa: E402
from aline_b200.utils.target_mask import create_target_mask  # noqa: E402


def timeit(fn, warm=2, it=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def model_for(dx, n_theta, mode, d=32):
    torch.manual_seed(123)
    return Aline(Embedder(dx, 1, d, 128, n_theta, mode), Encoder(d, 128, d // 8, 0.0, 3), OutputHead(dx, 1, d, 128)).cuda().eval()


def cuda_batch(task, B):
    hb = task.sample_batch(B)
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in hb.items()}


def run(name, task, model, B, steps, L=None, target_mask=None, spce_it=3):
    hb = cuda_batch(task, B)
    out = {}

    def roll():
        b = AttrDict(dict(hb))
        if target_mask is not None:
            b.target_mask = target_mask
        out["b"] = model.rollout(b, steps)

    ms = timeit(roll)
    rec = {"B": B, "design_steps": steps, "rollout_ms": ms, "design_steps_per_s": B * steps / ms * 1e3}
    if L:
        b = out["b"]
        x, y = task.unnormalise_design(b.context_x), b.context_y
        th = task.sample_theta((L + 1, B)).cuda()
        th[0] = hb["target_all"].reshape(th[0].shape)
        ms2 = timeit(lambda: spce.spce_history(task.log_likelihood, y, x, th, seq=None, skip_rows=1), warm=1, it=spce_it)
        rec.update({"L": L, "history_points": x.shape[1], "spce_ms": ms2, "prior_samples_per_s": L * B / ms2 * 1e3,
                    "likelihood_evals_per_s": L * B * x.shape[1] / ms2 * 1e3})
        del th
    print(name, rec, file=sys.stderr)
    return rec


def want(tag):
    """`python tools/bench_configs.py cfg5 gp` runs only the named sections; no arguments = everything."""
    return len(sys.argv) < 2 or tag in sys.argv[1:]


def main():
    res = {"gpu": torch.cuda.get_device_name(0), "precision": "bf16 candidate stream (default)", "when": time.strftime("%Y-%m-%d"), "round": 2}
    if want("cfg1"):
        res["cfg1_location_B1000_nq200_T30_L1e4"] = run("cfg1", HiddenLocation(n_query_init=200, design_scale=1),
                                                       model_for(2, 2, "theta"), 1000, 29, L=10_000)
    if want("cfg2"):
        res["cfg2_location_B200_nq2000_T35_L1e6"] = run("cfg2", HiddenLocation(n_query_init=2000, design_scale=1),
                                                       model_for(2, 2, "theta"), 200, 34, L=1_000_000)
    if want("cfg3"):
        res["cfg3_ces_B20_nq2000_T15_L1e7"] = run("cfg3", CESTask(n_context_init=1, n_query_init=2000),
                                                 model_for(6, 5, "theta"), 20, 14, L=10_000_000, spce_it=1)
    gp = GPTask(dim_x=2, embedding_type="mix", n_context_init=1, n_query_init=200, n_target_theta=3, n_target_data=100,
                design_scale=5)
    masks = {"attend_theta": create_target_mask("split", "mix", 100, 3, None, None, None, None, "theta"),
             "attend_data": create_target_mask("split", "mix", 100, 3, None, None, None, None, "data"),
             "all": create_target_mask("all", "mix", 100, 3), "none": torch.zeros(103, dtype=torch.bool)}
    torch.set_default_device("cuda")
    try:
        for tag, tm in (masks.items() if want("cfg4") else ()):
            res["cfg4_gpmix_B200_nq200_T50_" + tag] = run("cfg4 " + tag, gp, model_for(2, 3, "mix"), 200, 50, target_mask=tm)
    finally:
        torch.set_default_device("cpu")
    if not want("gp"):
        return finish(res)
    t0 = time.perf_counter()
    torch.set_default_device("cuda")
    try:
        gp.sample_batch(200)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gp.sample_batch(200)
        torch.cuda.synchronize()
    finally:
        torch.set_default_device("cpu")
    res["cfg4_gp_sample_batch_200x301_ms"] = (time.perf_counter() - t0) * 1e3
    # the GP draw itself: our kernel (kernel matrix + in-smem Cholesky + L z + noise, one block per matrix) against the
    # reference's arithmetic on the same GPU with library kernels: K built with torch ops, batched torch.linalg.cholesky
    # (cuSOLVER potrfBatched), bmm for L z   (tasks/gaussian_process.py:391-415, batched instead of its python loop)
    from aline_b200 import gp as gpk
    Bg, N = 200, 301
    xg = (torch.rand(Bg, N, 2, device="cuda") * 10 - 5)
    ls = torch.rand(Bg, 2, device="cuda") * 1.9 + 0.1
    sc = torch.rand(Bg, device="cuda") * 0.9 + 0.1
    kt = torch.zeros(Bg, dtype=torch.int32, device="cuda")           # rbf
    z, eps = torch.randn(Bg, N, device="cuda"), torch.randn(Bg, N, device="cuda")
    res["gp_kernel_200x301_ms"] = timeit(lambda: gpk.gp_sample(xg, ls, sc, kt, z, eps, 1e-5, 0.01, check=False), warm=2, it=10)

    def torch_gp():
        d = (xg.unsqueeze(2) - xg.unsqueeze(1)) / ls[:, None, None, :]
        K = sc[:, None, None] * torch.exp(-0.5 * (d ** 2).sum(-1)) + 1e-5 * torch.eye(N, device="cuda")
        Lc = torch.linalg.cholesky(K)
        return torch.bmm(Lc, z.unsqueeze(-1)).squeeze(-1) + 0.01 * eps

    res["gp_torch_linalg_cholesky_200x301_ms"] = timeit(torch_gp, warm=2, it=10)
    finish(res)


def finish(res):
    for tag, m in (("FFTT", [False, False, True, True]), ("TTFF", [True, True, False, False])):
        if not want("cfg5"):
            break
        res["cfg5_psychometric_B200_nq200_T30_mask_" + tag] = run(
            "cfg5", PsychometricTask(n_context_init=1, n_query_init=200, design_scale=5), model_for(1, 4, "theta"), 200, 30,
            target_mask=torch.tensor(m))
        res["cfg5_psychometric_d64_h8_B200_nq200_T30_mask_" + tag] = run(
            "cfg5 d64", PsychometricTask(n_context_init=1, n_query_init=200, design_scale=5), model_for(1, 4, "theta", d=64),
            200, 30, target_mask=torch.tensor(m))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
