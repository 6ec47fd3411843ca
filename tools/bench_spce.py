"""Micro-benchmark of the sPCE kernels at cfg2 / cfg3 sizes (CUDA events, L2 exceeded by the inputs)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from aline_b200 import spce  # noqa: E402
from aline_b200.tasks import HiddenLocation, CESTask  # noqa: E402


def timeit(fn, warm=2, it=5):
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


def main():
    out = {}
    L, B, T = 1_000_000, 200, 35
    task = HiddenLocation(design_scale=1)
    th = torch.rand(L + 1, B, 1, 2, device="cuda")
    x = torch.rand(B, T, 2, device="cuda")
    d2 = ((x - th[0]) ** 2).sum(-1, keepdim=True)                   # outcomes simulated from theta_0 = row 0
    y = torch.log(0.1 + 1.0 / (1e-4 + d2)) + 0.5 * torch.randn(B, T, 1, device="cuda")
    seq = torch.zeros(L + 1, B, device="cuda")
    ms = timeit(lambda: spce.spce_step(task.log_likelihood, y[:, 0], x[:, 0], th, seq))
    by = (L + 1) * B * (4 * 2 + 8)
    out["loc_step_ms"] = ms
    out["loc_step_GBs"] = by / ms / 1e6
    ms = timeit(lambda: spce.spce_history(task.log_likelihood, y, x, th, seq=None))
    out["loc_hist35_ms"] = ms
    out["loc_hist35_prior_samples_per_s"] = L * B / ms * 1e3
    out["loc_hist35_lik_evals_per_s"] = L * B * T / ms * 1e3
    out["loc_hist35_last_only_ms"] = timeit(lambda: spce.spce_history(task.log_likelihood, y, x, th, seq=None, last_only=True))
    del th, seq
    if "--loc-only" in sys.argv:
        print(json.dumps(out))
        return
    L, B, T = 2_000_000, 20, 15
    task = CESTask()
    th = task.sample_theta((L + 1, B)).cuda()
    x = (torch.rand(B, T, 6) * 100).cuda()
    th0 = th[0]
    y = task.forward(x.cpu(), th0.cpu().unsqueeze(1)).cuda()
    seq = torch.zeros(L + 1, B, device="cuda")
    ms = timeit(lambda: spce.spce_step(task.log_likelihood, y[:, 0], x[:, 0], th, seq))
    out["ces_step_ms_L2e6"] = ms
    out["ces_step_GBs"] = (L + 1) * B * (4 * 5 + 8) / ms / 1e6
    seq.zero_()
    ms = timeit(lambda: spce.spce_history(task.log_likelihood, y, x, th, seq=None), warm=1, it=3)
    out["ces_hist15_ms_L2e6"] = ms
    out["ces_hist15_prior_samples_per_s"] = L * B / ms * 1e3
    print(json.dumps(out))


if __name__ == "__main__":
    main()
