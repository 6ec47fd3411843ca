"""cfg2 bound evaluation (B=200, T=35, L=1e6) end to end through compute_EIG_from_history: torch-generator draws
(sample_theta + fused pass reading thetas) vs device-side Philox draws inside the fused pass."""
import json
import sys

import torch

sys.path.insert(0, ".")
from aline_b200.tasks import HiddenLocation  # noqa: E402
from aline_b200.utils.eval import compute_EIG_from_history  # noqa: E402


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


B, T, L = 200, 35, 1_000_000
task = HiddenLocation(design_scale=1)
torch.manual_seed(0)
theta0 = torch.rand(B, 1, 2, device="cuda")
x = torch.rand(B, T, 2, device="cuda")
y = torch.log(0.1 + 1.0 / (1e-4 + ((x - theta0) ** 2).sum(-1, keepdim=True))) + 0.5 * torch.randn(B, T, 1, device="cuda")
out = {}
with torch.device("cuda"):
    out["torch_prior_ms"] = timeit(lambda: compute_EIG_from_history(task, theta0, x, y, L=L, batch_size=B, stepwise=True))
out["device_prior_ms"] = timeit(lambda: compute_EIG_from_history(task, theta0, x, y, L=L, batch_size=B, stepwise=True,
                                                                prior="device", seed=1))
print(json.dumps(out))
