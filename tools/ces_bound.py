"""One cfg3 CES bound (B=20, T=15, L=1e7 contrastive rows, fused-history kernel) on random histories, for ncu captures:
    python tools/ces_bound.py [L]"""
import sys

import torch

sys.path.insert(0, ".")
from aline_b200 import prior as dprior, spce  # noqa: E402
from aline_b200.tasks import CESTask  # noqa: E402

L = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
B, T = 20, 15
dev = torch.device("cuda")
task = CESTask(n_context_init=1, n_query_init=T)
b = dprior.sample_batch_device(task, B, seed=5, device=dev)
x, y = task.unnormalise_design(b["query_x"]), b["query_y"]
rows = dprior.sample_theta_device(task, L + 1, B, seed=6, device=dev)
rows[0] = b["target_all"].reshape(B, -1)
for _ in range(2):
    m, s, lp0 = spce.spce_history(task.log_likelihood, y, x, rows, seq=None, skip_rows=1, check=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
m, s, lp0 = spce.spce_history(task.log_likelihood, y, x, rows, seq=None, skip_rows=1, check=False)
e1.record()
torch.cuda.synchronize()
pce = spce.lse_combine(m, s, lp0)
print("ok", L, "ms", e0.elapsed_time(e1), "pce[0,-1]", float(pce[0][0, -1]) if isinstance(pce, tuple) else float(pce.flatten()[-1]))
