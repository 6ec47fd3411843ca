"""GP prior draws (cfg4: 200 draws of N = 301 points, 2-D): aline_gp_sample against the reference's arithmetic with library
kernels on the same GPU (K from torch ops, batched torch.linalg.cholesky = cuSOLVER potrfBatched, bmm for L z)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from aline_b200 import gp as gpk  # noqa: E402


def timeit(fn, warm=3, it=20):
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


torch.manual_seed(0)
out = {}
for Bg, N in ((200, 301), (148, 301), (1000, 301), (200, 203)):
    xg = (torch.rand(Bg, N, 2, device="cuda") * 10 - 5)
    ls = torch.rand(Bg, 2, device="cuda") * 1.9 + 0.1
    sc = torch.rand(Bg, device="cuda") * 0.9 + 0.1
    for name, kt_v in (("rbf", 0), ("matern52", 3)):
        kt = torch.full((Bg,), kt_v, dtype=torch.int32, device="cuda")
        z, eps = torch.randn(Bg, N, device="cuda"), torch.randn(Bg, N, device="cuda")

        def torch_gp():
            d = (xg.unsqueeze(2) - xg.unsqueeze(1)) / ls[:, None, None, :]
            sq = (d ** 2).sum(-1)
            if kt_v == 0:
                K = sc[:, None, None] * torch.exp(-0.5 * sq)
            else:
                r = sq.sqrt() * 2.2360679774997898
                K = sc[:, None, None] * (1 + r + r * r / 3) * torch.exp(-r)
            K = K + 1e-5 * torch.eye(N, device="cuda")
            Lc = torch.linalg.cholesky(K)
            return torch.bmm(Lc, z.unsqueeze(-1)).squeeze(-1) + 0.01 * eps

        y = gpk.gp_sample(xg, ls, sc, kt, z, eps, 1e-5, 0.01)
        y_ref = torch_gp()
        rel = float(((y - y_ref).norm(dim=1) / y_ref.norm(dim=1)).max())
        out[f"B{Bg}_N{N}_{name}"] = {"aline_gp_sample_ms": timeit(lambda: gpk.gp_sample(xg, ls, sc, kt, z, eps, 1e-5, 0.01, check=False)),
                                     "torch_linalg_cholesky_ms": timeit(torch_gp), "max_rel_diff_of_draws": rel}
        print(Bg, N, name, out[f"B{Bg}_N{N}_{name}"], file=sys.stderr)
print(json.dumps(out, indent=1))
