"""query_tc3 alone at several batch sizes / key counts (tail and hand-over experiments):
    python tools/bench_query_tc3.py B1,B2,... n_c1,n_c2,..."""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from aline_b200 import _lib, rollout as ro  # noqa: E402
from aline_b200.model import Aline, Embedder, Encoder, OutputHead  # noqa: E402
from bench_query import timeit  # noqa: E402

Bs = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [200]
ncs = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [18]
nq = 2000
torch.manual_seed(123)
model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128)).cuda().eval()
pm = model.packed()
slots, n_sel = ro.target_slots(2, None, "cuda")
_lib.set_option("query_tc4", 0)
for B in Bs:
    qx = torch.rand(B, nq, 2, device="cuda")
    eq = ro.embed_queries(pm, qx)
    for n_c in ncs:
        cx, cy = torch.rand(B, n_c, 2, device="cuda"), torch.randn(B, n_c, 1, device="cuda")
        nk = n_c + n_sel
        tc_kv = ro.alloc_tc_kv(pm, B, nk, "cuda")
        kv, _ = ro.ctx_stack(pm, cx, cy, n_c, None, slots, n_sel, want_z=False, tc_kv=tc_kv)
        alive = torch.ones((B, nq), dtype=torch.uint8, device="cuda")
        alive[:, : n_c - 1] = 0
        us = timeit(lambda: ro.query_stream(pm, eq, alive, kv, nk, precision="bf16", tc_kv=tc_kv), warm=5, it=50)
        print(f"B={B} keys={nk} tc3_us={us:.1f} us_per_100_rollouts={us / B * 100:.1f}")
