"""One candidate-query launch shape for ncu captures: python tools/one_query.py <n_c> <tc4|tc3> [B] [nq]"""
import sys

import torch

sys.path.insert(0, ".")
from aline_b200 import _lib, rollout as ro  # noqa: E402

_lib.set_option("query_tc4", 1)
from aline_b200.model import Aline, Embedder, Encoder, OutputHead  # noqa: E402

n_c = int(sys.argv[1]) if len(sys.argv) > 1 else 18
which = sys.argv[2] if len(sys.argv) > 2 else "tc4"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 200
nq = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
torch.manual_seed(123)
model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128)).cuda().eval()
pm = model.packed()
qx = torch.rand(B, nq, 2, device="cuda")
eq, eq_rm = ro.embed_queries(pm, qx, row_major=True)
slots, n_sel = ro.target_slots(2, None, "cuda")
cx, cy = torch.rand(B, n_c, 2, device="cuda"), torch.randn(B, n_c, 1, device="cuda")
nk = n_c + n_sel
tc_kv = ro.alloc_tc_kv(pm, B, nk, "cuda")
kv, _ = ro.ctx_stack(pm, cx, cy, n_c, None, slots, n_sel, want_z=False, tc_kv=tc_kv)
alive = torch.ones((B, nq), dtype=torch.uint8, device="cuda")
alive[:, : n_c - 1] = 0
for _ in range(4):
    ro.query_stream(pm, eq, alive, kv, nk, precision="bf16", tc_kv=tc_kv, eq_rm=eq_rm if which == "tc4" else None)
torch.cuda.synchronize()
print("ok")
