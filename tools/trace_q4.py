"""Per-phase clock stamps of the two-threads-per-row candidate stream (instrumented build):
    ALINE_BUILD_TRACE=1 python -m aline_b200.build
    ALINE_B200_LIB=aline_b200/lib/libaline_b200_trace.so python tools/trace_q4.py [n_c]
Stamps of two epilogue warps' mma_phase: 0 epilogue done, 1 TMEM stores complete, 2 proxy fence done, 3 arrived on the
group's "operands ready" barrier, 5 MMA completion observed; of the issuer: 6 a group's operands seen ready, 7 its MMAs
issued + committed."""
import ctypes
import subprocess
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from aline_b200 import _lib, rollout as ro  # noqa: E402
from aline_b200.model import Aline, Embedder, Encoder, OutputHead  # noqa: E402

n_c = int(sys.argv[1]) if len(sys.argv) > 1 else 18
B, nq = 200, 2000
torch.manual_seed(123)
model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128)).cuda().eval()
pm = model.packed()
_lib.set_option("query_tc4", 1)
qx = torch.rand(B, nq, 2, device="cuda")
eq, eq_rm = ro.embed_queries(pm, qx, row_major=True)
slots, n_sel = ro.target_slots(2, None, "cuda")
cx, cy = torch.rand(B, n_c, 2, device="cuda"), torch.randn(B, n_c, 1, device="cuda")
nk = n_c + n_sel
tc_kv = ro.alloc_tc_kv(pm, B, nk, "cuda")
kv, _ = ro.ctx_stack(pm, cx, cy, n_c, None, slots, n_sel, want_z=False, tc_kv=tc_kv)
for _ in range(3):
    ro.query_stream(pm, eq, None, kv, nk, precision="bf16", tc_kv=tc_kv, eq_rm=eq_rm)
torch.cuda.synchronize()
N = 3 * 4096
buf = (ctypes.c_longlong * N)()
lib = _lib.lib()
assert lib.aline_debug_q4_trace(buf, N) == 0
a = np.frombuffer(buf, dtype=np.int64)
phase_names = ["Q", "S", "PV", "O", "F", "Z"] * 3 + ["A1", "A2"]
for who, lo in (("epilogue warp 0 (c=0)", 0), ("epilogue warp 5 (c=1)", 4096)):
    v = a[lo:lo + 4096]
    v = v[v != 0]
    t, slot = v // 8, v % 8
    print(f"== {who}: {len(v)} stamps")
    n_ph = len(v) // 5
    rows = []
    for p in range(min(n_ph, 80)):
        tt = t[5 * p:5 * p + 5]
        if list(slot[5 * p:5 * p + 5]) != [0, 1, 2, 3, 5]:
            print("unexpected slot order at", p, slot[5 * p:5 * p + 5])
            break
        prev_end = t[5 * p - 1] if p else tt[0]
        rows.append([tt[0] - prev_end, tt[1] - tt[0], tt[2] - tt[1], tt[3] - tt[2], tt[4] - tt[3]])
    rows = np.array(rows)
    print("phase  epilogue  st_wait  fence   arrive  mma_wait   total")
    for p, r in enumerate(rows[:40]):
        print(f"{phase_names[p % 20]:>4s} {r[0]:9d} {r[1]:8d} {r[2]:6d} {r[3]:8d} {r[4]:9d} {r.sum():8d}")
    if len(rows) > 20:
        m = rows[20:]
        print("mean over later phases:", m.mean(0).round(0), "total", m.sum(1).mean().round(0))
v = a[2 * 4096:]
v = v[v != 0]
t, slot = v // 8, v % 8
print(f"== issuer: {len(v)} stamps")
d_issue = [t[i + 1] - t[i] for i in range(0, len(v) - 1, 2) if slot[i] == 6 and slot[i + 1] == 7]
d_gap = [t[i + 2] - t[i + 1] for i in range(0, len(v) - 2, 2) if slot[i + 1] == 7 and slot[i + 2] == 6]
print("issue (ready seen -> committed): mean", np.mean(d_issue).round(0), "max", np.max(d_issue), " first 24:", d_issue[:24])
print("idle between phases: mean", np.mean(d_gap).round(0))
