"""Candidate-query stream at the cfg2 launch shape (B=200 x 2000 candidates): the two fast tcgen05 kernels (query_tc4: two
threads per row; query_tc3: one) timed alone with CUDA events for several context lengths, plus their agreement.
    python tools/bench_query.py [B] [nq]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from aline_b200 import _lib, rollout as ro  # noqa: E402
from aline_b200.model import Aline, Embedder, Encoder, OutputHead  # noqa: E402


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
    return e0.elapsed_time(e1) / it * 1e3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    torch.manual_seed(123)
    model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128)).cuda().eval()
    pm = model.packed()
    qx = torch.rand(B, nq, 2, device="cuda")
    eq, eq_rm = ro.embed_queries(pm, qx, row_major=True)
    slots, n_sel = ro.target_slots(2, None, "cuda")
    res = {}
    _lib.set_option("query_tc4", 1)            # eq_rm given -> two threads per row at every key count
    for n_c in (1, 14, 18, 30, 35, 46):
        cx, cy = torch.rand(B, n_c, 2, device="cuda"), torch.randn(B, n_c, 1, device="cuda")
        nk = n_c + n_sel
        tc_kv = ro.alloc_tc_kv(pm, B, nk, "cuda")
        kv, _ = ro.ctx_stack(pm, cx, cy, n_c, None, slots, n_sel, want_z=False, tc_kv=tc_kv)
        alive = torch.ones((B, nq), dtype=torch.uint8, device="cuda")
        alive[:, : n_c - 1] = 0
        l4, _ = ro.query_stream(pm, eq, alive, kv, nk, precision="bf16", tc_kv=tc_kv, eq_rm=eq_rm)
        l3, _ = ro.query_stream(pm, eq, alive, kv, nk, precision="bf16", tc_kv=tc_kv)
        l32, _ = ro.query_stream(pm, eq, alive, kv, nk, precision="fp32")
        torch.cuda.synchronize()
        live = alive.bool()
        r = {"keys": nk,
             "tc4_us": timeit(lambda: ro.query_stream(pm, eq, alive, kv, nk, precision="bf16", tc_kv=tc_kv, eq_rm=eq_rm)),
             "tc3_us": timeit(lambda: ro.query_stream(pm, eq, alive, kv, nk, precision="bf16", tc_kv=tc_kv)),
             "max_abs_tc4_vs_fp32": float((l4 - l32)[live].abs().max()),
             "max_abs_tc3_vs_fp32": float((l3 - l32)[live].abs().max()),
             "dead_are_minus_inf": bool(torch.isinf(l4[~live]).all()) if (~live).any() else True}
        res[f"n_c={n_c}"] = r
        print(n_c, r, file=sys.stderr)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
