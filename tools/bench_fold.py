"""Candidate-query stream (query_tc3, four warpgroups) with and without the folded operands (aline_set_option
"query_fold"): us per launch at the cfg2 launch shape and the error of each against the fp32 kernel.
    python tools/bench_fold.py [B] [nq]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from aline_b200 import _lib, rollout as ro  # noqa: E402
from aline_b200.model import Aline, Embedder, Encoder, OutputHead  # noqa: E402
from tools.bench_query import timeit  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    torch.manual_seed(123)
    model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128)).cuda().eval()
    pm = model.packed()
    qx = torch.rand(B, nq, 2, device="cuda")
    eq, eq_rm = ro.embed_queries(pm, qx, row_major=True)
    slots, n_sel = ro.target_slots(2, None, "cuda")
    _lib.set_option("query_tc4", 0)
    res = {}
    for n_c in (1, 7, 14, 18, 30, 35, 46):
        cx, cy = torch.rand(B, n_c, 2, device="cuda"), torch.randn(B, n_c, 1, device="cuda")
        nk = n_c + n_sel
        tc_kv = ro.alloc_tc_kv(pm, B, nk, "cuda")
        kv, _ = ro.ctx_stack(pm, cx, cy, n_c, None, slots, n_sel, want_z=False, tc_kv=tc_kv)
        alive = torch.ones((B, nq), dtype=torch.uint8, device="cuda")
        alive[:, : n_c - 1] = 0
        live = alive.bool()
        l32, _ = ro.query_stream(pm, eq, alive, kv, nk, precision="fp32")
        r = {"keys": nk}
        outs = {}
        for fold in (0, 1):
            _lib.set_option("query_fold", fold)
            lg, _ = ro.query_stream(pm, eq, alive, kv, nk, precision="bf16", tc_kv=tc_kv)
            torch.cuda.synchronize()
            outs[fold] = lg
            r[f"fold{fold}_us"] = timeit(lambda: ro.query_stream(pm, eq, alive, kv, nk, precision="bf16", tc_kv=tc_kv))
            d = (lg - l32)[live]
            r[f"fold{fold}_max_abs_vs_fp32"] = float(d.abs().max())
            r[f"fold{fold}_rms_vs_fp32"] = float(d.pow(2).mean().sqrt())
            ls = torch.log_softmax(lg.masked_fill(~live, -float("inf")), -1) - torch.log_softmax(l32.masked_fill(~live, -float("inf")), -1)
            r[f"fold{fold}_max_abs_logsoftmax_vs_fp32"] = float(ls[live].abs().max())
            r[f"fold{fold}_dead_minus_inf"] = bool(torch.isinf(lg[~live]).all()) if (~live).any() else True
        if nk > 32:                                  # the two-threads-per-row kernel, for the 33-48 key rule
            _lib.set_option("query_tc4", 1)
            r["tc4_us"] = timeit(lambda: ro.query_stream(pm, eq, alive, kv, nk, precision="bf16", tc_kv=tc_kv, eq_rm=eq_rm))
            _lib.set_option("query_tc4", 0)
        res[f"n_c={n_c}"] = r
        print(n_c, r, file=sys.stderr)
    _lib.set_option("query_fold", -1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
