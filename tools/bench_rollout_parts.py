"""Per-kernel timings of one rollout step at cfg2 sizes, for several context lengths."""
import json
import sys

import torch

sys.path.insert(0, ".")
from aline_b200 import rollout as ro  # noqa: E402
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
    return e0.elapsed_time(e1) / it * 1e3       # us


def main():
    torch.manual_seed(0)
    import os
    B, nq = int(os.environ.get("PARTS_B", 200)), int(os.environ.get("PARTS_NQ", 2000))      # cfg2 by default
    model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128)).cuda().eval()
    pm = model.packed()
    qx = torch.rand(B, nq, 2, device="cuda")
    eq = ro.embed_queries(pm, qx)
    slots, n_sel = ro.target_slots(2, None, "cuda")
    out = {}
    for n_c in [int(v) for v in os.environ.get('PARTS_NC', '2,14,18,30,35').split(',')]:
        cx, cy = torch.rand(B, 40, 2, device="cuda"), torch.randn(B, 40, 1, device="cuda")
        nk = n_c + n_sel
        tc_kv = ro.alloc_tc_kv(pm, B, nk, "cuda")
        kv = torch.empty((3, B, 40 + n_sel, 2, 32), device="cuda")
        r = {}
        r["ctx_us"] = timeit(lambda: ro.ctx_stack(pm, cx, cy, n_c, None, slots, n_sel, kv=kv, kv_slots=40 + n_sel, want_z=False))
        r["ctx_tckv_us"] = timeit(lambda: ro.ctx_stack(pm, cx, cy, n_c, None, slots, n_sel, kv=kv, kv_slots=40 + n_sel, want_z=False, tc_kv=tc_kv))
        if "PARTS_FAST_ONLY" not in os.environ:
            r["q_fp32_us"] = timeit(lambda: ro.query_stream(pm, eq, None, kv, nk, precision="fp32"), it=5)
            r["q_tc_ffma_attn_us"] = timeit(lambda: ro.query_stream(pm, eq, None, kv, nk, precision="bf16"))
        r["q_tc_tc_attn_us"] = timeit(lambda: ro.query_stream(pm, eq, None, kv, nk, precision="bf16", tc_kv=tc_kv))
        logits, _ = ro.query_stream(pm, eq, None, kv, nk, precision="bf16", tc_kv=tc_kv)
        r["select_us"] = timeit(lambda: ro.select(logits, want_zt=False))
        out[f"n_c={n_c}"] = {k: round(v, 1) for k, v in r.items()}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
