#!/usr/bin/env python
"""Headline benchmark: one `eval_boed` batch of the location-finding final evaluation (BASELINE.json configs[1]):
a T=35 rollout (34 design steps) of B=200 trajectories over 2000 candidates with a random-init ALINE model,
followed by the step-wise sPCE / sNMC bounds over L = 1e6 contrastive prior draws.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  `value` = prior samples per second over the whole step (rollout included),
inputs resident in HBM; `e2e` = the same through the public API (`model.rollout` + `compute_EIG_from_history`)
with the batch coming from pinned host memory and the bounds copied back; `components` carries the two metrics of
BASELINE.json separately.  N > 1: every rank evaluates its own batch of rollouts (weak scaling, the outer loop of
`eval_boed` dealt over ranks) and the bounds are all-gathered.  `--impl reference` times the CPU oracle
(`oracle/aline_oracle.py`, a port of the reference's algorithm incl. its dense N x N attention) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(workload="location_finding eval-final (cfg2): B=200, n_query=2000, T=35 (34 design steps), "
                    "sPCE/sNMC step-wise, L=1e6", B=200, n_query=2000, T=35, L=1_000_000, K=1, dim_x=2)
METRIC = "rollout design-steps/sec + sPCE prior samples/sec at 1/2/4/8 B200"
UNIT = "prior-samples/s over the whole eval step (rollout + sPCE); see components"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed regions: NVML in-process (sub-millisecond per sample),
    nvidia-smi as the fallback."""

    _REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
                ("sw_power_cap", 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.max_sm, self.reasons, self.stop_flag = index, [], None, set(), False
        self.region, self.by_region = "resident", {}
        self.handle, self.nvml = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:  # noqa: BLE001
            self.handle = None

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        f = [s.strip() for s in out.split(",")]
        if len(f) >= 6 and f[0].isdigit():
            self.sm.append(int(f[0]))
            self.max_sm = int(f[1]) if f[1].isdigit() else self.max_sm
            for i, (name, _) in enumerate(self._REASONS):
                if f[2 + i] == "Active":
                    self.reasons.add(name)

    def run(self):
        while not self.stop_flag:
            try:
                if self.handle is not None:
                    mhz = int(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
                    self.by_region.setdefault(self.region, []).append(mhz)
                    if self.region in ("resident", "e2e"):
                        self.sm.append(mhz)
                    bits = int(self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                    for name, bit in self._REASONS:
                        if bits & bit:
                            self.reasons.add(name)
                    time.sleep(0.02)      # 50 Hz: NVML queries contend with the launch path at kHz rates
                else:
                    self._sample_smi()
                    time.sleep(0.02)
            except Exception:  # noqa: BLE001
                time.sleep(0.02)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["unavailable"]}
        sm = sorted(self.sm)
        med = {k: sorted(v)[len(v) // 2] for k, v in self.by_region.items() if v}
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": "nvml" if self.handle is not None else "nvidia-smi",
                "sm_mhz_by_region": med}


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu capture (profiles/r1_traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))[kernel]["bytes"]
    except Exception:  # noqa: BLE001
        return None


def structured_flops_per_rollout(B, nq0, n_c0, steps, d=32, ff=128, dx=2, dy=1, n_t=2, n_sel=2, nl=3):
    """Algorithmic FLOPs of the structured forward, summed over the rollout; GMM-on-query excluded."""
    tot = 0.0
    for t in range(steps):
        n_c, n_q = n_c0 + t, nq0 - t
        emb = (n_c + n_q) * 2 * (dx * ff + ff * d) + n_c * 2 * (dy * ff + ff * d)
        per_c = 6 * d * d + 4 * d * n_c + 2 * d * d + 4 * d * ff
        per_t = 2 * d * d + 4 * d * n_c + 2 * d * d + 4 * d * ff
        per_q = 2 * d * d + 4 * d * (n_c + n_sel) + 2 * d * d + 4 * d * ff
        enc = nl * (n_c * per_c + n_t * per_t + n_sel * 4 * d * d + n_q * per_q)
        acq = n_q * 2 * (d * ff + ff)
        tot += B * (emb + enc + acq)
    return tot


def query_stream_flops(B, n_q, n_keys, d=32, ff=128, nl=3):
    """Algorithmic FLOPs of one query_stream launch (live candidates only)."""
    per_q = 2 * d * d + 4 * d * n_keys + 2 * d * d + 4 * d * ff
    return B * n_q * (nl * per_q + 2 * (d * ff + ff))


# ------------------------------------------------------------------ CPU oracle arm ----
def cpu_sample(rollout_B=8, rollout_steps=4, L_sample=150_000):
    """Oracle timed on a bounded sample of the cfg2 workload (~10-20 s of CPU work on 16 cores); linear extrapolation to
    the full step."""
    from oracle import aline_oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _util import load_golden, state_dict_of
    torch.manual_seed(123)
    sd = state_dict_of(load_golden("rollout_location"))           # random-init reference weights (seed 123)
    B, nq, T, L = CFG["B"], CFG["n_query"], CFG["T"], CFG["L"]
    cx, qx = torch.rand(rollout_B, 1, 2), torch.rand(rollout_B, nq, 2)
    batch = dict(context_x=cx, context_y=torch.randn(rollout_B, 1, 1), query_x=qx, query_y=torch.randn(rollout_B, nq, 1),
                 target_all=torch.rand(rollout_B, 2, 1))
    t0 = time.perf_counter()
    O.rollout(sd, batch, rollout_steps, "theta", 4, dense=True)    # dense N x N attention: the reference's cost model
    t_roll = time.perf_counter() - t0
    x, y = torch.rand(B, T, 2), torch.randn(B, T, 1)
    thetas = torch.rand(L_sample + 1, B, 1, 2)
    t0 = time.perf_counter()
    O.spce_history(O.location_log_likelihood, y, x, thetas, stepwise=True)
    t_spce = time.perf_counter() - t0
    full_roll = t_roll * (B * (T - 1)) / (rollout_B * rollout_steps)
    full_spce = t_spce * L / L_sample
    return dict(t_roll=t_roll, t_spce=t_spce, full=full_roll + full_spce, full_roll=full_roll, full_spce=full_spce,
                sample=f"rollout {rollout_B} of {B} trajectories x {rollout_steps} of {T - 1} steps (dense attention over "
                       f"N=2003 tokens) + sPCE L={L_sample} of {L} x B={B} x T={T}; linear extrapolation to the full step")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # all the host threads the box offers (torchrun exports OMP_NUM_THREADS=1, which would throttle the baseline)
    torch.set_num_threads(os.cpu_count() or 1)
    B, T, L = CFG["B"], CFG["T"], CFG["L"]
    for _ in range(args.warmup):
        cpu_sample(rollout_B=1, rollout_steps=1, L_sample=2000)
    t0 = time.perf_counter()
    fulls, last = [], None
    for _ in range(args.steps):
        last = cpu_sample()
        fulls.append(last["full"])
    wall = time.perf_counter() - t0
    full = sum(fulls) / len(fulls)
    value = L * B / full
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": full * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {k: CFG[k] for k in ("workload",)},
        "components": {"rollout_design_steps_per_s": B * (T - 1) / last["full_roll"],
                       "spce_prior_samples_per_s": L * B / last["full_spce"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": last["sample"], "measured_wall_s": wall},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ B200 arm ----
def run_native(args):
    import torch.distributed as dist
    from aline_b200 import kernel_launches, spce
    from aline_b200.attrdict import AttrDict
    from aline_b200.model import Aline, Embedder, Encoder, OutputHead
    from aline_b200.tasks import HiddenLocation
    from aline_b200.utils.eval import compute_EIG_from_history
    from aline_b200 import rollout as ro

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything libraries print there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, nq, T, L = CFG["B"], CFG["n_query"], CFG["T"], CFG["L"]
    steps_T = T - 1

    torch.manual_seed(123)
    model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128))
    model = model.to(dev).eval()
    model.precision = args.precision
    torch.manual_seed(1000 + rank)
    task = HiddenLocation(n_query_init=nq, design_scale=1)
    task.to(dev)
    host_batches = []
    for _ in range(2):                                     # synthetic task draws, generated on the host
        hb = task.sample_batch(B)
        host_batches.append({k: hb[k].contiguous().pin_memory() for k in
                             ("context_x", "context_y", "query_x", "query_y", "target_all")})
    h2d_bytes = sum(v.numel() * 4 for v in host_batches[0].values())
    res_batch = {k: v.to(dev) for k, v in host_batches[0].items()}
    with torch.device(dev):
        rows = task.sample_theta((L + 1, B))               # [L+1, B, 1, 2] resident contrastive draws (1.6 GB > L2)
    pce_host = torch.empty((B, T), dtype=torch.float32).pin_memory()
    nmc_host = torch.empty((B, T), dtype=torch.float32).pin_memory()
    out_keep = {}

    def step_resident():
        b = AttrDict({k: v for k, v in res_batch.items()})
        b.target_theta = b.target_all
        out = model.rollout(b, steps_T)
        theta_0 = b.target_all.reshape(B, 1, 2)
        x, y = task.unnormalise_design(out.context_x), out.context_y
        rows[0] = theta_0
        m, s, lp0 = spce.spce_history(task.log_likelihood, y, x, rows, seq=None, skip_rows=1)
        pl, nl = spce.lse_combine(m, s, lp0)
        out_keep["pce"], out_keep["nmc"] = math.log(L + 1) - pl, math.log(L) - nl

    def step_e2e(i):
        hb = host_batches[i % 2]
        b = AttrDict({k: v.to(dev, non_blocking=True) for k, v in hb.items()})
        b.target_theta = b.target_all
        out = model.rollout(b, steps_T)
        theta_0 = b.target_all.reshape(B, 1, 2)
        with torch.device(dev):
            pce, nmc = compute_EIG_from_history(task, theta_0, task.unnormalise_design(out.context_x), out.context_y,
                                                L=L, batch_size=B, stepwise=True)
        pce_host.copy_(pce, non_blocking=True)
        nmc_host.copy_(nmc, non_blocking=True)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, n):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = kernel_launches()
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, kernel_launches() - l0

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = timed(lambda i: step_resident(), args.steps)
    sampler.region = "e2e"
    for i in range(2):
        step_e2e(i)
    ms_e2e, _ = timed(step_e2e, args.steps)
    sampler.region = "components"

    # component timings (separate timed loops, same inputs)
    def only_rollout(i):
        b = AttrDict({k: v for k, v in res_batch.items()})
        out_keep["roll"] = model.rollout(b, steps_T)

    only_rollout(0)
    ms_roll, _ = timed(only_rollout, args.steps)
    x = task.unnormalise_design(out_keep["roll"].context_x)
    y = out_keep["roll"].context_y

    def only_spce(i):
        spce.spce_history(task.log_likelihood, y, x, rows, seq=None, skip_rows=1)

    ms_spce, _ = timed(only_spce, args.steps)

    def spce_with_draw(i):          # the bound evaluation INCLUDING the L contrastive prior draws
        with torch.device(dev):
            compute_EIG_from_history(task, res_batch["target_all"].reshape(B, 1, 2), x, y, L=L, batch_size=B, stepwise=True)

    spce_with_draw(0)
    ms_spce_draw, _ = timed(spce_with_draw, args.steps)

    # dominant-kernel roofline: query_stream at the rollout's mid step, timed alone with CUDA events
    pm = model.packed()
    mid = steps_T // 2
    n_c = 1 + mid
    eq = ro.embed_queries(pm, res_batch["query_x"])
    slots, n_sel = ro.target_slots(2, None, dev)
    tc_kv = None
    if ro.use_tensor_cores(pm, model.precision, n_c + n_sel) and n_c + n_sel <= pm.tc_fast_max_keys:
        tc_kv = ro.alloc_tc_kv(pm, B, n_c + n_sel, dev)
    kv, _ = ro.ctx_stack(pm, out_keep["roll"].context_x, out_keep["roll"].context_y, n_c, None, slots, n_sel,
                         want_z=False, tc_kv=tc_kv)
    alive = torch.ones((B, nq), dtype=torch.uint8, device=dev)
    alive[:, :mid] = 0

    def only_query(i):
        ro.query_stream(pm, eq, alive, kv, n_c + n_sel, precision=model.precision, tc_kv=tc_kv)

    for i in range(3):
        only_query(i)
    ms_q, _ = timed(only_query, 10)
    ms_q /= 10
    seq = torch.zeros((L + 1, B), device=dev)

    def only_spce_step(i):
        spce.spce_step(task.log_likelihood, y[:, 0], x[:, 0], rows, seq)

    for i in range(2):
        only_spce_step(i)
    ms_s1, _ = timed(only_spce_step, 5)
    ms_s1 /= 5
    sampler.stop_flag = True

    pk = peaks()
    ms_step = ms / args.steps
    value = world * L * B / (ms_step * 1e-3)
    e2e_value = world * L * B / (ms_e2e / args.steps * 1e-3)
    q_flops = query_stream_flops(B, nq - mid, n_c + n_sel)
    q_tf = q_flops / (ms_q * 1e-3) / 1e12
    step_bytes = (L + 1) * B * (4 * 2 + 8)
    hist_bytes = (L + 1) * B * 4 * 2
    clocks = sampler.summary()
    clk_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965        # MUFU peak at the clock measured under load
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16 operands / f32 accumulate (candidate stream), f32 elsewhere"
        if model.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": CFG["workload"], "per_gpu": "one eval batch (B=200 rollouts, L=1e6 draws) per rank",
                   "weights": "random-init ALINE d=32 ff=128 h=4 3 layers (seed 123)",
                   "l2": "thetas [1000001,200,1,2] fp32 = 1.6 GB per pass exceed the 126 MB L2; rollout working set "
                         "(~60 MB) is L2-resident by design and re-used across the 34 dependent steps"},
        "components": {
            "rollout_design_steps_per_s": world * B * steps_T / (ms_roll / args.steps * 1e-3),
            "rollout_ms": ms_roll / args.steps,
            "spce_prior_samples_per_s": world * L * B / (ms_spce / args.steps * 1e-3),
            "spce_ms": ms_spce / args.steps,
            "spce_incl_theta_sampling_ms": ms_spce_draw / args.steps,
            "spce_incl_theta_sampling_prior_samples_per_s": world * L * B / (ms_spce_draw / args.steps * 1e-3),
            "spce_likelihood_evals_per_s": world * L * B * T / (ms_spce / args.steps * 1e-3)},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 2 * B * T * 4,
                "note": "thetas are drawn on the device inside compute_EIG_from_history, as the reference does"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": "query_tc3_kernel<4> (tcgen05 bf16 x bf16 -> fp32 in TMEM, softmax probabilities / MLP "
                               "activations as TMEM A operands, 4 tiles in flight per SM; candidate tokens through 3 encoder "
                               "layers + acquisition MLP), mid-rollout launch" if model.precision == "bf16"
                     else "query_stream_kernel<32> (fp32 FFMA), mid-rollout launch", "bound": "tensor", "achieved": q_tf,
                     "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": q_tf / pk["tf_sust"],
                     "traffic": ncu_traffic("query_tc3_kernel") if model.precision == "bf16" else None,
                     "peak_source": pk["src"] + ", sustained bf16 (kernel timed inside a long step)",
                     "launch_ms": ms_q, "algorithmic_flops_per_launch": q_flops,
                     "share_of_step": (ms_q * steps_T) / ms_step},
        "rooflines": [
            {"kernel": "spce_step_tma_kernel (EIGStepLoss.step drop-in, one launch: cp.async.bulk ring, in-place update, "
                       "bulk store, fused theta_0 row + merge; timed through the C-ABI call)", "bound": "hbm",
             "achieved": step_bytes / (ms_s1 * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
             "frac": step_bytes / (ms_s1 * 1e-3) / 1e9 / pk["hbm"], "launch_ms": ms_s1,
             "algorithmic_bytes_per_launch": step_bytes, "traffic": ncu_traffic("spce_step_tma_kernel")},
            {"kernel": "spce_fast_loc12x2 x3 passes of 12 history points + cold theta_0 pass (fused history, shifted "
                       "accumulation, packed fp32x2 pairs, one MUFU reciprocal per four evaluations; MUFU/issue-bound: "
                       "2.25 MUFU per likelihood evaluation, XU pipe 16 lanes/clk/SM -- HBM shown for reference)",
             "bound": "hbm", "achieved": hist_bytes / (ms_spce / args.steps * 1e-3) / 1e9,
             "peak": pk["hbm"], "unit": "GB/s", "frac": hist_bytes / (ms_spce / args.steps * 1e-3) / 1e9 / pk["hbm"],
             "algorithmic_bytes_per_eval": hist_bytes,
             "mufu_bound": {"mufu_per_evaluation": 2.25, "evaluations": L * B * T,
                            "achieved_mufu_per_s": 2.25 * L * B * T / (ms_spce / args.steps * 1e-3),
                            "peak_mufu_per_s": 16 * 148 * clk_mhz * 1e6,
                            "frac": 2.25 * L * B * T / (ms_spce / args.steps * 1e-3) / (16 * 148 * clk_mhz * 1e6)}}],
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        c = cpu_sample()
        line["cpu_baseline"] = {"value": L * B / c["full"], "unit": UNIT, "cores": torch.get_num_threads(),
                                "kind": "port", "sample": c["sample"],
                                "rollout_design_steps_per_s": B * steps_T / c["full_roll"],
                                "spce_prior_samples_per_s": L * B / c["full_spce"]}
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    if rank == 0:
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    os.close(json_fd)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="candidate-query stream: tcgen05 bf16 (default) or fp32 FFMA validation mode")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
