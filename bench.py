#!/usr/bin/env python
"""Headline benchmark: the location-finding FINAL EVALUATION of the reference (README.md:45, config/eval/bed.yaml:7-11;
BASELINE.json configs[1]) -- `eval_boed` with M = 2000 outer rollouts in mini-batches of B = 200, 2000 candidate designs,
T = 35 (34 design steps) and the step-wise sPCE / sNMC bounds over L = 1e6 contrastive prior draws, random-init ALINE.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one whole evaluation (M x L prior samples).  `value` = prior samples per
second with the simulated batches and the contrastive draws resident in HBM; `e2e` = the same evaluation through the
public API (`eval_boed(..., batches=<pinned host batches>)`: host-to-device copies of every mini-batch, contrastive
thetas drawn inside, the summary copied back).  N > 1 is STRONG scaling of that fixed workload: the M rollouts are dealt
to the ranks as balanced slices (`aline_b200.utils.eval.rank_chunks`: 250 per rank at N = 8), no collective on the data
path, and the per-rollout bounds are combined by one NCCL all-gather inside the timed region.  `components` carries
BASELINE.json's two metrics separately (one mini-batch, stand-alone), the serial (unpipelined) loop, the L-sharded CES
bound of configs[2] (B = 20, L = 1e7, contrastive rows split over the ranks, one all-gather of the (max, sum-exp)
partials), and -- at N = 1 -- the UNMODIFIED reference on the same GPU (`reference_cuda_eager_ms`: PyTorch eager + cuBLAS,
fp32, TF32 off, from baseline/_ref).

`--impl reference` times the unmodified reference (baseline/_ref, installed by baseline/install_ref.py) on the box's
host cores: its own `get_traces` + `compute_EIG_from_history` at the full n_query / T / L of the workload on a bounded
sample of the rollouts (2 of the 2000), no extrapolation -- the metric is a rate.  Without baseline/_ref it falls back to
the oracle port and says so (`cpu_baseline.kind`).
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

if "--impl" in sys.argv and "reference" in sys.argv:
    # the reference decides `device = "cuda" if torch.cuda.is_available()` inside its constructors (model/encoder.py:80,
    # tasks/base_task.py:27): its CPU path only runs unmodified when no GPU is visible to the process
    os.environ["CUDA_VISIBLE_DEVICES"] = ""

import torch  # noqa: E402

CFG = dict(workload="location_finding eval-final (cfg2): eval_boed M=2000 rollouts in mini-batches of B=200, n_query=2000, "
                    "T=35 (34 design steps), sPCE/sNMC step-wise, L=1e6",
           M=2000, B=200, n_query=2000, T=35, L=1_000_000, K=1, dim_x=2)
CFG3 = dict(B=20, n_query=2000, T=15, L=10_000_000)
METRIC = "rollout design-steps/sec + sPCE prior samples/sec at 1/2/4/8 B200"
UNIT = "prior-samples/s over the whole evaluation (rollouts + sPCE); see components"
REF_SAMPLE_B = 2


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
    """dram bytes per launch of `kernel` from the committed ncu captures (profiles/r2_traffic.json, else round 1's)."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))[kernel]["bytes"]
        except Exception:  # noqa: BLE001
            continue
    return None


def query_stream_flops(B, n_q, n_keys, d=32, ff=128, nl=3):
    """Algorithmic FLOPs of one query_stream launch (live candidates only; SURVEY.md 8d structured count)."""
    per_q = 2 * d * d + 4 * d * n_keys + 2 * d * d + 4 * d * ff
    return B * n_q * (nl * per_q + 2 * (d * ff + ff))


# ------------------------------------------------------------------ reference arm (CPU) ----
def _ref_available():
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_loader
    return ref_loader if ref_loader.available() else None


def ref_cpu_step(tiny=False):
    """One bounded sample of the workload on the host cores: the unmodified reference's get_traces +
    compute_EIG_from_history (baseline/_ref) for REF_SAMPLE_B of the M rollouts at the workload's full n_query, T and L.
    Falls back to the oracle port when baseline/_ref is absent.  Returns timings and a description."""
    B = 1 if tiny else REF_SAMPLE_B
    nq, T, L = (64, 3, 2000) if tiny else (CFG["n_query"], CFG["T"], CFG["L"])
    rl = _ref_available()
    if rl is not None:
        R = rl.load()
        torch.manual_seed(123)
        model = R.Aline(R.Embedder(2, 1, 32, 128, 2, "theta"), R.Encoder(32, 128, 4, 0.0, 3),
                        R.OutputHead(2, 1, 32, 128)).eval()
        task = R.HiddenLocation(n_query_init=nq, design_scale=1)
        t0 = time.perf_counter()
        theta_0, x, y = R.eval.get_traces(model, task, T=T - 1, batch_size=B)
        t1 = time.perf_counter()
        pce, _ = R.eval.compute_EIG_from_history(task, theta_0, x, y, L=L, batch_size=B, stepwise=True)
        t2 = time.perf_counter()
        assert tuple(pce.shape) == (B, T)
        kind, src = "reference", "baseline/_ref (unmodified reference files: utils/eval.py get_traces + compute_EIG_from_history)"
    else:
        from oracle import aline_oracle as O
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from _util import load_golden, state_dict_of
        torch.manual_seed(123)
        sd = state_dict_of(load_golden("rollout_location"))
        batch = dict(context_x=torch.rand(B, 1, 2), context_y=torch.randn(B, 1, 1), query_x=torch.rand(B, nq, 2),
                     query_y=torch.randn(B, nq, 1), target_all=torch.rand(B, 2, 1))
        t0 = time.perf_counter()
        r = O.rollout(sd, batch, T - 1, "theta", 4, dense=True)
        t1 = time.perf_counter()
        thetas = torch.rand(L + 1, B, 1, 2)
        O.spce_history(O.location_log_likelihood, r["batch"]["context_y"], r["batch"]["context_x"], thetas, stepwise=True)
        t2 = time.perf_counter()
        kind, src = "port", "oracle/aline_oracle.py (baseline/_ref not installed)"
    return dict(t_roll=t1 - t0, t_spce=t2 - t1, t=t2 - t0, B=B, nq=nq, T=T, L=L, kind=kind, source=src,
                sample=f"{B} of the {CFG['M']} rollouts at the full n_query={nq}, T={T} ({T - 1} design steps, dense "
                       f"attention over N={nq + 3} tokens) + step-wise sPCE/sNMC at the full L={L} for those {B} "
                       "histories; rates measured on the sample, no extrapolation")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # all the host threads the box offers (torchrun exports OMP_NUM_THREADS=1, which would throttle the baseline)
    torch.set_num_threads(os.cpu_count() or 1)
    for _ in range(args.warmup):
        ref_cpu_step(tiny=True)
    t0 = time.perf_counter()
    runs = [ref_cpu_step() for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    last = runs[-1]
    t_step = sum(r["t"] for r in runs) / len(runs)
    B, T, L = last["B"], last["T"], last["L"]
    value = L * B / t_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": CFG["workload"] + f" -- bounded sample per step: {last['sample']}"},
        "components": {"rollout_design_steps_per_s": B * (T - 1) / (sum(r["t_roll"] for r in runs) / len(runs)),
                       "spce_prior_samples_per_s": L * B / (sum(r["t_spce"] for r in runs) / len(runs))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": last["kind"],
                         "source": last["source"], "sample": last["sample"], "measured_wall_s": wall},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def cpu_baseline_subprocess():
    """The cpu_baseline leg of the native arm: one bounded reference step in a child process that sees no GPU (the
    reference picks its device from torch.cuda.is_available())."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "OMP_NUM_THREADS"):
        env.pop(k, None)
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                              "--warmup", "1"], capture_output=True, text=True, timeout=900, env=env)
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
        d = json.loads(line)
        cb = d["cpu_baseline"]
        cb.update(rollout_design_steps_per_s=d["components"]["rollout_design_steps_per_s"],
                  spce_prior_samples_per_s=d["components"]["spce_prior_samples_per_s"])
        return cb
    except Exception as exc:  # noqa: BLE001
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": f"failed: {exc}"}


# ------------------------------------------------------------------ reference on the same GPU (cuda eager) ----
def reference_cuda_eager(dev):
    """The UNMODIFIED reference with every tensor on the B200 (torch.device context == train_aline.py:189's
    set_default_device): one mini-batch of the workload -- get_traces (B = 200, 2000 candidates, 34 steps; the dense
    [800, 2003, 2003] fp32 scores = 12.8 GB per layer) + compute_EIG_from_history (L = 1e6, step-wise).  fp32, TF32 off
    (torch's default).  Timed with CUDA events after one warm-up run."""
    rl = _ref_available()
    if rl is None:
        return None
    R = rl.load()
    B, nq, T, L = CFG["B"], CFG["n_query"], CFG["T"], CFG["L"]
    out = {}
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        with torch.device(dev):
            torch.manual_seed(123)
            model = R.Aline(R.Embedder(2, 1, 32, 128, 2, "theta"), R.Encoder(32, 128, 4, 0.0, 3),
                            R.OutputHead(2, 1, 32, 128)).eval()
            task = R.HiddenLocation(n_query_init=nq, design_scale=1)
            for rep in range(2):
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                torch.cuda.synchronize(dev)
                e[0].record()
                theta_0, x, y = R.eval.get_traces(model, task, T=T - 1, batch_size=B)
                e[1].record()
                pce, nmc = R.eval.compute_EIG_from_history(task, theta_0, x, y, L=L, batch_size=B, stepwise=True)
                e[2].record()
                torch.cuda.synchronize(dev)
                out = {"rollout_ms": e[0].elapsed_time(e[1]), "spce_ms": e[1].elapsed_time(e[2]),
                       "batch_ms": e[0].elapsed_time(e[2]), "pce_final_mean": float(pce[:, -1].mean())}
                del theta_0, x, y, pce, nmc
        del model, task
        torch.cuda.empty_cache()
        out["peak_mem_gb"] = torch.cuda.max_memory_allocated(dev) / 1e9
    except Exception as exc:  # noqa: BLE001
        return {"error": f"{type(exc).__name__}: {exc}"[:300]}
    return out


# ------------------------------------------------------------------ B200 arm ----
def run_native(args):
    import torch.distributed as dist
    from aline_b200 import kernel_launches, spce
    from aline_b200.attrdict import AttrDict
    from aline_b200.model import Aline, Embedder, Encoder, OutputHead
    from aline_b200.tasks import HiddenLocation, CESTask
    from aline_b200.utils import eval as ev
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
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    M, B, nq, T, L = CFG["M"], CFG["B"], CFG["n_query"], CFG["T"], CFG["L"]
    steps_T = T - 1
    chunks = ev.rank_chunks(M, B, rank, world)              # this rank's mini-batches: (global offset, size)
    n_local = sum(sz for _, sz in chunks)

    torch.manual_seed(123)
    model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128))
    model = model.to(dev).eval()
    model.precision = args.precision
    torch.manual_seed(1000 + rank)
    task = HiddenLocation(n_query_init=nq, design_scale=1)
    task.to(dev)
    keys = ("context_x", "context_y", "query_x", "query_y", "target_all")
    host_batches = []
    for _, sz in chunks:                                   # synthetic task draws, generated on the host
        hb = task.sample_batch(sz)
        host_batches.append({k: hb[k].contiguous().pin_memory() for k in keys})
    h2d_bytes = sum(v.numel() * 4 for hb in host_batches for v in hb.values())
    res_batches = [{k: v.to(dev) for k, v in hb.items()} for hb in host_batches]
    rows_by_size = {}
    with torch.device(dev):
        for sz in sorted({sz for _, sz in chunks}):
            rows_by_size[sz] = task.sample_theta((L + 1, sz))   # [L+1, sz, 1, 2] resident contrastive draws (>= 1 GB > L2)

    def roll(c):
        b = AttrDict(dict(res_batches[c]))
        out = model.rollout(b, steps_T)
        return task.unnormalise_design(out.context_x), out.context_y, b.target_all.reshape(-1, 1, 2)

    def bound(c, x, y, theta_0):
        rows = rows_by_size[x.shape[0]]
        rows[0] = theta_0
        m, s, lp0 = spce.spce_history(task.log_likelihood, y, x, rows, seq=None, skip_rows=1)
        pl, nl = spce.lse_combine(m, s, lp0)
        return math.log(L + 1) - pl, math.log(L) - nl

    def eval_resident(overlap=True):
        pipe = ev._TwoStage(dev, overlap and len(chunks) > 1)
        ps, ns = [], []
        for c in range(len(chunks)):
            x, y, th0 = pipe.rollout(lambda: roll(c))
            p, n = pipe.bound(lambda: bound(c, x, y, th0))
            ps.append(p)
            ns.append(n)
        pipe.join()
        pce, nmc = torch.cat(ps, 0), torch.cat(ns, 0)
        if world > 1:                                      # the one collective: NCCL all-gather of the [M, T] bounds
            pce, nmc = ev.gather_rows(dist, pce, M), ev.gather_rows(dist, nmc, M)
        return pce, nmc

    summary_keep = {}

    def eval_e2e(i):
        with torch.device(dev):
            summary_keep["res"] = ev.eval_boed(model, task, T=steps_T, L=L, M=M, batch_size=B, stepwise=True,
                                               verbose=False, batches=host_batches)

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
        eval_resident()
    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = timed(lambda i: eval_resident(), args.steps)
    sampler.region = "e2e"
    for i in range(2):
        eval_e2e(i)
    ms_e2e, _ = timed(eval_e2e, args.steps)
    sampler.region = "components"
    n_comp = max(3, min(args.steps, 10))
    eval_resident(overlap=False)
    ms_serial, _ = timed(lambda i: eval_resident(overlap=False), n_comp)

    # component timings: ONE mini-batch of B = 200, stand-alone (BASELINE.json's two metrics)
    one = {k: v[:B].contiguous() for k, v in res_batches[0].items()} if res_batches[0]["query_x"].shape[0] >= B else None
    if one is None:
        hb = task.sample_batch(B)
        one = {k: hb[k].to(dev).contiguous() for k in keys}
    with torch.device(dev):
        rows_one = rows_by_size.get(B)
        if rows_one is None:
            rows_one = task.sample_theta((L + 1, B))
    rows_one[0] = one["target_all"].reshape(B, 1, 2)       # row 0 = the theta_0 these histories were simulated from
    out_keep = {}

    def only_rollout(i):
        out_keep["roll"] = model.rollout(AttrDict(dict(one)), steps_T)

    for i in range(3):
        only_rollout(i)
    ms_roll, _ = timed(only_rollout, n_comp)
    x = task.unnormalise_design(out_keep["roll"].context_x)
    y = out_keep["roll"].context_y

    def only_spce(i):
        spce.spce_history(task.log_likelihood, y, x, rows_one, seq=None, skip_rows=1)

    only_spce(0)
    ms_spce, _ = timed(only_spce, n_comp)

    def spce_with_draw(i):          # the bound evaluation INCLUDING the L contrastive prior draws
        with torch.device(dev):
            ev.compute_EIG_from_history(task, one["target_all"].reshape(B, 1, 2), x, y, L=L, batch_size=B, stepwise=True)

    spce_with_draw(0)
    ms_spce_draw, _ = timed(spce_with_draw, n_comp)

    # dominant-kernel roofline: query_stream at the rollout's mid step, timed alone with CUDA events
    pm = model.packed()
    mid = steps_T // 2
    n_c = 1 + mid
    eq, eq_rm = ro.embed_queries(pm, one["query_x"], row_major=True)
    slots, n_sel = ro.target_slots(2, None, dev)
    tc_kv = None
    if ro.use_tensor_cores(pm, model.precision, n_c + n_sel) and n_c + n_sel <= pm.tc_fast_max_keys:
        tc_kv = ro.alloc_tc_kv(pm, B, n_c + n_sel, dev)
    kv, _ = ro.ctx_stack(pm, out_keep["roll"].context_x, out_keep["roll"].context_y, n_c, None, slots, n_sel,
                         want_z=False, tc_kv=tc_kv)
    alive = torch.ones((B, nq), dtype=torch.uint8, device=dev)
    alive[:, :mid] = 0

    def only_query(i):
        ro.query_stream(pm, eq, alive, kv, n_c + n_sel, precision=model.precision, tc_kv=tc_kv, eq_rm=eq_rm)

    def only_query_tc4(i):          # the two-threads-per-row kernel forced at this key count, for the record
        ro.query_stream(pm, eq, alive, kv, n_c + n_sel, precision=model.precision, tc_kv=tc_kv, eq_rm=eq_rm)

    from aline_b200 import _lib as alib
    for i in range(3):
        only_query(i)
    ms_q, _ = timed(only_query, 10)
    ms_q /= 10
    alib.set_option("query_tc4", 1)
    for i in range(3):
        only_query_tc4(i)
    ms_q3, _ = timed(only_query_tc4, 10)
    ms_q3 /= 10
    alib.set_option("query_tc4", -1)
    seq = torch.zeros((L + 1, B), device=dev)

    def only_spce_step(i):
        spce.spce_step(task.log_likelihood, y[:, 0], x[:, 0], rows_one, seq)

    for i in range(2):
        only_spce_step(i)
    ms_s1, _ = timed(only_spce_step, 5)
    ms_s1 /= 5
    del seq

    # configs[2]: CES final evaluation, B = 20, L = 1e7 contrastive rows SHARDED over the ranks (one all-gather of the
    # per-(b,t) (max, sum-exp) partials), 14-step rollout over 2000 candidates
    ces = {}
    try:
        torch.manual_seed(77)                              # same histories on every rank (the L axis is what is split)
        cmodel = Aline(Embedder(6, 1, 32, 128, 5, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(6, 1, 32, 128))
        cmodel = cmodel.to(dev).eval()
        cmodel.precision = args.precision
        from aline_b200 import prior as dprior
        ctask = CESTask(n_context_init=1, n_query_init=CFG3["n_query"])
        cb = dprior.sample_batch_device(ctask, CFG3["B"], seed=77, device=dev)     # Philox: same batch on every rank
        cb = {k: cb[k].contiguous() for k in keys}
        cT = CFG3["T"] - 1

        def ces_roll(i):
            out_keep["ces"] = cmodel.rollout(AttrDict(dict(cb)), cT)

        for i in range(3):
            ces_roll(i)
        ms_croll, _ = timed(ces_roll, n_comp)
        cx, cy = ctask.unnormalise_design(out_keep["ces"].context_x), out_keep["ces"].context_y
        cth0 = cb["target_all"].reshape(CFG3["B"], -1)
        lo, hi = spce.shard_rows(CFG3["L"], rank, world)   # this rank's contrastive rows, keyed by their GLOBAL index
        crows = dprior.sample_theta_device(ctask, hi - lo + 1, CFG3["B"], seed=78, row_offset=lo, device=dev)
        crows[0] = cth0

        def ces_bound(i):
            m, s, lp0 = spce.spce_history(ctask.log_likelihood, cy, cx, crows, seq=None, skip_rows=1, check=False)
            if world > 1:
                m, s = spce.all_gather_partials(m, s)
            out_keep["ces_pce"] = spce.lse_combine(m, s, lp0)

        ces_bound(0)
        ms_cb, _ = timed(ces_bound, n_comp)
        ces = {"workload": "CES eval-final (cfg3): B=20, n_query=2000, T=15, L=1e7 sharded over the ranks",
               "rollout_ms": ms_croll / n_comp,
               "rollout_design_steps_per_s": CFG3["B"] * cT / (ms_croll / n_comp * 1e-3),
               "spce_lsharded_ms": ms_cb / n_comp, "rows_per_rank": hi - lo,
               "spce_prior_samples_per_s": CFG3["L"] * CFG3["B"] / (ms_cb / n_comp * 1e-3),
               "spce_likelihood_evals_per_s": CFG3["L"] * CFG3["B"] * CFG3["T"] / (ms_cb / n_comp * 1e-3),
               "collective": "NCCL all-gather of the [B,T,2] (max, sum-exp) partials" if world > 1 else None}
        del crows
    except Exception as exc:  # noqa: BLE001
        ces = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    sampler.stop_flag = True

    ref_gpu = None
    if world == 1 and not args.no_ref_cuda:
        del rows_by_size, rows_one
        torch.cuda.empty_cache()
        ref_gpu = reference_cuda_eager(dev)

    pk = peaks()
    ms_step = ms / args.steps
    ms_e2e_step = ms_e2e / args.steps
    value = M * L / (ms_step * 1e-3)
    e2e_value = M * L / (ms_e2e_step * 1e-3)
    q_flops = query_stream_flops(B, nq - mid, n_c + n_sel)
    q_tf = q_flops / (ms_q * 1e-3) / 1e12
    step_bytes = (L + 1) * B * (4 * 2 + 8)
    hist_bytes = (L + 1) * B * 4 * 2
    ms_spce1, ms_roll1 = ms_spce / n_comp, ms_roll / n_comp
    clocks = sampler.summary()
    clk_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965        # MUFU peak at the clock measured under load
    n_batches_equiv = M / B
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16 operands / f32 accumulate (candidate stream), f32 elsewhere"
        if model.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": CFG["workload"],
                   "per_gpu": f"{n_local} of the {M} rollouts as {len(chunks)} mini-batches of "
                              f"{sorted({sz for _, sz in chunks})} (rank 0); bounds all-gathered (NCCL) in the timed region"
                   if world > 1 else f"all {M} rollouts as {len(chunks)} mini-batches of {B}",
                   "pipeline": "mini-batch i+1's rollout (CUDA graph, high-priority stream) overlaps mini-batch i's sPCE",
                   "weights": "random-init ALINE d=32 ff=128 h=4 3 layers (seed 123)",
                   "l2": "thetas [1000001,B,1,2] fp32 = 8 MB x B per pass exceed the 126 MB L2; rollout working set "
                         "(~60 MB) is L2-resident by design and re-used across the 34 dependent steps"},
        "components": {
            "ms_per_minibatch_of_200": ms_step / n_batches_equiv * world,
            "serial_loop_ms_per_step": ms_serial / n_comp,
            "pipeline_gain": (ms_serial / n_comp) / ms_step,
            "rollout_design_steps_per_s": B * steps_T / (ms_roll1 * 1e-3),
            "rollout_ms": ms_roll1,
            "spce_prior_samples_per_s": L * B / (ms_spce1 * 1e-3),
            "spce_ms": ms_spce1,
            "spce_incl_theta_sampling_ms": ms_spce_draw / n_comp,
            "spce_incl_theta_sampling_prior_samples_per_s": L * B / (ms_spce_draw / n_comp * 1e-3),
            "spce_likelihood_evals_per_s": L * B * T / (ms_spce1 * 1e-3),
            "query_launch_ms": ms_q, "query_tc4_two_threads_per_row_launch_ms": ms_q3,
            "cfg3_ces": ces},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e_step, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4 * T * 4,
                "note": "eval_boed(batches=pinned host mini-batches): per-rank bytes; thetas are drawn on the device inside "
                        "compute_EIG_from_history, as the reference does; the result read back is the reference's summary "
                        "(pce/nmc mean and error per step)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": "query_tc3_kernel<4, fold> (tcgen05 bf16 x bf16 -> fp32 in TMEM, softmax probabilities / MLP "
                               "activations as TMEM A operands, 4 tiles in flight per SM, MMAs issued by each warpgroup's first "
                               "warp from uniform-register descriptors; query / output projections folded into the key / value "
                               "operands the context kernel emits: 13 MMA phases per tile; candidate tokens through 3 encoder "
                               "layers + acquisition MLP), mid-rollout launch (20 keys)"
                     if model.precision == "bf16"
                     else "query_stream_kernel<32> (fp32 FFMA), mid-rollout launch", "bound": "tensor", "achieved": q_tf,
                     "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": q_tf / pk["tf_sust"],
                     "traffic": ncu_traffic("query_tc3_kernel") if model.precision == "bf16" else None,
                     "traffic_warm_l2": ncu_traffic("query_tc3_kernel_warm_l2") if model.precision == "bf16" else None,
                     "traffic_note": "traffic = ncu default (caches flushed before each replay: the 51 MB of candidate "
                                     "embeddings come from DRAM); traffic_warm_l2 = the same launch with --cache-control "
                                     "none, i.e. as inside the rollout, where the embeddings stay L2-resident across the 34 "
                                     "steps; algorithmic bytes per launch ~4.8 MB (2 floats in, 1 logit out per candidate)",
                     "peak_source": pk["src"] + ", sustained bf16 (kernel timed inside a long step)",
                     "launch_ms": ms_q, "algorithmic_flops_per_launch": q_flops,
                     "share_of_minibatch": (ms_q * steps_T) / (ms_roll1 + ms_spce1)},
        "rooflines": [
            {"kernel": "spce_step_tma_kernel (EIGStepLoss.step drop-in, one launch: cp.async.bulk ring, in-place update, "
                       "bulk store, fused theta_0 row + merge; timed through the C-ABI call)", "bound": "hbm",
             "achieved": step_bytes / (ms_s1 * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
             "frac": step_bytes / (ms_s1 * 1e-3) / 1e9 / pk["hbm"], "launch_ms": ms_s1,
             "algorithmic_bytes_per_launch": step_bytes, "traffic": ncu_traffic("spce_step_tma_kernel")},
            {"kernel": "spce_fast_loc_onepass_kernel (fused-history sPCE, location K=1, D=2): whole history in one pass over "
                       "the thetas, 4 rows per thread, shifted accumulation, packed fp32x2 pairs, one MUFU reciprocal per "
                       "eight evaluations; MUFU-bound: 2.125 MUFU per likelihood evaluation, XU pipe 16 lanes/clk/SM -- HBM "
                       "shown for reference",
             "bound": "hbm", "achieved": hist_bytes / (ms_spce1 * 1e-3) / 1e9,
             "peak": pk["hbm"], "unit": "GB/s", "frac": hist_bytes / (ms_spce1 * 1e-3) / 1e9 / pk["hbm"],
             "algorithmic_bytes_per_eval": hist_bytes,
             "mufu_bound": {"mufu_per_evaluation": 2.125, "evaluations": L * B * T,
                            "achieved_mufu_per_s": 2.125 * L * B * T / (ms_spce1 * 1e-3),
                            "peak_mufu_per_s": 16 * 148 * clk_mhz * 1e6,
                            "frac": 2.125 * L * B * T / (ms_spce1 * 1e-3) / (16 * 148 * clk_mhz * 1e6)}}],
    }
    if "spce_lsharded_ms" in ces:
        ces_bytes = (CFG3["L"] // world + 1) * CFG3["B"] * 4 * 5          # this rank's thetas (5 floats per draw), read once
        line["rooflines"].append({
            "kernel": "spce_stream_kernel<CesLik,18> (cfg3 CES bound, all 15 history points in one pass over this rank's "
                      "contrastive rows): ISSUE-bound -- ~200 warp instructions per likelihood evaluation (six 2^(rho log2 x) "
                      "with hi+lo products, two accurate log2, censored sigmoid-normal log-prob), ncu issue slots 81 % busy "
                      "(profiles/r2_ces_spce_after_ncu.txt); HBM shown for reference",
            "bound": "hbm", "achieved": ces_bytes / (ces["spce_lsharded_ms"] * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
            "frac": ces_bytes / (ces["spce_lsharded_ms"] * 1e-3) / 1e9 / pk["hbm"], "launch_ms": ces["spce_lsharded_ms"],
            "algorithmic_bytes_per_launch": ces_bytes, "traffic": None,
            "issue_bound": {"warp_instructions_per_evaluation": 199, "issue_active_pct_ncu": 80.8,
                            "evaluations_per_s": ces["spce_likelihood_evals_per_s"] / world}})
    if ref_gpu is not None:
        line["components"]["reference_cuda_eager"] = ref_gpu
        if "batch_ms" in ref_gpu:
            line["components"]["reference_cuda_eager_ms"] = ref_gpu["batch_ms"]
            line["components"]["speedup_vs_reference_cuda_eager"] = {
                "resident": ref_gpu["batch_ms"] / (ms_step / n_batches_equiv),
                "e2e": ref_gpu["batch_ms"] / (ms_e2e_step / n_batches_equiv),
                "rollout": ref_gpu["rollout_ms"] / ms_roll1, "spce": ref_gpu["spce_ms"] / (ms_spce_draw / n_comp),
                "note": "per mini-batch of 200 rollouts x L=1e6, same B200, reference unmodified (baseline/_ref)"}
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_subprocess()
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
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference-on-the-same-GPU leg")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="candidate-query stream: tcgen05 bf16 (default) or fp32 FFMA validation mode")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
