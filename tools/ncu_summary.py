"""Summarise ncu artefacts brought back in gpurun_out/ into small text files under profiles/.
    python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches.txt
    python tools/ncu_summary.py full gpurun_out/rollout_r1.ncu-rep profiles/r1_rollout_fp32_ncu.txt
"""
import collections
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum']
STALL = 'smsp__average_warps_issue_stalled_'


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows:
        name = r[4].split('(')[0][:90]
        v = float(r[-1].replace(',', ''))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(dst, 'w') as f:
        f.write(f"# {src}: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised; compare SHARES)\n")
        f.write(f"# launches {len(rows)}, total {tot / 1e6:.3f} ms\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t / tot * 100:6.2f}%  n={n:4d}  avg={t / n / 1e3:10.1f} us  {k}\n")


def full(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, 'w') as f:
        f.write(f"# {src}: ncu --set full --clock-control none --import-source on (selected raw metrics per launch)\n")
        for r in rows[2:]:
            f.write('----\n')
            for w in WANT:
                if w in idx:
                    f.write(f"{w} = {r[idx[w]][:110]} {units[idx[w]]}\n")
            for h in hdr:
                if h.startswith(STALL) and h.endswith('_per_issue_active.ratio'):
                    v = r[idx[h]]
                    try:
                        if float(v) >= 0.1:
                            f.write(f"stall {h[len(STALL):-len('_per_issue_active.ratio')]} = {v}\n")
                    except ValueError:
                        pass


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
