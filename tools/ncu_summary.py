#!/usr/bin/env python
"""Summarise an ncu launch list (csv) and/or a full .ncu-rep into a short text block for profiles/."""
import csv, subprocess, sys
from collections import defaultdict

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.max',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_op_global_red.sum',
        'smsp__inst_executed_op_global_ld.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'smsp__average_warp_latency_issue_stalled_barrier.ratio', 'sm__sass_thread_inst_executed_op_dfma_pred_on.sum',
        'sm__sass_thread_inst_executed_op_dmul_pred_on.sum', 'sm__sass_thread_inst_executed_op_dadd_pred_on.sum']


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = defaultdict(list)
    for r in rows[1:]:
        agg[r[ki][:70]].append(float(r[vi].replace(',', '')))
    print(f"# launch list {path} (gpu__time_duration, cold-cache, serialised)")
    for k, v in agg.items():
        print(f"{k:72s} n={len(v):3d} mean_us={sum(v)/len(v)/1e3:10.1f}")


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    print(f"# full capture {path}")
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        print("kernel:", name[:80])
        for i, h in enumerate(hdr):
            if h in KEYS or h.startswith('smsp__average_warps_issue_stalled') or (h.startswith('smsp__warp_issue_stalled') and h.endswith('_per_warp_active.pct')):
                print(f"  {h:80s} {rows[1][i]:>12s} {r[i]}")


if __name__ == '__main__':
    for a in sys.argv[1:]:
        (launches if a.endswith('.csv') else full)(a)
