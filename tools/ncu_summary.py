#!/usr/bin/env python3
"""Summarise an .ncu-rep: key metrics + stall breakdown per kernel launch (reads `ncu --page raw --csv`)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__grid_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
for r in rows[2:]:
    for k in keys:
        if k in hdr:
            print(f"{k:70s} {r[hdr.index(k)]}")
    st = [(h, float(r[hdr.index(h)].replace(',', ''))) for h in hdr
          if h.startswith('smsp__pcsamp_warps_issue_stalled_') and 'not_issued' not in h and r[hdr.index(h)]]
    tot = sum(v for _, v in st) or 1
    for h, v in sorted(st, key=lambda x: -x[1])[:10]:
        print(f"   stall {h.replace('smsp__pcsamp_warps_issue_stalled_', ''):30s} {100 * v / tot:5.1f}%")
    print('---')
