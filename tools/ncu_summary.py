#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into a small JSON for profiles/: python tools/ncu_summary.py in.ncu-rep out.json"""
import csv, json, subprocess, sys, io
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__grid_size', 'launch__block_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.max', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    d = {k: (r[hdr.index(k)] + ' ' + units[hdr.index(k)]).strip() for k in KEYS if k in hdr}
    for i, h in enumerate(hdr):
        if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 2.0:
                d[h.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('smsp__average_warp_latency_issue_stalled_', 'stall_')] = round(v, 1)
    out.append(d)
json.dump(out, open(sys.argv[2], 'w'), indent=1)
for d in out:
    print(json.dumps(d, indent=1))
