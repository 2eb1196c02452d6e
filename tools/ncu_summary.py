#!/usr/bin/env python
"""ncu_summary.py <report.ncu-rep> — key metrics of each profiled launch (reads `ncu --page raw --csv`)."""
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__inst_executed.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.per_cycle_active',
    'smsp__inst_executed.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'smsp__inst_executed_op_shared_ld.sum', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
]
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
print('launches:', [d[col['Kernel Name']] for d in data])
for k in KEYS:
    if k in col:
        print(f'{k:75s} {units[col[k]]:14s}', [d[col[k]] for d in data])
stall = [h for h in hdr if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('_not_issued.ratio') is False and h.endswith('.ratio')]
vals = sorted(((float(data[-1][col[h]] or 0), h) for h in stall), reverse=True)
print('top stall reasons (warp latency per issued instr, last launch):')
for v, h in vals[:10]:
    print(f'   {v:8.3f}  {h}')
