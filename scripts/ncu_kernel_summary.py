"""Prints the key metrics + stall samples of the first kernel in an ncu report exported with
`ncu -i X.ncu-rep --page raw --csv > X.csv`.   usage: python scripts/ncu_kernel_summary.py X.csv [row]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
r = rows[2 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_fmaheavy.sum']
for w in want:
    if w in hdr:
        print("%-72s %s %s" % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
for i, h in enumerate(hdr):
    if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h and r[i] not in ('0', ''):
        print("%-72s %s" % (h, r[i]))
