"""Summarises an `ncu -i X.ncu-rep --page raw --csv` dump: one block per distinct kernel with the metrics the profiles/
summaries quote.  Usage: python tools/ncu_summary.py raw.csv "title line" > profiles/NAME.txt"""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__waves_per_multiprocessor',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed.sum']

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
print(sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
print()
seen = set()
for r in rows[2:]:
    name = r[idx['Kernel Name']].replace('void ', '').replace('unnamed>::', '').split('(')[0]
    if name in seen:
        continue
    seen.add(name)
    print('== ' + name)
    for w in WANT + stall:
        if w in idx:
            print(f"{w:95s} {r[idx[w]]:>20s} {units[idx[w]]}")
    print()
