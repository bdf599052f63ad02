"""Print the headline metrics of every kernel in an ncu report (raw page CSV).
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'launch__shared_mem_per_block_dynamic', 'smsp__average_warp_latency_per_inst_issued.ratio']
for r in rows[2:]:
    print('----')
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print(f"{k:72s} {r[i]:>22s} {units[i]}")
    names = [h for h in hdr if 'average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
    vals = sorted([(float(r[hdr.index(n)].replace(',', '') or 0), n) for n in names], reverse=True)[:8]
    for v, n in vals:
        print(f"   stall {v:7.2f}  {n.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')}")
