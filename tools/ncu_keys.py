"""Key metrics of an `ncu --page raw --csv` export (one row per captured kernel launch)."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for v in rows[2:]:
    print('==', v[hdr.index('Kernel Name')][:90])
    for h, u, x in zip(hdr, units, v):
        stall = 'smsp__average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h
        if h in KEYS or (stall and float(x or 0) >= 0.2):
            print('  %-95s %-8s %s' % (h.replace('smsp__average_warps_issue_stalled_', 'stall_'), u, x))
