#!/usr/bin/env python3
"""Print the handful of ncu raw-page metrics we steer by, for every launch in a .ncu-rep."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed.sum','lts__t_sector_hit_rate.pct','sm__cycles_elapsed.avg','lts__t_sectors.sum','lts__t_requests.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','sm__cycles_active.avg','lts__cycles_elapsed.avg','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum', 'dram__cycles_elapsed.avg.per_second','lts__cycles_elapsed.avg.per_second','sm__cycles_elapsed.avg.per_second']
for r in rows[2:]:
    print('---')
    for k in keys:
        if k in hdr: print(' ', k, units[hdr.index(k)], r[hdr.index(k)])
    st = []
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            v = float(r[i])
            if v > 0.25: st.append((v, h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
    print('  stalls/issue:', ' '.join('%s=%.2f' % (n, v) for v, n in sorted(st, reverse=True)))
