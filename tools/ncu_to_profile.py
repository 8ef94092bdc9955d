#!/usr/bin/env python3
"""Turn .ncu-rep files (ncu --set full) into the JSON summary committed under profiles/: one dict per launch with
the metrics the roofline numbers in bench.py / DESIGN.md are read from, plus every stall ratio.
usage: ncu_to_profile.py out.json rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv, io, json, subprocess, sys

KEEP = ['launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.avg', 'sm__cycles_active.avg', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum']
out = []
for rep in sys.argv[2:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {"report": rep.split("/")[-1], "Kernel Name": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                d[h + (" [%s]" % units[i] if units[i] else "")] = r[i]
        out.append(d)
json.dump(out, open(sys.argv[1], "w"), indent=1)
print("wrote", sys.argv[1], len(out), "launches")
