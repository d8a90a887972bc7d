"""Dev tool: pick the metrics the profiles quote out of `ncu -i X.ncu-rep --page raw --csv` output.

    ncu -i X.ncu-rep --page raw --csv > X.csv ; python tools/ncu_pick.py X.csv
"""
import csv
import json
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]

rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr = next(r for r in rows if "Kernel Name" in r)
units = rows[rows.index(hdr) + 1]
for r in rows[rows.index(hdr) + 2:]:
    if len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    out = {"kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "block": d.get("Block Size")}
    for k in WANT:
        if k in d:
            out[k] = d[k] + " " + units[hdr.index(k)]
    print(json.dumps(out))
