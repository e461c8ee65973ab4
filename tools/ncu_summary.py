#!/usr/bin/env python
"""Key `ncu --set full` metrics per captured launch:  python tools/ncu_summary.py report.ncu-rep [json-out]"""
import csv, io, json, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "launch__shared_mem_per_block_dynamic"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
res = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    e = {"kernel": d.get("Kernel Name")}
    for k in WANT:
        if k in d:
            e[k] = d[k] + " " + units[hdr.index(k)]
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            e[k.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "")] = d[k]
    res.append(e)
if len(sys.argv) > 2:
    json.dump(res, open(sys.argv[2], "w"), indent=1)
for e in res:
    print(json.dumps(e, indent=0))
