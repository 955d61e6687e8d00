#!/usr/bin/env python
"""Summarise an .ncu-rep (`ncu --set full`) as text + JSON: per kernel launch the duration, launch shape, issue / pipe
utilisation, DRAM bytes and the stall reasons per issue slot.  The JSON (profiles/*_traffic.json) is what bench.py reads its
`roofline.traffic` from: a number measured by a committed capture, not a literal.

    python tools/ncu_summary.py gpurun_out/x/prof.ncu-rep [--json out.json] [--label text]"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warp_latency_per_inst_issued.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_active.avg",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    rep = sys.argv[1]
    jout = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    label = sys.argv[sys.argv.index("--label") + 1] if "--label" in sys.argv else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    out = []
    print("# %s" % label)
    for r in data:
        name = r[col["Kernel Name"]]
        print("Kernel Name  %s" % name[:110])
        rec = {"kernel": name}
        for m in METRICS:
            if m in col:
                print("  %-78s %s %s" % (m, r[col[m]], units[col[m]]))
                try:
                    rec[m] = float(r[col[m]].replace(",", "")) * UNIT_SCALE.get(units[col[m]], 1.0)
                except ValueError:
                    pass
        issued = None
        for n in hdr:
            if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio"):
                key = n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
                try:
                    v = float(r[col[n]])
                except ValueError:
                    continue
                if v >= 0.02:
                    print("  stall/issue %-30s %.6f" % (key, v))
                    rec["stall_" + key] = v
        if "dram__bytes_read.sum" in rec and "dram__bytes_write.sum" in rec:
            rec["dram_bytes"] = rec["dram__bytes_read.sum"] + rec["dram__bytes_write.sum"]
            print("  dram bytes per launch (read + write)                                           %.4e" % rec["dram_bytes"])
        out.append(rec)
        print()
    if jout:
        json.dump({"source": label, "launches": out}, open(jout, "w"), indent=1)


if __name__ == "__main__":
    main()
