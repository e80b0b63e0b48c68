#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep [regex] -- the metrics profiles/*_ncu_summary.txt quote, one block per captured launch
(reads `ncu -i REPORT --page raw --csv`).

tools/ncu_summary.py --facts REPORT.ncu-rep REGEX WORKLOAD FP64_INST_PER_UPDATE FP64_SOURCE  -- one JSON object for
profiles/shell_halos_ncu_facts.json (what bench.py's roofline reports as measured: DRAM bytes per launch, pipe utilisations)."""
import csv
import io
import os
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_red.sum",
    "lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_atom.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def facts():
    import json
    rep, pat, workload, fp64, src = sys.argv[2], re.compile(sys.argv[3]), sys.argv[4], float(sys.argv[5]), sys.argv[6]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    r = next(r for r in rows[2:] if pat.search(r[col["Kernel Name"]]))

    def f(name, scale=1.0):
        unit = rows[1][col[name]]
        v = float(r[col[name]].replace(",", ""))
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(unit, 1.0)
        return v * mult * scale
    print(json.dumps({
        "workload": workload, "source": "profiles/" + os.path.basename(rep).replace(".ncu-rep", "_ncu_summary.txt"),
        "dram_bytes_per_launch": f("dram__bytes_read.sum") + f("dram__bytes_write.sum"),
        "kernel_ms_under_ncu": f("gpu__time_duration.sum") * (1.0 if rows[1][col["gpu__time_duration.sum"]] == "ms" else 1e-6),
        "fp64_inst_per_update": fp64, "fp64_inst_source": src,
        "fp64_pipe_active_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_inst_executed": f("smsp__inst_executed.sum"),
        "l2_red_sectors": f("lts__t_sectors_srcunit_tex_op_red.sum"),
        "registers_per_thread": f("launch__registers_per_thread")}))


def main():
    if sys.argv[1] == "--facts":
        return facts()
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if pat and not pat.search(name):
            continue
        print("# launch id %s  %s  grid %s block %s" % (r[col["ID"]], name[:110], r[col.get("Grid Size", 0)],
                                                        r[col.get("Block Size", 0)]))
        for m in METRICS:
            if m in col:
                print("%-90s%-18s%s" % (m, units[col[m]], r[col[m]]))
        print()


if __name__ == "__main__":
    main()
