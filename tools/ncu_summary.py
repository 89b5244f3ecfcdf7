#!/usr/bin/env python3
"""Condense an .ncu-rep (read here, without a GPU) into the markdown table kept under profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rN_ncu_summary.md"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):      # already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [k for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k]
    print("# ncu summary of `%s`\n" % rep)
    print("Captured with `ncu --set full --clock-control none --import-source on` under gpurun; read with "
          "`ncu -i ... --page raw --csv` (tools/ncu_summary.py).  Per-launch values.\n")
    for r in rows[2:]:
        print("## %s\n" % r[idx["Kernel Name"]])
        print("| metric | value |\n|---|---|")
        for k in WANT:
            if k in idx:
                print("| `%s` | %s %s |" % (k, r[idx[k]], units[idx[k]]))
        st = sorted(((float(r[idx[k]].replace(",", "")), k) for k in stalls), reverse=True)[:6]
        print("| top stall reasons (warps per issue) | %s |" % ", ".join(
            "%s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
            for v, k in st))
        print()


if __name__ == "__main__":
    main()
