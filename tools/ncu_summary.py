#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.txt
"""
import csv
import io
import re
import subprocess
import sys
import collections

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__warps_eligible.avg.per_cycle_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["ncu --set full --clock-control none summary of %s" % rep, ""]
    for r in data:
        lines.append("kernel: %s  grid %s block %s" % (r[idx["Kernel Name"]][:110], r[idx.get("Grid Size", 0)], r[idx.get("Block Size", 0)]))
        for k in KEYS:
            if k in idx:
                lines.append("  %-86s %14s %s" % (k, r[idx[k]], units[idx[k]]))
        lines.append("")
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                          capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(sass)))
    for k, r in enumerate(srows):
        if r and r[0] == "Address":
            h = {x: i for i, x in enumerate(r)}
            body = []
            for rr in srows[k + 1:]:
                if rr and rr[0] == "Kernel Name":
                    break
                if len(rr) > h["Instructions Executed"]:
                    body.append(rr)
            ops = collections.Counter()
            tot = 0
            for rr in body:
                m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", rr[h["Source"]])
                op = m.group(2).split(".")[0] if m else "?"
                c = int(rr[h["Instructions Executed"]] or 0)
                ops[op] += c
                tot += c
            lines.append("SASS opcode mix of the first profiled launch (share of %d executed warp instructions):" % tot)
            lines.append("  " + "  ".join("%s %.1f%%" % (o, 100.0 * c / tot) for o, c in ops.most_common(14)))
            break
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
