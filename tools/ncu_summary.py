#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the small tracked artefacts under profiles/:
   <out>.metrics.csv  key raw metrics per captured launch
   <out>.opcodes.csv  executed warp instructions and stall samples per SASS opcode
and optionally update profiles/traffic.json (dram bytes per launch for bench.py's roofline.traffic)."""
import collections
import csv
import json
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    workload = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    with open(out + ".metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}:{r[kn][:40]}" for i, r in enumerate(data)])
        for m in KEEP:
            if m in hdr:
                i = hdr.index(m)
                w.writerow([m, units[i]] + [r[i] for r in data])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    hidx = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
    if hidx:
        h = srows[hidx[0]]
        blk = srows[hidx[0] + 1: hidx[1] - 1 if len(hidx) > 1 else None]
        ia, isrc, ismp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
        tot, samp = collections.Counter(), collections.Counter()
        for r in blk:
            if len(r) <= ia or not r[ia].isdigit():
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
            op = (m.group(2) if m else r[isrc][:16]).split(".")[0]
            tot[op] += int(r[ia]); samp[op] += int(r[ismp])
        with open(out + ".opcodes.csv", "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["opcode", "warp_instructions_executed", "pct", "stall_samples", "pct_samples"])
            T, S = sum(tot.values()) or 1, sum(samp.values()) or 1
            for op, c in tot.most_common():
                w.writerow([op, c, f"{100*c/T:.2f}", samp[op], f"{100*samp[op]/S:.2f}"])
    if workload:
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        vals = [float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]] for r in data]
        path = "profiles/traffic.json"
        try:
            t = json.load(open(path))
        except Exception:
            t = {}
        t[workload] = sum(vals) / len(vals)
        t[workload + "_source"] = out + ".metrics.csv"
        du = hdr.index("gpu__time_duration.sum")
        us = {"usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "nsecond": 1e-3, "ns": 1e-3, "second": 1e6, "s": 1e6}
        iso = t.setdefault("isolated_launch_us", {})
        iso[workload] = sum(float(r[du]) * us[units[du]] for r in data) / len(data)   # serialised, cold-cache launches under ncu
        json.dump(t, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
