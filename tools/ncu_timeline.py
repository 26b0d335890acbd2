#!/usr/bin/env python
"""PM-sampling time series of one captured launch (ncu --set full) -> small CSV under profiles/.
usage: ncu_timeline.py <rep> <out.csv> [launch_row]"""
import csv
import subprocess
import sys

NAMES = {
    "dram_read_pct": "FBSP.TriageCompute.dram__read_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_write_pct": "FBSP.TriageCompute.dram__write_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_inst_issue_pct": "TPC.TriageCompute.sm__inst_executed_realtime.avg.pct_of_peak_sustained_elapsed",
    "lts_tex_pct": "LTS.TriageCompute.lts__t_sector_throughput_srcunit_tex.avg.pct_of_peak_sustained_elapsed",
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    row = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    csv.field_size_limit(10 ** 9)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-metric-instances", "values"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, data = rows[0], rows[2:]

    def series(name):
        v = data[row][hdr.index(name)]
        return [float(x) for x in v[v.index("(") + 1:v.rindex(")")].split(";")]

    ser = {k: series(v) for k, v in NAMES.items()}
    inst = ser["sm_inst_issue_pct"]
    first = next(i for i, x in enumerate(inst) if x > 0)
    last = max(i for i, x in enumerate(ser["dram_write_pct"]) if x > 0.5)
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["sample_index_about_1us"] + list(ser))
        for i in range(first - 2, last + 3):
            w.writerow([i - first] + [ser[k][i] for k in ser])
    mid = range(first + 20, last - 15)
    print("samples %d, plateau dram read+write pct %.1f, issue pct %.1f" % (
        last - first, sum(ser["dram_write_pct"][i] + ser["dram_read_pct"][i] for i in mid) / len(mid),
        sum(inst[i] for i in mid) / len(mid)))


if __name__ == "__main__":
    main()
