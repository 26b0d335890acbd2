#!/usr/bin/env python
"""Fixed vs per-frame cost of one synchronous SyncTest tick: the same world at check_distance 1, 2, 4, 8 (a tick is
1 Load + d Save + (d+1) Advance), kernel duration from the device trace, least-squares line.  usage: frame_cost_fit.py [entities ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import bench  # noqa: E402
import sync_sweep  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [100_000]
    for n in sizes:
        pts = []
        for d in (1, 2, 4, 8):
            bench.WORKLOADS[f"fit_{n}_{d}"] = (n, d, 9)
            r = sync_sweep.one(f"fit_{n}_{d}", {}, K=200)
            pts.append((d, r["kernel"]["kernel_us_median"], r["sync_us_per_tick"]))
        import numpy as np
        x = np.array([p[0] for p in pts], dtype=float)
        for name, col in (("kernel_us", 1), ("sync_call_us", 2)):
            y = np.array([p[col] for p in pts])
            slope, icpt = np.polyfit(x + 1, y, 1)   # frames advanced per tick = d + 1
            print(json.dumps({"entities": n, "what": name, "points": {int(a): float(b) for a, b in zip(x, y)},
                              "per_frame_us": float(slope), "fixed_us": float(icpt)}), flush=True)


if __name__ == "__main__":
    main()
