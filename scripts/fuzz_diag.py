#!/usr/bin/env python
"""Diagnostic driver for tests/test_gpu_fuzz_requests.py: runs every seed on both paths, reports the first diverging step
with the world's configuration and a state diff (GPU box only)."""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_fuzz_requests as fz  # noqa: E402
from bevy_ggrs_b200 import capi  # noqa: E402


def state_diff(eng, orc, cols):
    rows = min(eng.row_count(), orc.row_count())
    out = {"rows": (eng.row_count(), orc.row_count())}
    ae, ao = eng.read_alive(0, rows).astype(bool), orc.read_alive(0, rows).astype(bool)
    out["alive_diff_rows"] = np.flatnonzero(ae != ao)[:10].tolist()
    for c in cols:
        vo, ho = orc.read_component_alive(c, 0, rows)
        he = eng.has_component(c, 0, rows).astype(bool)
        ve = eng.read_component(c, 0, rows)
        both = he & ho.astype(bool)
        bad = np.flatnonzero(both & (ve != vo).any(axis=1))
        out[f"col{c}"] = {"presence_diff": np.flatnonzero(he != ho.astype(bool))[:10].tolist(), "value_diff_rows": bad[:10].tolist(),
                          "sample": (ve[bad[0]].tolist(), vo[bad[0]].tolist()) if bad.size else None}
    return out


def main():
    for flags in (0, capi.BGR_CFG_FORCE_STEPWISE):
        for seed in range(12):
            rng = np.random.default_rng(1000 + seed)
            eng, orc, cols, sizes, optional, depth = fz._make_worlds(rng, flags)
            try:
                fz._drive(eng, orc, cols, sizes, optional, rng, flags, seed)
                print(f"flags={flags} seed={seed}: ok", flush=True)
            except AssertionError as ex:
                msg = str(ex).splitlines()[0][:160]
                print(f"flags={flags} seed={seed}: FAIL {msg}\n   sizes={sizes} optional={optional} depth={depth} n={orc.row_count()}", flush=True)
                try:
                    print("   diff:", state_diff(eng, orc, cols), flush=True)
                except Exception:
                    traceback.print_exc()
            except Exception:
                traceback.print_exc()
            eng.close(); orc.close()


if __name__ == "__main__":
    main()
