#!/usr/bin/env python
"""particles.bin for oracle/ref_harness: n x (10 f32 Transform | 3 f32 Velocity | u64 Ttl), the SAME seeded population
the parity tests use (bevy_ggrs_b200.stress.synth_particles).   python tests/golden/gen_reference_inputs.py out.bin n seed ttl_lo ttl_hi"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bevy_ggrs_b200.stress import synth_particles  # noqa: E402


def main():
    out, n, seed, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3], 0), int(sys.argv[4]), int(sys.argv[5])
    tf, vel, ttl = synth_particles(n, seed, lo, hi)
    rec = np.zeros(n, dtype=[("tf", "<f4", 10), ("vel", "<f4", 3), ("ttl", "<u8")])
    rec["tf"], rec["vel"], rec["ttl"] = tf, vel, ttl
    assert rec.dtype.itemsize == 60
    rec.tofile(out)


if __name__ == "__main__":
    main()
