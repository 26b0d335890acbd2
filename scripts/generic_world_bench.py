#!/usr/bin/env python
"""Tick time of a NON-bundle world on the generic one-launch program (generic_program.cuh) next to the stepwise path and
to the particles bundle at the same entity count: Score (optional, checksummed, +1 per frame), Health (optional,
saturating-sub + despawn), Tag (12 B, checksummed) — the presence world of tests/test_gpu_component_presence.py — driven
by a SyncTest session with check_distance 8.   usage: generic_world_bench.py <entities> <ticks>"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bevy_ggrs_b200 import capi  # noqa: E402
from bevy_ggrs_b200.engine import Engine  # noqa: E402
from bevy_ggrs_b200.session import SAVE, SyncTestSession  # noqa: E402
from bevy_ggrs_b200.stress import populate, register_particles, synth_particles  # noqa: E402


def presence_world(n, flags):
    OPT = capi.BGR_STRATEGY_OPTIONAL
    w = Engine(max_entities=n, max_depth=9, flags=flags)
    score = w.rollback_component("Score", 4, capi.BGR_STRATEGY_COPY | OPT)
    health = w.rollback_component("Health", 4, capi.BGR_STRATEGY_CLONE | OPT)
    tag = w.rollback_component("Tag", 12, capi.BGR_STRATEGY_COPY)
    for c, ln in ((score, 4), (tag, 12), (health, 4)):
        w.checksum_component(c, 0, ln)
    w.add_system(capi.BGR_SYS_U32_ADD, [score], [0, 1])
    w.add_system(capi.BGR_SYS_U32_SATSUB_DESPAWN, [health], [0, 1])
    w.build()
    w.spawn(n)
    rng = np.random.default_rng(5)
    w.write_component(score, 0, rng.integers(0, 1000, n, dtype=np.uint32))
    w.write_component(health, 0, rng.integers(100000, 200000, n, dtype=np.uint32))
    w.write_component(tag, 0, rng.integers(0, 2**32, (n, 3), dtype=np.uint32))
    for r in range(0, n, 7):
        w.remove_component(score, r) if r % 2 else None
    return w


def particles_world(n, flags):
    w = Engine(max_entities=n, max_depth=9, flags=flags)
    cols = register_particles(w)
    w.build()
    populate(w, cols, *synth_particles(n, 1, 100000, 100000))
    return w


def particles_world_optional(n, flags):
    """the particles bundle with Velocity and Ttl registered optional (fused kernel MODE 2), a few components removed"""
    OPT = capi.BGR_STRATEGY_OPTIONAL
    w = Engine(max_entities=n, max_depth=9, flags=flags)
    t = w.rollback_component("Transform", 40, capi.BGR_STRATEGY_CLONE)
    v = w.rollback_component("Velocity", 12, capi.BGR_STRATEGY_COPY | OPT)
    l = w.rollback_component("Ttl", 8, capi.BGR_STRATEGY_COPY | OPT)
    w.checksum_component(v, 0, 12, capi.BGR_HASH_FLAG_ASSERT_FINITE_F32)
    w.checksum_component(t, 0, 12, capi.BGR_HASH_FLAG_ASSERT_FINITE_F32)
    w.add_system(capi.BGR_SYS_PARTICLES_UPDATE, [t, v])
    w.add_system(capi.BGR_SYS_PARTICLES_DESPAWN, [l])
    w.build()
    tf, vel, ttl = synth_particles(n, 1, 100000, 100000)
    w.spawn(n)
    w.write_component(t, 0, tf); w.write_component(v, 0, vel); w.write_component(l, 0, ttl)
    for r in range(0, min(n, 2000), 7):
        w.remove_component(v if r % 2 else l, r)
    return w


def run(w, ticks):
    sess = SyncTestSession(2, 8, 9, input_delay=2)
    vecs = []
    for t in range(ticks + 12):
        for h in range(2):
            sess.add_local_input(h, 0)
        reqs = sess.advance_frame()
        for r in reqs:
            if r.kind == SAVE:
                sess.save_cell(r.frame, 0)
        vecs.append((capi.make_session_info(sess.info()), capi.make_requests(reqs), len(reqs)))
    for info, arr, n in vecs[:12]:
        w.submit_prepared(info, arr, n); w.collect()
    l0 = w.launch_count()
    traced = w.last_path_fused()
    if traced:
        w.trace_enable(ticks + 4)
    t0 = time.perf_counter()
    for info, arr, n in vecs[12:]:
        w.submit_prepared(info, arr, n); w.collect()
    dt = (time.perf_counter() - t0) / ticks
    out = {"sync_us_per_tick": dt * 1e6, "launches_per_tick": (w.launch_count() - l0) / ticks, "one_launch": w.last_path_fused(),
           "specialised_kernel": w.generic_specialised()}  # True: the registration's own NVRTC-compiled kernel ran
    if traced:
        tr = w.trace_read(ticks + 4)
        out["kernel_us_median"] = float(np.median((tr[:, 1].astype(np.int64) - tr[:, 0].astype(np.int64)) / 1e3))
        w.trace_enable(0)
    # pipelined: up to four request vectors un-collected (the same vectors again: a SyncTest tick is Load + re-simulate)
    if os.environ.get("GENERIC_BENCH_PIPELINED"):
        more = []
        for t in range(ticks):
            for h in range(2):
                sess.add_local_input(h, 0)
            reqs = sess.advance_frame()
            for r in reqs:
                if r.kind == SAVE:
                    sess.save_cell(r.frame, 0)
            more.append((capi.make_session_info(sess.info()), capi.make_requests(reqs), len(reqs)))
        w.synchronize()
        t0 = time.perf_counter()
        inflight = 0
        for info, arr, n in more:
            w.submit_prepared(info, arr, n)
            inflight += 1
            if inflight == 4:
                w.collect(); inflight -= 1
        while inflight:
            w.collect(); inflight -= 1
        out["pipelined_us_per_tick"] = (time.perf_counter() - t0) / ticks * 1e6
    return out


def main():
    n, ticks = int(sys.argv[1]), int(sys.argv[2])
    out = {"entities": n, "ticks": ticks, "check_distance": 8}
    for name, make, flags in (("presence_world_generic_program", presence_world, 0),
                              ("presence_world_stepwise", presence_world, capi.BGR_CFG_FORCE_STEPWISE),
                              ("particles_bundle", particles_world, 0),
                              ("particles_bundle_with_optional_columns", particles_world_optional, 0)):
        w = make(n, flags)
        out[name] = run(w, ticks)
        w.close()
    out["optional_bundle_vs_bundle"] = out["particles_bundle_with_optional_columns"]["sync_us_per_tick"] / out["particles_bundle"]["sync_us_per_tick"]
    out["generic_vs_bundle"] = out["presence_world_generic_program"]["sync_us_per_tick"] / out["particles_bundle"]["sync_us_per_tick"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
