"""Per-entity component presence (BGR_STRATEGY_OPTIONAL) against the oracle: the four-way match of
ComponentSnapshotPlugin::load (component_snapshot.rs:99-115), save / checksum / system queries that only see entities
having the component, across tile boundaries and through a SyncTest-shaped run."""
import numpy as np
import pytest

from bevy_ggrs_b200 import capi
from bevy_ggrs_b200.capi import BgrError
from bevy_ggrs_b200.engine import Engine
from bevy_ggrs_b200.session import ADVANCE, LOAD, SAVE, Request
from oracle_backend import OracleWorld

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("generic_kernel")]
NOSESS = (capi.BGR_SESSION_NONE, 0, 0, 0)
OPT = capi.BGR_STRATEGY_OPTIONAL


def _pair(n, depth=8, flags=0):
    """Score (optional, checksummed, +1 per frame), Health (optional, satsub-despawn), Tag (always present, checksummed)."""
    worlds, cols = [], None
    for w in (Engine(max_entities=n + 8, max_depth=depth, flags=flags), OracleWorld()):
        score = w.rollback_component("Score", 4, capi.BGR_STRATEGY_COPY | OPT)
        health = w.rollback_component("Health", 4, capi.BGR_STRATEGY_CLONE | OPT)
        tag = w.rollback_component("Tag", 12, capi.BGR_STRATEGY_COPY)
        w.checksum_component(score, 0, 4)
        w.checksum_component(tag, 0, 12)
        w.checksum_component(health, 0, 4)
        w.add_system(capi.BGR_SYS_U32_ADD, [score], [0, 1])
        w.add_system(capi.BGR_SYS_U32_SATSUB_DESPAWN, [health], [0, 1])
        w.build()
        w.spawn(n)
        rng = np.random.default_rng(5)
        w.write_component(score, 0, rng.integers(0, 1000, n, dtype=np.uint32))
        w.write_component(health, 0, rng.integers(3, 40, n, dtype=np.uint32))
        w.write_component(tag, 0, rng.integers(0, 2**32, (n, 3), dtype=np.uint32))
        worlds.append(w)
        cols = (score, health, tag)
    return worlds[0], worlds[1], cols


def _same(eng, orc, cols, n):
    alive_e = eng.read_alive(0, n).astype(bool)
    for c in cols:
        vo, ho = orc.read_component_alive(c, 0, n)
        he = eng.has_component(c, 0, n).astype(bool)
        assert np.array_equal(he, ho.astype(bool)), f"presence of column {c}"
        assert np.array_equal(eng.read_component(c, 0, n)[he], vo[he]), f"values of column {c}"
    assert np.array_equal(alive_e, orc.read_alive(0, n).astype(bool))


@pytest.mark.parametrize("flags", [0, capi.BGR_CFG_FORCE_STEPWISE])
def test_optional_columns_take_the_generic_path_and_match_the_oracle(flags):
    """flags = 0: the generic ONE-launch program (generic_program.cuh); FORCE_STEPWISE: one launch per request and system."""
    n = 1300  # three tiles, the last one partial
    eng, orc, cols = _pair(n, flags=flags)
    score, health, tag = cols
    both = lambda f, *a: [getattr(w, f)(*a) for w in (eng, orc)]
    rows = [0, 1, 511, 512, 513, 1023, 1024, 1299]
    for r in rows[::2]:
        both("remove_component", score, r)
    both("remove_component", health, 512)
    l0 = eng.launch_count()
    a, b = both("handle_requests", NOSESS, [Request(SAVE, 0), Request(ADVANCE, 0, [0]), Request(SAVE, 1), Request(ADVANCE, 0, [0])])
    assert a == b and len(a) == 2
    assert eng.last_path_fused() == (flags == 0)
    assert (eng.launch_count() - l0 == 1) == (flags == 0)
    _same(eng, orc, cols, n)
    # change presence after the snapshots: every arm of the four-way match is hit by the Load below
    for r in rows[1::2]:
        both("remove_component", score, r)                       # (None, Some) -> insert
    for r in rows[::2][:2]:
        both("insert_component", score, r, np.uint32(4242))      # (Some, None) -> remove
    both("insert_component", health, 512, np.uint32(7))
    _same(eng, orc, cols, n)
    a, b = both("handle_requests", NOSESS, [Request(LOAD, 1), Request(ADVANCE, 0, [0]), Request(SAVE, 2)])
    assert a == b
    _same(eng, orc, cols, n)
    # peek: the snapshot of frame 0 holds Score exactly for the rows that had it then
    vals, had = eng.peek(0, score, 0, n)
    vo, ho = orc.peek(0, score, 0, n)
    assert np.array_equal(had.astype(bool), ho.astype(bool)) and 0 < ho.sum() < n
    assert np.array_equal(vals[ho.astype(bool)], vo[ho.astype(bool)])


@pytest.mark.parametrize("flags", [0, capi.BGR_CFG_FORCE_STEPWISE])
def test_synctest_shaped_run_with_presence_changes_between_ticks(flags):
    """Load(f-d), d x (Advance, Save) every tick, with removals / insertions applied between ticks (a system outside
    GgrsSchedule): rollbacks undo them for the resimulated frames exactly as in the oracle; entities also die (Health)."""
    n, d = 700, 4
    SESS = (capi.BGR_SESSION_SYNCTEST, 8, d, 0)  # max_prediction 8: the ring keeps 8 frames, confirmed = frame - d
    eng, orc, cols = _pair(n, flags=flags)
    score, health, tag = cols
    both = lambda f, *a: [getattr(w, f)(*a) for w in (eng, orc)]
    rng = np.random.default_rng(11)
    frame = 0
    for tick in range(14):
        reqs = []
        if tick >= d:
            reqs.append(Request(LOAD, frame - d))
            for k in range(d):
                reqs += [Request(ADVANCE, 0, [0]), Request(SAVE, frame - d + k + 1)] if k < d - 1 else [Request(ADVANCE, 0, [0])]
        reqs += [Request(SAVE, frame), Request(ADVANCE, 0, [0])]
        l0 = eng.launch_count()
        a, b = both("handle_requests", SESS, reqs)
        assert a == b, f"tick {tick}"
        assert (eng.launch_count() - l0 == 1) == (flags == 0)     # the whole request vector is ONE launch on the default path
        frame += 1
        alive = orc.read_alive(0, n).astype(bool)
        for r in rng.choice(np.flatnonzero(alive), 5, replace=False):
            col = (score, health)[int(rng.integers(2))]
            if orc.has_component(col, int(r), 1)[0]:
                both("remove_component", col, int(r))
            else:
                both("insert_component", col, int(r), np.uint32(rng.integers(5, 50)))
        _same(eng, orc, cols, n)
    assert 0 < orc.read_alive(0, n).sum() < n  # Health ran out for some entities inside the run


@pytest.mark.parametrize("grid", ["2", "5"])
def test_presence_world_with_many_tiles_per_block(monkeypatch, grid):
    """40 tiles on 2 / 5 blocks of the generic program (BGR_TUNE_GRID): the per-row presence bits, the optional columns'
    checksums and the despawns must come out like the oracle's when a block walks through many tiles."""
    monkeypatch.setenv("BGR_TUNE_GRID", grid)
    n, d = 20_000, 3
    SESS = (capi.BGR_SESSION_SYNCTEST, 8, d, 0)
    eng, orc, cols = _pair(n)
    score, health, tag = cols
    both = lambda f, *a: [getattr(w, f)(*a) for w in (eng, orc)]
    for r in range(0, n, 37):
        both("remove_component", score, r)
    frame = 0
    for tick in range(9):
        reqs = []
        if tick >= d:
            reqs.append(Request(LOAD, frame - d))
            for k in range(d):
                reqs += [Request(ADVANCE, 0, [0]), Request(SAVE, frame - d + k + 1)] if k < d - 1 else [Request(ADVANCE, 0, [0])]
        reqs += [Request(SAVE, frame), Request(ADVANCE, 0, [0])]
        l0 = eng.launch_count()
        a, b = both("handle_requests", SESS, reqs)
        assert a == b, f"tick {tick}"
        assert eng.launch_count() - l0 == 1 and eng.last_path_fused()
        frame += 1
    _same(eng, orc, cols, n)
    assert 0 < orc.read_alive(0, n).sum() < n


@pytest.mark.timeout(180)
@pytest.mark.parametrize("tiledep", ["0", "1", "2"])
@pytest.mark.parametrize("n", [3000, 120_000])
def test_pipelined_submits_on_the_generic_program(monkeypatch, generic_kernel, tiledep, n):
    """Four request vectors in flight (bgr_submit_requests / bgr_collect) on a world that is not the bundle.  With
    BGR_TUNE_JIT_TILEDEP=1 consecutive launches of the generated kernel overlap on the GPU — work item i of tick k+1 starts
    when item i of tick k has signalled (=2: also on synchronous calls).  Every tick's checksums, the final world and a
    snapshot equal the oracle's.  A dependency bug hangs or corrupts: bounded by pytest-timeout."""
    from bevy_ggrs_b200.session import SyncTestSession
    monkeypatch.setenv("BGR_TUNE_JIT_TILEDEP", tiledep)
    d, n_ticks = 3, 30 if n < 100_000 else 14
    eng, orc, cols = _pair(n, depth=8)
    score, health, tag = cols
    for w in (eng, orc):
        for r in range(0, n, 41):
            w.remove_component(score, r)
    sess = SyncTestSession(2, d, 8, input_delay=2)
    vectors = []
    for t in range(n_ticks):
        sess.add_local_input(0, 0); sess.add_local_input(1, 0)
        reqs = sess.advance_frame()
        for r in reqs:
            if r.kind == SAVE:
                sess.save_cell(r.frame, 0)
        vectors.append(reqs)
    got, want, inflight = [], [], 0
    for v in vectors:
        eng.submit_requests(sess.info(), v)
        inflight += 1
        if inflight == 4:
            got += eng.collect()
            inflight -= 1
        want += orc.handle_requests(sess.info(), v)
    while inflight:
        got += eng.collect()
        inflight -= 1
    assert got == want and len(got) >= n_ticks
    assert eng.last_path_fused()
    _same(eng, orc, cols, n)
    f = eng.snapshot_frames()[-1]
    pe, po = eng.peek(f, tag, 0, n), orc.peek(f, tag, 0, n)
    m = po[1].astype(bool)
    assert np.array_equal(pe[1].astype(bool), m) and np.array_equal(pe[0][m], po[0][m])
    # synchronous calls after the pipelined ones: the chain state must not leak
    nxt = eng.rollback_frame_count() + 1
    a, b = [w.handle_requests(NOSESS, [Request(ADVANCE, 0, [0, 0]), Request(SAVE, nxt)]) for w in (eng, orc)]
    assert a == b
    _same(eng, orc, cols, n)
    eng.close(); orc.close()


def test_build_compiles_the_registrations_own_kernel(generic_kernel):
    """BGR_TUNE_JIT=2: bgr_build hands the registration to NVRTC (csrc/jit.hpp) and every request vector then runs on
    that kernel; =0: the interpreter kernel.  A registration the specialised kernel does not cover (a checksum over a byte
    range that is not whole words) keeps the interpreter without failing."""
    eng, orc, cols = _pair(600)
    assert eng.generic_specialised() == (generic_kernel != "interpreter")
    a, b = [w.handle_requests(NOSESS, [Request(SAVE, 0), Request(ADVANCE, 0, [0]), Request(SAVE, 1)]) for w in (eng, orc)]
    assert a == b and eng.last_path_fused()
    eng.close(); orc.close()
    odd = Engine(max_entities=64, max_depth=4)
    c = odd.rollback_component("Odd", 7, capi.BGR_STRATEGY_COPY)
    odd.checksum_component(c, 1, 5)
    odd.build()
    assert not odd.generic_specialised()
    odd.close()


def test_presence_api_errors():
    eng, orc, (score, health, tag) = _pair(10)
    with pytest.raises(BgrError):
        eng.remove_component(tag, 0)          # not registered optional
    with pytest.raises(BgrError):
        eng.remove_component(score, 10_000)   # row out of range
    e2 = Engine(max_entities=4, max_depth=2)
    for i in range(capi.BGR_MAX_OPTIONAL_COLUMNS):
        e2.rollback_component(f"C{i}", 4, OPT)
    with pytest.raises(BgrError) as ei:
        e2.rollback_component("C8", 4, OPT)
    assert ei.value.status == capi.BGR_ERR_CAPACITY


def _particles_pair(n, depth=8, flags=0):
    """The stress-test bundle with Velocity and Ttl registered optional (Transform always present)."""
    from bevy_ggrs_b200.stress import synth_particles
    worlds, cols = [], None
    for w in (Engine(max_entities=n, max_depth=depth, flags=flags), OracleWorld()):
        t = w.rollback_component("Transform", 40, capi.BGR_STRATEGY_CLONE)
        v = w.rollback_component("Velocity", 12, capi.BGR_STRATEGY_COPY | OPT)
        l = w.rollback_component("Ttl", 8, capi.BGR_STRATEGY_COPY | OPT)
        w.checksum_component(v, 0, 12, capi.BGR_HASH_FLAG_ASSERT_FINITE_F32)
        w.checksum_component(t, 0, 12, capi.BGR_HASH_FLAG_ASSERT_FINITE_F32)
        w.add_system(capi.BGR_SYS_PARTICLES_UPDATE, [t, v])
        w.add_system(capi.BGR_SYS_PARTICLES_DESPAWN, [l])
        w.build()
        w.spawn(n)
        tf, vel, ttl = synth_particles(n, 17, 4, 30, z_fraction=0.2)
        w.write_component(t, 0, tf); w.write_component(v, 0, vel); w.write_component(l, 0, ttl)
        worlds.append(w)
        cols = (t, v, l)
    return worlds[0], worlds[1], cols


@pytest.mark.parametrize("n", [700, 5000])
def test_particles_bundle_with_optional_columns_stays_on_the_one_launch_path(n):
    """Optional Velocity / Ttl on the particles bundle: update_particles only moves entities that have both Transform
    and Velocity, despawn_particles only ages entities that have a Ttl, the Velocity checksum only covers entities
    that have one, Load re-inserts / removes per entity — all inside the ONE fused launch per request vector, bit for
    bit like the oracle (and like the generic kernels)."""
    d = 4
    SESS = (capi.BGR_SESSION_SYNCTEST, 8, d, 0)
    eng, orc, cols = _particles_pair(n)
    stp = _particles_pair(n, flags=capi.BGR_CFG_FORCE_STEPWISE)[0]
    t, v, l = cols
    rng = np.random.default_rng(23)
    frame = 0
    l0 = eng.launch_count()
    ticks = 14
    for tick in range(ticks):
        reqs = []
        if tick >= d:
            reqs.append(Request(LOAD, frame - d))
            for k in range(d):
                reqs += [Request(ADVANCE, 0, [0]), Request(SAVE, frame - d + k + 1)] if k < d - 1 else [Request(ADVANCE, 0, [0])]
        reqs += [Request(SAVE, frame), Request(ADVANCE, 0, [0])]
        before = eng.launch_count()
        a, b, c = [w.handle_requests(SESS, reqs) for w in (eng, orc, stp)]
        assert eng.launch_count() - before == 1                 # ONE launch per request vector
        assert a == b == c, f"tick {tick}"
        assert eng.last_path_fused() and not stp.last_path_fused()
        frame += 1
        alive = orc.read_alive(0, n).astype(bool)
        for r in rng.choice(np.flatnonzero(alive), 6, replace=False):
            col = (v, l)[int(rng.integers(2))]
            if orc.has_component(col, int(r), 1)[0]:
                for w in (eng, orc, stp):
                    w.remove_component(col, int(r))
            else:
                val = np.array([1.5, -2.5, 0.0], np.float32) if col == v else np.uint64(rng.integers(3, 20))
                for w in (eng, orc, stp):
                    w.insert_component(col, int(r), val)
        _same(eng, orc, cols, n)
        _same(stp, orc, cols, n)
    assert 0 < orc.read_alive(0, n).sum() < n                   # Ttl ran out for some entities inside the run
    vals, had = eng.peek(frame - 2, v, 0, n)
    vo, ho = orc.peek(frame - 2, v, 0, n)
    assert np.array_equal(had.astype(bool), ho.astype(bool)) and 0 < ho.sum() < n
    assert np.array_equal(vals[ho.astype(bool)], vo[ho.astype(bool)])
    for w in (eng, orc, stp):
        w.close()


@pytest.mark.parametrize("flags", [0, capi.BGR_CFG_FORCE_STEPWISE])
def test_load_shrinks_a_world_that_grew_between_ticks_on_the_generic_program(flags):
    """Entities spawned by the host between ticks (`commands.spawn((…, Rollback))` outside GgrsSchedule) after a
    snapshot was taken: loading that snapshot takes them out again (EntitySnapshotPlugin::load despawns entities the
    snapshot does not know, entity.rs:55-99) — across a tile boundary, on the one-launch generic program and stepwise."""
    n0, extra = 700, 900          # 700 rows = 2 tiles, 1600 rows = 4 tiles
    worlds = []
    for w in (Engine(max_entities=n0 + extra, max_depth=8, flags=flags), OracleWorld()):
        s_ = w.rollback_component("Score", 4, capi.BGR_STRATEGY_COPY | OPT)
        h_ = w.rollback_component("Health", 4, capi.BGR_STRATEGY_CLONE | OPT)
        t_ = w.rollback_component("Tag", 12, capi.BGR_STRATEGY_COPY)
        for c, ln in ((s_, 4), (t_, 12), (h_, 4)):
            w.checksum_component(c, 0, ln)
        w.add_system(capi.BGR_SYS_U32_ADD, [s_], [0, 1])
        w.add_system(capi.BGR_SYS_U32_SATSUB_DESPAWN, [h_], [0, 1])
        w.build()
        w.spawn(n0)
        rng = np.random.default_rng(3)
        w.write_component(s_, 0, rng.integers(0, 1000, n0, dtype=np.uint32))
        w.write_component(h_, 0, rng.integers(3, 40, n0, dtype=np.uint32))
        w.write_component(t_, 0, rng.integers(0, 2**32, (n0, 3), dtype=np.uint32))
        worlds.append(w)
        cols = (s_, h_, t_)
    eng, orc = worlds
    score, health, tag = cols
    both = lambda f, *a: [getattr(w, f)(*a) for w in (eng, orc)]
    a, b = both("handle_requests", NOSESS, [Request(SAVE, 0), Request(ADVANCE, 0, [0])])
    assert a == b
    rng = np.random.default_rng(4)
    new_score, new_health = rng.integers(0, 99, extra, dtype=np.uint32), rng.integers(5, 50, extra, dtype=np.uint32)
    new_tag = rng.integers(0, 2**32, (extra, 3), dtype=np.uint32)
    for w in (eng, orc):
        first = w.spawn(extra)
        assert first == n0
        w.write_component(score, first, new_score); w.write_component(health, first, new_health); w.write_component(tag, first, new_tag)
    a, b = both("handle_requests", NOSESS, [Request(SAVE, 1), Request(ADVANCE, 0, [0]), Request(SAVE, 2), Request(ADVANCE, 0, [0])])
    assert a == b and eng.row_count() == orc.row_count() == n0 + extra
    _same(eng, orc, cols, n0 + extra)
    l0 = eng.launch_count()
    a, b = both("handle_requests", NOSESS, [Request(LOAD, 0), Request(ADVANCE, 0, [0]), Request(SAVE, 1), Request(ADVANCE, 0, [0])])
    assert a == b
    assert (eng.launch_count() - l0 == 1) == (flags == 0)
    assert eng.row_count() == orc.row_count() == n0           # RollbackOrdered was rolled back with the snapshot
    assert eng.active_count() == orc.active_count() <= n0
    _same(eng, orc, cols, n0)
    a, b = both("handle_requests", NOSESS, [Request(SAVE, 2), Request(ADVANCE, 0, [0])])   # and the world keeps working
    assert a == b
    eng.close(); orc.close()
