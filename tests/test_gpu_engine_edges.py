"""Edge cases of the C ABI on a real GPU: empty / ragged worlds, capacity limits, the reference's panics as
status codes with the same text, atomic validation of a request vector, registration errors."""
import numpy as np
import pytest

from bevy_ggrs_b200 import capi
from bevy_ggrs_b200.capi import BgrError
from bevy_ggrs_b200.engine import Engine
from bevy_ggrs_b200.session import ADVANCE, LOAD, SAVE, Request, SyncTestSession
from bevy_ggrs_b200.stress import populate, register_particles, synth_particles
from oracle_backend import OracleWorld

pytestmark = pytest.mark.gpu
NOSESS = (capi.BGR_SESSION_NONE, 0, 0, 0)


def _pair(n, max_depth=8, flags=0, seed=1, ttl=(5, 50)):
    eng, orc = Engine(max_entities=max(n, 1), max_depth=max_depth, flags=flags), OracleWorld()
    out = []
    for w in (eng, orc):
        cols = register_particles(w)
        w.build()
        if n:
            populate(w, cols, *synth_particles(n, seed, *ttl))
        out.append(cols)
    return eng, orc, out[0]


@pytest.mark.parametrize("flags", [0, capi.BGR_CFG_FORCE_STEPWISE])
def test_empty_world_checksum_is_entity_part_and_zero_entity_component_parts(flags):
    eng, orc, _ = _pair(0, flags=flags)
    reqs = [Request(SAVE, 0), Request(ADVANCE, 0, [0]), Request(SAVE, 1), Request(LOAD, 0), Request(ADVANCE, 0, [0])]
    a, b = eng.handle_requests(NOSESS, reqs), orc.handle_requests(NOSESS, reqs)
    assert a == b and len(a) == 2
    # entity part (0, 0) ^ two component parts with zero entities (seahash(0u64))
    import ctypes as C
    lib = capi.load_library()
    z16, z8 = C.create_string_buffer(bytes(16), 16), C.create_string_buffer(bytes(8), 8)
    assert a[0][1] == lib.bgr_seahash(z16, 16)  # the two identical component parts cancel under XOR
    assert lib.bgr_seahash(z8, 8) == 0x1CDEE552A46D795F


@pytest.mark.parametrize("n", [1, 31, 33, 511, 512, 513, 1025])
@pytest.mark.parametrize("flags", [0, capi.BGR_CFG_FORCE_STEPWISE])
def test_ragged_row_counts_around_warp_and_tile_boundaries(n, flags):
    eng, orc, cols = _pair(n, flags=flags, seed=n)
    sess_e, sess_o = SyncTestSession(1, 3, 8), SyncTestSession(1, 3, 8)
    for _ in range(10):
        for s, w, keep in ((sess_e, eng, []), (sess_o, orc, [])):
            s.add_local_input(0, 0)
            cs = w.handle_requests(s.info(), s.advance_frame())
            for f, c in cs:
                s.save_cell(f, c)
            keep.append(cs)
        assert eng.snapshot_frames() == orc.snapshot_frames()
    alive = orc.read_alive(0, n).astype(bool)
    assert np.array_equal(eng.read_alive(0, n).astype(bool), alive)
    for c in cols:
        assert np.array_equal(eng.read_component(c, 0, n)[alive], orc.read_component(c, 0, n)[alive])
    assert eng.save_world() == orc.save_world()


def test_rollback_to_missing_frame_is_the_reference_panic_and_executes_nothing():
    eng, orc, cols = _pair(100)
    eng.handle_requests(NOSESS, [Request(SAVE, 0), Request(ADVANCE, 0, [0]), Request(SAVE, 1)])
    before = eng.read_component(cols[0], 0, 100).copy()
    frames, fc = eng.snapshot_frames(), eng.rollback_frame_count()
    with pytest.raises(BgrError, match="Could not rollback to 99: no snapshot at that moment could be found.") as ei:
        eng.handle_requests(NOSESS, [Request(ADVANCE, 0, [0]), Request(SAVE, 2), Request(LOAD, 99), Request(ADVANCE, 0, [0])])
    assert ei.value.status == capi.BGR_ERR_NO_SNAPSHOT
    # the vector was validated first: nothing ran, nothing was committed
    assert eng.snapshot_frames() == frames and eng.rollback_frame_count() == fc
    assert np.array_equal(eng.read_component(cols[0], 0, 100), before)


@pytest.mark.parametrize("flags", [0, capi.BGR_CFG_FORCE_STEPWISE])
def test_non_finite_translation_raises_the_hasher_assertion(flags):
    eng, _, cols = _pair(64, flags=flags)
    tf = eng.read_component(cols[0], 0, 64).view(np.float32).copy()
    tf[17, 1] = np.inf
    eng.write_component(cols[0], 0, tf)
    with pytest.raises(BgrError, match="Hashing is not stable for NaN f32 values.") as ei:
        eng.save_world()
    assert ei.value.status == capi.BGR_ERR_NON_FINITE
    # a dead entity is not hashed: despawn it and the save goes through
    eng.despawn(17)
    eng.save_world()


def test_ring_capacity_and_depth_semantics():
    eng, orc, _ = _pair(10, max_depth=3)
    sess = (capi.BGR_SESSION_P2P, 3, 0, -1)
    reqs = []
    for f in range(6):
        reqs += [Request(SAVE, f), Request(ADVANCE, 0, [0, 0])]
    assert eng.handle_requests(sess, reqs) == orc.handle_requests(sess, reqs)
    assert eng.snapshot_frames() == orc.snapshot_frames() == [5, 4, 3]           # depth 3: oldest evicted
    assert eng.peek(2, 0, 0, 10) is None and eng.peek(3, 0, 0, 10) is not None
    # MaxPredictionWindow larger than the slots allocated at build time: loud capacity error, not silent eviction
    with pytest.raises(BgrError) as ei:
        eng.handle_requests((capi.BGR_SESSION_P2P, 8, 0, -1), [Request(SAVE, 6), Request(ADVANCE, 0, [0, 0])] * 1 + [Request(SAVE, 7)])
    assert ei.value.status == capi.BGR_ERR_CAPACITY


def test_i32_wraparound_frames_through_the_engine_ring():
    eng, _, _ = _pair(8)
    I32_MAX, I32_MIN = 2**31 - 1, -(2**31)
    for f in (I32_MAX - 1, I32_MAX, I32_MIN):     # mod.rs:480-493: MIN after MAX is a forward step
        eng.set_rollback_frame_count(f)
        eng.save_world()
    assert eng.snapshot_frames() == [I32_MIN, I32_MAX, I32_MAX - 1]
    eng.set_rollback_frame_count(I32_MAX)          # mod.rs:497-508: pushing MAX again evicts MIN as a future frame
    eng.save_world()
    assert eng.snapshot_frames() == [I32_MAX, I32_MAX - 1]


def test_registration_and_argument_errors():
    eng = Engine(max_entities=16, max_depth=4)
    c = eng.rollback_component("Health", 4)
    with pytest.raises(BgrError) as ei:
        eng.checksum_component(c, 2, 8)            # range exceeds the element
    assert ei.value.status == capi.BGR_ERR_INVALID_ARGUMENT
    with pytest.raises(BgrError):
        eng.add_system(capi.BGR_SYS_PARTICLES_UPDATE, [c, c])      # wrong component sizes
    with pytest.raises(BgrError):
        eng.add_system(999, [c])
    with pytest.raises(BgrError) as ei:
        eng.handle_requests(NOSESS, [Request(SAVE, 0)])             # before bgr_build
    assert ei.value.status == capi.BGR_ERR_STATE
    eng.build()
    with pytest.raises(BgrError) as ei:
        eng.rollback_component("Late", 4)                           # registration after build
    assert ei.value.status == capi.BGR_ERR_STATE
    with pytest.raises(BgrError) as ei:
        eng.spawn(17)
    assert ei.value.status == capi.BGR_ERR_CAPACITY
    with pytest.raises(BgrError) as ei:
        eng.handle_requests(NOSESS, [Request(ADVANCE, 0, [0])] * (capi.BGR_MAX_REQUESTS + 1))
    assert ei.value.status == capi.BGR_ERR_CAPACITY
    first = eng.spawn(16)
    assert first == 0 and eng.row_count() == 16 and eng.active_count() == 16
    eng.despawn(3)
    assert eng.active_count() == 15


def test_p2p_trace_matches_oracle_request_for_request():
    from bevy_ggrs_b200.session import P2PTraceSession
    eng, orc, cols = _pair(5000, seed=77, ttl=(3, 25))
    se, so = P2PTraceSession(2, 8, 2, seed=0xB200), P2PTraceSession(2, 8, 2, seed=0xB200)
    for t in range(60):
        for s in (se, so):
            s.add_local_input(0, (1 << 5) if t % 4 == 0 else 0)
        assert eng.handle_requests(se.info(), se.advance_frame()) == orc.handle_requests(so.info(), so.advance_frame())
        assert eng.snapshot_frames() == orc.snapshot_frames()
        assert eng.confirmed_frame_count() == orc.confirmed_frame_count()
    alive = orc.read_alive(0, 5000).astype(bool)
    assert np.array_equal(eng.read_alive(0, 5000).astype(bool), alive)
    for c in cols:
        assert np.array_equal(eng.read_component(c, 0, 5000)[alive], orc.read_component(c, 0, 5000)[alive])


@pytest.mark.parametrize("flags", [0, capi.BGR_CFG_FORCE_STEPWISE])
def test_async_download_sees_the_world_at_begin_and_overlaps_later_ticks(flags):
    """bgr_download_begin is ordered after the submits before it and is not disturbed by the submits after it: the
    mirrored Transform.translation / Velocity / Ttl bytes equal the oracle's world at that tick, bit for bit."""
    n = 3000  # ragged: 5 full tiles + a partial one
    eng, orc, cols = _pair(n, flags=flags, seed=9, ttl=(3, 40))
    t, v, l = cols
    tick = lambda f: [Request(SAVE, f), Request(ADVANCE, f, [0, 0])]
    host = {name: eng.host_alloc(n, nb) for name, nb in (("tr", 12), ("vel", 12), ("ttl", 8), ("rot", 16))}
    for f in range(3):
        eng.submit_requests(NOSESS, tick(f))
        orc.handle_requests(NOSESS, tick(f))
    tickets = [eng.download_begin(t, 0, 12, 0, n, host["tr"]), eng.download_begin(v, 0, 12, 0, n, host["vel"]),
               eng.download_begin(l, 0, 8, 0, n, host["ttl"]), eng.download_begin(t, 12, 16, 0, n, host["rot"])]
    with pytest.raises(BgrError) as ei:  # BGR_MAX_DOWNLOADS in flight
        eng.download_begin(t, 0, 12, 0, n, host["tr"])
    assert ei.value.status == capi.BGR_ERR_STATE
    tr, alive = orc.read_component_alive(t, 0, n)  # dead rows' stale bytes are not observable: compare the live ones
    alive = alive.astype(bool)
    assert 0 < alive.sum() < n
    want = {"tr": tr[:, :12].copy(), "vel": orc.read_component(v, 0, n).copy(),
            "ttl": orc.read_component(l, 0, n).copy(), "rot": tr[:, 12:28].copy()}
    for f in range(3, 6):  # later ticks run while the copies are in flight and must not leak into them
        eng.submit_requests(NOSESS, tick(f))
        orc.handle_requests(NOSESS, tick(f))
    for k in tickets:
        eng.download_wait(k)
    for name in want:
        assert np.array_equal(host[name][alive], want[name][alive]), name
    for _ in range(6):
        eng.collect()
    # a sub-range, after the in-flight slots were released
    k = eng.download_begin(t, 4, 8, 513, 1000, host["ttl"])
    eng.download_wait(k)
    sub, sub_alive = orc.read_component_alive(t, 513, 1000)
    assert np.array_equal(host["ttl"][:1000][sub_alive.astype(bool)], sub[:, 4:12][sub_alive.astype(bool)])
    with pytest.raises(BgrError):
        eng.download_wait(k)  # already waited for
    with pytest.raises(BgrError):
        eng.download_begin(t, 2, 8, 0, 10, host["ttl"])  # unaligned field range
    with pytest.raises(BgrError):
        eng.download_begin(t, 0, 12, n - 5, 10, host["tr"])  # beyond the spawned rows
    eng.close()


def test_reads_between_submits_keep_their_results_queued_for_collect():
    """An entry point that touches the world waits for the submitted request vectors but must not swallow their
    results: a later bgr_collect still returns every checksum, in order (desync detection has no gaps)."""
    n = 2000
    eng, orc, cols = _pair(n, seed=3, ttl=(4, 30))
    tick = lambda f: [Request(SAVE, f), Request(ADVANCE, f, [0, 0])]
    want = []
    for f in range(3):
        eng.submit_requests(NOSESS, tick(f))
        want += orc.handle_requests(NOSESS, tick(f))
    with pytest.raises(BgrError) as ei:          # the synchronous call would return the wrong vector's checksums
        eng.handle_requests(NOSESS, tick(3))
    assert ei.value.status == capi.BGR_ERR_STATE and "bgr_collect" in str(ei.value)
    alive = orc.read_alive(0, n).astype(bool)    # a read in the middle: waits for the GPU, results stay queued
    assert np.array_equal(eng.read_alive(0, n).astype(bool), alive)
    assert np.array_equal(eng.read_component(cols[0], 0, n)[alive], orc.read_component(cols[0], 0, n)[alive])
    got = []
    for _ in range(3):
        got += eng.collect()
    assert got == want and len(got) == 3
    with pytest.raises(BgrError):
        eng.collect()                            # nothing left
    assert eng.handle_requests(NOSESS, tick(3)) == orc.handle_requests(NOSESS, tick(3))
    eng.close()


def test_non_finite_status_survives_a_drain():
    eng, orc, cols = _pair(600, seed=4)
    bad = eng.read_component(cols[1], 0, 1).copy()
    bad.view(np.float32)[0, 1] = np.inf
    eng.write_component(cols[1], 7, bad)
    eng.submit_requests(NOSESS, [Request(SAVE, 0)])
    eng.read_alive(0, 10)                        # drains
    with pytest.raises(BgrError) as ei:
        eng.collect()
    assert ei.value.status == capi.BGR_ERR_NON_FINITE
    eng.close()


def test_sharded_engines_refuse_dynamic_spawning():
    """A shard appends rows locally: a newborn's RollbackOrdered index would collide with the next shard's range and
    every shard would draw the same ParticleRng stream — refused instead of silently diverging from one GPU."""
    for flags, base in ((capi.BGR_CFG_SHARDED, 0), (0, 4096)):
        eng = Engine(max_entities=4096, max_depth=4, flags=flags, order_base=base)
        with pytest.raises(BgrError) as ei:
            register_particles(eng, spawn_rate=8)
        assert ei.value.status == capi.BGR_ERR_UNSUPPORTED
        eng.close()
    eng = Engine(max_entities=4096, max_depth=4, flags=capi.BGR_CFG_SHARDED, order_base=0)
    cols = register_particles(eng)
    eng.build()
    populate(eng, cols, *synth_particles(1000, 1, 5, 50))       # the initial population is fine
    eng.handle_requests(NOSESS, [Request(SAVE, 0), Request(ADVANCE, 0, [0])])
    with pytest.raises(BgrError) as ei:
        eng.spawn(10)
    assert ei.value.status == capi.BGR_ERR_UNSUPPORTED
    eng.close()


def test_trace_records_one_interval_per_fused_launch():
    eng, orc, cols = _pair(50_000, seed=5, ttl=(500, 500))
    tick = lambda f: [Request(SAVE, f), Request(ADVANCE, f, [0, 0])]
    eng.handle_requests(NOSESS, tick(0))
    eng.trace_enable(16)
    for f in range(1, 6):
        eng.handle_requests(NOSESS, tick(f))
    tr = eng.trace_read(16)
    assert tr.shape == (5, 4) and np.all(tr[:, 2] >= tr[:, 1])
    assert np.all(tr[:, 1] > tr[:, 0]) and np.all(np.diff(tr[:, 0].astype(np.int64)) > 0)
    assert np.all((tr[:, 1] - tr[:, 0]) < 5_000_000)             # a 50k-entity tick is microseconds, not milliseconds
    eng.trace_enable(0)
    eng.close()
