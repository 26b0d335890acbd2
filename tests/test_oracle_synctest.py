"""The reference's SyncTest integration tests (tests/synctest.rs, tests/component_rollback.rs) restated
on the oracle through the plugin mirror: the oracle must pass the reference's own self-consistency
checks before it is trusted as the parity checker."""
import numpy as np
import pytest

from bevy_ggrs_b200 import capi
from bevy_ggrs_b200.plugin import (App, GgrsPlugin, GgrsSchedule, LocalInputs, ReadInputs, Session, Startup,
                                   SyncTestMismatch, System)
from bevy_ggrs_b200.session import SyncTestSession
from oracle_backend import OracleWorld


def input_system(app):
    app.insert_resource(LocalInputs({h: 0 for h in app.local_players.handles}))


def base_synctest_app(check_distance, backend=None):
    """tests/common/mod.rs:44-54"""
    app = App(backend or OracleWorld())
    app.insert_resource(Session.SyncTest(SyncTestSession(1, check_distance)))
    app.add_plugins(GgrsPlugin())
    app.add_systems(ReadInputs, input_system)
    return app


def test_copy_strategy_rolls_back_component_data():
    """component_rollback.rs:33-64 (Score += 1 per frame; score == RollbackFrameCount; no mismatch)."""
    app = base_synctest_app(2)
    score = app.rollback_component_with_copy("Score", 4)
    app.checksum_component_with_hash(score)
    app.add_systems(GgrsSchedule, System(capi.BGR_SYS_U32_ADD, [score], [0, 1]))
    app.add_systems(Startup, lambda a: a.world.write_component(score, a.world.spawn(1), np.zeros(1, np.uint32)))

    def boom(ev):
        raise AssertionError(f"SyncTestMismatch {ev}")
    app.add_observer(SyncTestMismatch, boom)
    for _ in range(20):
        app.update()
    frame = app.rollback_frame_count()
    val = app.world.read_component(score, 0, 1).view(np.uint32)[0, 0]
    assert frame == 19  # the first bevy update has zero delta
    assert val == frame


def test_despawn_and_rollback_does_not_panic():
    """synctest.rs:59-75: Health 10, -1 per frame, despawn at 0, check_distance 5, 60 updates."""
    app = base_synctest_app(5)
    health = app.rollback_component_with_copy("Health", 4)
    app.add_systems(GgrsSchedule, System(capi.BGR_SYS_U32_SATSUB_DESPAWN, [health], [0, 1]))
    app.add_systems(Startup, lambda a: a.world.write_component(health, a.world.spawn(1), np.full(1, 10, np.uint32)))
    for _ in range(60):
        app.update()
    assert app.world.active_count() == 0


def test_synctest_mismatch_fires_on_non_determinism():
    """synctest.rs:83-125: a counter that is not rolled back is written into a checksummed component."""
    app = base_synctest_app(2)
    counter = app.rollback_component_with_copy("Counter", 4)
    app.checksum_component_with_hash(counter)
    app.add_systems(GgrsSchedule, System(capi.BGR_SYS_U32_STORE_CALL_COUNT, [counter], [0]))
    app.add_systems(Startup, lambda a: a.world.spawn(1))
    detected = []
    app.add_observer(SyncTestMismatch, lambda ev: detected.append(ev))
    for _ in range(10):
        app.update()
    assert detected, "SyncTestMismatch should have fired due to non-deterministic game logic"


def test_synctest_prunes_confirmed_snapshots():
    """synctest.rs:129-153: ConfirmedFrameCount advances and frame 0 is pruned."""
    app = base_synctest_app(5)
    c = app.rollback_component_with_copy("FrameCounter", 4)
    app.add_systems(GgrsSchedule, System(capi.BGR_SYS_U32_ADD, [c], [0, 1]))
    app.add_systems(Startup, lambda a: a.world.spawn(1))
    for _ in range(20):
        app.update()
    assert app.confirmed_frame_count() > 0
    assert app.world.peek(0, c, 0, 1) is None
    assert 0 not in app.world.snapshot_frames()


def test_rollback_to_missing_frame_reports_reference_panic_text():
    w = OracleWorld()
    c = w.rollback_component("X", 4)
    w.spawn(1)
    w.save_world()
    w.set_rollback_frame_count(99)
    with pytest.raises(Exception, match="Could not rollback to 99"):
        w.load_world()


def test_box_game_synctest_c1_plumbing():
    """BASELINE config C1: box_game SyncTest, 2 players, check_distance 8 (needs max_prediction 9),
    input_delay 2 (box_game_synctest.rs:25-28).  Oracle-only (move_cube_system uses powf)."""
    w = OracleWorld()
    app = App(w)
    app.insert_resource(Session.SyncTest(SyncTestSession(2, 8, 9, input_delay=2)))
    app.add_plugins(GgrsPlugin())
    seq = [0b0001, 0b1000, 0b0101, 0, 0b0010, 0b1010]
    app.add_systems(ReadInputs, lambda a: a.insert_resource(
        LocalInputs({h: seq[(a.ticks + 2 * h) % len(seq)] for h in a.local_players.handles})))
    vel = app.rollback_component_with_copy("Velocity", 12)
    tf = app.rollback_component_with_clone("Transform", 40)
    fc = w.rollback_resource("FrameCount", b"\0\0\0\0", checksum=True)
    app.add_systems(GgrsSchedule, System(capi.BGR_SYS_BOX_MOVE, [tf, vel]))
    app.add_systems(GgrsSchedule, System(100, [], [fc]))  # increase_frame_system (oracle-only resource system)

    def setup(a):
        first = a.world.spawn(2)
        t = np.zeros((2, 10), np.float32)
        r = 5.0 / 4.0
        for h in range(2):
            rot = np.float32(h) / np.float32(2) * np.float32(2.0) * np.float32(np.pi)
            t[h, 0] = r * np.cos(rot); t[h, 1] = 0.1; t[h, 2] = r * np.sin(rot)
            t[h, 6] = 1.0; t[h, 7:10] = 1.0
        a.world.write_component(tf, first, t)
    app.add_systems(Startup, setup)
    bad = []
    app.add_observer(SyncTestMismatch, lambda ev: bad.append(ev))
    for _ in range(60):
        app.update()
    assert not bad
    assert app.rollback_frame_count() == 59
    import struct
    assert struct.unpack("<I", w.read_resource(fc, 4))[0] == 59
    t = w.read_component(tf, 0, 2).view(np.float32)
    assert np.all(np.abs(t[:, [0, 2]]) <= 2.4 + 1e-6) and np.all(np.isfinite(t))
    v = w.read_component(vel, 0, 2).view(np.float32)
    assert np.all(np.linalg.norm(v, axis=1) <= 3.0 + 1e-5)
    assert np.any(v != 0)


def test_catch_up_vectors_merged_equal_tick_by_tick_on_the_oracle():
    """handle_requests is a plain loop over the request vector (schedule_systems.rs:222-269), so several ticks' vectors
    concatenated into one call must leave the same world, ring and checksums as one call per tick — the property the
    engine's catch-up batches (one fused launch for the merged vector) are tested against on the GPU."""
    import numpy as np
    from bevy_ggrs_b200.session import SAVE, SyncTestSession
    from bevy_ggrs_b200.stress import populate, register_particles, synth_particles
    from oracle_backend import OracleWorld
    n, d, maxp, n_ticks, group = 400, 4, 8, 18, 3
    worlds = []
    for _ in range(2):
        w = OracleWorld()
        cols = register_particles(w, spawn_rate=15, spawn_ttl=6)
        w.build()
        populate(w, cols, *synth_particles(n, 5, 3, 30))
        worlds.append((w, cols))
    sess = SyncTestSession(2, d, maxp, input_delay=2)
    vectors = []
    for t in range(n_ticks):
        sess.add_local_input(0, (1 << 4) if t % 4 == 1 else 0)
        sess.add_local_input(1, 0)
        reqs = sess.advance_frame()
        for r in reqs:
            if r.kind == SAVE:
                sess.save_cell(r.frame, 0)
        vectors.append(reqs)
    (a, cols), (b, _) = worlds
    got_a, got_b = [], []
    for g in range(0, n_ticks, group):
        got_a += a.handle_requests(sess.info(), [r for v in vectors[g:g + group] for r in v])
        for v in vectors[g:g + group]:
            got_b += b.handle_requests(sess.info(), v)
    assert got_a == got_b and len(got_a) > n_ticks
    assert a.row_count() == b.row_count() > n and a.snapshot_frames() == b.snapshot_frames()
    rows = a.row_count()
    for c in cols:
        va, ha = a.read_component_alive(c, 0, rows)
        vb, hb = b.read_component_alive(c, 0, rows)
        assert np.array_equal(ha, hb) and np.array_equal(va[ha.astype(bool)], vb[hb.astype(bool)])
