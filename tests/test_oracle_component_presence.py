"""ComponentSnapshotPlugin::load's four-way match (src/snapshot/component_snapshot.rs:99-115) on the oracle:
(Some, Some) update, (Some, None) remove, (None, Some) insert, (None, None) nothing — and the matching query
semantics of save (:72-76) and of the checksum (component_checksum.rs:77-90): only entities that have the component."""
import numpy as np

from bevy_ggrs_b200 import capi
from bevy_ggrs_b200.session import ADVANCE, LOAD, SAVE, Request
from oracle_backend import OracleWorld

NOSESS = (capi.BGR_SESSION_NONE, 0, 0, 0)


def _world(n=6):
    w = OracleWorld()
    score = w.rollback_component("Score", 4, capi.BGR_STRATEGY_COPY | capi.BGR_STRATEGY_OPTIONAL)
    w.checksum_component(score, 0, 4)
    w.add_system(capi.BGR_SYS_U32_ADD, [score], [0, 1])
    w.build()
    w.spawn(n)
    w.write_component(score, 0, np.arange(10, 10 + n, dtype=np.uint32))
    return w, score


def test_load_updates_removes_inserts_or_leaves_alone():
    w, score = _world()
    w.remove_component(score, 1)                      # absent when the snapshot is taken
    w.remove_component(score, 4)
    (f0, ck0), = w.handle_requests(NOSESS, [Request(SAVE, 0)])
    w.handle_requests(NOSESS, [Request(ADVANCE, 0, [0])])       # rows that have Score: +1
    w.remove_component(score, 2)                      # (None, Some) at load time -> insert
    w.insert_component(score, 1, np.uint32(99))       # (Some, None) at load time -> remove
    vals, has = w.read_component_alive(score, 0, 6)
    assert has.tolist() == [1, 1, 0, 1, 0, 1]
    assert vals.view(np.uint32).ravel()[[0, 1, 3, 5]].tolist() == [11, 99, 14, 16]
    w.handle_requests(NOSESS, [Request(LOAD, 0)])
    vals, has = w.read_component_alive(score, 0, 6)
    assert has.tolist() == [1, 0, 1, 1, 0, 1]         # 0: update, 1: remove, 2: insert, 3: update, 4: nothing
    assert vals.view(np.uint32).ravel()[[0, 2, 3, 5]].tolist() == [10, 12, 13, 15]
    (f1, ck1), = w.handle_requests(NOSESS, [Request(SAVE, 0)])
    assert (f0, ck0) == (f1, ck1)                     # same frame, same world -> same checksum


def test_checksum_only_covers_entities_that_have_the_component():
    a, sa = _world()
    b, sb = _world()
    b.write_component(sb, 3, np.array([777], dtype=np.uint32))  # differs only in a row that is about to lose Score
    a.remove_component(sa, 3)
    b.remove_component(sb, 3)
    assert a.handle_requests(NOSESS, [Request(SAVE, 0)]) == b.handle_requests(NOSESS, [Request(SAVE, 0)])
    c, sc = _world()
    assert c.handle_requests(NOSESS, [Request(SAVE, 0)]) != a.handle_requests(NOSESS, [Request(SAVE, 0)])


def test_app_synctest_rollback_reinserts_and_removes_optional_components():
    """Through the plugin mirror (App) on the oracle backend — the same scenario tests/cpp/test_host_mirror.cpp runs on
    the engine: a component removed by code outside GgrsSchedule comes back with the next rollback (resimulated from the
    snapshot, so Score == RollbackFrameCount), one inserted there disappears again, and SyncTest never sees a mismatch."""
    from bevy_ggrs_b200.plugin import (App, GgrsPlugin, GgrsSchedule, LocalInputs, ReadInputs, Session, Startup,
                                       SyncTestMismatch, System)
    from bevy_ggrs_b200.session import SyncTestSession
    w = OracleWorld()
    app = App(w)
    app.add_plugins(GgrsPlugin())
    app.add_systems(ReadInputs, lambda a: a.insert_resource(LocalInputs({h: 0 for h in a.local_players.handles})))
    score = app.rollback_optional_component_with_copy("Score", 4)
    health = app.rollback_optional_component_with_copy("Health", 4)
    app.checksum_component_with_hash(score)
    app.add_systems(GgrsSchedule, System(capi.BGR_SYS_U32_ADD, [score], [0, 1]))

    def startup(a):
        a.world.spawn(2)
        a.world.remove_component(health, 1)
    app.add_systems(Startup, startup)
    app.insert_resource(Session.SyncTest(SyncTestSession(1, 3, 8, input_delay=0)))
    mism = []
    app.add_observer(SyncTestMismatch, lambda ev: mism.append(ev))
    for _ in range(10):
        app.update()
    assert w.has_component(score, 0, 2).tolist() == [1, 1] and w.has_component(health, 0, 2).tolist() == [1, 0]
    w.remove_component(score, 0)
    w.insert_component(health, 1, np.uint32(77))
    assert w.has_component(score, 0, 2).tolist() == [0, 1] and w.has_component(health, 0, 2).tolist() == [1, 1]
    app.update()
    assert w.has_component(score, 0, 2).tolist() == [1, 1] and w.has_component(health, 0, 2).tolist() == [1, 0]
    vals = w.read_component(score, 0, 2).view(np.uint32).ravel()
    assert vals[0] == vals[1] == app.rollback_frame_count()
    assert not mism
