"""Host side table for non-POD rollback components (Sprite: particles.rs:191) on the CPU: the table rides on an oracle
world and must track an optional POD column of a second oracle world holding the same handle ids."""
import pytest

from bevy_ggrs_b200.host_components import HostComponents
from bevy_ggrs_b200.session import LOAD, SAVE, Request
from host_components_util import Sprite, make_world, run_side_table_against_handle_column
from oracle_backend import OracleWorld


def test_side_table_tracks_an_optional_pod_column_through_synctest_rollbacks():
    w = OracleWorld()
    st = run_side_table_against_handle_column(w, n=300, d=4, ticks=16)
    assert st["rolled_back"] == 12 and st["inserted"] > 0 and st["removed"] > 0
    assert 0 < st["alive"] < 300           # Health ran out for some entities inside the run: their sprites went with them
    assert 0 < st["sprites"] <= st["alive"]
    assert st["snapshots"] == st["ring"]   # the table's ring holds exactly the frames the world's ring holds
    w.close()


def test_side_table_errors_follow_the_reference():
    w = OracleWorld()
    make_world(w, 4, 8, False)
    t = HostComponents(w)
    c = t.register("Sprite")
    with pytest.raises(ValueError):
        t.register("Sprite")
    with pytest.raises(KeyError):
        t.insert(c, 99, Sprite(1, None))   # no such entity
    t.insert(c, 1, Sprite(7, None))
    w.handle_requests((0, 0, 0, 0), [Request(SAVE, 0)])
    t.handle_requests([Request(SAVE, 0)])
    with pytest.raises(RuntimeError, match="Could not rollback to 5"):
        t.handle_requests([Request(LOAD, 5)])
    w.close()


def test_plugin_registers_a_non_pod_component_and_rolls_it_back():
    """`app.rollback_component_with_clone("Sprite")` (no size: not plain bytes) through the plugin mirror with a SyncTest
    session: what Startup inserted is in every snapshot and survives every rollback; an edit made outside GgrsSchedule
    is undone by the next tick's Load of an older frame — the reference's behaviour for any rollback component."""
    from bevy_ggrs_b200 import capi
    from bevy_ggrs_b200.plugin import App, GgrsPlugin, LocalInputs, ReadInputs, Session, Startup
    from bevy_ggrs_b200.session import SyncTestSession

    app = App(OracleWorld())
    app.insert_resource(Session.SyncTest(SyncTestSession(1, 2)))
    app.add_plugins(GgrsPlugin())
    app.add_systems(ReadInputs, lambda a: a.insert_resource(LocalInputs({h: 0 for h in a.local_players.handles})))
    score = app.rollback_component_with_copy("Score", 4)
    sprite = app.rollback_component_with_clone("Sprite")
    image = object()

    def startup(a):
        a.world.spawn(3)
        for r in range(3):
            a.host_components.insert(sprite, r, Sprite(10 + r, image))
    app.add_systems(Startup, startup)
    for _ in range(6):
        app.update()
    t = app.host_components
    assert [(r, s.handle_id) for r, s in t.items(sprite)] == [(0, 10), (1, 11), (2, 12)]
    assert all(s.image is image for _, s in t.items(sprite))
    t.remove(sprite, 1)
    t.insert(sprite, 2, Sprite(99, image))
    assert [(r, s.handle_id) for r, s in t.items(sprite)] == [(0, 10), (2, 99)]
    app.update()    # Load(frame - 2) restores the snapshot taken before the edit, the re-simulation does not redo it
    assert [(r, s.handle_id) for r, s in t.items(sprite)] == [(0, 10), (1, 11), (2, 12)]
    assert sorted(sprite.snapshots) == sorted(app.world.snapshot_frames())
    app.world.close()
