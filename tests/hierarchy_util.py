"""tests/hierarchy.rs restated on the plugin mirror (any backend): parent / child / grandchild entities that all carry
`Rollback`, linked by `ChildOf`.  Here `ChildOf` is an optional 8-byte rollback column holding the parent's
RollbackOrdered index: engine rows are stable across rollback (a despawned-and-restored entity comes back as the same
row), so the link needs neither ChildOfSnapshotPlugin's entity remapping (childof_snapshot.rs) nor MapEntities
(component_map.rs) — it is saved / restored / presence-tracked like any other optional POD column."""
import struct

import numpy as np

from bevy_ggrs_b200 import capi
from bevy_ggrs_b200.plugin import (App, GgrsPlugin, GgrsSchedule, LocalInputs, ReadInputs, ResourceSystem, Session, Startup,
                                   SyncTestMismatch, System)
from bevy_ggrs_b200.session import SyncTestSession


def frame_counter(res):  # hierarchy.rs:47-49
    res["FrameCounter"][:] = struct.pack("<H", (struct.unpack("<H", res["FrameCounter"])[0] + 1) & 0xFFFF)


def build_app(backend, levels, with_delete_system):
    """levels = 2: parent -> child; 3: parent -> child -> grandchild.  Returns (app, columns, mismatches, state)."""
    app = App(backend)
    app.insert_resource(Session.SyncTest(SyncTestSession(1, 2)))      # common::synctest_session(2)
    app.add_plugins(GgrsPlugin())
    state = {"delete": False}

    def input_system(a):  # hierarchy.rs:17-27: 1 on the frame the delete message was sent
        v = 1 if state["delete"] else 0
        state["delete"] = False
        a.insert_resource(LocalInputs({0: v}))
    app.add_systems(ReadInputs, input_system)
    markers = [app.rollback_optional_component_with_copy(n, 1) for n in ("ParentEntity", "ChildEntity", "GrandchildEntity")[:levels]]
    child_of = app.rollback_optional_component_with_copy("ChildOf", 8)
    app.rollback_resource_with_copy("FrameCounter", bytes(2))
    app.add_systems(GgrsSchedule, ResourceSystem(frame_counter))
    if with_delete_system:  # delete_child_system, hierarchy.rs:36-45
        app.add_systems(GgrsSchedule, System(capi.BGR_SYS_DESPAWN_ON_INPUT, [child_of], [0, 1]))

    def setup(a):  # commands.spawn((ParentEntity, Rollback)).with_children(|p| p.spawn((ChildEntity, Rollback)) ...)
        w = a.world
        first = w.spawn(levels)
        for lvl in range(levels):
            row = first + lvl
            for k, m in enumerate(markers):
                if k != lvl:
                    w.remove_component(m, row)
            if lvl == 0:
                w.remove_component(child_of, row)
            else:
                w.write_component(child_of, row, np.array([first + lvl - 1], dtype=np.uint64))   # ChildOf(parent)
    app.add_systems(Startup, setup)
    bad = []
    app.add_observer(SyncTestMismatch, lambda ev: bad.append(ev))
    return app, markers, child_of, bad, state


def snapshot_of_world(app, markers, child_of, levels):
    w = app.world
    alive = w.read_alive(0, levels).astype(bool)
    has_marker = [w.has_component(m, 0, levels).astype(bool) for m in markers]
    has_link = w.has_component(child_of, 0, levels).astype(bool)
    link = w.read_component(child_of, 0, levels).view(np.uint64)[:, 0]
    return alive, has_marker, has_link, link


def run_recursive_hierarchy(backend):
    """hierarchy.rs:55-122 recursive_hierarchy_is_preserved_through_rollback"""
    app, markers, child_of, bad, _ = build_app(backend, 3, with_delete_system=False)
    for _ in range(20):
        app.update()
    alive, has_marker, has_link, link = snapshot_of_world(app, markers, child_of, 3)
    assert not bad
    assert alive.tolist() == [True, True, True]                                        # all three levels still exist
    assert [h.tolist() for h in has_marker] == [[True, False, False], [False, True, False], [False, False, True]]
    assert has_link.tolist() == [False, True, True]                                    # child and grandchild kept their ChildOf link
    assert link[1] == 0 and link[2] == 1                                               # ... to the right parents
    assert struct.unpack("<H", app.resources["FrameCounter"])[0] == app.rollback_frame_count() == 19
    return app


def run_hierarchy_with_deletion(backend):
    """hierarchy.rs:125-188 hierarchy: the child is despawned inside GgrsSchedule on one frame; rollbacks re-simulate that
    frame; afterwards the child is gone for good and the parent still exists."""
    app, markers, child_of, bad, state = build_app(backend, 2, with_delete_system=True)
    app.update()
    alive, has_marker, has_link, link = snapshot_of_world(app, markers, child_of, 2)
    assert alive.tolist() == [True, True] and has_link.tolist() == [False, True] and link[1] == 0   # the world is set up
    app.update()
    state["delete"] = True            # app.world_mut().resource_mut::<Messages<DeleteChildEntityMessage>>().write(..)
    seen_alive_again = False
    for _ in range(5):                # enough updates for rollbacks across the deletion frame
        app.update()
    alive, has_marker, has_link, link = snapshot_of_world(app, markers, child_of, 2)
    assert not bad
    assert alive.tolist() == [True, False]                       # "Child exists after deletion" / "Parent doesn't exist"
    assert (has_marker[1] & alive).sum() == 0 and (has_link & alive).sum() == 0
    assert has_marker[0][0]
    return app


def run_reference_survives_despawn_and_restore(world):
    """What MapEntities exists for (component_map.rs:1-5): "After a rollback, some entities may have been recreated with
    new Entity IDs" — so components holding an Entity must be remapped.  Here a reference is the target's
    RollbackOrdered index (its row), and a Load brings a despawned entity back AS THE SAME ROW: the reference is valid
    again without any fix-up.  Parent (row 0) is despawned on frame 1, a Load of frame 0 restores it; the child's ChildOf
    still names row 0, which is alive again with its own data."""
    from bevy_ggrs_b200.session import ADVANCE, LOAD, SAVE, Request
    NOSESS = (capi.BGR_SESSION_NONE, 0, 0, 0)
    parent = world.rollback_component("ParentEntity", 4, capi.BGR_STRATEGY_COPY | capi.BGR_STRATEGY_OPTIONAL)
    child_of = world.rollback_component("ChildOf", 8, capi.BGR_STRATEGY_COPY | capi.BGR_STRATEGY_OPTIONAL)
    world.checksum_component(parent, 0, 4)
    world.checksum_component(child_of, 0, 8)
    world.add_system(capi.BGR_SYS_DESPAWN_ON_INPUT, [parent], [0, 1])     # despawn the PARENT when player 0 presses 1
    world.build()
    first = world.spawn(3)
    world.write_component(parent, first, np.array([0xAAAA], dtype=np.uint32))
    for row in (first + 1, first + 2):
        world.remove_component(parent, row)
        world.write_component(child_of, row, np.array([first], dtype=np.uint64))
    world.remove_component(child_of, first)
    cs = []
    cs += world.handle_requests(NOSESS, [Request(SAVE, 0), Request(ADVANCE, 0, [1]), Request(SAVE, 1), Request(ADVANCE, 0, [0])])
    assert world.read_alive(0, 3).tolist() == [0, 1, 1]                   # the parent is gone, its children dangle
    cs += world.handle_requests(NOSESS, [Request(LOAD, 0)])
    assert world.read_alive(0, 3).tolist() == [1, 1, 1]                   # restored as the SAME row ...
    assert world.read_component(parent, 0, 1).view(np.uint32)[0, 0] == 0xAAAA
    link = world.read_component(child_of, 0, 3).view(np.uint64)[:, 0]
    assert link[1] == first and link[2] == first                          # ... so the references are valid again as they are
    assert world.has_component(child_of, 0, 3).tolist() == [0, 1, 1]
    cs += world.handle_requests(NOSESS, [Request(ADVANCE, 0, [0]), Request(SAVE, 1)])
    assert cs[1][0] == cs[2][0] == 1 and cs[1][1] != cs[2][1]            # frame 1 without / with the parent: different checksums
    return cs
