"""tests/resource_lifecycle.rs:42-117 on the plugin mirror's host-side resources: a resource inserted or removed
mid-session must be removed / re-inserted when a Load crosses the transition frame, otherwise the checksummed
FrameLog diverges during resimulation and SyncTestMismatch fires.  Pure host logic (+ any backend; the oracle here)."""
import struct

from bevy_ggrs_b200 import capi
from bevy_ggrs_b200.plugin import (App, GgrsPlugin, GgrsSchedule, LocalInputs, ReadInputs, ResourceSystem, Session,
                                   SyncTestMismatch)
from bevy_ggrs_b200.session import SyncTestSession
from oracle_backend import OracleWorld


def _base(check_distance):
    app = App(OracleWorld())
    app.insert_resource(Session.SyncTest(SyncTestSession(1, check_distance)))
    app.add_plugins(GgrsPlugin())
    app.add_systems(ReadInputs, lambda a: a.insert_resource(LocalInputs({h: 0 for h in a.local_players.handles})))
    app.world.rollback_component("Marker", 4)   # one registered column so that snapshots exist
    return app


def track_wallet(res):  # resource_lifecycle.rs:31-33
    log = struct.unpack("<I", res["FrameLog"])[0]
    res["FrameLog"][:] = struct.pack("<I", (log + (2 if "Wallet" in res else 1)) & 0xFFFFFFFF)


def _run(app, updates=20):
    bad = []
    app.add_observer(SyncTestMismatch, lambda ev: bad.append(ev))
    for _ in range(updates):
        app.update()
    return bad


def test_resource_inserted_mid_session_rolls_back():
    app = _base(4)
    app.rollback_resource_with_clone("Wallet").rollback_resource_with_clone("FrameLog", bytes(4))
    app.checksum_resource_with_hash("FrameLog")

    def insert_wallet_at_frame_3(res, frame):
        if frame == 3:
            res["Wallet"] = bytearray(struct.pack("<I", 100))
    app.add_systems(GgrsSchedule, ResourceSystem(insert_wallet_at_frame_3)).add_systems(GgrsSchedule, ResourceSystem(track_wallet))
    assert not _run(app)
    assert struct.unpack("<I", app.resources["Wallet"])[0] == 100


def test_resource_removed_mid_session_rolls_back():
    app = _base(4)
    app.rollback_resource_with_clone("Wallet", struct.pack("<I", 100)).rollback_resource_with_clone("FrameLog", bytes(4))
    app.checksum_resource_with_hash("FrameLog")

    def remove_wallet_at_frame_3(res, frame):
        if frame == 3:
            res.pop("Wallet", None)
    app.add_systems(GgrsSchedule, ResourceSystem(remove_wallet_at_frame_3)).add_systems(GgrsSchedule, ResourceSystem(track_wallet))
    assert not _run(app)
    assert "Wallet" not in app.resources


def test_unregistered_resource_change_is_detected_as_mismatch():
    """Negative control: a resource that is NOT registered for rollback leaks across Loads and the checksummed
    log diverges (the class of bug docs/pitfalls.md warns about) — the mismatch MUST fire."""
    app = _base(4)
    app.rollback_resource_with_clone("FrameLog", bytes(4)).checksum_resource_with_hash("FrameLog")

    def insert_wallet_at_frame_3(res, frame):
        if frame == 3:
            res["Wallet"] = bytearray(4)
    app.add_systems(GgrsSchedule, ResourceSystem(insert_wallet_at_frame_3)).add_systems(GgrsSchedule, ResourceSystem(track_wallet))
    assert _run(app)
