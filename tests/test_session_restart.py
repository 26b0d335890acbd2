"""run_ggrs_schedules without a session (schedule_systems.rs:70-79): "No session has been started yet, reset time data
and snapshots" — LocalPlayers::default(), RollbackFrameCount(0), ConfirmedFrameCount(-1), MaxPredictionWindow(8).
A session that is removed and re-inserted therefore starts from frame 0 again instead of inheriting the old counters."""
from bevy_ggrs_b200.plugin import App, GgrsPlugin, LocalInputs, ReadInputs, Session
from bevy_ggrs_b200.session import SyncTestSession
from oracle_backend import OracleWorld


def _app():
    app = App(OracleWorld())
    app.add_plugins(GgrsPlugin())
    app.add_systems(ReadInputs, lambda a: a.insert_resource(LocalInputs({h: 0 for h in a.local_players.handles})))
    app.world.rollback_component("Marker", 4)
    return app


def test_no_session_resets_the_frame_resources():
    app = _app()
    app.insert_resource(Session.SyncTest(SyncTestSession(2, 2)))
    for _ in range(12):
        app.update()
    assert app.rollback_frame_count() > 5 and app.confirmed_frame_count() >= 0
    assert app.local_players.handles == [0, 1]
    app.remove_resource(Session)
    app.update()                                   # takes the session-less branch
    assert app.rollback_frame_count() == 0
    assert app.confirmed_frame_count() == -1
    assert app.local_players.handles == []
    assert app._accumulator_ns == 0 and app._run_slow is False


def test_session_restart_runs_into_bevy_times_monotonicity_assert():
    """The session-less branch resets RollbackFrameCount but not Time<GgrsTime> (time.rs:100 only rolls it back with
    snapshots), so the first AdvanceFrame of a re-inserted session asks `Time::advance_to` (time.rs:73-75) for an earlier
    moment — bevy_time asserts "tried to move time backwards to an earlier elapsed moment".  The mirror reports the
    same failure instead of silently continuing with the old counters."""
    import pytest
    from oracle_backend import OracleError
    app = _app()
    app.insert_resource(Session.SyncTest(SyncTestSession(1, 2)))
    for _ in range(9):
        app.update()
    app.remove_resource(Session)
    app.update()
    assert app.rollback_frame_count() == 0
    app.insert_resource(Session.SyncTest(SyncTestSession(1, 2)))
    with pytest.raises(OracleError, match="move time backwards"):
        for _ in range(3):
            app.update()
