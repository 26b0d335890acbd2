"""Request-vector shapes produced by the session stand-ins — pure host logic."""
import pytest

from bevy_ggrs_b200.session import (ADVANCE, LOAD, SAVE, InvalidRequest, MismatchedChecksum, P2PTraceSession,
                                    SyncTestSession, count_advances)


def shape(reqs):
    return [(r.kind, r.frame) if r.kind != ADVANCE else (ADVANCE,) for r in reqs]


def drive(sess, ticks, checksum_of=lambda f, version: f):
    out = []
    for _ in range(ticks):
        for h in range(sess.num_players()):
            sess.add_local_input(h, 0)
        reqs = sess.advance_frame()
        for r in reqs:
            if r.kind == SAVE:
                sess.save_cell(r.frame, checksum_of(r.frame, 0))
        out.append(reqs)
    return out


def test_synctest_tick_shape():
    d = 3
    ticks = drive(SyncTestSession(1, d, 8), 10)
    # frames 0..d: plain [Save(f), Adv]
    for f in range(d + 1):
        assert shape(ticks[f]) == [(SAVE, f), (ADVANCE,)]
    # f > d: [Load(f-d), Adv, Save(f-d+1), Adv, ..., Save(f-1), Adv, Save(f), Adv]
    for f in range(d + 1, 10):
        want = [(LOAD, f - d), (ADVANCE,)]
        for k in range(f - d + 1, f):
            want += [(SAVE, k), (ADVANCE,)]
        want += [(SAVE, f), (ADVANCE,)]
        assert shape(ticks[f]) == want
        assert count_advances(ticks[f]) == d + 1
        assert sum(1 for r in ticks[f] if r.kind == SAVE) == d


def test_synctest_rejects_check_distance_ge_max_prediction():
    with pytest.raises(InvalidRequest):
        SyncTestSession(1, 8, 8)
    SyncTestSession(1, 8, 9)


def test_synctest_detects_mismatch():
    sess = SyncTestSession(1, 2, 8)
    version = {"n": 0}

    def drifting(frame, _):
        version["n"] += 1
        return frame * 1000 + version["n"]  # every re-save differs
    with pytest.raises(MismatchedChecksum) as ei:
        drive(sess, 10, drifting)
    assert ei.value.mismatched_frames


def test_synctest_input_delay():
    sess = SyncTestSession(2, 1, 8, input_delay=2)
    seen = []
    for t in range(6):
        sess.add_local_input(0, 10 + t)
        sess.add_local_input(1, 20 + t)
        reqs = sess.advance_frame()
        for r in reqs:
            if r.kind == SAVE:
                sess.save_cell(r.frame, 0)
        seen.append(list(reqs[-1].inputs))
    # the input given at frame f is used at frame f+2; the first two frames use the blank input
    assert seen[0] == [0, 0] and seen[1] == [0, 0]
    assert seen[2] == [10, 20] and seen[5] == [13, 23]


def test_p2p_trace_shape_and_determinism():
    a, b = P2PTraceSession(seed=0xB200), P2PTraceSession(seed=0xB200)
    depths = []
    for t in range(200):
        for s in (a, b):
            s.add_local_input(0, 0)
        ra, rb = a.advance_frame(), b.advance_frame()
        assert shape(ra) == shape(rb)
        L = a.last_rollback_depth
        depths.append(L)
        assert 0 <= L <= min(8, t)
        if L == 0:
            assert shape(ra) == [(SAVE, t), (ADVANCE,)]
        else:
            assert ra[0].kind == LOAD and ra[0].frame == t - L
            assert count_advances(ra) == L + 1
            assert [r.frame for r in ra if r.kind == SAVE] == list(range(t - L + 1, t + 1))
    assert 0.35 < depths.count(0) / len(depths) < 0.65
    assert max(depths) == 8
