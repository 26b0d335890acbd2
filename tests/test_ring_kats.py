"""The reference's own 11 GgrsSnapshots unit tests (src/snapshot/mod.rs:361-508), ported
one-to-one and run against BOTH the oracle ring and the engine's host-side ring (bgr_ring_*).
Plus depth defaults (mod.rs:107-116) and a randomized equivalence sweep between the two."""
import random

import pytest

from ring_adapters import I32_MAX, I32_MIN, EngineRing, OracleRing, RollbackPanic

IMPLS = [OracleRing, EngineRing]


@pytest.fixture(params=IMPLS, ids=lambda c: c.__name__)
def snap_with_depth(request):
    return lambda depth: request.param(depth)


def test_default_depth_is_default_fps():
    # Default depth = DEFAULT_FPS = 60 (mod.rs:112, lib.rs:58)
    assert OracleRing().depth() == 60
    assert EngineRing().depth() == 60


# --- push ---
def test_push_evicts_oldest_when_depth_exceeded(snap_with_depth):  # mod.rs:364-376
    s = snap_with_depth(3)
    for i in range(5):
        s.push(i, i)
    assert s.peek(0) is None
    assert s.peek(1) is None
    assert s.peek(2) == 2
    assert s.peek(3) == 3
    assert s.peek(4) == 4


def test_push_older_frame_discards_newer(snap_with_depth):  # mod.rs:379-390
    s = snap_with_depth(8)
    s.push(5, 50)
    s.push(6, 60)
    s.push(7, 70)
    s.push(5, 99)
    assert s.peek(5) == 99
    assert s.peek(6) is None
    assert s.peek(7) is None


def test_push_same_frame_replaces(snap_with_depth):  # mod.rs:393-399
    s = snap_with_depth(8)
    s.push(3, 10)
    s.push(3, 20)
    assert s.peek(3) == 20


# --- confirm ---
def test_confirm_prunes_older_frames(snap_with_depth):  # mod.rs:404-418
    s = snap_with_depth(8)
    for i in range(6):
        s.push(i, i)
    s.confirm(3)
    assert s.peek(0) is None and s.peek(1) is None and s.peek(2) is None
    assert s.peek(3) == 3  # confirm is an exclusive lower bound
    assert s.peek(4) == 4
    assert s.peek(5) == 5


def test_confirm_beyond_all_frames_empties_storage(snap_with_depth):  # mod.rs:421-431
    s = snap_with_depth(8)
    for i in range(4):
        s.push(i, i)
    s.confirm(100)
    for i in range(4):
        assert s.peek(i) is None


def test_confirm_on_empty_does_not_panic(snap_with_depth):  # mod.rs:434-438
    s = snap_with_depth(8)
    s.confirm(5)


# --- rollback ---
def test_rollback_to_existing_frame(snap_with_depth):  # mod.rs:443-451
    s = snap_with_depth(8)
    for i in range(5):
        s.push(i, i * 10)
    s.rollback(2)
    assert s.get() == 20


def test_rollback_discards_newer_frames(snap_with_depth):  # mod.rs:454-464
    s = snap_with_depth(8)
    for i in range(5):
        s.push(i, i)
    s.rollback(2)
    assert s.peek(3) is None
    assert s.peek(4) is None
    assert s.peek(2) == 2


def test_rollback_missing_frame_panics(snap_with_depth):  # mod.rs:467-473
    s = snap_with_depth(8)
    s.push(0, 0)
    with pytest.raises(RollbackPanic, match="Could not rollback to 99"):
        s.rollback(99)


def test_get_on_empty_panics(snap_with_depth):  # mod.rs:226-230
    s = snap_with_depth(8)
    with pytest.raises(RollbackPanic, match="no snapshot available"):
        s.get()


# --- i32 wraparound ---
def test_push_wraps_i32_max_to_min_retains_history(snap_with_depth):  # mod.rs:480-493
    s = snap_with_depth(8)
    s.push(I32_MAX - 2, 1)
    s.push(I32_MAX - 1, 2)
    s.push(I32_MAX, 3)
    s.push(I32_MIN, 4)
    assert s.peek(I32_MAX - 2) == 1
    assert s.peek(I32_MAX - 1) == 2
    assert s.peek(I32_MAX) == 3
    assert s.peek(I32_MIN) == 4


def test_push_max_after_min_evicts_min_as_future(snap_with_depth):  # mod.rs:497-508
    s = snap_with_depth(8)
    s.push(I32_MIN, 1)
    s.push(I32_MAX, 2)
    assert s.peek(I32_MIN) is None, "i32::MIN should be evicted as a future frame"
    assert s.peek(I32_MAX) == 2


def test_randomized_equivalence_oracle_vs_engine_ring():
    """10k random push/confirm/rollback/set_depth ops: identical observable state."""
    rng = random.Random(0xB200)
    for trial in range(20):
        depth = rng.randint(1, 12)
        a, b = OracleRing(depth), EngineRing(depth)
        frame = rng.choice([0, 5, I32_MAX - 20])
        known = set()
        for step in range(500):
            op = rng.random()
            if op < 0.55:
                frame = frame + 1 if frame < I32_MAX else I32_MIN
                v = rng.getrandbits(32)
                a.push(frame, v); b.push(frame, v)
                known.add(frame)
            elif op < 0.70 and known:
                f = rng.choice(sorted(known))
                v = rng.getrandbits(32)
                a.push(f, v); b.push(f, v)
                frame = f
            elif op < 0.80:
                f = frame - rng.randint(0, 6)
                if f >= I32_MIN:
                    a.confirm(f); b.confirm(f)
            elif op < 0.92 and known:
                f = rng.choice(sorted(known))
                ra = rb = None
                try:
                    a.rollback(f)
                except RollbackPanic as e:
                    ra = str(e)
                try:
                    b.rollback(f)
                except RollbackPanic as e:
                    rb = str(e)
                assert ra == rb
                if ra is None:
                    frame = f
            else:
                d = rng.randint(1, 12)
                a.set_depth(d); b.set_depth(d)
            for f in list(known)[-24:]:
                assert a.peek(f) == b.peek(f), (trial, step, f)
