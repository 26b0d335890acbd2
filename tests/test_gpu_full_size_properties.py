"""-m gpu, BASELINE.json's full sizes: 1M-entity worlds checked through size-independent properties of the domain
(the oracle needs ~1.5 s per 1M-entity save, so only one short run is compared with it directly):

  * SyncTest self-consistency — every re-saved frame reports the checksum first recorded for it (what ggrs checks)
  * save -> N x advance -> load is a round trip: every column and the alive mask come back bit for bit
  * the checksum is linear under entity-range sharding — folding the XOR partials of two half-worlds
    (order_base 0 and E/2) gives exactly the whole world's checksum
  * fused / stepwise / skip-unchanged-planes engines agree checksum for checksum
"""
import numpy as np
import pytest

from bevy_ggrs_b200 import capi
from bevy_ggrs_b200.engine import Engine, fold_partials
from bevy_ggrs_b200.session import ADVANCE, LOAD, SAVE, Request, SyncTestSession
from bevy_ggrs_b200.stress import populate, register_particles, synth_particles
from oracle_backend import OracleWorld

pytestmark = pytest.mark.gpu
E = 1_000_000
NOSESS = (capi.BGR_SESSION_NONE, 0, 0, 0)


def _engine(n, pop, max_depth=9, flags=0, order_base=0, lo=0):
    eng = Engine(max_entities=n, max_depth=max_depth, flags=flags, order_base=order_base)
    cols = register_particles(eng)
    eng.build()
    tf, vel, ttl = pop
    populate(eng, cols, tf[lo:lo + n], vel[lo:lo + n], ttl[lo:lo + n])
    return eng, cols


def _synctest(eng, d, ticks):
    sess = SyncTestSession(2, d, d + 1, input_delay=2)
    hist = []
    for t in range(ticks):
        for h in range(2):
            sess.add_local_input(h, 0)
        cs = eng.handle_requests(sess.info(), sess.advance_frame())   # raises MismatchedChecksum on a desync
        for f, c in cs:
            sess.save_cell(f, c)
        hist += cs
    return hist


def test_1m_synctest_d8_is_self_consistent_and_paths_agree():
    pop = synth_particles(E, 0xB200, 6, 40)        # particles die inside the run
    runs = []
    for flags in (0, capi.BGR_CFG_SKIP_UNCHANGED_PLANES, capi.BGR_CFG_FORCE_STEPWISE):
        eng, _ = _engine(E, pop, flags=flags)
        hist = _synctest(eng, 8, 13 if flags != capi.BGR_CFG_FORCE_STEPWISE else 11)
        first = {}
        for f, c in hist:
            assert first.setdefault(f, c) == c      # SyncTest property
        runs.append(first)
        eng.close()
    for f, c in runs[2].items():
        assert runs[0][f] == c
    assert runs[0] == runs[1]


def test_1m_save_advance_load_round_trip():
    pop = synth_particles(E, 7, 3, 12)
    eng, cols = _engine(E, pop, max_depth=4)
    eng.handle_requests(NOSESS, [Request(SAVE, 0)])
    before = [eng.read_component(c, 0, E) for c in cols] + [eng.read_alive(0, E)]
    eng.handle_requests(NOSESS, [Request(ADVANCE, 0, [0, 0])] * 8 + [Request(SAVE, 8)])
    assert eng.active_count() < E                   # some died
    moved = eng.read_component(cols[0], 0, E)
    assert not np.array_equal(moved, before[0])
    eng.handle_requests(NOSESS, [Request(LOAD, 0)])
    after = [eng.read_component(c, 0, E) for c in cols] + [eng.read_alive(0, E)]
    for a, b in zip(before, after):
        assert np.array_equal(a, b)
    assert eng.rollback_frame_count() == 0 and eng.snapshot_frames() == [0]
    eng.close()


def test_1m_checksum_is_linear_under_entity_range_sharding():
    pop = synth_particles(E, 99, 4, 30)
    reqs = [Request(SAVE, 0)] + [Request(ADVANCE, 0, [0, 0]), Request(SAVE, 0)] * 5
    for i, r in enumerate(r for r in reqs if r.kind == SAVE):
        r.frame = i
    whole, _ = _engine(E, pop, max_depth=8)
    want = whole.handle_requests(NOSESS, reqs)
    whole.close()
    half = E // 2
    shards = [_engine(half, pop, max_depth=8, flags=capi.BGR_CFG_SHARDED, order_base=k * half, lo=k * half)[0] for k in range(2)]
    parts = []
    for s in shards:
        s.handle_requests(NOSESS, reqs)
        parts.append(s.last_partials())
    got = []
    for a, b in zip(*parts):
        q = capi.bgr_partial()
        q.frame, q.n_columns, q.active, q.total = a.frame, a.n_columns, a.active + b.active, a.total + b.total
        for c in range(capi.BGR_MAX_CHECKSUM_COLUMNS):
            q.xor_[c] = a.xor_[c] ^ b.xor_[c]
        got.append((a.frame, fold_partials(q)))
    assert got == want and len(got) == 6
    for s in shards:
        s.close()


def test_1m_short_run_against_the_oracle():
    """One direct comparison at full size: 2 plain ticks + 2 rollback ticks (d=2) on 1M entities."""
    pop = synth_particles(E, 5, 2, 9)
    eng, cols = _engine(E, pop, max_depth=4)
    orc = OracleWorld(save_threads=8)
    ocols = register_particles(orc)
    populate(orc, ocols, *pop)
    se, so = SyncTestSession(1, 2, 3), SyncTestSession(1, 2, 3)
    for _ in range(5):
        for s, w in ((se, eng), (so, orc)):
            s.add_local_input(0, 0)
        ce = eng.handle_requests(se.info(), se.advance_frame())
        co = orc.handle_requests(so.info(), so.advance_frame())
        assert ce == co
        for f, c in ce:
            se.save_cell(f, c); so.save_cell(f, c)
    alive = orc.read_alive(0, E).astype(bool)
    assert np.array_equal(eng.read_alive(0, E).astype(bool), alive) and 0 < alive.sum() < E
    for c in cols:
        assert np.array_equal(eng.read_component(c, 0, E)[alive], orc.read_component(c, 0, E)[alive])
    eng.close(); orc.close()


def test_4m_d16_ring_of_17_slots():
    """A larger world and a deeper window (4M entities x 17 slots = 4.4 GB): self-consistent SyncTest."""
    n = 4_000_000
    pop = synth_particles(n, 21, 400, 400)
    eng, _ = _engine(n, pop, max_depth=17)
    hist = _synctest(eng, 16, 20)
    first = {}
    for f, c in hist:
        assert first.setdefault(f, c) == c
    assert eng.last_path_fused() and eng.active_count() == n
    eng.close()
