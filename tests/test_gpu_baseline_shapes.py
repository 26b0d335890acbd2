"""-m gpu: BASELINE.json's configurations at their FULL size, compared with a CPU oracle checksum for checksum and
(at the end) column for column — not only through self-consistency:

  C3  stress_test 1M entities, 16-frame window, checksum every frame      vs oracle/world.hpp   (2 rollback ticks)
  C4  simulated 2-peer P2P session, 1M entities, max_prediction 8          vs oracle/world.hpp   (10 ticks of the trace)
  C5  10M entities, 32-frame speculative rollback                          vs oracle/soa_baseline.hpp, the second oracle
      that tests/test_oracle_soa_baseline.py proves equal to world.hpp checksum for checksum at <= 100k entities
      (world.hpp itself needs ~15 s per 10M-entity save)

The faithful oracle costs ~1.5 s per 1M-entity SaveGameState (it rebuilds the reference's per-frame hash maps), which
bounds how many ticks each case runs."""
import numpy as np
import pytest

from bevy_ggrs_b200 import capi
from bevy_ggrs_b200.engine import Engine
from bevy_ggrs_b200.session import LOAD, P2PTraceSession, SyncTestSession
from bevy_ggrs_b200.stress import populate, register_particles, synth_particles
from oracle_backend import OracleWorld, SoaWorld

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _engine(n, pop, maxp):
    eng = Engine(max_entities=n, max_depth=maxp)
    cols = register_particles(eng)
    eng.build()
    populate(eng, cols, *pop)
    return eng, cols


def _oracle(pop):
    orc = OracleWorld(save_threads=8)
    cols = register_particles(orc)
    populate(orc, cols, *pop)
    return orc


def _drive_pair(sess_e, sess_o, eng, orc, stop, inputs=lambda t, h: (1 << 5) if (t + h) % 3 == 0 else 0, max_ticks=200):
    """Identical sessions on both worlds; every tick's checksums must be equal.  `stop(tick, requests)` ends the run."""
    n_cs = rollbacks = 0
    for t in range(max_ticks):
        for s in (sess_e, sess_o):
            for h in range(s.num_players()):
                s.add_local_input(h, inputs(t, h))
        re, ro = sess_e.advance_frame(), sess_o.advance_frame()
        ce, co = eng.handle_requests(sess_e.info(), re), orc.handle_requests(sess_o.info(), ro)
        assert ce == co, f"tick {t}"
        for f, c in ce:
            sess_e.save_cell(f, c); sess_o.save_cell(f, c)
        n_cs += len(ce)
        rollbacks += 1 if (re and re[0].kind == LOAD) else 0
        if stop(t, re, rollbacks):
            break
    return n_cs, rollbacks


def _same_state(eng, cols, n, alive_o, cols_o):
    alive = alive_o.astype(bool)
    assert np.array_equal(eng.read_alive(0, n).astype(bool), alive)
    for c, want in zip(cols, cols_o):
        got = eng.read_component(c, 0, n)
        assert np.array_equal(got[alive], want[alive])


def test_c3_1m_entities_16_frame_window_against_the_oracle():
    n, d = 1_000_000, 16
    pop = synth_particles(n, 0xC3, 12, 60)                       # entities die inside the run
    eng, cols = _engine(n, pop, d + 1)
    orc = _oracle(pop)
    n_cs, rollbacks = _drive_pair(SyncTestSession(2, d, d + 1, input_delay=2), SyncTestSession(2, d, d + 1, input_delay=2),
                                  eng, orc, stop=lambda t, reqs, rb: rb >= 2)
    assert rollbacks == 2 and n_cs >= 2 * d + 16
    assert eng.last_path_fused()
    alive = orc.read_alive(0, n)
    assert 0 < alive.sum() < n
    _same_state(eng, cols, n, alive, [orc.read_component(c, 0, n) for c in cols])
    assert eng.snapshot_frames() == orc.snapshot_frames()
    eng.close(); orc.close()


def test_c4_1m_entities_p2p_trace_against_the_oracle():
    n, maxp = 1_000_000, 8
    pop = synth_particles(n, 0xC4, 4, 30)
    eng, cols = _engine(n, pop, maxp)
    orc = _oracle(pop)
    n_cs, rollbacks = _drive_pair(P2PTraceSession(2, maxp, 2, seed=0xB200), P2PTraceSession(2, maxp, 2, seed=0xB200),
                                  eng, orc, stop=lambda t, reqs, rb: t >= 9)
    assert n_cs >= 10 and rollbacks >= 2                         # the trace rolls back on about half of the ticks
    alive = orc.read_alive(0, n)
    assert 0 < alive.sum() < n
    _same_state(eng, cols, n, alive, [orc.read_component(c, 0, n) for c in cols])
    assert eng.snapshot_frames() == orc.snapshot_frames() and eng.confirmed_frame_count() == orc.confirmed_frame_count()
    eng.close(); orc.close()


def test_c5_10m_entities_32_frame_window_against_the_soa_oracle():
    n, d = 10_000_000, 32
    pop = synth_particles(n, 0xC5, 20, 120)
    eng, cols = _engine(n, pop, d + 1)                           # 34 images x 610 MB = 20.7 GB of HBM
    soa = SoaWorld(*pop, depth=d + 1)
    n_cs, rollbacks = _drive_pair(SyncTestSession(2, d, d + 1, input_delay=2), SyncTestSession(2, d, d + 1, input_delay=2),
                                  eng, soa, stop=lambda t, reqs, rb: rb >= 2)
    assert rollbacks == 2 and n_cs >= 2 * d + 32
    assert eng.last_path_fused()
    tf, vel, ttl, alive = soa.columns()
    assert 0 < alive.sum() < n
    _same_state(eng, cols, n, alive, [tf.view(np.uint8).reshape(n, 40), vel.view(np.uint8).reshape(n, 12),
                                      ttl.view(np.uint8).reshape(n, 8)])
    eng.close(); soa.close()
