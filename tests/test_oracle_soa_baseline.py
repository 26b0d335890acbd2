"""The optimised CPU SoA bar must compute exactly what the faithful restatement computes."""
import numpy as np

from bevy_ggrs_b200.session import SAVE, P2PTraceSession, SyncTestSession
from bevy_ggrs_b200.stress import populate, register_particles, synth_particles
from oracle_backend import OracleWorld, SoaWorld


def _drive(sess, worlds, ticks):
    hist = [[] for _ in worlds]
    for t in range(ticks):
        for h in range(sess.num_players()):
            sess.add_local_input(h, 0)
        reqs = sess.advance_frame()
        for i, w in enumerate(worlds):
            cs = w.handle_requests(sess.info(), reqs)
            hist[i] += cs
        for f, c in hist[0][-sum(1 for r in reqs if r.kind == SAVE):]:
            sess.save_cell(f, c)
    return hist


def test_soa_bar_matches_faithful_restatement_synctest_and_p2p():
    for make in (lambda: SyncTestSession(2, 6, 8, input_delay=2), lambda: P2PTraceSession(2, 8, 2, seed=5)):
        n = 3000
        tf, vel, ttl = synth_particles(n, 9, 3, 40)
        orc = OracleWorld()
        cols = register_particles(orc)
        populate(orc, cols, tf, vel, ttl)
        soa = SoaWorld(tf, vel, ttl, depth=8, threads=3)
        h = _drive(make(), [orc, soa], 30)
        assert h[0] == h[1] and len(h[0]) > 30
        stf, svel, sttl, salive = soa.columns()
        otf, oalive = orc.read_component_alive(cols[0], 0, n)
        m = oalive.astype(bool)
        assert np.array_equal(salive.astype(bool), m)
        assert np.array_equal(stf.view(np.uint8).reshape(n, 40)[m], otf[m])
        soa.close(); orc.close()


def test_soa_bar_matches_faithful_restatement_at_100k_entities():
    """The size at which tests/test_gpu_baseline_shapes.py starts trusting the SoA bar as a second oracle (C5, 10M):
    100k entities, the headline window d = 8, entities dying inside the rollback window, every thread count path."""
    n, d = 100_000, 8
    tf, vel, ttl = synth_particles(n, 0x5A, 6, 25)
    orc = OracleWorld(save_threads=4)
    cols = register_particles(orc)
    populate(orc, cols, tf, vel, ttl)
    soa = SoaWorld(tf, vel, ttl, depth=d + 1, threads=7)        # a thread count that does not divide the rows
    h = _drive(SyncTestSession(2, d, d + 1, input_delay=2), [orc, soa], d + 5)
    assert h[0] == h[1] and len(h[0]) > 3 * d
    stf, svel, sttl, salive = soa.columns()
    m = orc.read_alive(0, n).astype(bool)
    assert 0 < m.sum() < n and np.array_equal(salive.astype(bool), m)
    for c, got in zip(cols, (stf.view(np.uint8).reshape(n, 40), svel.view(np.uint8).reshape(n, 12), sttl.view(np.uint8).reshape(n, 8))):
        assert np.array_equal(got[m], orc.read_component(c, 0, n)[m])
    soa.close(); orc.close()
