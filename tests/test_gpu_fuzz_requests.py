"""-m gpu: randomized differential test — random schemas (sizes, optional columns, byte-range checksums), random
compiled systems, random populations, random VALID request vectors (plain ticks, rollbacks of random depth into the
snapshots that exist, spectator-style catch-up runs), random host edits between vectors (remove / insert of optional
components, spawns), and the occasional invalid rollback.  After every vector: checksums, frame resources and ring
contents equal the oracle's; periodically every column, presence bit and the alive set.  Each seed runs on the default
one-launch path and on the stepwise path."""
import numpy as np
import pytest

from bevy_ggrs_b200 import capi
from bevy_ggrs_b200.capi import BgrError
from bevy_ggrs_b200.engine import Engine
from bevy_ggrs_b200.session import ADVANCE, LOAD, SAVE, Request
from oracle_backend import OracleError, OracleWorld

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300), pytest.mark.usefixtures("generic_kernel")]
NOSESS = (capi.BGR_SESSION_NONE, 0, 0, 0)
OPT = capi.BGR_STRATEGY_OPTIONAL


def _make_worlds(rng, flags):
    n = int(rng.integers(1, 1400))
    depth = int(rng.integers(2, 9))
    n_cols = int(rng.integers(1, 5))
    sizes = [int(rng.choice([1, 4, 4, 8, 12, 16, 40, 7])) for _ in range(n_cols)]
    optional = [bool(rng.random() < 0.5) for _ in range(n_cols)]
    worlds = [Engine(max_entities=n + 64, max_depth=depth + 1, flags=flags), OracleWorld()]
    cols = []
    for w in worlds:
        cols = [w.rollback_component(f"C{i}", sizes[i], (capi.BGR_STRATEGY_COPY | OPT) if optional[i] else capi.BGR_STRATEGY_CLONE)
                for i in range(n_cols)]
    # checksums: random byte ranges (aligned and unaligned)
    cks = []
    for i in range(n_cols):
        if rng.random() < 0.7:
            off = int(rng.integers(0, sizes[i]))
            ln = int(rng.integers(1, sizes[i] - off + 1))
            if rng.random() < 0.5 and sizes[i] >= 4:
                off, ln = 0, sizes[i] - sizes[i] % 4
            cks.append((i, off, max(1, ln)))
    for w in worlds:
        for i, off, ln in cks:
            w.checksum_component(cols[i], off, ln)
    # systems on columns that have a u32 field
    systems = []
    for i in range(n_cols):
        if sizes[i] >= 4 and sizes[i] % 4 == 0 and rng.random() < 0.8:
            off = 4 * int(rng.integers(0, sizes[i] // 4))
            kind = rng.choice([capi.BGR_SYS_U32_ADD, capi.BGR_SYS_U32_SATSUB_DESPAWN, capi.BGR_SYS_U32_STORE_CALL_COUNT])
            if kind == capi.BGR_SYS_U32_STORE_CALL_COUNT:
                systems.append((int(kind), [cols[i]], [off]))
            else:
                systems.append((int(kind), [cols[i]], [off, int(rng.integers(1, 4))]))
    if optional and rng.random() < 0.5:
        i = int(rng.integers(0, n_cols))
        systems.append((capi.BGR_SYS_DESPAWN_ON_INPUT, [cols[i]], [0, 3]))
    for w in worlds:
        for sid, c, p in systems:
            w.add_system(sid, c, p)
        w.build()
        w.set_depth(depth)
    data = [rng.integers(0, 256, (n, sizes[i]), dtype=np.uint8) for i in range(n_cols)]
    for i in range(n_cols):   # u32 fields small enough that SATSUB despawns some entities inside the run
        if sizes[i] % 4 == 0:
            data[i].view(np.uint32)[:] = rng.integers(1, 30, (n, sizes[i] // 4), dtype=np.uint32)
    removes = [(i, int(r)) for i in range(n_cols) if optional[i] for r in rng.choice(n, size=min(n, int(rng.integers(0, 20))), replace=False)]
    for w in worlds:
        w.spawn(n)
        for i in range(n_cols):
            w.write_component(cols[i], 0, data[i])
        for i, r in removes:
            w.remove_component(cols[i], r)
    return worlds[0], worlds[1], cols, sizes, optional, depth


def _compare_state(eng, orc, cols):
    rows = eng.row_count()
    assert rows == orc.row_count()
    alive = orc.read_alive(0, rows).astype(bool)
    assert np.array_equal(eng.read_alive(0, rows).astype(bool), alive)
    for c in cols:
        vo, ho = orc.read_component_alive(c, 0, rows)
        he = eng.has_component(c, 0, rows).astype(bool)
        assert np.array_equal(he, ho.astype(bool)), f"presence of column {c}"
        assert np.array_equal(eng.read_component(c, 0, rows)[he], vo[he]), f"values of column {c}"


def _drive(eng, orc, cols, sizes, optional, rng, flags, seed, n_vectors=40, input_hi=5, insert_value=None, allow_host_spawn=True):
    frame = 0  # RollbackFrameCount of both worlds
    for step in range(n_vectors):
        frames = orc.snapshot_frames()
        assert eng.snapshot_frames() == frames
        choice = rng.random()
        reqs = []
        inp = lambda: [int(rng.integers(0, input_hi))]
        if choice < 0.45 or not frames:                       # a plain tick
            reqs = [Request(SAVE, frame), Request(ADVANCE, 0, inp())]
            frame += 1
        elif choice < 0.85:                                   # a rollback into a snapshot that exists, then resimulation
            g = int(rng.choice(frames))
            reqs = [Request(LOAD, g)]
            f = g
            for k in range(frame - g):
                if k > 0:
                    reqs.append(Request(SAVE, f))
                reqs.append(Request(ADVANCE, 0, inp()))
                f += 1
            reqs += [Request(SAVE, f), Request(ADVANCE, 0, inp())]
            frame = f + 1
        else:                                                 # spectator-style catch-up: advances only
            k = int(rng.integers(1, 4))
            reqs = [Request(ADVANCE, 0, inp()) for _ in range(k)]
            frame += k
        a, b = eng.handle_requests(NOSESS, reqs), orc.handle_requests(NOSESS, reqs)
        assert a == b, f"seed {seed} step {step}"
        assert eng.rollback_frame_count() == orc.rollback_frame_count() == frame
        if flags == 0:
            assert eng.last_path_fused()
        # host edits between vectors
        rows = orc.row_count()
        alive = np.flatnonzero(orc.read_alive(0, rows))
        opt_cols = [i for i, o in enumerate(optional) if o]
        if opt_cols and alive.size and rng.random() < 0.5:
            for r in rng.choice(alive, size=min(3, alive.size), replace=False):
                i = int(rng.choice(opt_cols))
                if orc.has_component(cols[i], int(r), 1)[0]:
                    for w in (eng, orc):
                        w.remove_component(cols[i], int(r))
                else:
                    val = insert_value(i, rng) if insert_value else rng.integers(1, 200, sizes[i], dtype=np.uint8)
                    for w in (eng, orc):
                        w.insert_component(cols[i], int(r), val)
        if allow_host_spawn and rng.random() < 0.1 and rows < eng.max_entities - 8:
            k = int(rng.integers(1, 6))
            vals = [rng.integers(1, 40, (k, s), dtype=np.uint8) for s in sizes]
            for w in (eng, orc):
                first = w.spawn(k)
                for i, c in enumerate(cols):
                    w.write_component(c, first, vals[i])
        if step % 8 == 7:
            _compare_state(eng, orc, cols)
    _compare_state(eng, orc, cols)
    # Last: an invalid rollback.  The reference panics ("Could not rollback to ...", mod.rs:209-212) after popping every
    # snapshot on its way — the app is dead at that point, so nothing after it is compared; the engine reports the same
    # text as a status and has executed nothing (its ring is untouched).
    before = eng.snapshot_frames()
    with pytest.raises(BgrError) as ee:
        eng.handle_requests(NOSESS, [Request(LOAD, frame + 1000)])
    with pytest.raises(OracleError) as eo:
        orc.handle_requests(NOSESS, [Request(LOAD, frame + 1000)])
    assert ee.value.status == capi.BGR_ERR_NO_SNAPSHOT and str(ee.value) == str(eo.value)
    assert eng.snapshot_frames() == before and eng.rollback_frame_count() == frame


@pytest.mark.parametrize("flags", [0, capi.BGR_CFG_FORCE_STEPWISE])
@pytest.mark.parametrize("seed", list(range(12)))
def test_random_worlds_and_request_vectors_match_the_oracle(seed, flags):
    rng = np.random.default_rng(1000 + seed)
    eng, orc, cols, sizes, optional, depth = _make_worlds(rng, flags)
    _drive(eng, orc, cols, sizes, optional, rng, flags, seed)
    eng.close(); orc.close()


@pytest.mark.parametrize("flags", [0, capi.BGR_CFG_FORCE_STEPWISE])
@pytest.mark.parametrize("seed", list(range(8)))
def test_random_request_vectors_on_the_particles_bundle_match_the_oracle(seed, flags):
    """The stress-test bundle under the same random driver: random optional flags on Velocity / Ttl (fused kernel MODE 2),
    an optional extra passive column of odd size, spawn_particles on random inputs (INPUT_SPAWN = 1 << 4) when no column
    is optional, particles dying inside the window, random rollbacks / catch-up runs / invalid rollbacks."""
    from bevy_ggrs_b200.stress import synth_particles
    rng = np.random.default_rng(5000 + seed)
    n = int(rng.integers(1, 3000))
    depth = int(rng.integers(2, 9))
    opt_v, opt_l = bool(rng.random() < 0.4), bool(rng.random() < 0.4)
    extra = int(rng.choice([0, 0, 5, 16]))
    spawn = (not opt_v and not opt_l) and rng.random() < 0.6
    rate = int(rng.integers(1, 40))
    worlds = [Engine(max_entities=n + 20000, max_depth=depth + 1, flags=flags), OracleWorld()]   # room for every possible spawn
    for w in worlds:
        t = w.rollback_component("Transform", 40, capi.BGR_STRATEGY_CLONE)
        v = w.rollback_component("Velocity", 12, capi.BGR_STRATEGY_COPY | (OPT if opt_v else 0))
        l = w.rollback_component("Ttl", 8, capi.BGR_STRATEGY_COPY | (OPT if opt_l else 0))
        cols, sizes, optional = [t, v, l], [40, 12, 8], [False, opt_v, opt_l]
        if extra:
            cols.append(w.rollback_component("Extra", extra, capi.BGR_STRATEGY_COPY)); sizes.append(extra); optional.append(False)
        w.checksum_component(v, 0, 12, capi.BGR_HASH_FLAG_ASSERT_FINITE_F32)
        w.checksum_component(t, 0, 12, capi.BGR_HASH_FLAG_ASSERT_FINITE_F32)
        if spawn:
            w.add_system(capi.BGR_SYS_PARTICLES_SPAWN, [t, v, l], [rate, 6, 123, 0])
        w.add_system(capi.BGR_SYS_PARTICLES_UPDATE, [t, v])
        w.add_system(capi.BGR_SYS_PARTICLES_DESPAWN, [l])
        w.build()
        w.set_depth(depth)
        tf, vel, ttl = synth_particles(n, 900 + seed, 2, 25, z_fraction=0.3)
        w.spawn(n)
        w.write_component(t, 0, tf); w.write_component(v, 0, vel); w.write_component(l, 0, ttl)
        if extra:
            w.write_component(cols[3], 0, np.random.default_rng(seed).integers(0, 256, (n, extra), dtype=np.uint8))
    eng, orc = worlds

    def insert_value(i, r):
        if i == 1:
            return np.array([1.5, -2.5, 0.25], np.float32).view(np.uint8)
        return np.array([int(r.integers(2, 20))], np.uint64).view(np.uint8)
    _drive(eng, orc, cols, sizes, optional, rng, flags, seed, n_vectors=32, input_hi=32, insert_value=insert_value, allow_host_spawn=False)
    eng.close(); orc.close()
