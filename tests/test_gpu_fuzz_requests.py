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

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
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


@pytest.mark.parametrize("flags", [0, capi.BGR_CFG_FORCE_STEPWISE])
@pytest.mark.parametrize("seed", list(range(12)))
def test_random_worlds_and_request_vectors_match_the_oracle(seed, flags):
    rng = np.random.default_rng(1000 + seed)
    eng, orc, cols, sizes, optional, depth = _make_worlds(rng, flags)
    frame = 0  # RollbackFrameCount of both worlds
    n_vectors = 40
    for step in range(n_vectors):
        frames = orc.snapshot_frames()
        assert eng.snapshot_frames() == frames
        choice = rng.random()
        reqs = []
        if choice < 0.45 or not frames:                       # a plain tick
            reqs = [Request(SAVE, frame), Request(ADVANCE, 0, [int(rng.integers(0, 5))])]
            frame += 1
        elif choice < 0.85:                                   # a rollback into a snapshot that exists, then resimulation
            g = int(rng.choice(frames))
            reqs = [Request(LOAD, g)]
            f = g
            for k in range(frame - g):
                if k > 0:
                    reqs.append(Request(SAVE, f))
                reqs.append(Request(ADVANCE, 0, [int(rng.integers(0, 5))]))
                f += 1
            reqs += [Request(SAVE, f), Request(ADVANCE, 0, [int(rng.integers(0, 5))])]
            frame = f + 1
        elif choice < 0.95:                                   # spectator-style catch-up: advances only
            k = int(rng.integers(1, 4))
            reqs = [Request(ADVANCE, 0, [int(rng.integers(0, 5))]) for _ in range(k)]
            frame += k
        else:                                                 # an invalid rollback: same panic text, nothing executed
            with pytest.raises(BgrError) as ee:
                eng.handle_requests(NOSESS, [Request(LOAD, frame + 1000)])
            with pytest.raises(OracleError) as eo:
                orc.handle_requests(NOSESS, [Request(LOAD, frame + 1000)])
            assert ee.value.status == capi.BGR_ERR_NO_SNAPSHOT and str(ee.value) == str(eo.value)
            orc.set_rollback_frame_count(frame)               # the reference had already set the frame when it panicked
            eng.set_rollback_frame_count(frame)
            continue
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
                    val = rng.integers(1, 200, sizes[i], dtype=np.uint8)
                    for w in (eng, orc):
                        w.insert_component(cols[i], int(r), val)
        if rng.random() < 0.1 and rows < eng.max_entities - 8:
            k = int(rng.integers(1, 6))
            vals = [rng.integers(1, 40, (k, s), dtype=np.uint8) for s in sizes]
            for w in (eng, orc):
                first = w.spawn(k)
                for i, c in enumerate(cols):
                    w.write_component(c, first, vals[i])
        if step % 8 == 7:
            _compare_state(eng, orc, cols)
    _compare_state(eng, orc, cols)
    eng.close(); orc.close()
