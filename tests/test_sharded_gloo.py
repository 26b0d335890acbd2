"""world_size-2 gloo test of the N>1 path: two entity-range shards (oracle worlds with order_base
offsets standing in for two GPUs' engines) exchange checksum partials with all_gather and fold them with
the product's bgr_fold_partials; the result must equal the unsharded world's checksum, every frame."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, seed, ticks, d, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from bevy_ggrs_b200.session import SAVE, SyncTestSession
    import numpy as np
    from bevy_ggrs_b200.sharded import PARTIAL_DTYPE, all_fold, all_fold_array, shard_range
    from bevy_ggrs_b200.stress import populate, register_particles, synth_particles
    from oracle_backend import OracleWorld

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    first, count = shard_range(n_total, rank, world)
    tf, vel, ttl = synth_particles(n_total, seed, 3, 14)          # same global population on every rank
    w = OracleWorld(order_base=first)
    cols = register_particles(w)
    populate(w, cols, tf[first:first + count], vel[first:first + count], ttl[first:first + count])
    sess = SyncTestSession(1, d, 8)
    got = []
    raw = []
    for t in range(ticks):
        sess.add_local_input(0, 0)
        reqs = sess.advance_frame()
        # run request by request so that the partials of every Save can be captured
        for r in reqs:
            w.handle_requests(sess.info(), [r])
            if r.kind == SAVE:
                part = w.last_partial()
                raw.append((part.frame, part.n_columns, part.active, part.total, tuple(part.xor_[c] for c in range(6))))
                folded = all_fold([part])
                frame, cs = folded[0]
                sess.save_cell(r.frame, cs)
                got.append((r.frame, cs))
    # the batched, vectorised fold used by bench.py must agree with the per-save fold
    arr = np.array(raw, dtype=PARTIAL_DTYPE)
    assert all_fold_array(arr) == got
    if rank == 0:
        q.put(got)
    dist.barrier()
    dist.destroy_process_group()


def test_two_shards_fold_to_the_unsharded_checksum():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bevy_ggrs_b200.session import SAVE, SyncTestSession
    from bevy_ggrs_b200.stress import populate, register_particles, synth_particles
    from oracle_backend import OracleWorld

    n_total, seed, ticks, d = 1001, 77, 18, 4
    # unsharded truth
    w = OracleWorld()
    cols = register_particles(w)
    populate(w, cols, *synth_particles(n_total, seed, 3, 14))
    sess = SyncTestSession(1, d, 8)
    want = []
    for t in range(ticks):
        sess.add_local_input(0, 0)
        reqs = sess.advance_frame()
        cs = w.handle_requests(sess.info(), reqs)
        for f, c in cs:
            sess.save_cell(f, c)
        want += cs

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, seed, ticks, d, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == want
    assert len(got) > ticks  # several saves per tick once rollbacks start


def test_shard_range_partitions_exactly():
    from bevy_ggrs_b200.sharded import shard_range
    for total in (0, 1, 7, 8, 1001, 10_000_000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0
            for a, b in zip(spans, spans[1:]):
                assert a[0] + a[1] == b[0]
            assert spans[-1][0] + spans[-1][1] == total
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
