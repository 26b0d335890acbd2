"""world_size-2 (and 3) CPU test of the in-engine cross-shard exchange (bevy_ggrs_b200/csrc/shard_group.hpp).

Two processes stand in for two GPUs' engines: each runs an oracle shard (order_base offset), publishes the raw partials
of every SaveGameState into the group's shared segment with bgr_group_publish — the CPU stand-in for what the fused
kernel's last block stores there — and bgr_group_collect waits for every rank, checks that all shards executed the
same request vector, and folds.  The folded checksums must equal the unsharded world's, on every rank, every frame.
torch.distributed (gloo) is only the launcher-side rendezvous that hands every rank the group's unique name."""
import ctypes as C
import os
import socket
import sys
import uuid

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, seed, ticks, d, q, diverge):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from bevy_ggrs_b200 import capi
    from bevy_ggrs_b200.session import SAVE, SyncTestSession
    from bevy_ggrs_b200.sharded import shard_range
    from bevy_ggrs_b200.stress import populate, register_particles, synth_particles
    from oracle_backend import OracleWorld

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    names = [f"bgr_test_{uuid.uuid4().hex}" if rank == 0 else None]
    dist.broadcast_object_list(names, src=0)
    lib = capi.load_library()
    grp = lib.bgr_group_join(names[0].encode(), rank, world, 2, 20000)
    assert grp, lib.bgr_last_error().decode()

    first, count = shard_range(n_total, rank, world)
    tf, vel, ttl = synth_particles(n_total, seed, 3, 14)
    w = OracleWorld(order_base=first)
    cols = register_particles(w)
    populate(w, cols, tf[first:first + count], vel[first:first + count], ttl[first:first + count])
    sess = SyncTestSession(1, d, 8)
    got, gseq, err = [], 0, None
    out = (capi.bgr_checksum * capi.BGR_MAX_REQUESTS)()
    nout = C.c_uint32()
    for t in range(ticks):
        sess.add_local_input(0, 0)
        reqs = sess.advance_frame()
        parts = []
        for r in reqs:  # request by request so that the raw partials of every Save can be captured
            w.handle_requests(sess.info(), [r])
            if r.kind == SAVE:
                parts.append(w.last_partial())
        if diverge and rank == 1 and t == 5:
            parts[0].frame += 1000      # this shard was handed a different request vector
        arr = (capi.bgr_partial * max(1, len(parts)))(*parts)
        gseq += 1
        assert lib.bgr_group_publish(grp, gseq, arr, len(parts)) == 0, lib.bgr_last_error().decode()
        st = lib.bgr_group_collect(grp, gseq, out, capi.BGR_MAX_REQUESTS, C.byref(nout))
        if st != 0:
            err = lib.bgr_last_error().decode()
            break
        assert nout.value == len(parts)
        for i in range(nout.value):
            assert out[i].has_checksum == 1 and out[i].hi == 0
            sess.save_cell(out[i].frame, out[i].lo)
            got.append((out[i].frame, out[i].lo))
    q.put((rank, got, err))
    if not diverge:
        dist.barrier()
    lib.bgr_group_leave(grp)
    dist.destroy_process_group()


def _unsharded(n_total, seed, ticks, d):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bevy_ggrs_b200.session import SyncTestSession
    from bevy_ggrs_b200.stress import populate, register_particles, synth_particles
    from oracle_backend import OracleWorld
    w = OracleWorld()
    cols = register_particles(w)
    populate(w, cols, *synth_particles(n_total, seed, 3, 14))
    sess = SyncTestSession(1, d, 8)
    want = []
    for t in range(ticks):
        sess.add_local_input(0, 0)
        cs = w.handle_requests(sess.info(), sess.advance_frame())
        for f, c in cs:
            sess.save_cell(f, c)
        want += cs
    return want


def _run(world, n_total, seed, ticks, d, diverge=False):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, seed, ticks, d, q, diverge)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return {r: (got, err) for r, got, err in res}


@pytest.mark.parametrize("world", [2, 3])
def test_group_fold_equals_the_unsharded_checksum_on_every_rank(world):
    n_total, seed, ticks, d = 1001, 77, 30, 4      # 30 ticks > 8 result buffers: the reuse hand-shake is exercised
    want = _unsharded(n_total, seed, ticks, d)
    res = _run(world, n_total, seed, ticks, d)
    for r in range(world):
        got, err = res[r]
        assert err is None
        assert got == want
    assert len(want) > ticks


def test_group_detects_shards_that_executed_different_request_vectors():
    res = _run(2, 200, 5, 10, 2, diverge=True)
    errs = [res[r][1] for r in range(2)]
    assert all(e and "different request vectors" in e for e in errs)
