"""Entity-range sharding across GPUs: the only cross-shard step of the hot path.

Each rank owns a contiguous range of rollback entities (its own engine, its own ring; state never
moves).  Per saved frame a shard produces raw partials — the XOR of per-entity hashes per checksummed
column (component_checksum.rs:81-90, *before* the final `result.hash()` at :93), its live-row count and
its RollbackOrdered length.  NCCL has no XOR reduction, so the partials are all-gathered (a few hundred
bytes) and folded locally: XOR the column words, sum the counts, then `bgr_fold_partials` applies the
three scalar hashes of component_checksum.rs:93-95, entity_checksum.rs:35-43 and checksum.rs:88-99.

Works with any torch.distributed backend: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from . import capi
from .engine import fold_partials

_COLS = capi.BGR_MAX_CHECKSUM_COLUMNS


def pack_partials(partials: Sequence["capi.bgr_partial"]) -> np.ndarray:
    """(k, COLS+2) int64: column XORs (bit pattern), active, total."""
    buf = np.zeros((len(partials), _COLS + 2), dtype=np.uint64)
    for i, p in enumerate(partials):
        for c in range(_COLS):
            buf[i, c] = p.xor_[c]
        buf[i, _COLS] = p.active
        buf[i, _COLS + 1] = p.total
    return buf.view(np.int64)


def combine_gathered(gathered: np.ndarray, partials: Sequence["capi.bgr_partial"]) -> List[Tuple[int, int]]:
    """gathered: (world, k, COLS+2) uint64 -> [(frame, checksum_u128)] after the cross-shard fold."""
    g = gathered.view(np.uint64)
    out = []
    for i, p in enumerate(partials):
        q = capi.bgr_partial()
        q.frame, q.n_columns = p.frame, p.n_columns
        q.active = int(g[:, i, _COLS].sum())
        q.total = int(g[:, i, _COLS + 1].sum())
        for c in range(_COLS):
            q.xor_[c] = int(np.bitwise_xor.reduce(g[:, i, c]))
        out.append((p.frame, fold_partials(q)))
    return out


def all_fold(partials: Sequence["capi.bgr_partial"], device=None, group=None) -> List[Tuple[int, int]]:
    """all_gather the shard partials and fold them; every rank returns the same checksums."""
    import torch
    import torch.distributed as dist

    if len(partials) == 0:
        return []
    world = dist.get_world_size(group)
    t = torch.from_numpy(pack_partials(partials))
    if device is not None:
        t = t.to(device)
    g = torch.empty((world * t.shape[0], t.shape[1]), dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(g, t, group=group)
    return combine_gathered(g.cpu().numpy().reshape(world, t.shape[0], t.shape[1]), partials)


# numpy views of the C structs (include/bevy_ggrs_b200.h)
PARTIAL_DTYPE = np.dtype([("frame", "<i4"), ("n_columns", "<u4"), ("active", "<u8"), ("total", "<u8"), ("xor_", "<u8", (_COLS,))])
CHECKSUM_DTYPE = np.dtype([("frame", "<i4"), ("has_checksum", "<u4"), ("lo", "<u8"), ("hi", "<u8")])
assert PARTIAL_DTYPE.itemsize == 24 + 8 * _COLS and CHECKSUM_DTYPE.itemsize == 24


class PartialBuffer:
    """Preallocated array of bgr_partial that bgr_collect_partials fills in place (no per-tick allocation)."""

    def __init__(self, capacity: int):
        self.arr = np.zeros(capacity, dtype=PARTIAL_DTYPE)
        self.n = 0
        self._lib = capi.load_library()
        self._cnt = __import__("ctypes").c_uint32()

    def collect_from(self, engine) -> int:
        import ctypes as C
        room = self.arr.shape[0] - self.n
        st = self._lib.bgr_collect_partials(engine._h, self.arr.ctypes.data + self.n * PARTIAL_DTYPE.itemsize, room, C.byref(self._cnt))
        if st != capi.BGR_OK:
            raise capi.BgrError(st, self._lib.bgr_last_error().decode())
        self.n += min(room, self._cnt.value)
        return self._cnt.value

    def take(self) -> np.ndarray:
        out = self.arr[: self.n].copy()
        self.n = 0
        return out


def all_fold_array(partials: np.ndarray, device=None, group=None):
    """Vectorised cross-shard fold of a PARTIAL_DTYPE array: one all_gather for the whole batch.
    Returns [(frame, checksum_u128)]."""
    import torch
    import torch.distributed as dist

    k = partials.shape[0]
    if k == 0:
        return []
    world = dist.get_world_size(group)
    buf = np.empty((k, _COLS + 2), dtype=np.uint64)
    buf[:, :_COLS] = partials["xor_"]
    buf[:, _COLS] = partials["active"]
    buf[:, _COLS + 1] = partials["total"]
    t = torch.from_numpy(buf.view(np.int64))
    if device is not None:
        t = t.to(device)
    g = torch.empty((world * k, _COLS + 2), dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(g, t, group=group)
    gh = g.cpu().numpy().view(np.uint64).reshape(world, k, _COLS + 2)
    comb = partials.copy()
    comb["xor_"] = np.bitwise_xor.reduce(gh[:, :, :_COLS], axis=0)
    comb["active"] = gh[:, :, _COLS].sum(axis=0)
    comb["total"] = gh[:, :, _COLS + 1].sum(axis=0)
    out = np.zeros(k, dtype=CHECKSUM_DTYPE)
    lib = capi.load_library()
    st = lib.bgr_fold_partials_n(comb.ctypes.data, k, out.ctypes.data)
    if st != capi.BGR_OK:
        raise capi.BgrError(st, lib.bgr_last_error().decode())
    return [(int(f), (int(h) << 64) | int(l)) for f, l, h in zip(out["frame"], out["lo"], out["hi"])]


def shard_range(total_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous entity range [first, first+count) of `rank` (SURVEY §8e partitioning)."""
    base, rem = divmod(total_rows, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)
