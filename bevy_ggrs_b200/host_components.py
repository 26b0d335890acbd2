"""Host-side side table for rollback components that are NOT plain data.

The particles example registers ``rollback_component_with_clone::<Sprite>()`` (examples/stress_tests/particles.rs:191):
``Sprite`` holds an ``Arc`` asset handle, so its bytes cannot live in an HBM column.  Such components stay on the host
(in a real Bevy app: in the ECS, on bevy_ggrs' own ``ComponentSnapshotPlugin``); this module is the mirror of that host
path for the Python plugin, driven by the SAME request vector the engine executes:

* Save  (component_snapshot.rs:66-90):  ``snapshot[rollback] = strategy.store(component)`` for every rollback entity that
  has the component; the snapshot goes into a per-type ``GgrsComponentSnapshots`` ring keyed by frame.
* Load  (component_snapshot.rs:92-123): the four-way match per live rollback entity —
  (Some, Some) overwrite, (Some, None) remove, (None, Some) insert, (None, None) nothing.
* entities: the row index of the engine IS the ``RollbackOrdered`` index and rows are never reused, so the table is keyed
  by row.  Whether an entity exists is the engine's business (the alive mask rides with every HBM snapshot); a value whose
  entity is dead is unobservable: ``get`` / ``items`` consult the live alive mask, and dead entries are pruned after every
  request vector.  A snapshot taken after an in-window despawn may still carry the dead entity's value; it can never be
  seen, because the frame it belongs to restores the entity as dead (rows are not reused).

Nothing here is data-parallel and nothing runs on the GPU: it is the caller side of the hot path.
"""
from __future__ import annotations

import copy
from typing import Any, Callable, Dict, Iterator, Optional, Sequence, Tuple

import numpy as np

from .session import LOAD, SAVE, Request


class HostColumn:
    """One non-POD rollback component type: live values by row + per-frame snapshots."""

    def __init__(self, type_name: str, clone: Callable[[Any], Any] = copy.copy):
        self.type_name = type_name
        self.clone = clone
        self.live: Dict[int, Any] = {}
        self.snapshots: Dict[int, Dict[int, Any]] = {}

    def _store(self) -> Dict[int, Any]:
        c = self.clone
        return {row: c(v) for row, v in self.live.items()}


class HostComponents:
    """All host columns of one world.  ``world`` needs ``read_alive(first, n)``, ``row_count()`` and
    ``snapshot_frames()`` — the engine and the oracle backend both have them."""

    def __init__(self, world):
        self.world = world
        self.columns: Dict[str, HostColumn] = {}
        self._alive: Optional[np.ndarray] = None
        self._alive_key = None     # (row_count, active_count) the cached mask belongs to

    # ---- registration: rollback_component_with_clone::<T>() for a T that is not Copy-able bytes ----
    def register(self, type_name: str, clone: Callable[[Any], Any] = copy.copy) -> HostColumn:
        if type_name in self.columns:
            raise ValueError(f"{type_name} is already registered for rollback")
        col = HostColumn(type_name, clone)
        self.columns[type_name] = col
        return col

    # ---- ECS access outside GgrsSchedule: commands.entity(e).insert(t) / .remove::<T>() / Query<&T> ----
    def _alive_mask(self) -> np.ndarray:
        key = (self.world.row_count(), self.world.active_count())  # host-side counters: spawn / despawn change them
        if self._alive is None or key != self._alive_key:
            n = key[0]
            self._alive = np.asarray(self.world.read_alive(0, n)).astype(bool) if n else np.zeros(0, dtype=bool)
            self._alive_key = key
        return self._alive

    def invalidate(self) -> None:
        """The set of live entities may have changed without the counters showing it (a request vector ran)."""
        self._alive = None

    def _exists(self, row: int) -> bool:
        m = self._alive_mask()
        return 0 <= row < len(m) and bool(m[row])

    def insert(self, col: HostColumn, row: int, value: Any) -> None:
        if not self._exists(row):
            raise KeyError(f"entity of row {row} does not exist")
        col.live[row] = value

    def remove(self, col: HostColumn, row: int) -> None:
        if not self._exists(row):
            raise KeyError(f"entity of row {row} does not exist")
        col.live.pop(row, None)

    def get(self, col: HostColumn, row: int) -> Optional[Any]:
        return col.live.get(row) if self._exists(row) else None

    def items(self, col: HostColumn) -> Iterator[Tuple[int, Any]]:
        m = self._alive_mask()
        return ((r, v) for r, v in sorted(col.live.items()) if r < len(m) and m[r])

    # ---- the host half of handle_requests for these columns ----
    def handle_requests(self, requests: Sequence[Request]) -> None:
        """Replay the request vector the engine just executed.  Call AFTER ``world.handle_requests``."""
        for r in requests:
            if r.kind == SAVE:
                for col in self.columns.values():
                    col.snapshots[r.frame] = col._store()
            elif r.kind == LOAD:
                for col in self.columns.values():
                    snap = col.snapshots.get(r.frame)
                    if snap is None:  # mod.rs:209-212
                        raise RuntimeError(f"Could not rollback to {r.frame}: no snapshot at that moment could be found.")
                    c = col.clone
                    col.live = {row: c(v) for row, v in snap.items()}  # overwrite / insert / remove in one assignment
            # ADVANCE: no compiled GgrsSchedule system touches a host column (a game that moves sprites inside the
            # schedule keeps that component on bevy_ggrs' own path)
        self.invalidate()
        # GgrsSnapshots::push / confirm discard what the engine's ring discarded (mod.rs:144-199)
        kept = set(self.world.snapshot_frames())
        m = self._alive_mask() if any(c.live for c in self.columns.values()) else None
        for col in self.columns.values():
            col.snapshots = {f: s for f, s in col.snapshots.items() if f in kept}
            if m is not None:  # the components of despawned entities are gone with them
                col.live = {r: v for r, v in col.live.items() if r < len(m) and m[r]}
