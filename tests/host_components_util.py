"""Shared driver for the host side-table tests (CPU: oracle world; GPU: engine): the table of opaque handles must behave
exactly like an OPTIONAL 8-byte POD column of the oracle that carries the handle ids through the same request vectors —
i.e. like ComponentSnapshotPlugin::save / load with the four-way match (component_snapshot.rs:66-123)."""
import numpy as np

from bevy_ggrs_b200 import capi
from bevy_ggrs_b200.host_components import HostComponents
from bevy_ggrs_b200.session import ADVANCE, LOAD, SAVE, Request
from oracle_backend import OracleWorld

OPT = capi.BGR_STRATEGY_OPTIONAL


class Sprite:
    """Stand-in for bevy's Sprite: holds a shared handle (the Arc) — not plain bytes."""

    def __init__(self, handle_id, image):
        self.handle_id, self.image = handle_id, image

    def __copy__(self):  # Clone: a new Sprite sharing the same asset
        return Sprite(self.handle_id, self.image)


def make_world(w, n, depth, with_handle_column):
    health = w.rollback_component("Health", 4, capi.BGR_STRATEGY_CLONE | OPT)
    tag = w.rollback_component("Tag", 12, capi.BGR_STRATEGY_COPY)
    handle = w.rollback_component("SpriteHandle", 8, capi.BGR_STRATEGY_COPY | OPT) if with_handle_column else None
    w.checksum_component(tag, 0, 12)
    w.checksum_component(health, 0, 4)
    w.add_system(capi.BGR_SYS_U32_SATSUB_DESPAWN, [health], [0, 1])
    w.build()
    w.spawn(n)
    rng = np.random.default_rng(5)
    w.write_component(health, 0, rng.integers(3, 30, n, dtype=np.uint32))
    w.write_component(tag, 0, rng.integers(0, 2**32, (n, 3), dtype=np.uint32))
    return health, tag, handle


def run_side_table_against_handle_column(product_world, n=300, d=4, ticks=16, seed=3):
    """`product_world`: the world the side table rides on (engine or a second oracle).  The oracle twin carries the same
    handles in an optional POD column.  Returns counters for the caller's assertions."""
    orc = OracleWorld()
    make_world(product_world, n, 8, False)
    _, _, hcol = make_world(orc, n, 8, True)
    image = object()  # the shared asset every clone must keep pointing at
    table = HostComponents(product_world)
    sprite = table.register("Sprite")
    rng = np.random.default_rng(seed)
    # initial sprites on 2/3 of the entities
    for r in range(n):
        if r % 3:
            table.insert(sprite, r, Sprite(1000 + r, image))
            orc.insert_component(hcol, r, np.uint64(1000 + r))
        else:
            orc.remove_component(hcol, r)
    SESS = (capi.BGR_SESSION_SYNCTEST, 8, d, 0)
    frame, next_id, stats = 0, 5000, {"inserted": 0, "removed": 0, "rolled_back": 0}

    def compare(where):
        alive = np.asarray(orc.read_alive(0, n)).astype(bool)
        assert np.array_equal(np.asarray(product_world.read_alive(0, n)).astype(bool), alive), where
        has = np.asarray(orc.has_component(hcol, 0, n)).astype(bool) & alive
        vals = np.asarray(orc.read_component(hcol, 0, n)).view(np.uint64).reshape(-1)
        got = dict(table.items(sprite))
        assert sorted(got) == list(np.flatnonzero(has)), where
        for r, s in got.items():
            assert s.handle_id == int(vals[r]) and s.image is image, (where, r)
            assert table.get(sprite, r) is s
        for r in np.flatnonzero(~has)[:20]:
            assert table.get(sprite, int(r)) is None
        return len(got)

    for tick in range(ticks):
        reqs = []
        if tick >= d:
            reqs.append(Request(LOAD, frame - d))
            for k in range(d):
                reqs += [Request(ADVANCE, 0, [0]), Request(SAVE, frame - d + k + 1)] if k < d - 1 else [Request(ADVANCE, 0, [0])]
            stats["rolled_back"] += 1
        reqs += [Request(SAVE, frame), Request(ADVANCE, 0, [0])]
        a = product_world.handle_requests(SESS, reqs)
        b = orc.handle_requests(SESS, reqs)
        assert a == b, f"tick {tick}"
        table.handle_requests(reqs)
        frame += 1
        compare(f"after tick {tick}")
        # code outside GgrsSchedule changes sprites; the next tick's rollback undoes it for the re-simulated frames
        alive = np.flatnonzero(np.asarray(orc.read_alive(0, n)).astype(bool))
        for r in rng.choice(alive, min(6, len(alive)), replace=False):
            r = int(r)
            if orc.has_component(hcol, r, 1)[0]:
                if rng.integers(2):
                    table.remove(sprite, r); orc.remove_component(hcol, r); stats["removed"] += 1
                else:  # overwrite with another asset id
                    table.insert(sprite, r, Sprite(next_id, image)); orc.insert_component(hcol, r, np.uint64(next_id)); next_id += 1
            else:
                table.insert(sprite, r, Sprite(next_id, image)); orc.insert_component(hcol, r, np.uint64(next_id)); next_id += 1
                stats["inserted"] += 1
        compare(f"after edits of tick {tick}")
    stats["alive"] = int(np.asarray(orc.read_alive(0, n)).sum())
    stats["sprites"] = compare("end")
    stats["snapshots"] = sorted(sprite.snapshots)
    stats["ring"] = sorted(product_world.snapshot_frames())
    orc.close()
    return stats
