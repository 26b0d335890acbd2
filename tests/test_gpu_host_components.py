"""Host side table for non-POD rollback components riding on the ENGINE: entity existence comes
from the alive mask in HBM; the table must track the oracle's optional handle column through SyncTest rollbacks with
in-window despawns, on the one-launch path and on the stepwise path."""
import pytest

from bevy_ggrs_b200 import capi
from bevy_ggrs_b200.engine import Engine
from bevy_ggrs_b200.host_components import HostComponents
from host_components_util import Sprite, make_world, run_side_table_against_handle_column

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("generic_kernel")]


@pytest.mark.parametrize("flags", [0, capi.BGR_CFG_FORCE_STEPWISE])
@pytest.mark.parametrize("n", [300, 1300])
def test_side_table_on_the_engine_tracks_the_oracle_handle_column(flags, n):
    eng = Engine(max_entities=n + 8, max_depth=8, flags=flags)
    st = run_side_table_against_handle_column(eng, n=n, d=4, ticks=16)
    assert st["rolled_back"] == 12 and st["inserted"] > 0 and st["removed"] > 0
    assert 0 < st["alive"] < n and 0 < st["sprites"] <= st["alive"]
    assert st["snapshots"] == st["ring"]
    eng.close()


def test_despawn_outside_the_schedule_takes_the_sprite_with_it():
    eng = Engine(max_entities=16, max_depth=8)
    make_world(eng, 4, 8, False)
    t = HostComponents(eng)
    c = t.register("Sprite")
    for r in range(4):
        t.insert(c, r, Sprite(r, None))
    eng.despawn(2)
    assert t.get(c, 2) is None and [r for r, _ in t.items(c)] == [0, 1, 3]
    with pytest.raises(KeyError):
        t.insert(c, 2, Sprite(9, None))
    first = eng.spawn(1)   # rows are never reused: the new entity is a new row without a sprite
    assert first == 4 and t.get(c, 4) is None
    eng.close()
