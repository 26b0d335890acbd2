"""tests/hierarchy.rs on the GPU engine through the plugin mirror, on the generic one-launch program and stepwise, with
every checksum compared to the oracle's run of the same app."""
import pytest

from bevy_ggrs_b200 import capi
from bevy_ggrs_b200.engine import Engine
from hierarchy_util import build_app, run_hierarchy_with_deletion, run_recursive_hierarchy
from oracle_backend import OracleWorld

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("generic_kernel")]


@pytest.mark.parametrize("flags", [0, capi.BGR_CFG_FORCE_STEPWISE])
def test_recursive_hierarchy_is_preserved_through_rollback(flags):
    app = run_recursive_hierarchy(Engine(max_entities=8, max_depth=8, flags=flags))
    assert app.world.last_path_fused() == (flags == 0)


@pytest.mark.parametrize("flags", [0, capi.BGR_CFG_FORCE_STEPWISE])
def test_hierarchy_child_deleted_inside_the_schedule_stays_deleted(flags):
    app = run_hierarchy_with_deletion(Engine(max_entities=8, max_depth=8, flags=flags))
    assert app.world.last_path_fused() == (flags == 0)


def test_hierarchy_checksums_match_the_oracle_tick_for_tick():
    eng_app, _, _, _, st_e = build_app(Engine(max_entities=8, max_depth=8), 2, with_delete_system=True)
    orc_app, _, _, _, st_o = build_app(OracleWorld(), 2, with_delete_system=True)
    total = 0
    for i in range(12):
        if i == 3:
            st_e["delete"] = st_o["delete"] = True
        eng_app.update(); orc_app.update()
        assert eng_app.last_checksums == orc_app.last_checksums, i
        total += len(eng_app.last_checksums)
    assert total > 12 and eng_app.world.active_count() == orc_app.world.active_count() == 1


@pytest.mark.parametrize("flags", [0, capi.BGR_CFG_FORCE_STEPWISE])
def test_entity_reference_survives_despawn_and_restore_without_mapping(flags):
    from hierarchy_util import run_reference_survives_despawn_and_restore
    got = run_reference_survives_despawn_and_restore(Engine(max_entities=8, max_depth=8, flags=flags))
    want = run_reference_survives_despawn_and_restore(OracleWorld())
    assert got == want and len(got) == 3
