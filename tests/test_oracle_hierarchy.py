"""tests/hierarchy.rs on the oracle through the plugin mirror (CPU): see tests/hierarchy_util.py."""
from hierarchy_util import run_hierarchy_with_deletion, run_recursive_hierarchy
from oracle_backend import OracleWorld


def test_recursive_hierarchy_is_preserved_through_rollback():
    run_recursive_hierarchy(OracleWorld())


def test_hierarchy_child_deleted_inside_the_schedule_stays_deleted():
    run_hierarchy_with_deletion(OracleWorld())


def test_entity_reference_survives_despawn_and_restore_without_mapping():
    from hierarchy_util import run_reference_survives_despawn_and_restore
    cs = run_reference_survives_despawn_and_restore(OracleWorld())
    assert len(cs) == 3
