"""tests/hierarchy.rs on the oracle through the plugin mirror (CPU): see tests/hierarchy_util.py."""
from hierarchy_util import run_hierarchy_with_deletion, run_recursive_hierarchy
from oracle_backend import OracleWorld


def test_recursive_hierarchy_is_preserved_through_rollback():
    run_recursive_hierarchy(OracleWorld())


def test_hierarchy_child_deleted_inside_the_schedule_stays_deleted():
    run_hierarchy_with_deletion(OracleWorld())
