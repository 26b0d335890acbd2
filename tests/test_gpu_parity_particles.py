"""-m gpu: the CUDA path through the C ABI vs the oracle, bit-exact (checksums, f32 bit patterns of
Transform/Velocity, Ttl, alive mask, frame counters, ring contents)."""
import numpy as np
import pytest

from bevy_ggrs_b200 import capi
from parity_util import run_particles_synctest_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d,ticks", [(1, 2, 12), (33, 2, 12), (1000, 7, 24), (4097, 4, 16), (50_000, 8, 14)])
def test_synctest_fused_matches_oracle(n, d, ticks):
    r = run_particles_synctest_pair(n, d, ticks, seed=1234 + n)
    assert r["fused"]
    assert r["n_checksums"] > 0
    assert r["checksums_equal"]
    assert r["state_equal"]
    assert r["mismatch_events"] == (0, 0)
    assert r["frames"][0] == r["frames"][1] == ticks
    assert r["ring"][0] == r["ring"][1]
    assert r["confirmed"][0] == r["confirmed"][1]


@pytest.mark.parametrize("n,d,ticks", [(257, 3, 14), (10_000, 8, 12)])
def test_synctest_stepwise_matches_oracle(n, d, ticks):
    r = run_particles_synctest_pair(n, d, ticks, seed=99, flags=capi.BGR_CFG_FORCE_STEPWISE)
    assert not r["fused"]
    assert r["checksums_equal"] and r["state_equal"]
    assert r["ring"][0] == r["ring"][1]


@pytest.mark.parametrize("flags", [0, capi.BGR_CFG_FORCE_STEPWISE])
def test_despawn_inside_the_rollback_window(flags):
    """ttl ~ U(1, 2*window): particles die inside the window, are resurrected by Load and die again
    (entity.rs:55-99 reconcile + despawn_particles).  active count, entity checksum part and the
    alive mask must track the oracle exactly."""
    r = run_particles_synctest_pair(5000, 6, 30, seed=5, ttl_lo=1, ttl_hi=12, flags=flags)
    assert r["checksums_equal"] and r["state_equal"]
    assert r["active"][0] == r["active"][1]
    assert r["active"][0] == 0  # everything is dead after 30 frames


def test_fused_and_stepwise_agree_checksum_for_checksum():
    a = run_particles_synctest_pair(20_000, 8, 12, seed=42)
    b = run_particles_synctest_pair(20_000, 8, 12, seed=42, flags=capi.BGR_CFG_FORCE_STEPWISE)
    assert a["checksums"] == b["checksums"]
    # one launch per tick on the fused path (+0 for setup): stepwise needs dozens
    assert a["launches"] == 12
    assert b["launches"] > 10 * a["launches"]


@pytest.mark.parametrize("flags", [0, capi.BGR_CFG_FORCE_STEPWISE])
def test_spawn_particles_inside_the_rollback_window(flags):
    """spawn_particles.run_if(spawn_pressed) with the rolled-back ParticleRng — rows are born
    (and, with ttl 9, die) inside the rollback window; Load shrinks RollbackOrdered and the resimulation must
    re-spawn the same particles.  Row count, alive mask, columns and every checksum track the oracle."""
    r = run_particles_synctest_pair(300, 6, 36, seed=11, ttl_lo=3, ttl_hi=40, flags=flags, spawn_rate=40,
                                    spawn_ttl=9, startup_burst=True)
    assert r["fused"] == (flags == 0)
    assert r["rows"][0] == r["rows"][1] > 300 + 40 * 10
    assert r["checksums_equal"] and r["state_equal"]
    assert r["active"][0] == r["active"][1]
    assert r["mismatch_events"] == (0, 0)
    assert r["ring"][0] == r["ring"][1]
