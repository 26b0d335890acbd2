"""The C++ host mirror of the plugin API (bevy_ggrs_b200/host/bevy_ggrs.hpp) and the reference's
integration tests written against it (tests/cpp/test_host_mirror.cpp)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _binary():
    """The test binary is produced by __graft_entry__.build().  Never rebuild the engine library from inside a
    test process (other tests have it dlopen'ed); only (re)link the small C++ test program if it is missing."""
    import __graft_entry__ as g
    out = os.path.join(ROOT, "tests", "cpp", "test_host_mirror")
    if os.path.exists(out) and os.path.exists(g.LIB):
        return out
    if not os.path.exists(g.LIB):
        g.build_engine()
    return g.build_host_mirror_tests()


def test_host_mirror_builds_and_refuses_without_gpu():
    """CPU box: the C++ layer compiles against the header, links the .so, and the engine refuses to
    start without a device (no CPU fallback).  On a GPU box the same invocation just starts."""
    r = subprocess.run([_binary(), "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert ("refused" in r.stdout) or ("engine started" in r.stdout)


@pytest.mark.gpu
def test_reference_integration_tests_through_cpp_host_mirror():
    r = subprocess.run([_binary()], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all host-mirror tests passed" in r.stdout
