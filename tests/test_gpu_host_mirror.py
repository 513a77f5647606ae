"""Runs tests/cpp/test_host_mirror.cpp: the reference's integration tests and doctests written against the
C++ host-side mirror of its API (ring-zk_b200/host/ring_zk.hpp), linked against the C-ABI library."""
import subprocess

import pytest

import __graft_entry__ as ge

pytestmark = pytest.mark.gpu


def test_cpp_host_mirror_replays_reference_tests():
    exe = ge.build_host_mirror_test()
    res = subprocess.run([exe, "3"], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "HOST_MIRROR PASSED" in res.stdout
