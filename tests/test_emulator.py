"""Runs the host lane emulator (tests/cpp/emu_check.cpp): the per-lane kernel code of
ring-zk_b200/csrc/rzk_vm_exec.cuh compiled with g++ and checked against the CPU oracle for
every program shape the engine launches (commit, Open, Linear, 3-prime product sums,
tampered / out-of-range inputs).  CPU-only."""
import subprocess

import __graft_entry__ as ge


def test_emulator_matches_oracle():
    exe = ge.build_emulator()
    res = subprocess.run([exe, "6"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "EMU_CHECK PASSED" in res.stdout
