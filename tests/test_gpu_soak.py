"""GPU: differential soak (tools/soak.py) -- random batch sizes around every launch-shape boundary; the integer (with
and without phase mixing), FP64 and hybrid commitment kernels must agree with each other, the rotation-kernel response
with the NTT response, honest proofs must verify, and Sum / Linear proofs in their default lowering (chunked three-prime
epilogue, shared challenge image, one-launch product sums) must agree with the plain lowering run by the generic interpreter."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_differential_soak():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "soak.py"), "24", "11"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "SOAK PASSED" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
