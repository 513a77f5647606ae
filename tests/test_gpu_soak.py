"""GPU: differential soak (tools/soak.py) -- random batch sizes around every launch-shape boundary; the split-key
commitment program (with and without phase mixing) and the generic two-prime program must agree with each other, the
rotation-kernel response with the NTT response, Open verify by signed rotations with the NTT-domain product (honest and
tampered), and Sum / Linear proofs in their default lowering (chunked three-prime epilogue, shared challenge image,
one-launch product sums) must agree with the plain lowering run by the generic interpreter (RZK_TEST_LOWERING)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_differential_soak():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "soak.py"), "24", "11"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "SOAK PASSED" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
