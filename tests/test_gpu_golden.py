"""GPU: the CUDA engine (through the C ABI) reproduces the committed golden vectors bit for bit."""
import importlib

import pytest

import golden_check as gc

pytestmark = pytest.mark.gpu
engine = importlib.import_module("ring-zk_b200.engine")
UB = engine.unpack_bitmap


class EngineAdapter:
    def __init__(self, e, B):
        self.e, self.B = e, B

    def commit(self, x, r):
        c, ok = self.e.commit(x, r)
        return c, UB(ok, self.B)

    def open_commit(self, x, r, y):
        c, t, ok = self.e.open_commit(x, r, y)
        return c, t, UB(ok, self.B)

    def open_respond(self, y, r, d):
        return self.e.open_respond(y, r, d)

    def open_verify(self, z, t, c1, d):
        return UB(self.e.open_verify(z, t, c1, d), self.B)

    def linear_commit(self, *a):
        return self.e.linear_commit(*a)

    def linear_respond(self, *a):
        return self.e.linear_respond(*a)

    def linear_verify(self, *a):
        return UB(self.e.linear_verify(*a), self.B)

    def sum_commit(self, *a):
        return self.e.sum_commit(*a)

    def sum_respond(self, *a):
        return self.e.sum_respond(*a)

    def sum_verify(self, *a):
        return UB(self.e.sum_verify(*a), self.B)


@pytest.mark.parametrize("lowering", ["", "generic", "norotw", "generic,nosparse,norot,nodimg,nofuse,nosegments"])
def test_engine_matches_golden_vectors(lowering, monkeypatch):
    """default lowering (compile-time programs, rotation kernels and rotation sums), the same programs through the generic
    interpreter, Linear / Sum first equations without rotation sums, and the plain one (generic interpreter, NTT products only)"""
    monkeypatch.setenv("RZK_TEST_LOWERING", lowering)
    G = gc.load()
    e = engine.Engine(N=512, device=0)
    try:
        e.set_key_blocks(G["a1p"], G["a2p"])
        assert gc.check(G, EngineAdapter(e, G["x"].shape[0])) == 2
    finally:
        e.close()
