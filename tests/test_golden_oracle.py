"""CPU: the C oracle reproduces the committed golden vectors (tests/golden/ringzk_n512.npz, produced by the
independent pure-Python big-int restatement, tests/golden/make_golden.py) for commit / Open / Linear / Sum."""
import numpy as np

from oracle import oracle as orc
import golden_check as gc


class OracleAdapter:
    def __init__(self, o):
        self.o = o

    def commit(self, x, r):
        c, ok = self.o.commit_batch(x, r)
        return c, ok.astype(bool)

    def open_commit(self, x, r, y):
        c, t, ok = self.o.open_commit_batch(x, r, y)
        return c, t, ok.astype(bool)

    def open_respond(self, y, r, d):
        return self.o.open_respond_batch(y, r, d)

    def open_verify(self, z, t, c1, d):
        return self.o.open_verify_batch(z, t, c1, d).astype(bool)

    def linear_commit(self, *a):
        return self.o.linear_commit_batch(*a)

    def linear_respond(self, *a):
        return self.o.linear_respond_batch(*a)

    def linear_verify(self, *a):
        return self.o.linear_verify_batch(*a).astype(bool)

    def sum_commit(self, *a):
        return self.o.sum_commit_batch(*a)

    def sum_respond(self, *a):
        return self.o.sum_respond_batch(*a)

    def sum_verify(self, *a):
        return self.o.sum_verify_batch(*a).astype(bool)


def test_c_oracle_matches_golden_vectors():
    G = gc.load()
    o = orc.Oracle(orc.Params(N=512), G["a1p"], G["a2p"])
    assert gc.check(G, OracleAdapter(o)) == 2


def test_golden_file_is_self_describing():
    G = gc.load()
    assert int(G["seed"]) == 20261018 and G["x"].shape == (2, 1, 512) and G["s_cs"].shape == (2, 2, 2, 512)
    half = (3515337053 - 1) // 2
    for k in ("c", "t", "l_u", "s_u", "s_xp"):
        assert np.abs(G[k].astype(np.int64)).max() <= half          # canonical centred residues
