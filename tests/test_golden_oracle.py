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


def test_flat_export_round_trips_and_c_oracle_matches_the_raw_products():
    """tests/golden/ringzk_n512.bin (the file the Rust-side golden test reads, shim/src/golden_vectors.rs) holds exactly the
    npz's arrays plus raw ring products; the C oracle reproduces those products and the representative pins too."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("export_flat", os.path.join(os.path.dirname(gc.GOLDEN), "export_flat.py"))
    ef = importlib.util.module_from_spec(spec); spec.loader.exec_module(ef)
    F, G = ef.load_flat(), gc.load()
    for k, v in G.items():
        a = F[k]
        assert (np.asarray(v).astype(np.int64).reshape(a.shape) == a.astype(np.int64)).all(), k
    o = orc.Oracle(orc.Params(N=512), G["a1p"], G["a2p"])
    a, b = G["a1p"][0, 0].astype(np.int64), G["a2p"][0, 0].astype(np.int64)
    assert (o.poly_mul(a, b) == F["p_ab"]).all()
    assert (o.poly_mul(a, G["r"][0, 1].astype(np.int64)) == F["p_ar"]).all()
    assert (o.poly_mul(a, G["d"][0].astype(np.int64)) == F["p_ad"]).all()
    hi = F["p_hi"].astype(np.int64)
    assert (o.poly_mul(hi, hi) == F["p_hh"]).all()
    assert (o.center(a + b) == F["p_a_plus_b"]).all() and (o.center(a - b) == F["p_a_minus_b"]).all()
    assert (o.center(F["rep_in"].astype(np.int64)) == F["rep_out"]).all()
    assert F["rep_out"][0] == -1757668526          # ZqI64::<3515337053>::from(1757668527).into() per SURVEY 8(c)
