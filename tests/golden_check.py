"""Shared checker for tests/golden/ringzk_n512.npz (made by tests/golden/make_golden.py with the pure-Python
big-int restatement of the reference).  `impl` is anything with the batched protocol methods of
oracle.oracle.Oracle or ring-zk_b200.engine.Engine; both are driven through the adapter below."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ringzk_n512.npz")


def load():
    return dict(np.load(GOLDEN))


def same(a, b, what):
    a, b = np.asarray(a).astype(np.int64), np.asarray(b).astype(np.int64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    assert (a == b).all(), f"{what}: {int((a != b).sum())} coefficients differ"


def check(G, A):
    """A: adapter with commit/open_*/linear_*/sum_* returning plain arrays and bool arrays."""
    B = G["x"].shape[0]
    # Open
    c, t, ok = A.open_commit(G["x"], G["r"], G["y"])
    same(c, G["c"], "open c"); same(t, G["t"], "open t"); assert ok.all()
    z = A.open_respond(G["y"], G["r"], G["d"])
    same(z, G["z"], "open z")
    c1 = np.ascontiguousarray(G["c"][:, :1])
    assert (A.open_verify(G["z"], G["t"], c1, G["d"]) == G["open_ok"]).all() and G["open_ok"].all()
    zb = G["z"].copy(); zb[:, 1, 7] += 1
    assert (A.open_verify(zb, G["t"], c1, G["d"]) == G["open_bad"]).all() and not G["open_bad"].any()
    cc, okc = A.commit(G["x"], G["r"])
    same(cc, G["c"], "commit c"); assert okc.all()
    # Linear
    L = A.linear_commit(G["g"], G["x"], G["rp"], G["r"], G["y"], G["yp"])
    for k in ("gx", "cp", "c", "t", "tp", "u"):
        same(L[k], G["l_" + k], "linear " + k)
    lz, lzp = A.linear_respond(G["y"], G["yp"], G["r"], G["rp"], G["d"])
    same(lz, G["l_z"], "linear z"); same(lzp, G["l_zp"], "linear zp")
    args = (G["l_z"], G["l_zp"], G["l_c"], G["l_cp"], G["g"], G["l_t"], G["l_tp"])
    assert (A.linear_verify(*args, G["l_u"], G["d"]) == G["l_ok"]).all() and G["l_ok"].all()
    ub = G["l_u"].copy(); ub[:, 0, 9] += 1
    assert (A.linear_verify(*args, ub, G["d"]) == G["l_bad"]).all() and not G["l_bad"].any()
    # Sum
    S = A.sum_commit(G["gs"], G["xs"], G["rps"], G["rs"], G["ys"], G["yps"])
    for k in ("xp", "cp", "cs", "ts", "tp", "u"):
        same(S[k], G["s_" + k], "sum " + k)
    zs, zp = A.sum_respond(G["ys"], G["yps"], G["rs"], G["rps"], G["d"])
    same(zs, G["s_zs"], "sum zs"); same(zp, G["s_zp"], "sum zp")
    a2 = (G["s_zs"], G["s_zp"], G["s_cs"], G["s_cp"])
    rest = (G["s_ts"], G["s_tp"], G["s_u"], G["d"])
    assert (A.sum_verify(*a2, G["gs"], *rest) == G["s_ok"]).all() and G["s_ok"].all()
    gb = G["gs"].copy(); gb[:, 1, 3] += 1
    assert (A.sum_verify(*a2, gb, *rest) == G["s_bad"]).all() and not G["s_bad"].any()
    return B
