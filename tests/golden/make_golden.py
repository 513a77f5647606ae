"""Generates tests/golden/ringzk_n512.npz: seeded inputs and every intermediate of the three protocols at
N = 512, Params::default(), computed by the pure-Python big-int restatement of the reference
(oracle/pyref.py: schoolbook negacyclic products on Python ints, the reference's own operation order,
/root/reference/src/{mat,commit}.rs and src/prove/{open,linear,sum}.rs).

The reference itself (Rust + the un-vendored crate poly-ring-xnp1) cannot be built or imported in this
image and holds no golden vectors for ring products, so these vectors pin the two CPU
oracles and the CUDA engine to one independently written implementation, not to a run of the crate.

    python tests/golden/make_golden.py        # ~2 minutes of pure-Python arithmetic; output is committed
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402

synth = importlib.import_module("ring-zk_b200.synth")

N, B, T, SEED = 512, 2, 2, 20261018


def to_mat(a):
    return [[list(map(int, row))] for row in a]


def vec(a):
    return [list(map(int, row)) for row in a]


def arr(m):
    """pyref Mat (rows x 1, m[i] = [poly]) or list of polys (m[i] = poly) -> [rows][N] int64 array"""
    out = np.zeros((len(m), N), np.int64)
    for i, e in enumerate(m):
        p = e[0] if e and isinstance(e[0], list) else e
        out[i, :len(p)] = p
    return out


def main():
    s = synth.Synth(SEED, N=N)
    a1p, a2p = s.key()
    P = pyref.Params(N=N)
    ck = pyref.CommitmentKey(P, [[list(map(int, p)) for p in row] for row in a1p],
                             [[list(map(int, p)) for p in row] for row in a2p])
    x, r, y, d = s.message(B, ragged=True), s.small(B), s.gaussian(B), s.challenge(B)
    g, rp, yp = s.scalar(B), s.small(B), s.gaussian(B)
    gs, xs = s.scalar(B, T), s.uniform_q(B, T, 1)
    rs, ys = s.small(B, T), s.gaussian(B, T)
    rps, yps = s.small(B), s.gaussian(B)
    out = dict(a1p=a1p, a2p=a2p, x=x, r=r, y=y, d=d, g=g, rp=rp, yp=yp, gs=gs, xs=xs, rs=rs, ys=ys, rps=rps, yps=yps)
    res = {k: [] for k in ("c", "t", "z", "open_ok", "open_bad",
                           "l_gx", "l_cp", "l_c", "l_t", "l_tp", "l_u", "l_z", "l_zp", "l_ok", "l_bad",
                           "s_xp", "s_cp", "s_cs", "s_ts", "s_tp", "s_u", "s_zs", "s_zp", "s_ok", "s_bad")}
    for i in range(B):
        di = list(map(int, d[i]))
        # Open (open.rs)
        ok, c, t = pyref.open_commit(ck, P, vec(x[i]), to_mat(r[i]), to_mat(y[i]))
        z = pyref.open_respond(P, to_mat(y[i]), to_mat(r[i]), di)
        c1, _ = pyref.c1_c2(c, P)
        res["c"].append(arr(c)); res["t"].append(arr(t)); res["z"].append(arr(z))
        res["open_ok"].append(ok and pyref.open_verify(ck, P, z, t, c1, di))
        zb = [[list(p[0])] for p in z]; zb[1][0][7] += 1
        res["open_bad"].append(pyref.open_verify(ck, P, zb, t, c1, di))
        # Linear (linear.rs)
        gi = list(map(int, g[i]))
        L = pyref.linear_commit(ck, P, gi, vec(x[i]), to_mat(rp[i]), to_mat(r[i]), to_mat(y[i]), to_mat(yp[i]))
        lz, lzp = pyref.linear_respond(P, to_mat(y[i]), to_mat(yp[i]), to_mat(r[i]), to_mat(rp[i]), di)
        for k, v in (("l_gx", L["gx"]), ("l_cp", L["cp"]), ("l_c", L["c"]), ("l_t", L["t"]), ("l_tp", L["tp"]),
                     ("l_u", L["u"]), ("l_z", lz), ("l_zp", lzp)):
            res[k].append(arr(v))
        res["l_ok"].append(L["ok"] and pyref.linear_verify(ck, P, lz, lzp, L["c"], L["cp"], gi, L["t"], L["tp"], L["u"], di))
        ub = [[list(p[0])] for p in L["u"]]; ub[0][0][9] += 1
        res["l_bad"].append(pyref.linear_verify(ck, P, lz, lzp, L["c"], L["cp"], gi, L["t"], L["tp"], ub, di))
        # Sum, T terms (sum.rs)
        gsi = [list(map(int, gs[i, k])) for k in range(T)]
        S = pyref.sum_commit(ck, P, gsi, [vec(xs[i, k]) for k in range(T)], to_mat(rps[i]),
                             [to_mat(rs[i, k]) for k in range(T)], [to_mat(ys[i, k]) for k in range(T)], to_mat(yps[i]))
        szs, szp = pyref.sum_respond(P, [to_mat(ys[i, k]) for k in range(T)], to_mat(yps[i]),
                                     [to_mat(rs[i, k]) for k in range(T)], to_mat(rps[i]), di)
        res["s_xp"].append(arr(S["xp"])); res["s_cp"].append(arr(S["cp"]))
        res["s_cs"].append(np.stack([arr(c_) for c_ in S["cs"]])); res["s_ts"].append(np.stack([arr(t_) for t_ in S["ts"]]))
        res["s_tp"].append(arr(S["tp"])); res["s_u"].append(arr(S["u"]))
        res["s_zs"].append(np.stack([arr(z_) for z_ in szs])); res["s_zp"].append(arr(szp))
        res["s_ok"].append(S["ok"] and pyref.sum_verify(ck, P, szs, szp, S["cs"], S["cp"], gsi, S["ts"], S["tp"], S["u"], di))
        gb = [list(p) for p in gsi]; gb[1][3] += 1
        res["s_bad"].append(pyref.sum_verify(ck, P, szs, szp, S["cs"], S["cp"], gb, S["ts"], S["tp"], S["u"], di))
        print("item", i, "done", flush=True)
    for k, v in res.items():
        a = np.stack([np.asarray(e) for e in v])
        out[k] = a.astype(np.int32) if a.dtype == np.int64 else a
    assert all(out["open_ok"]) and all(out["l_ok"]) and all(out["s_ok"])
    assert not any(out["open_bad"]) and not any(out["l_bad"]) and not any(out["s_bad"])
    np.savez_compressed(os.path.join(HERE, "ringzk_n512.npz"), seed=SEED, **out)
    print("wrote", os.path.join(HERE, "ringzk_n512.npz"))


if __name__ == "__main__":
    main()
