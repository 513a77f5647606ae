"""Pins the CPU oracle (oracle/) against everything the reference's own tests pin,
and against the independently written big-int twin (oracle/pyref.py).

Reference fixtures used (paths relative to /root/reference):
  src/params.rs:144-150      sigma(1024) == 21780
  src/polynomial.rs:105-121  norms of [1,-2,3,-4] are 10 / 5 / 4
  src/mat.rs:243-406         Mat dot/add/sub/componentwise_mul == the same expression on polynomials
  src/commit.rs:161-170      honest openings verify, swapped openings fail
  tests/test.rs:11-93        honest Open/Linear/Sum transcripts verify at N=16 (4-term sum)
  README.md:32-55            the README Open flow at N=512, x = [1,2,3,4]
"""
import importlib

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import pyref

synth = importlib.import_module("ring-zk_b200").synth


def test_sigma_pin():
    # params.rs:144-150
    o = orc.Oracle(orc.Params(N=1024))
    assert o.sigma() == 21780
    assert pyref.Params(N=1024).standard_deviation(1024) == 21780
    o = orc.Oracle(orc.Params(N=512))
    assert o.sigma() == 15444
    assert o.commit_bound() == 1359072 and o.verify_bound() == 679536


def test_norm_pins():
    # polynomial.rs:105-121
    o = orc.Oracle(orc.Params(N=4))
    p = [1, -2, 3, -4]
    assert o.norm2(p) == 5
    assert pyref.norm_2(p) == 5 and pyref.norm_1(p) == 10 and pyref.norm_infinity(p) == 4


def test_center_representative():
    # canonical centred residue; -1 reads back as -1 (commit.rs:100-105)
    o = orc.Oracle(orc.Params(N=4))
    q = o.P.q
    assert o.center(np.array([-1]))[0] == -1
    assert o.center(np.array([1757668527]))[0] == -1757668526
    assert o.center(np.array([q]))[0] == 0
    assert pyref.center(1757668527, q) == -1757668526


def test_mat_ops_small_literals():
    # mat.rs:243-268, 389-406: N = 4, tiny literals; products fit far below q so the
    # i32 expectations of the reference equal the centred residues here.
    P = orc.Params(N=4)
    o = orc.Oracle(P)
    a00, a01 = [1, 2, 3, 0], [4, 5, 6, 0]
    b00, b10 = [1, 2, 0, 0], [3, 4, 0, 0]
    c = o.mat_dot(np.array([[a00, a01]]), np.array([[b00], [b10]]))
    # (1+2x+3x^2)(1+2x) + (4+5x+6x^2)(3+4x) mod x^4+1
    exp = np.array([1 + 12, 2 + 2 + 16 + 15, 4 + 3 + 20 + 18, 6 + 24], np.int64)
    assert (c[0, 0] == exp).all()
    # wrap-around sign: x^3 * x = -1
    assert (o.poly_mul([0, 0, 0, 1], [0, 1, 0, 0]) == np.array([-1, 0, 0, 0])).all()
    e = [1, 2, 3, 0]
    cm = o.mat_cmul(np.array([[a00, a01]]), e)
    assert (cm[0, 0] == o.poly_mul(a00, e)).all() and (cm[0, 1] == o.poly_mul(a01, e)).all()
    assert (o.mat_add(np.array([[a00, a01]]), np.array([[a00, a01]]))[0, 1] == 2 * np.array(a01)).all()
    assert (o.mat_sub(np.array([[a00, a01]]), np.array([[a00, a01]])) == 0).all()


def _to_mat(arr):
    """[rows][N] ndarray -> pyref Mat (rows x 1)."""
    return [[list(map(int, row))] for row in arr]


def _vec(arr):
    return [list(map(int, row)) for row in arr]


@pytest.mark.parametrize("N", [16, 64])
def test_c_oracle_equals_python_twin(N):
    """Two independently written restatements agree on every intermediate of all three protocols."""
    P = orc.Params(N=N)
    PP = pyref.Params(N=N)
    s = synth.Synth(1234 + N, N=N)
    a1p, a2p = s.key()
    o = orc.Oracle(P, a1p, a2p)
    ck = pyref.CommitmentKey(PP, [[list(map(int, p)) for p in row] for row in a1p],
                             [[list(map(int, p)) for p in row] for row in a2p])
    B, T = 3, 4
    x = s.message(B, ragged=True)
    r, rp = s.small(B), s.small(B)
    y, yp = s.gaussian(B), s.gaussian(B)
    d = s.challenge(B)
    g = s.scalar(B)
    # open
    c, t, ok = o.open_commit_batch(x, r, y)
    z = o.open_respond_batch(y, r, d)
    v = o.open_verify_batch(z, t, c[:, :P.l], d)
    for i in range(B):
        ok_i, c_i, t_i = pyref.open_commit(ck, PP, _vec(x[i]), _to_mat(r[i]), _to_mat(y[i]))
        assert ok_i == bool(ok[i])
        assert c_i == _to_mat(c[i]) and t_i == _vec(t[i])
        z_i = pyref.open_respond(PP, _to_mat(y[i]), _to_mat(r[i]), list(map(int, d[i])))
        assert z_i == _to_mat(z[i])
        assert pyref.open_verify(ck, PP, z_i, t_i, c_i[:PP.l], list(map(int, d[i]))) == bool(v[i]) == True
    # linear
    lc = o.linear_commit_batch(g, x, rp, r, y, yp)
    lz, lzp = o.linear_respond_batch(y, yp, r, rp, d)
    lv = o.linear_verify_batch(lz, lzp, lc["c"], lc["cp"], g, lc["t"], lc["tp"], lc["u"], d)
    for i in range(B):
        ref = pyref.linear_commit(ck, PP, list(map(int, g[i])), _vec(x[i]), _to_mat(rp[i]), _to_mat(r[i]),
                                  _to_mat(y[i]), _to_mat(yp[i]))
        assert ref["gx"] == _vec(lc["gx"][i]) and ref["cp"] == _to_mat(lc["cp"][i])
        assert ref["c"] == _to_mat(lc["c"][i]) and ref["t"] == _vec(lc["t"][i])
        assert ref["tp"] == _vec(lc["tp"][i]) and ref["u"] == _to_mat(lc["u"][i])
        zz, zzp = pyref.linear_respond(PP, _to_mat(y[i]), _to_mat(yp[i]), _to_mat(r[i]), _to_mat(rp[i]),
                                       list(map(int, d[i])))
        assert zz == _to_mat(lz[i]) and zzp == _to_mat(lzp[i])
        assert pyref.linear_verify(ck, PP, zz, zzp, ref["c"], ref["cp"], list(map(int, g[i])), ref["t"],
                                   ref["tp"], ref["u"], list(map(int, d[i]))) == bool(lv[i]) == True
    # sum, T terms
    gs = s.scalar(B, T)
    xs = s.uniform_q(B, T, P.l)
    rs, ys = s.small(B, T), s.gaussian(B, T)
    sc = o.sum_commit_batch(gs, xs, rp, rs, ys, yp)
    zs, zp = o.sum_respond_batch(ys, yp, rs, rp, d)
    sv = o.sum_verify_batch(zs, zp, sc["cs"], sc["cp"], gs, sc["ts"], sc["tp"], sc["u"], d)
    for i in range(B):
        ref = pyref.sum_commit(ck, PP, _vec(gs[i]), [_vec(xx) for xx in xs[i]], _to_mat(rp[i]),
                               [_to_mat(rr) for rr in rs[i]], [_to_mat(yy) for yy in ys[i]], _to_mat(yp[i]))
        assert ref["xp"] == _vec(sc["xp"][i]) and ref["cp"] == _to_mat(sc["cp"][i])
        assert ref["cs"] == [_to_mat(cc) for cc in sc["cs"][i]]
        assert ref["ts"] == [_vec(tt) for tt in sc["ts"][i]] and ref["tp"] == _vec(sc["tp"][i])
        assert ref["u"] == _to_mat(sc["u"][i])
        rzs, rzp = pyref.sum_respond(PP, [_to_mat(yy) for yy in ys[i]], _to_mat(yp[i]),
                                     [_to_mat(rr) for rr in rs[i]], _to_mat(rp[i]), list(map(int, d[i])))
        assert rzs == [_to_mat(zz) for zz in zs[i]] and rzp == _to_mat(zp[i])
        assert pyref.sum_verify(ck, PP, rzs, rzp, ref["cs"], ref["cp"], _vec(gs[i]), ref["ts"], ref["tp"],
                                ref["u"], list(map(int, d[i]))) == bool(sv[i]) == True


def test_product_counts():
    """6 / 9+3+4 / 26+6+17 / (14T+12)+(3T+3)+(9T+8) ring products per unit."""
    N, T = 16, 4
    P = orc.Params(N=N)
    s = synth.Synth(7, N=N)
    o = orc.Oracle(P, *s.key())
    x, r, rp, y, yp, d, g = s.message(1), s.small(1), s.small(1), s.gaussian(1), s.gaussian(1), s.challenge(1), s.scalar(1)
    o.product_count()
    o.commit(x[0], r[0]);                           assert o.product_count() == 6
    c, t, _ = o.open_commit_batch(x, r, y, nthreads=1); assert o.product_count() == 9
    z = o.open_respond_batch(y, r, d, nthreads=1);  assert o.product_count() == 3
    o.open_verify_batch(z, t, c[:, :1], d, nthreads=1); assert o.product_count() == 4
    lc = o.linear_commit_batch(g, x, rp, r, y, yp, nthreads=1); assert o.product_count() == 26
    lz, lzp = o.linear_respond_batch(y, yp, r, rp, d, nthreads=1); assert o.product_count() == 6
    o.linear_verify_batch(lz, lzp, lc["c"], lc["cp"], g, lc["t"], lc["tp"], lc["u"], d, nthreads=1)
    assert o.product_count() == 17
    gs, xs, rs, ys = s.scalar(1, T), s.uniform_q(1, T, 1), s.small(1, T), s.gaussian(1, T)
    sc = o.sum_commit_batch(gs, xs, rp, rs, ys, yp, nthreads=1); assert o.product_count() == 14 * T + 12
    zs, zp = o.sum_respond_batch(ys, yp, rs, rp, d, nthreads=1); assert o.product_count() == 3 * T + 3
    o.sum_verify_batch(zs, zp, sc["cs"], sc["cp"], gs, sc["ts"], sc["tp"], sc["u"], d, nthreads=1)
    assert o.product_count() == 9 * T + 8


def test_readme_open_flow_n512():
    # README.md:32-55 / commit.rs:66-78 at N=512 with x = [1,2,3,4]
    N = 512
    P = orc.Params(N=N)
    s = synth.Synth(2026, N=N)
    o = orc.Oracle(P, *s.key())
    x = np.zeros((1, 1, N), np.int64); x[0, 0, :4] = [1, 2, 3, 4]
    r, y, d = s.small(1), s.gaussian(1), s.challenge(1)
    c, t, ok = o.open_commit_batch(x, r, y)
    assert ok[0] == 1
    assert o.commitment_verify(c[0], x[0], r[0])                      # commit.rs:77
    z = o.open_respond_batch(y, r, d)
    assert o.open_verify_batch(z, t, c[:, :1], d)[0] == 1             # README.md:54
    # commit.rs:165-170: second message, swapped openings fail
    x2 = np.zeros((1, 1, N), np.int64); x2[0, 0, :4] = [4, 5, 6, 7]
    r2 = s.small(1)
    ok2, c2 = o.commit(x2[0], r2[0])
    assert ok2 and o.commitment_verify(c2, x2[0], r2[0])
    assert not o.commitment_verify(c2, x[0], r[0])
    assert not o.commitment_verify(c[0], x2[0], r2[0])


def test_negative_cases_n16():
    """Negative tests the reference lacks: tampering flips the verify bit."""
    N = 16
    P = orc.Params(N=N)
    s = synth.Synth(99, N=N)
    o = orc.Oracle(P, *s.key())
    B = 8
    x, r, y, d = s.message(B), s.small(B), s.gaussian(B), s.challenge(B)
    c, t, _ = o.open_commit_batch(x, r, y)
    z = o.open_respond_batch(y, r, d)
    assert o.open_verify_batch(z, t, c[:, :1], d).all()
    for name in ("z", "t", "c", "d"):
        zz, tt, cc, dd = z.copy(), t.copy(), c.copy(), d.astype(np.int64).copy()
        {"z": zz, "t": tt, "c": cc, "d": dd}[name][:, ..., 3] += 1
        assert not o.open_verify_batch(zz, tt, cc[:, :1], dd).any(), name
    # oversized z fails the norm check (params.rs:112-118) even if the equation were to hold
    zbig = z.copy(); zbig[:, 0, 0] = o.verify_bound() + 1
    assert not o.open_verify_batch(zbig, t, c[:, :1], d).any()
    # randomised opening f (commit.rs:203-207): f*c == A r' + f*[0;x] holds for r' = f*r when norms allow
    f = np.zeros(N, np.int64); f[1] = 1          # f = X, a unit of the challenge space
    rf = np.stack([o.poly_mul(r[0, j], f) for j in range(P.k)])
    assert o.commitment_verify(c[0], x[0], rf, f)
    assert not o.commitment_verify(c[0], x[0], r[0], f)
