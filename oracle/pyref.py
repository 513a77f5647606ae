"""Pure-Python (big-int) twin of the C oracle, written to read like the reference.

TEST INFRASTRUCTURE ONLY.  It exists so that the C oracle is itself checked by an
independently written restatement (tests/test_oracle_pins.py) and to generate the
fixtures in tests/golden/ (tests/golden/make_golden.py).  Same parity status as
ringzk_oracle.h: "parity unpinned" for ring products (poly-ring-xnp1 0.3 is not in
/root/reference), pinned for everything the reference's own tests pin.

Polynomials are Python lists of N ints (canonical centred residues); a Mat is a
list of rows, each a list of polynomials, exactly like mat.rs:12-17.
"""
from __future__ import annotations

from dataclasses import dataclass
from math import isqrt

import numpy as np


@dataclass(frozen=True)
class Params:
    """params.rs:18-36; `Q` is the const generic of ZqI64<Q>, `N` the ring degree."""
    N: int = 512
    Q: int = 3515337053
    b: int = 1
    n: int = 1
    k: int = 3
    l: int = 1
    kappa: int = 36

    @property
    def q(self):  # Params.q field, params.rs:126
        return self.Q // 2

    # params.rs:94-98
    def standard_deviation(self, deg_n):
        return self.b * (11 * self.kappa) * isqrt(self.k * deg_n)

    # params.rs:102-108
    def check_commit_constraint(self, r):
        bound = 4 * self.standard_deviation(self.N) * isqrt(self.N)
        return all(norm_2(p) <= bound for row in r for p in row)

    # params.rs:112-118
    def check_verify_constraint(self, r):
        bound = 2 * self.standard_deviation(self.N) * isqrt(self.N)
        return all(norm_2(p) <= bound for row in r for p in row)


def center(v, Q):
    half = (Q - 1) // 2
    r = v % Q
    return r - Q if r > half else r


# ---- Polynomial<ZqI64<Q>, N> (crate poly-ring-xnp1) ----
def poly(coeffs, P):
    c = [center(int(v), P.Q) for v in coeffs]
    assert len(c) <= P.N
    return c + [0] * (P.N - len(c))


def p_zero(P):
    return [0] * P.N


def p_one(P):
    return [1] + [0] * (P.N - 1)


def p_add(a, b, P):
    return [center(x + y, P.Q) for x, y in zip(a, b)]


def p_sub(a, b, P):
    return [center(x - y, P.Q) for x, y in zip(a, b)]


def p_mul(a, b, P):
    """Negacyclic product in Z_Q[X]/(X^N+1), exact big-int arithmetic."""
    N = P.N
    full = np.convolve(np.array(a, dtype=object), np.array(b, dtype=object))
    out = [0] * N
    for i, v in enumerate(full):
        if i < N:
            out[i] += int(v)
        else:
            out[i - N] -= int(v)
    return [center(v, P.Q) for v in out]


# polynomial.rs:60-73
def norm_2(p):
    return isqrt(sum(int(c) * int(c) for c in p))


# polynomial.rs:49-57, 76-87 (test-only in the reference)
def norm_1(p):
    return sum(abs(int(c)) for c in p)


def norm_infinity(p):
    return max(abs(int(c)) for c in p)


# ---- Mat (mat.rs) ----
def m_dim(A):
    return len(A), (len(A[0]) if A else 0)


def m_dot(A, B, P):  # mat.rs:95-115
    m, n = m_dim(A)
    n2, p = m_dim(B)
    assert n == n2
    out = [[p_zero(P) for _ in range(p)] for _ in range(m)]
    for i in range(m):
        for j in range(p):
            for k in range(n):
                out[i][j] = p_add(out[i][j], p_mul(A[i][k], B[k][j], P), P)
    return out


def m_add(A, B, P):  # mat.rs:122-140
    assert m_dim(A) == m_dim(B)
    return [[p_add(a, b, P) for a, b in zip(ra, rb)] for ra, rb in zip(A, B)]


def m_sub(A, B, P):  # mat.rs:147-165
    assert m_dim(A) == m_dim(B)
    return [[p_sub(a, b, P) for a, b in zip(ra, rb)] for ra, rb in zip(A, B)]


def m_cmul(A, e, P):  # mat.rs:168-178
    return [[p_mul(a, e, P) for a in row] for row in A]


def m_from_vec(v):  # mat.rs:46-50
    return [[p] for p in v]


def m_to_vec(A):  # mat.rs:56-64
    assert all(len(r) == 1 for r in A)
    return [r[0] for r in A]


def m_split_rows(A, r):  # mat.rs:203-213
    m = len(A)
    assert r <= m
    return A[: m - r], A[m - r:]


# ---- commitment scheme (commit.rs) ----
class CommitmentKey:
    def __init__(self, P, a1p, a2p):
        """a1p: n x (k-n) polys, a2p: l x (k-n-l) polys (commit.rs:40-41, 52-53)."""
        n, k, l = P.n, P.k, P.l
        self.a1 = [[(p_one(P) if j == i else p_zero(P)) for j in range(n)] + list(a1p[i]) for i in range(n)]
        self.a2 = [[p_zero(P) for _ in range(n)] + [(p_one(P) if j == i else p_zero(P)) for j in range(l)]
                   + list(a2p[i]) for i in range(l)]
        assert all(len(r) == k for r in self.a1 + self.a2)

    def commit(self, x, r, P):
        """commit.rs:88-128 with r supplied. x: list of l polys, r: k x 1 Mat."""
        assert P.l == len(x)
        ok = P.check_commit_constraint(r)
        a = self.a1 + self.a2
        z = [[p_zero(P)] for _ in range(P.n)] + m_from_vec(x)
        c = m_add(m_dot(a, r, P), z, P)
        return ok, c


def commitment_verify(c, x, r, f, ck, P):  # commit.rs:173-210
    if not P.check_commit_constraint(r):
        return False
    a = ck.a1 + ck.a2
    z = [[p_zero(P)] for _ in range(P.n)] + m_from_vec(x)
    if f is not None:
        return m_cmul(c, f, P) == m_add(m_dot(a, r, P), m_cmul(z, f, P), P)
    return m_add(m_dot(a, r, P), z, P) == c


def c1_c2(c, P):  # commit.rs:213-218
    return m_split_rows(c, P.n)


# ---- Open proof (prove/open.rs) ----
def open_commit(ck, P, x, r, y):  # open.rs:80-103
    ok, c = ck.commit(x, r, P)
    t = m_to_vec(m_dot(ck.a1, y, P))
    return ok, c, t


def open_respond(P, y, r, d):  # open.rs:107-117
    return m_add(y, m_cmul(r, d, P), P)


def open_verify(ck, P, z, t, c1, d):  # open.rs:162-174
    if not P.check_verify_constraint(z):
        return False
    lhs = m_dot(ck.a1, z, P)
    rhs = m_add(m_from_vec(t), m_cmul(c1, d, P), P)
    return lhs == rhs


# ---- Linear proof (prove/linear.rs) ----
def linear_commit(ck, P, g, x, rp, r, y, yp):  # linear.rs:82-140
    gx = [p_mul(xi, g, P) for xi in x]
    okp, cp = ck.commit(gx, rp, P)
    ok, c = ck.commit(x, r, P)
    t = m_to_vec(m_dot(ck.a1, y, P))
    tp = m_to_vec(m_dot(ck.a1, yp, P))
    u = m_sub(m_cmul(m_dot(ck.a2, y, P), g, P), m_dot(ck.a2, yp, P), P)
    return dict(ok=ok and okp, gx=gx, cp=cp, c=c, t=t, tp=tp, u=u)


def linear_respond(P, y, yp, r, rp, d):  # linear.rs:144-158
    return open_respond(P, y, r, d), open_respond(P, yp, rp, d)


def linear_verify(ck, P, z, zp, c, cp, g, t, tp, u, d):  # linear.rs:213-250
    c1, c2 = c1_c2(c, P)
    c1p, c2p = c1_c2(cp, P)
    if not P.check_verify_constraint(z):
        return False
    if not P.check_verify_constraint(zp):
        return False
    if m_dot(ck.a1, z, P) != m_add(m_from_vec(t), m_cmul(c1, d, P), P):
        return False
    if m_dot(ck.a1, zp, P) != m_add(m_from_vec(tp), m_cmul(c1p, d, P), P):
        return False
    lhs = m_sub(m_cmul(m_dot(ck.a2, z, P), g, P), m_dot(ck.a2, zp, P), P)
    rhs = m_add(m_cmul(m_sub(m_cmul(c2, g, P), c2p, P), d, P), u, P)
    return lhs == rhs


# ---- Sum proof (prove/sum.rs) ----
def _reduce_add(mats, P):
    acc = mats[0]
    for m in mats[1:]:
        acc = m_add(acc, m, P)
    return acc


def sum_commit(ck, P, gs, xs, rp, rs, ys, yp):  # sum.rs:99-178
    assert gs and len(gs) == len(xs)
    xp = m_to_vec(_reduce_add([m_cmul(m_from_vec(x), g, P) for x, g in zip(xs, gs)], P))
    okp, cp = ck.commit(xp, rp, P)
    oks, cs = zip(*[ck.commit(x, r, P) for x, r in zip(xs, rs)])
    ts = [m_to_vec(m_dot(ck.a1, y, P)) for y in ys]
    tp = m_to_vec(m_dot(ck.a1, yp, P))
    u = m_sub(_reduce_add([m_cmul(m_dot(ck.a2, y, P), g, P) for g, y in zip(gs, ys)], P),
              m_dot(ck.a2, yp, P), P)
    return dict(ok=okp and all(oks), xp=xp, cp=cp, cs=list(cs), ts=ts, tp=tp, u=u)


def sum_respond(P, ys, yp, rs, rp, d):  # sum.rs:182-200
    return [open_respond(P, y, r, d) for y, r in zip(ys, rs)], open_respond(P, yp, rp, d)


def sum_verify(ck, P, zs, zp, cs, cp, gs, ts, tp, u, d):  # sum.rs:257-320
    css = [c1_c2(c, P) for c in cs]
    c1p, c2p = c1_c2(cp, P)
    if not all(P.check_verify_constraint(z) for z in zs):
        return False
    if not P.check_verify_constraint(zp):
        return False
    if len(zs) != len(ts) and len(zs) != len(css):  # sum.rs:273 (sic)
        return False
    lhs = [m_dot(ck.a1, z, P) for z in zs]
    rhs = [m_add(m_from_vec(t), m_cmul(c1, d, P), P) for (c1, _), t in zip(css, ts)]
    if lhs != rhs:
        return False
    if m_dot(ck.a1, zp, P) != m_add(m_from_vec(tp), m_cmul(c1p, d, P), P):
        return False
    lhs = m_sub(_reduce_add([m_cmul(m_dot(ck.a2, z, P), g, P) for z, g in zip(zs, gs)], P),
                m_dot(ck.a2, zp, P), P)
    rhs = m_add(m_cmul(m_sub(_reduce_add([m_cmul(c2, g, P) for (_, c2), g in zip(css, gs)], P), c2p, P), d, P),
                u, P)
    return lhs == rhs
