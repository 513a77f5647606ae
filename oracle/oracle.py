"""ctypes loader for the C oracle (oracle/ringzk_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py, never by the product package.
Parity status of the oracle itself: see ringzk_oracle.h ("parity unpinned" for
ring products; the pins that do exist are checked in tests/test_oracle_pins.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libringzk_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with the Makefile next to this file."""
    src = [os.path.join(_HERE, f) for f in ("ringzk_oracle.c", "ringzk_oracle.h", "Makefile")]
    stale = force or not os.path.exists(_LIB) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB) for s in src)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB


class _Params(C.Structure):
    _fields_ = [("q", C.c_int64), ("b", C.c_int64), ("N", C.c_int32), ("n", C.c_int32),
                ("k", C.c_int32), ("l", C.c_int32), ("kappa", C.c_int32)]


@dataclass(frozen=True)
class Params:
    """Params<ZqI64<Q>> + const generic N (params.rs:18-36, 121-138)."""
    N: int = 512
    q: int = 3515337053
    b: int = 1
    n: int = 1
    k: int = 3
    l: int = 1
    kappa: int = 36

    def c(self) -> _Params:
        return _Params(self.q, self.b, self.N, self.n, self.k, self.l, self.kappa)


_lib = None
_I64P = C.POINTER(C.c_int64)
_U8P = C.POINTER(C.c_uint8)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.rzko_sigma.restype = C.c_uint64
        _lib.rzko_commit_bound.restype = C.c_uint64
        _lib.rzko_verify_bound.restype = C.c_uint64
        _lib.rzko_norm2.restype = C.c_uint64
        _lib.rzko_product_count.restype = C.c_uint64
        _lib.rzko_center.restype = C.c_int64
        _lib.rzko_center.argtypes = [C.c_int64, C.c_int64]
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(_I64P)


def _u8(a):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_U8P)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


class Oracle:
    """Thin object wrapper: one set of params + expanded key."""

    def __init__(self, params: Params, a1p=None, a2p=None):
        self.P = params
        self._c = params.c()
        self.L = lib()
        N, n, k, l = params.N, params.n, params.k, params.l
        self.a1 = np.zeros((n, k, N), np.int64)
        self.a2 = np.zeros((l, k, N), np.int64)
        if a1p is not None:
            self.set_key(a1p, a2p)

    # ---- scalars ----
    def sigma(self):
        return int(self.L.rzko_sigma(C.byref(self._c)))

    def commit_bound(self):
        return int(self.L.rzko_commit_bound(C.byref(self._c)))

    def verify_bound(self):
        return int(self.L.rzko_verify_bound(C.byref(self._c)))

    def norm2(self, poly):
        return int(self.L.rzko_norm2(C.byref(self._c), _p(i64(poly))))

    def center(self, a):
        a = i64(a)
        q = self.P.q
        half = (q - 1) // 2
        r = np.fmod(a, q)
        r = np.where(r > half, r - q, r)
        r = np.where(r < -half, r + q, r)
        return r

    def product_count(self, reset=True):
        return int(self.L.rzko_product_count(1 if reset else 0))

    # ---- ring / matrix ops ----
    def poly_mul(self, a, b):
        out = np.zeros(self.P.N, np.int64)
        self.L.rzko_poly_mul(C.byref(self._c), _p(i64(a)), _p(i64(b)), _p(out))
        return out

    def mat_dot(self, A, B):
        A, B = i64(A), i64(B)
        m, n, _ = A.shape
        n2, p, _ = B.shape
        assert n == n2
        out = np.zeros((m, p, self.P.N), np.int64)
        self.L.rzko_mat_dot(C.byref(self._c), m, n, p, _p(A), _p(B), _p(out))
        return out

    def mat_add(self, A, B):
        A, B = i64(A), i64(B)
        out = np.zeros_like(A)
        self.L.rzko_mat_add(C.byref(self._c), A.shape[0], A.shape[1], _p(A), _p(B), _p(out))
        return out

    def mat_sub(self, A, B):
        A, B = i64(A), i64(B)
        out = np.zeros_like(A)
        self.L.rzko_mat_sub(C.byref(self._c), A.shape[0], A.shape[1], _p(A), _p(B), _p(out))
        return out

    def mat_cmul(self, A, e):
        A = i64(A)
        out = np.zeros_like(A)
        self.L.rzko_mat_cmul(C.byref(self._c), A.shape[0], A.shape[1], _p(A), _p(i64(e)), _p(out))
        return out

    # ---- key ----
    def set_key(self, a1p, a2p):
        """a1p [n][k-n][N], a2p [l][k-n-l][N] random blocks (commit.rs:40-41, 52-53)."""
        P = self.P
        a1p = i64(a1p).reshape(P.n, P.k - P.n, P.N)
        a2p = i64(a2p).reshape(P.l, P.k - P.n - P.l, P.N)
        self.L.rzko_key_expand(C.byref(self._c), _p(a1p), _p(a2p), _p(self.a1), _p(self.a2))

    # ---- single-item API ----
    def commit(self, x, r):
        P = self.P
        c = np.zeros((P.n + P.l, P.N), np.int64)
        ok = self.L.rzko_commit(C.byref(self._c), _p(self.a1), _p(self.a2), _p(i64(x)), _p(i64(r)), _p(c))
        return bool(ok), c

    def commitment_verify(self, c, x, r, f=None):
        return bool(self.L.rzko_commitment_verify(C.byref(self._c), _p(self.a1), _p(self.a2),
                                                  _p(i64(c)), _p(i64(x)), _p(i64(r)),
                                                  _p(i64(f)) if f is not None else None))

    # ---- batch API (arrays [B][...]) ----
    def commit_batch(self, x, r, nthreads=0):
        P = self.P
        x, r = i64(x), i64(r)
        B = x.shape[0]
        c = np.zeros((B, P.n + P.l, P.N), np.int64)
        ok = np.zeros(B, np.uint8)
        self.L.rzko_commit_batch(C.byref(self._c), _p(self.a1), _p(self.a2), C.c_size_t(B),
                                 _p(x), _p(r), _p(c), _u8(ok), nthreads)
        return c, ok

    def open_commit_batch(self, x, r, y, nthreads=0):
        P = self.P
        x, r, y = i64(x), i64(r), i64(y)
        B = x.shape[0]
        c = np.zeros((B, P.n + P.l, P.N), np.int64)
        t = np.zeros((B, P.n, P.N), np.int64)
        ok = np.zeros(B, np.uint8)
        self.L.rzko_open_commit_batch(C.byref(self._c), _p(self.a1), _p(self.a2), C.c_size_t(B),
                                      _p(x), _p(r), _p(y), _p(c), _p(t), _u8(ok), nthreads)
        return c, t, ok

    def open_respond_batch(self, y, r, d, nthreads=0):
        y, r, d = i64(y), i64(r), i64(d)
        B = y.shape[0]
        z = np.zeros_like(y)
        self.L.rzko_open_respond_batch(C.byref(self._c), C.c_size_t(B), _p(y), _p(r), _p(d), _p(z), nthreads)
        return z

    def open_verify_batch(self, z, t, c1, d, nthreads=0):
        z, t, c1, d = i64(z), i64(t), i64(c1), i64(d)
        B = z.shape[0]
        ok = np.zeros(B, np.uint8)
        self.L.rzko_open_verify_batch(C.byref(self._c), _p(self.a1), C.c_size_t(B),
                                      _p(z), _p(t), _p(c1), _p(d), _u8(ok), nthreads)
        return ok

    def linear_commit_batch(self, g, x, rp, r, y, yp, nthreads=0):
        P = self.P
        g, x, rp, r, y, yp = map(i64, (g, x, rp, r, y, yp))
        B = x.shape[0]
        gx = np.zeros((B, P.l, P.N), np.int64)
        cp = np.zeros((B, P.n + P.l, P.N), np.int64)
        c = np.zeros_like(cp)
        t = np.zeros((B, P.n, P.N), np.int64)
        tp = np.zeros_like(t)
        u = np.zeros((B, P.l, P.N), np.int64)
        ok = np.zeros(B, np.uint8)
        self.L.rzko_linear_commit_batch(C.byref(self._c), _p(self.a1), _p(self.a2), C.c_size_t(B),
                                        _p(g), _p(x), _p(rp), _p(r), _p(y), _p(yp),
                                        _p(gx), _p(cp), _p(c), _p(t), _p(tp), _p(u), _u8(ok), nthreads)
        return dict(gx=gx, cp=cp, c=c, t=t, tp=tp, u=u, ok=ok)

    def linear_respond_batch(self, y, yp, r, rp, d, nthreads=0):
        y, yp, r, rp, d = map(i64, (y, yp, r, rp, d))
        B = y.shape[0]
        z, zp = np.zeros_like(y), np.zeros_like(yp)
        self.L.rzko_linear_respond_batch(C.byref(self._c), C.c_size_t(B), _p(y), _p(yp), _p(r), _p(rp),
                                         _p(d), _p(z), _p(zp), nthreads)
        return z, zp

    def linear_verify_batch(self, z, zp, c, cp, g, t, tp, u, d, nthreads=0):
        z, zp, c, cp, g, t, tp, u, d = map(i64, (z, zp, c, cp, g, t, tp, u, d))
        B = z.shape[0]
        ok = np.zeros(B, np.uint8)
        self.L.rzko_linear_verify_batch(C.byref(self._c), _p(self.a1), _p(self.a2), C.c_size_t(B),
                                        _p(z), _p(zp), _p(c), _p(cp), _p(g), _p(t), _p(tp), _p(u), _p(d),
                                        _u8(ok), nthreads)
        return ok

    def sum_commit_batch(self, gs, xs, rp, rs, ys, yp, nthreads=0):
        P = self.P
        gs, xs, rp, rs, ys, yp = map(i64, (gs, xs, rp, rs, ys, yp))
        B, T = gs.shape[0], gs.shape[1]
        xp = np.zeros((B, P.l, P.N), np.int64)
        cp = np.zeros((B, P.n + P.l, P.N), np.int64)
        cs = np.zeros((B, T, P.n + P.l, P.N), np.int64)
        ts = np.zeros((B, T, P.n, P.N), np.int64)
        tp = np.zeros((B, P.n, P.N), np.int64)
        u = np.zeros((B, P.l, P.N), np.int64)
        ok = np.zeros(B, np.uint8)
        self.L.rzko_sum_commit_batch(C.byref(self._c), _p(self.a1), _p(self.a2), C.c_size_t(B), T,
                                     _p(gs), _p(xs), _p(rp), _p(rs), _p(ys), _p(yp),
                                     _p(xp), _p(cp), _p(cs), _p(ts), _p(tp), _p(u), _u8(ok), nthreads)
        return dict(xp=xp, cp=cp, cs=cs, ts=ts, tp=tp, u=u, ok=ok)

    def sum_respond_batch(self, ys, yp, rs, rp, d, nthreads=0):
        ys, yp, rs, rp, d = map(i64, (ys, yp, rs, rp, d))
        B, T = ys.shape[0], ys.shape[1]
        zs, zp = np.zeros_like(ys), np.zeros_like(yp)
        self.L.rzko_sum_respond_batch(C.byref(self._c), C.c_size_t(B), T, _p(ys), _p(yp), _p(rs), _p(rp),
                                      _p(d), _p(zs), _p(zp), nthreads)
        return zs, zp

    def sum_verify_batch(self, zs, zp, cs, cp, gs, ts, tp, u, d, nthreads=0):
        zs, zp, cs, cp, gs, ts, tp, u, d = map(i64, (zs, zp, cs, cp, gs, ts, tp, u, d))
        B, T = zs.shape[0], zs.shape[1]
        ok = np.zeros(B, np.uint8)
        self.L.rzko_sum_verify_batch(C.byref(self._c), _p(self.a1), _p(self.a2), C.c_size_t(B), T,
                                     _p(zs), _p(zp), _p(cs), _p(cp), _p(gs), _p(ts), _p(tp), _p(u), _p(d),
                                     _u8(ok), nthreads)
        return ok


def max_threads() -> int:
    return int(lib().rzko_max_threads())
