"""Host-side mirror of the reference's public API (ring_zk::*, /root/reference/src/lib.rs:5-24) on top
of the CUDA engine.  Same names, argument meaning and error behaviour as the Rust crate, with
`*_batch` entry points added alongside; every ring operation runs on the GPU through the C ABI
(ring-zk_b200/engine.py -> libringzk_b200.so).  The reference toolchain (Rust) is absent from this
image, so this mirror is what the parity tests drive; INTEGRATION.md shows the equivalent Rust shim.

Conventions
  * a polynomial is a numpy int32 array of N canonical centred coefficients; batches add a leading axis;
  * `rng` is a numpy Generator (the reference takes `&mut impl RngExt`); all randomness (r, y, d) is drawn
    here on the host with the reference's distributions and handed to the engine as inputs;
  * shape errors raise AssertionError where the reference panics through assert!
    (params.rs:71, commit.rs:95, sum.rs:105); verification returns bool.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from math import isqrt

import numpy as np

from . import engine as _engine

Q_DEFAULT = 3515337053


def _poly(coeffs, N, Q):
    """Polynomial::from_coeffs (params.rs:75,90): zero-padded, canonical centred residues."""
    c = np.asarray(list(coeffs), dtype=np.int64)
    assert c.size <= N, "more than N coefficients"
    half = (Q - 1) // 2
    c = np.fmod(c, Q)
    c = np.where(c > half, c - Q, c)
    c = np.where(c < -half, c + Q, c)
    out = np.zeros(N, np.int32)
    out[: c.size] = c
    return out


@dataclass
class Params:
    """params.rs:18-36.  `q` is the field of the reference (Q // 2); `Q` the modulus of ZqI64<Q>."""
    q: int = Q_DEFAULT // 2
    b: int = 1
    n: int = 1
    k: int = 3
    l: int = 1
    kappa: int = 36
    Q: int = Q_DEFAULT

    @staticmethod
    def default() -> "Params":  # params.rs:121-138
        return Params()

    def standard_deviation(self, deg_n: int) -> int:  # params.rs:94-98
        return self.b * (11 * self.kappa) * isqrt(self.k * deg_n)

    def generate_commitment_key(self, rng, N: int = 512) -> "CommitmentKey":  # params.rs:49-54
        return CommitmentKey.new(rng, self, N)

    def prepare_value(self, value, N: int = 512):  # params.rs:67-77
        assert len(value) == self.l
        return [_poly(v, N, self.Q) for v in value]

    def prepare_scalar(self, scalar, N: int = 512):  # params.rs:89-91
        return _poly(scalar, N, self.Q)

    # ---- samplers (polynomial.rs:14-44, challenge_space.rs:12-33), host side ----
    def sample_small(self, rng, shape, N):
        return rng.integers(-self.b, self.b + 1, size=tuple(shape) + (self.k, N)).astype(np.int8)

    def sample_gaussian(self, rng, shape, N):
        sigma = float(self.standard_deviation(N))
        return np.trunc(rng.normal(0.0, sigma, size=tuple(shape) + (self.k, N))).astype(np.int32)

    def sample_challenge(self, rng, B, N):
        nnz = min(self.kappa, N)
        d = np.zeros((B, N), np.int8)
        signs = (rng.integers(0, 2, size=(B, nnz)) * 2 - 1).astype(np.int8)
        pos = np.argsort(rng.random((B, N)), axis=1)[:, :nnz]
        np.put_along_axis(d, pos, signs, axis=1)
        return d


@dataclass
class Opening:  # commit.rs:223-235
    x: np.ndarray          # [l][N]
    r: np.ndarray          # [k][N] int8
    f: np.ndarray | None = None


@dataclass
class Commitment:  # commit.rs:135-141
    c: np.ndarray          # [(n+l)][N]

    def verify(self, opening: Opening, ck: "CommitmentKey", params: Params) -> bool:  # commit.rs:173-210
        f = None if opening.f is None else np.ascontiguousarray(opening.f, np.int8)[None]
        bm = ck.engine.commitment_verify(np.ascontiguousarray(self.c, np.int32)[None], opening.x[None].astype(np.int32),
                                         opening.r[None].astype(np.int8), f)
        return bool(_engine.unpack_bitmap(bm, 1)[0])

    @staticmethod
    def verify_batch(C, X, R, ck: "CommitmentKey", F=None):
        """Batched Commitment::verify: C [B][n+l][N], X [B][l][N], R [B][k][N] int8, F None or [B][N] int8."""
        B = C.shape[0]
        return _engine.unpack_bitmap(ck.engine.commitment_verify(np.ascontiguousarray(C, np.int32), np.ascontiguousarray(X, np.int32),
                                                                 np.ascontiguousarray(R, np.int8),
                                                                 None if F is None else np.ascontiguousarray(F, np.int8)), B)

    def c1_c2(self, params: Params):  # commit.rs:213-218
        m = self.c.shape[0]
        return self.c[: m - params.n], self.c[m - params.n:]


class CommitmentKey:  # commit.rs:19-60
    def __init__(self, a1, a2, params: Params, N: int, device: int = -1):
        self.a1, self.a2, self.N = np.asarray(a1, np.int64), np.asarray(a2, np.int64), N
        # the engine's sigma / norm bounds / exactness limits come from THESE parameters (not from Params::default())
        rp = _engine.RzkParams(params.Q, params.b, N, params.n, params.k, params.l, params.kappa)
        self.engine = _engine.Engine(N=N, device=device, params=rp)
        assert self.engine.sigma() == params.standard_deviation(N)
        self.engine.set_key(self.a1, self.a2)

    @staticmethod
    def new(rng, params: Params, N: int, device: int = -1) -> "CommitmentKey":
        n, k, l = params.n, params.k, params.l
        a1 = np.zeros((n, k, N), np.int64)
        a2 = np.zeros((l, k, N), np.int64)
        for i in range(n):
            a1[i, i, 0] = 1                                                   # diag(n, n, one)  commit.rs:39
            a1[i, n:] = rng.integers(-params.q, params.q + 1, size=(k - n, N))     # commit.rs:40-41
        for i in range(l):
            a2[i, n + i, 0] = 1                                               # commit.rs:50-51
            a2[i, n + l:] = rng.integers(-params.q, params.q + 1, size=(k - n - l, N))   # commit.rs:52-53
        return CommitmentKey(a1, a2, params, N, device)

    # CommitmentKey::commit (commit.rs:88-128) for a batch X [B][l][N]
    def commit_batch(self, rng, X, params: Params):
        X = np.ascontiguousarray(X, dtype=np.int32)
        assert X.ndim == 3 and X.shape[1] == params.l, "x.len() != l"          # commit.rs:95
        B = X.shape[0]
        r = params.sample_small(rng, (B,), self.N)
        c, ok = self.engine.commit(X, r)
        okb = _engine.unpack_bitmap(ok, B)
        while not okb.all():                                                   # commit.rs:98-107 redraw loop
            bad = np.nonzero(~okb)[0]
            r[bad] = params.sample_small(rng, (len(bad),), self.N)
            c2, ok2 = self.engine.commit(np.ascontiguousarray(X[bad]), np.ascontiguousarray(r[bad]))
            c[bad] = c2
            okb[bad] = _engine.unpack_bitmap(ok2, len(bad))
        return r, c

    def commit(self, rng, x, params: Params):
        assert len(x) == params.l
        r, c = self.commit_batch(rng, np.stack(x)[None], params)
        return Opening(np.stack(x), r[0], None), Commitment(c[0])


# --------------------------------------------------------------------------- Fiat-Shamir transcript prefix (docs/FIAT_SHAMIR.md)

def fs_key_digest(ck: "CommitmentKey") -> bytes:
    """SHAKE128-256 of the key's random blocks a11 || a12 || a22 as int32 LE canonical centred coefficients"""
    import hashlib
    h = hashlib.shake_128()
    half = (Q_DEFAULT - 1) // 2
    for p in (ck.a1[0, 1], ck.a1[0, 2], ck.a2[0, 2]):
        c = np.fmod(np.asarray(p, np.int64), Q_DEFAULT)
        c = np.where(c > half, c - Q_DEFAULT, np.where(c < -half, c + Q_DEFAULT, c))
        h.update(c.astype("<i4").tobytes())
    return h.digest(32)


def fs_prefix(tag: str, ck: "CommitmentKey", params: Params, T: int = 0, session: bytes = b"") -> bytes:
    """tag (32 bytes, zero padded) || key digest (32) || q u64 || N u32 || kappa u32 || T u32 || b u32 || session (8-byte aligned)"""
    import struct
    t = tag.encode()
    assert len(t) <= 32 and len(session) % 8 == 0
    return t.ljust(32, b"\0") + fs_key_digest(ck) + struct.pack("<QIIII", params.Q, ck.N, params.kappa, T, params.b) + session


FS_TAG_OPEN, FS_TAG_LINEAR, FS_TAG_SUM = "ring-zk/fs/open/v1", "ring-zk/fs/linear/v1", "ring-zk/fs/sum/v1"


# --------------------------------------------------------------------------- Open proof (prove/open.rs)

@dataclass
class OpenProofResponseContext:
    opening: Opening
    y: np.ndarray


@dataclass
class OpenProofCommitment:
    c: Commitment
    t: np.ndarray


@dataclass
class OpenProofVerificationContext:
    c1: np.ndarray
    t: np.ndarray
    d: np.ndarray


@dataclass
class OpenProofChallenge:
    d: np.ndarray


@dataclass
class OpenProofResponse:
    z: np.ndarray


class OpenProofProver:
    def __init__(self, ck: CommitmentKey, params: Params):  # open.rs:69
        self.ck, self.params = ck, params

    def commit_batch(self, rng, X):
        """open.rs:80-103 for X [B][l][N] -> dict(x, r, y, c, t)"""
        P, N = self.params, self.ck.N
        X = np.ascontiguousarray(X, dtype=np.int32)
        assert X.shape[1] == P.l
        B = X.shape[0]
        r = P.sample_small(rng, (B,), N)                     # ck.commit draws r first (open.rs:85)
        y = P.sample_gaussian(rng, (B,), N)                  # then y (open.rs:88-94)
        c, t, ok = self.ck.engine.open_commit(X, r, y)
        assert _engine.unpack_bitmap(ok, B).all(), "commit constraint failed (redraw r)"
        return dict(x=X, r=r, y=y, c=c, t=t)

    def create_response_batch(self, y, r, d):
        """open.rs:107-117"""
        return self.ck.engine.open_respond(np.ascontiguousarray(y), np.ascontiguousarray(r), np.ascontiguousarray(d))

    def prove_batch_fs(self, rng, X, session: bytes = b""):
        """Non-interactive Open proofs (docs/FIAT_SHAMIR.md; NOT in the reference): commit -> challenge from the transcript ->
        response, chained on the device without a host round trip.  Returns dict(x, r, y) (the openings, host) and the
        proofs dict(c, t, z) as device tensors (d is recomputed by the verifier)."""
        import torch
        P, N, e = self.params, self.ck.N, self.ck.engine
        X = np.ascontiguousarray(X, dtype=np.int32)
        assert X.shape[1] == P.l
        B = X.shape[0]
        r = P.sample_small(rng, (B,), N)
        y = P.sample_gaussian(rng, (B,), N)
        dev = torch.device("cuda", e.device)
        T = lambda a: torch.from_numpy(a).to(dev)
        E = lambda *sh, dt=torch.int32: torch.empty(sh, dtype=dt, device=dev)
        c, t, z, d = E(B, 2, N), E(B, 1, N), E(B, 3, N), E(B, N, dt=torch.int8)
        flags = torch.zeros(B, dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        e.open_prove_fs(T(X), T(r), T(y), fs_prefix(FS_TAG_OPEN, self.ck, P, session=session), c, t, d, z, flags, stream=st)
        assert not flags.any(), "commit constraint / range check failed (redraw r, y)"
        return dict(x=X, r=r, y=y), dict(c=c, t=t, z=z)

    def commit(self, rng, x):
        s = self.commit_batch(rng, np.stack(x)[None])
        return (OpenProofResponseContext(Opening(s["x"][0], s["r"][0]), s["y"][0]),
                OpenProofCommitment(Commitment(s["c"][0]), s["t"][0]))

    def create_response(self, context: OpenProofResponseContext, challenge: OpenProofChallenge):
        z = self.create_response_batch(context.y[None], context.opening.r[None], challenge.d[None])
        return OpenProofResponse(z[0])


class OpenProofVerifier:
    def __init__(self, ck: CommitmentKey, params: Params):  # open.rs:135
        self.ck, self.params = ck, params

    def generate_challenge_batch(self, rng, B):
        return self.params.sample_challenge(rng, B, self.ck.N)           # open.rs:148

    def verify_batch(self, z, t, c1, d):
        """open.rs:162-174 -> bool[B]"""
        B = z.shape[0]
        bm = self.ck.engine.open_verify(np.ascontiguousarray(z), np.ascontiguousarray(t),
                                        np.ascontiguousarray(c1), np.ascontiguousarray(d))
        return _engine.unpack_bitmap(bm, B)

    def verify_batch_fs(self, proofs, session: bytes = b""):
        """Verifies non-interactive Open proofs dict(c, t, z) (device tensors): d from the transcript, then open.rs:162-174."""
        import torch
        e = self.ck.engine
        c, t, z = proofs["c"], proofs["t"], proofs["z"]
        B = c.shape[0]
        d = torch.empty((B, self.ck.N), dtype=torch.int8, device=c.device)
        flags = torch.zeros(B, dtype=torch.int32, device=c.device)
        e.open_verify_fs(c, t, z, fs_prefix(FS_TAG_OPEN, self.ck, self.params, session=session), d, flags,
                         stream=torch.cuda.current_stream(c.device).cuda_stream)
        return (flags & 1).eq(0).cpu().numpy()

    def generate_challenge(self, rng, commitment: OpenProofCommitment):
        d = self.generate_challenge_batch(rng, 1)[0]
        c1, _ = commitment.c.c1_c2(self.params)
        return OpenProofVerificationContext(c1, commitment.t, d.copy()), OpenProofChallenge(d)

    def verify(self, response: OpenProofResponse, context: OpenProofVerificationContext) -> bool:
        return bool(self.verify_batch(response.z[None], context.t[None], context.c1[None], context.d[None])[0])


# --------------------------------------------------------------------------- Linear proof (prove/linear.rs)

@dataclass
class LinearProofResponseContext:
    opening: Opening
    opening_p: Opening
    y: np.ndarray
    yp: np.ndarray


@dataclass
class LinearProofCommitment:
    c: Commitment
    cp: Commitment
    g: np.ndarray
    t: np.ndarray
    tp: np.ndarray
    u: np.ndarray


@dataclass
class LinearProofVerificationContext:
    c: np.ndarray
    cp: np.ndarray
    g: np.ndarray
    t: np.ndarray
    tp: np.ndarray
    u: np.ndarray
    d: np.ndarray


@dataclass
class LinearProofChallenge:
    d: np.ndarray


@dataclass
class LinearProofResponse:
    z: np.ndarray
    zp: np.ndarray


class LinearProofProver:
    def __init__(self, ck: CommitmentKey, params: Params):  # linear.rs:71
        self.ck, self.params = ck, params

    def commit_batch(self, rng, G, X):
        """linear.rs:82-140 for G [B][N], X [B][l][N]"""
        P, N = self.params, self.ck.N
        G = np.ascontiguousarray(G, dtype=np.int32)
        X = np.ascontiguousarray(X, dtype=np.int32)
        assert X.shape[1] == P.l
        B = X.shape[0]
        rp = P.sample_small(rng, (B,), N)       # draw order of the reference: r', r, y, y' (linear.rs:96-115)
        r = P.sample_small(rng, (B,), N)
        y = P.sample_gaussian(rng, (B,), N)
        yp = P.sample_gaussian(rng, (B,), N)
        o = self.ck.engine.linear_commit(G, X, rp, r, y, yp)
        assert _engine.unpack_bitmap(o["ok"], B).all()
        o.update(g=G, x=X, r=r, rp=rp, y=y, yp=yp)
        return o

    def create_response_batch(self, y, yp, r, rp, d):
        """linear.rs:144-158"""
        return self.ck.engine.linear_respond(*(np.ascontiguousarray(a) for a in (y, yp, r, rp, d)))

    def commit(self, rng, g, x):
        s = self.commit_batch(rng, g[None], np.stack(x)[None])
        return (LinearProofResponseContext(Opening(s["x"][0], s["r"][0]), Opening(s["gx"][0], s["rp"][0]), s["y"][0], s["yp"][0]),
                LinearProofCommitment(Commitment(s["c"][0]), Commitment(s["cp"][0]), g, s["t"][0], s["tp"][0], s["u"][0]))

    def create_response(self, context: LinearProofResponseContext, challenge: LinearProofChallenge):
        z, zp = self.create_response_batch(context.y[None], context.yp[None], context.opening.r[None],
                                           context.opening_p.r[None], challenge.d[None])
        return LinearProofResponse(z[0], zp[0])


class LinearProofVerifier:
    def __init__(self, ck: CommitmentKey, params: Params):  # linear.rs:176
        self.ck, self.params = ck, params

    def verify_batch(self, z, zp, c, cp, g, t, tp, u, d):
        """linear.rs:213-250 -> bool[B]"""
        B = z.shape[0]
        bm = self.ck.engine.linear_verify(*(np.ascontiguousarray(a) for a in (z, zp, c, cp, g, t, tp, u, d)))
        return _engine.unpack_bitmap(bm, B)

    def generate_challenge(self, rng, commitment: LinearProofCommitment):
        d = self.params.sample_challenge(rng, 1, self.ck.N)[0]              # linear.rs:192
        return (LinearProofVerificationContext(commitment.c.c, commitment.cp.c, commitment.g, commitment.t,
                                               commitment.tp, commitment.u, d.copy()), LinearProofChallenge(d))

    def verify(self, response: LinearProofResponse, ctx: LinearProofVerificationContext) -> bool:
        return bool(self.verify_batch(response.z[None], response.zp[None], ctx.c[None], ctx.cp[None], ctx.g[None],
                                      ctx.t[None], ctx.tp[None], ctx.u[None], ctx.d[None])[0])


# --------------------------------------------------------------------------- Sum proof (prove/sum.rs)

@dataclass
class SumProofResponseContext:
    openings: list
    opening_p: Opening
    yp: np.ndarray
    ys: np.ndarray


@dataclass
class SumProofCommitment:
    cp: Commitment
    cs: list
    gs: np.ndarray
    tp: np.ndarray
    ts: np.ndarray
    u: np.ndarray


@dataclass
class SumProofVerificationContext:
    cp: np.ndarray
    cs: np.ndarray
    gs: np.ndarray
    ts: np.ndarray
    tp: np.ndarray
    u: np.ndarray
    d: np.ndarray


@dataclass
class SumProofChallenge:
    d: np.ndarray


@dataclass
class SumProofResponse:
    zp: np.ndarray
    zs: np.ndarray


class SumProofProver:
    def __init__(self, ck: CommitmentKey, params: Params):  # sum.rs:84
        self.ck, self.params = ck, params

    def commit_batch(self, rng, GS, XS):
        """sum.rs:99-178 for GS [B][T][N], XS [B][T][l][N]"""
        P, N = self.params, self.ck.N
        GS = np.ascontiguousarray(GS, dtype=np.int32)
        XS = np.ascontiguousarray(XS, dtype=np.int32)
        assert GS.ndim == 3 and GS.shape[1] > 0 and GS.shape[:2] == XS.shape[:2], "gs empty or gs.len() != xs.len()"  # sum.rs:105
        assert XS.shape[2] == P.l
        B, T = GS.shape[:2]
        rp = P.sample_small(rng, (B,), N)       # draw order: r', r_0.., y_0.., y' (sum.rs:116-142)
        rs = P.sample_small(rng, (B, T), N)
        ys = P.sample_gaussian(rng, (B, T), N)
        yp = P.sample_gaussian(rng, (B,), N)
        o = self.ck.engine.sum_commit(GS, XS, rp, rs, ys, yp)
        assert _engine.unpack_bitmap(o["ok"], B).all()
        o.update(gs=GS, xs=XS, rp=rp, rs=rs, ys=ys, yp=yp)
        return o

    def create_response_batch(self, ys, yp, rs, rp, d):
        """sum.rs:182-200"""
        return self.ck.engine.sum_respond(*(np.ascontiguousarray(a) for a in (ys, yp, rs, rp, d)))

    def commit(self, rng, gs, xs):
        assert len(gs) > 0 and len(gs) == len(xs)
        s = self.commit_batch(rng, np.stack(gs)[None], np.stack([np.stack(x) for x in xs])[None])
        T = len(gs)
        ctx = SumProofResponseContext([Opening(s["xs"][0, i], s["rs"][0, i]) for i in range(T)],
                                      Opening(s["xp"][0], s["rp"][0]), s["yp"][0], s["ys"][0])
        com = SumProofCommitment(Commitment(s["cp"][0]), [Commitment(s["cs"][0, i]) for i in range(T)], s["gs"][0],
                                 s["tp"][0], s["ts"][0], s["u"][0])
        return ctx, com

    def create_response(self, context: SumProofResponseContext, challenge: SumProofChallenge):
        rs = np.stack([o.r for o in context.openings])
        zs, zp = self.create_response_batch(context.ys[None], context.yp[None], rs[None], context.opening_p.r[None],
                                            challenge.d[None])
        return SumProofResponse(zp[0], zs[0])


class SumProofVerifier:
    def __init__(self, ck: CommitmentKey, params: Params):  # sum.rs:219
        self.ck, self.params = ck, params

    def verify_batch(self, zs, zp, cs, cp, gs, ts, tp, u, d):
        """sum.rs:257-320 -> bool[B]"""
        B = zs.shape[0]
        if zs.shape[1] != ts.shape[1] or zs.shape[1] != cs.shape[1]:       # sum.rs:273 / Vec inequality at sum.rs:289
            return np.zeros(B, bool)
        bm = self.ck.engine.sum_verify(*(np.ascontiguousarray(a) for a in (zs, zp, cs, cp, gs, ts, tp, u, d)))
        return _engine.unpack_bitmap(bm, B)

    def generate_challenge(self, rng, commitment: SumProofCommitment):
        d = self.params.sample_challenge(rng, 1, self.ck.N)[0]              # sum.rs:233
        cs = np.stack([c.c for c in commitment.cs])
        return (SumProofVerificationContext(commitment.cp.c, cs, commitment.gs, commitment.ts, commitment.tp,
                                            commitment.u, d.copy()), SumProofChallenge(d))

    def verify(self, response: SumProofResponse, ctx: SumProofVerificationContext) -> bool:
        return bool(self.verify_batch(response.zs[None], response.zp[None], ctx.cs[None], ctx.cp[None], ctx.gs[None],
                                      ctx.ts[None], ctx.tp[None], ctx.u[None], ctx.d[None])[0])
