"""CPU restatement of the reference's wire format -- TEST INFRASTRUCTURE (checker for ring-zk_b200's device packer, never shipped
or measured): what `bincode::serialize` (bincode 1.3.3, /root/reference/Cargo.toml dev-dependencies; default options: little
endian, fixed-width integers, u64 sequence lengths, u8 Option tags, struct fields in declaration order) produces for the
derive(Serialize) message structs of the crate, written field by field from the struct definitions:

    Mat { polynomials: Vec<Vec<Polynomial>> }                      /root/reference/src/mat.rs:11-17
    Commitment { c: Mat }                                          src/commit.rs:134-141
    Opening { x: Vec<Polynomial>, r: Mat, f: Option<Polynomial> }  src/commit.rs:222-235
    OpenProofCommitment { c: Commitment, t: Vec<Polynomial> }      src/prove/open.rs:190-198
    *ProofChallenge { d: Polynomial }                              src/prove/open.rs:213-219, linear.rs:309-315, sum.rs:375-381
    OpenProofResponse { z: Mat }                                   src/prove/open.rs:222-228
    LinearProofCommitment { c, cp, g, t, tp, u: Mat }              src/prove/linear.rs:271-285
    LinearProofResponse { z: Mat, zp: Mat }                        src/prove/linear.rs:318-325 (NO Serialize derive in the reference)
    SumProofCommitment { cp, cs: Vec<Commitment>, gs, tp, ts: Vec<Vec<Polynomial>>, u }   src/prove/sum.rs:342-355
    SumProofResponse { zp: Mat, zs: Vec<Mat> }                     src/prove/sum.rs:384-391

Pinned by the reference's own test (src/mat.rs:424-438): the 1 x 1 Mat of Polynomial::<i32, N>::new(vec![1, 2, 3]) is
8 + 8 + (8 + 3 * 4) = 36 bytes.  UNVERIFIED (the serde impls of Polynomial and ZqI64 live in poly-ring-xnp1, absent here):
whether trailing zero coefficients are stored (`trim`) and the width of a ZqI64 coefficient (`elem_bytes`, 8 = i64)."""
import struct

import numpy as np


def _u64(v):
    return struct.pack("<Q", int(v))


def poly(p, elem_bytes=8, trim=True):
    p = np.asarray(p).astype(np.int64)
    n = p.size
    if trim:
        nz = np.nonzero(p)[0]
        n = int(nz[-1]) + 1 if nz.size else 0
    body = p[:n].astype("<i8" if elem_bytes == 8 else "<i4").tobytes()
    return _u64(n) + body


def vec(ps, **kw):
    return _u64(len(ps)) + b"".join(poly(p, **kw) for p in ps)


def mat(rows, **kw):                      # rows x 1 matrix, as every Mat of the protocol messages is
    return _u64(len(rows)) + b"".join(_u64(1) + poly(p, **kw) for p in rows)


def commitment(c, **kw):
    return mat(c, **kw)


def opening(x, r, f=None, **kw):
    return vec(x, **kw) + mat(r, **kw) + (b"\x00" if f is None else b"\x01" + poly(f, **kw))


def open_commitment(c, t, **kw):
    return mat(c, **kw) + vec(t, **kw)


def challenge(d, **kw):
    return poly(d, **kw)


def open_response(z, **kw):
    return mat(z, **kw)


def linear_commitment(c, cp, g, t, tp, u, **kw):
    return mat(c, **kw) + mat(cp, **kw) + poly(g, **kw) + vec(t, **kw) + vec(tp, **kw) + mat(u, **kw)


def linear_response(z, zp, **kw):
    return mat(z, **kw) + mat(zp, **kw)


def sum_commitment(cp, cs, gs, tp, ts, u, **kw):
    T = len(gs)
    out = mat(cp, **kw) + _u64(T) + b"".join(mat(cs[i], **kw) for i in range(T))
    out += vec(gs, **kw) + vec(tp, **kw) + _u64(T) + b"".join(vec(ts[i], **kw) for i in range(T)) + mat(u, **kw)
    return out


def sum_response(zp, zs, **kw):
    return mat(zp, **kw) + _u64(len(zs)) + b"".join(mat(z, **kw) for z in zs)


def from_tokens(toks, streams, item, elem_bytes=8, trim=True):
    """The same bytes from the engine's token list (engine.wire_layout): pins rzk_wire_layout without a GPU."""
    out = b""
    for kind, s, p, v in toks:
        if kind == 1:
            out += _u64(v)
        elif kind == 3:
            out += bytes([v])
        elif kind == 2:
            a = np.asarray(streams[s][item])
            out += poly(a.reshape(-1, a.shape[-1])[p], elem_bytes=elem_bytes, trim=trim)
    return out
