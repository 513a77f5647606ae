"""Wire format of the messages: the CPU restatement against the reference's own 36-byte pin
(/root/reference/src/mat.rs:424-438), the engine's token lists against the restatement (no GPU: rzk_wire_layout is host
logic), and -- on the GPU -- the device packer / parser against both."""
import importlib
import struct

import numpy as np
import pytest

from oracle import wire_ref as W

engine = importlib.import_module("ring-zk_b200.engine")
N = 512


def _rand(rng, *shape, lo=-1757668526, hi=1757668526):
    return rng.integers(lo, hi + 1, size=shape + (N,), dtype=np.int64).astype(np.int32)


def _messages(rng, B, T):
    """streams per message kind (numbered as in include/ringzk_b200.h) with ragged trailing zeros, and the direct serialisers"""
    d = np.zeros((B, 1, N), np.int8)
    for b in range(B):
        pos = rng.choice(N, 36, replace=False)
        d[b, 0, pos] = rng.choice([-1, 1], 36)
    r = rng.integers(-1, 2, size=(B, 3, N)).astype(np.int8)
    c, cp, t, tp, u, g, x, z, zp = (_rand(rng, B, 2), _rand(rng, B, 2), _rand(rng, B, 1), _rand(rng, B, 1), _rand(rng, B, 1),
                                     _rand(rng, B, 1), _rand(rng, B, 1), _rand(rng, B, 3, lo=-70000, hi=70000), _rand(rng, B, 3, lo=-70000, hi=70000))
    cs, gs, ts, zs = _rand(rng, B, T, 2), _rand(rng, B, T), _rand(rng, B, T), _rand(rng, B, T, 3, lo=-70000, hi=70000)
    # trailing zeros of different lengths, a zero polynomial, a polynomial with a single leading coefficient
    t[0, 0, 100:] = 0; u[0, 0, :] = 0; x[0, 0, 1:] = 0; z[0, 1, 300:] = 0; cs[0, 0, 1, 7:] = 0; gs[0, T - 1, :] = 0
    if B > 1:
        c[1, 1, N - 1] = 0; zs[1, 0, 2, 1:] = 0; r[1, 2, :] = 0
    f = d[:, 0].copy()[:, None]
    K = engine
    return {
        K.MSG_COMMITMENT: ([c], lambda i, **kw: W.commitment(c[i], **kw)),
        K.MSG_OPENING: ([x, r], lambda i, **kw: W.opening(x[i], r[i], None, **kw)),
        K.MSG_OPENING_F: ([x, r, f], lambda i, **kw: W.opening(x[i], r[i], f[i, 0], **kw)),
        K.MSG_OPEN_COMMITMENT: ([c, t], lambda i, **kw: W.open_commitment(c[i], t[i], **kw)),
        K.MSG_CHALLENGE: ([d], lambda i, **kw: W.challenge(d[i, 0], **kw)),
        K.MSG_OPEN_RESPONSE: ([z], lambda i, **kw: W.open_response(z[i], **kw)),
        K.MSG_LINEAR_COMMITMENT: ([c, cp, g, t, tp, u], lambda i, **kw: W.linear_commitment(c[i], cp[i], g[i, 0], t[i], tp[i], u[i], **kw)),
        K.MSG_LINEAR_RESPONSE: ([z, zp], lambda i, **kw: W.linear_response(z[i], zp[i], **kw)),
        K.MSG_SUM_COMMITMENT: ([cp, cs, gs, tp, ts, u],
                               lambda i, **kw: W.sum_commitment(cp[i], cs[i], gs[i], tp[i], ts[i][:, None], u[i], **kw)),
        K.MSG_SUM_RESPONSE: ([zp, zs], lambda i, **kw: W.sum_response(zp[i], zs[i], **kw)),
    }


def test_reference_pin_36_bytes():
    """mat.rs:424-438: Mat { polynomials: vec![vec![Polynomial::<i32, N>::new(vec![1, 2, 3])]] } serialises to 36 bytes"""
    p = np.zeros(N, np.int32); p[:3] = [1, 2, 3]
    got = W.mat([p], elem_bytes=4, trim=True)
    assert len(got) == 36
    assert got == struct.pack("<QQQiii", 1, 1, 3, 1, 2, 3)
    assert len(W.mat([p], elem_bytes=4, trim=False)) == 8 + 8 + 8 + 4 * N          # the other reading of the dependency


def test_token_lists_match_the_struct_layouts():
    """rzk_wire_layout (host logic of the shared library) replayed in Python == the field-by-field serialisers"""
    rng = np.random.default_rng(5)
    B, T = 3, 4
    for kind, (streams, direct) in _messages(rng, B, T).items():
        toks = engine.wire_layout(kind, T)
        for i in range(B):
            for eb in (8, 4):
                for trim in (True, False):
                    assert W.from_tokens(toks, streams, i, elem_bytes=eb, trim=trim) == direct(i, elem_bytes=eb, trim=trim), (kind, i, eb, trim)
    with pytest.raises(engine.RzkError):
        engine.wire_layout(99)
    with pytest.raises(engine.RzkError):
        engine.wire_layout(engine.MSG_SUM_RESPONSE, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("elem_bytes,trim", [(8, True), (8, False), (4, True)])
def test_device_pack_and_unpack(elem_bytes, trim):
    import torch
    rng = np.random.default_rng(11)
    B, T = 37, 5
    e = engine.Engine(N=N, device=0)
    dev = torch.device("cuda:0")
    try:
        for kind, (streams, direct) in _messages(rng, B, T).items():
            ds = [torch.from_numpy(a).to(dev) for a in streams]
            data, off = e.wire_pack(kind, ds, T=T, elem_bytes=elem_bytes, trim=trim)
            off_h, data_h = off.cpu().numpy(), data.cpu().numpy().tobytes()
            want = [direct(i, elem_bytes=elem_bytes, trim=trim) for i in range(B)]
            assert off_h[0] == 0 and list(np.diff(off_h)) == [len(w) for w in want], kind
            assert data_h == b"".join(want), kind
            # parse back into fresh arrays
            outs = [torch.full_like(a, 77) for a in ds]
            flags = e.wire_unpack(kind, data, off, outs, T=T, elem_bytes=elem_bytes)
            assert not flags.any(), kind
            for a, b in zip(ds, outs):
                assert torch.equal(a, b), kind
            # malformed items are flagged, the others still parse: a wrong length word, a truncated item, an oversize polynomial
            bad = bytearray(data_h)
            bad[int(off_h[1])] ^= 0x04                                   # first u64 of item 1
            data_bad = torch.frombuffer(bad, dtype=torch.uint8).to(dev)
            off_bad = off.clone(); off_bad[B] -= 1                        # last item one byte short
            flags = e.wire_unpack(kind, data_bad, off_bad, outs, T=T, elem_bytes=elem_bytes).cpu().numpy()
            assert flags[1] == 1 and flags[B - 1] == 1 and not flags[[0] + list(range(2, B - 1))].any(), kind
            # offsets that run backwards or leave the buffer mark the item and read nothing of it
            off_bad = off.clone(); off_bad[3] = off_bad[2] - 8
            off_bad[B] = data.numel() + 4096
            flags = e.wire_unpack(kind, data, off_bad, outs, T=T, elem_bytes=elem_bytes).cpu().numpy()
            assert flags[2] == 1 and flags[B - 1] == 1 and not flags[[0, 1] + list(range(4, B - 1))].any(), kind
    finally:
        e.close()


@pytest.mark.gpu
def test_device_pack_reproduces_the_reference_pin():
    """the 36 bytes of mat.rs:434 from the device packer: a 1 x 1 Mat is the `u` field of the Linear commitment layout"""
    import torch
    e = engine.Engine(N=N, device=0)
    try:
        p = torch.zeros((1, 1, N), dtype=torch.int32, device="cuda:0"); p[0, 0, :3] = torch.tensor([1, 2, 3])
        toks = [(engine.WIRE_LEN, 0, 0, 1), (engine.WIRE_LEN, 0, 0, 1), (engine.WIRE_POLY, 0, 0, 0)]      # Mat, 1 x 1
        import ctypes as C
        ctoks = (engine.WireTok * 3)(*[engine.WireTok(*t) for t in toks])
        cs = (engine.WireStream * 1)(engine.WireStream(p.data_ptr(), 1, 0))
        off = torch.empty(2, dtype=torch.int64, device="cuda:0"); out = torch.zeros(64, dtype=torch.uint8, device="cuda:0")
        total = C.c_uint64(0)
        rc = e.L.rzk_wire_pack_dev(e.h, 1, ctoks, 3, cs, 1, 4, 1, out.data_ptr(), 64, off.data_ptr(), C.byref(total), 0)
        assert rc == 0 and total.value == 36
        assert out[:36].cpu().numpy().tobytes() == struct.pack("<QQQiii", 1, 1, 3, 1, 2, 3)
    finally:
        e.close()


@pytest.mark.gpu
def test_host_forms_of_pack_and_unpack():
    """rzk_wire_pack / rzk_wire_unpack (host arrays: what the Rust shim binds) == the field-by-field serialisers, and back"""
    rng = np.random.default_rng(12)
    B, T = 19, 3
    e = engine.Engine(N=N, device=0)
    try:
        for kind, (streams, direct) in _messages(rng, B, T).items():
            data, off = e.wire_pack_host(kind, streams, T=T)
            want = [direct(i, elem_bytes=8, trim=True) for i in range(B)]
            assert data == b"".join(want) and list(np.diff(off.astype(np.int64))) == [len(w) for w in want], kind
            outs = [np.full_like(a, 55) for a in streams]
            ok = e.wire_unpack_host(kind, data, off, outs, T=T)
            assert engine.unpack_bitmap(ok, B).all(), kind
            for a, b in zip(streams, outs):
                assert (a == b).all(), kind
            bad = bytearray(data); bad[int(off[2]) + 3] ^= 0x40
            ok = engine.unpack_bitmap(e.wire_unpack_host(kind, bytes(bad), off, outs, T=T), B)
            assert not ok[2] and ok.sum() == B - 1, kind
    finally:
        e.close()
