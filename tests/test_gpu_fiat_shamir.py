"""Fiat-Shamir challenges on the device (docs/FIAT_SHAMIR.md) against the hashlib restatement, and the chained
non-interactive Open proof (commit -> challenge -> response without a host round trip)."""
import importlib

import numpy as np
import pytest

from oracle import fs_ref

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("ring-zk_b200")
engine = importlib.import_module("ring-zk_b200.engine")
api = importlib.import_module("ring-zk_b200.api")
N = 512


def test_shake_and_sample_in_ball_match_hashlib():
    import torch
    rng = np.random.default_rng(3)
    dev = torch.device("cuda:0")
    e = engine.Engine(N=N, device=0)
    try:
        B = 301
        c = rng.integers(-1757668526, 1757668527, size=(B, 2, N)).astype(np.int32)
        t = rng.integers(-1757668526, 1757668527, size=(B, 1, N)).astype(np.int32)
        r8 = rng.integers(-128, 128, size=(B, 3, N)).astype(np.int8)
        c[0] = 0; t[0] = 0                                            # all-zero message
        for pre in (b"", b"ring-zk/fs/test".ljust(32, b"\0"), bytes(range(168)), bytes(200)):   # incl. a prefix of exactly one rate block
            for segs in ([c, t], [t], [c, r8, t]):
                d = torch.empty((B, N), dtype=torch.int8, device=dev)
                e.fs_challenge(pre, [torch.from_numpy(a).to(dev) for a in segs], d)
                torch.cuda.synchronize()
                got = d.cpu().numpy()
                for i in list(range(8)) + [B - 1]:
                    want = fs_ref.challenge(pre, [a[i] for a in segs], N, 36)
                    assert (got[i] == want).all(), (len(pre), len(segs), i)
                assert (np.abs(got).sum(axis=1) == 36).all() and (np.abs(got) <= 1).all()
        with pytest.raises(engine.RzkError):
            e.fs_challenge(b"1234567", [torch.from_numpy(t).to(dev)], d)        # prefix not a multiple of 8 bytes
    finally:
        e.close()


def test_challenges_are_spread_over_the_ball():
    """positions and signs of 4096 challenges: every position is hit about kappa/N of the time, signs are balanced"""
    import torch
    rng = np.random.default_rng(4)
    dev = torch.device("cuda:0")
    e = engine.Engine(N=N, device=0)
    try:
        B = 4096
        t = torch.from_numpy(rng.integers(-10**9, 10**9, size=(B, 1, N)).astype(np.int32)).to(dev)
        d = torch.empty((B, N), dtype=torch.int8, device=dev)
        e.fs_challenge(b"", [t], d)
        g = d.cpu().numpy().astype(np.int64)
        hits = np.abs(g).sum(axis=0)                                   # Binomial(4096, 36/512): mean 288, sd 16.4
        assert abs(hits.mean() - 288) < 1e-9 and hits.min() > 288 - 6 * 17 and hits.max() < 288 + 6 * 17
        assert abs(g.sum()) < 6 * np.sqrt(B * 36)
        assert len({row.tobytes() for row in g}) == B                  # no two transcripts share a challenge
    finally:
        e.close()


def test_non_interactive_open_proofs():
    import torch
    rng = np.random.default_rng(8)
    params = api.Params.default()
    ck = params.generate_commitment_key(rng, N)
    prover, verifier = api.OpenProofProver(ck, params), api.OpenProofVerifier(ck, params)
    B = 257
    X = rng.integers(-params.q, params.q + 1, size=(B, 1, N)).astype(np.int32)
    secrets, proofs = prover.prove_batch_fs(rng, X)
    assert verifier.verify_batch_fs(proofs).all()
    # the same transcript through the interactive entry points and the host restatement of the challenge
    pre = api.fs_prefix(api.FS_TAG_OPEN, ck, params)
    c, t, z = (proofs[k].cpu().numpy() for k in ("c", "t", "z"))
    d_ref = np.stack([fs_ref.challenge(pre, [c[i], t[i]], N, params.kappa) for i in range(0, B, 37)])
    z_ref = prover.create_response_batch(secrets["y"][::37], secrets["r"][::37], d_ref)
    assert (z_ref == z[::37]).all()
    assert verifier.verify_batch(z[::37], t[::37], c[::37, :1], d_ref).all()
    # tampering with any part of a proof, or another session / key digest, invalidates it
    for key, idx in (("c", (3, 0, 5)), ("c", (4, 1, 500)), ("t", (5, 0, 0)), ("z", (6, 2, 9))):
        bad = {k: v.clone() for k, v in proofs.items()}
        bad[key][idx] += 1
        ok = verifier.verify_batch_fs(bad)
        assert not ok[idx[0]] and ok.sum() == B - 1, key
    assert not verifier.verify_batch_fs(proofs, session=b"session1").any()
    _, proofs2 = prover.prove_batch_fs(rng, X, session=b"session1")
    assert verifier.verify_batch_fs(proofs2, session=b"session1").all()
    ck.engine.close()


def test_host_entry_points_match_the_device_chain():
    """rzk_open_prove_fs_batch / rzk_open_verify_fs_batch (host pointers, chunked pipeline: what the Rust shim binds) against the
    interactive host entry points driven with the hashlib challenge, over more than one pipeline chunk"""
    rng = np.random.default_rng(21)
    s = pkg.synth.Synth(21, N=N)
    e = engine.Engine(N=N, device=0)
    try:
        e.set_key_blocks(*s.key())
        B = 8192 + 77
        x, r, y = s.message(B), s.small(B), s.gaussian(B)
        pre = b"ring-zk/fs/open/v1".ljust(32, b"\0") + bytes(range(32)) + bytes(24)
        o = e.open_prove_fs_host(x, r, y, pre)
        assert engine.unpack_bitmap(o["ok"], B).all()
        c, t, _ = e.open_commit(x, r, y)
        assert (o["c"] == c).all() and (o["t"] == t).all()
        idx = [0, 1, 8191, 8192, B - 1] + list(rng.integers(0, B, 12))
        for i in idx:
            assert (o["d"][i] == fs_ref.challenge(pre, [c[i], t[i]], N, 36)).all(), i
        assert (o["z"] == e.open_respond(y, r, o["d"])).all()
        assert engine.unpack_bitmap(e.open_verify_fs_host(o["c"], o["t"], o["z"], pre), B).all()
        zt = o["z"].copy(); zt[5, 0, 3] += 1; zt[B - 1, 2, 511] -= 1
        v = engine.unpack_bitmap(e.open_verify_fs_host(o["c"], o["t"], zt, pre), B)
        assert not v[5] and not v[B - 1] and v.sum() == B - 2
        assert not engine.unpack_bitmap(e.open_verify_fs_host(o["c"], o["t"], o["z"], pre[:-8] + b"other-id"), B).any()
    finally:
        e.close()
