"""GPU parity tests: the CUDA engine (through the C ABI, host-buffer entry points) against the
CPU oracle on identical seeded inputs, bit-exact.  Run with `-m gpu` on the B200 box.

Cases follow the reference's own tests (tests/test.rs:11-93: honest Open / Linear / Sum
transcripts with ragged messages, 4-term sums; commit.rs:161-170: swapped openings) plus the
negative tests the reference lacks (tampered z / t / c / d / g / u, oversized z, bad r).
"""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

engine = importlib.import_module("ring-zk_b200.engine")
synth = importlib.import_module("ring-zk_b200.synth")
from oracle import oracle as orc  # noqa: E402  (checker only)

N = 512
UB = engine.unpack_bitmap


@pytest.fixture(scope="module")
def setup():
    s = synth.Synth(42, N=N)
    a1p, a2p = s.key()
    eng = engine.Engine(N=N, device=0)
    eng.set_key_blocks(a1p, a2p)
    o = orc.Oracle(orc.Params(N=N), a1p, a2p)
    yield eng, o, s
    eng.close()


def test_bounds(setup):
    eng, o, _ = setup
    assert eng.sigma() == o.sigma() == 15444
    assert eng.commit_bound() == o.commit_bound() == 1359072
    assert eng.verify_bound() == o.verify_bound() == 679536
    assert eng.small_limit() > 20 * 15444


@pytest.mark.parametrize("B,ragged", [(1, False), (7, True), (200, True)])
def test_commit(setup, B, ragged):
    eng, o, s = setup
    x, r = s.message(B, ragged=ragged), s.small(B)
    c, ok = eng.commit(x, r)
    c_o, ok_o = o.commit_batch(x, r)
    assert (c == c_o).all()
    assert (UB(ok, B) == ok_o.astype(bool)).all() and ok_o.all()


def test_commit_readme_example(setup):
    # README.md:32-55 / commit.rs:66-78: x = [1,2,3,4]
    eng, o, s = setup
    x = np.zeros((1, 1, N), np.int32); x[0, 0, :4] = [1, 2, 3, 4]
    r = s.small(1)
    c, ok = eng.commit(x, r)
    assert o.commitment_verify(c[0], x[0], r[0]) and UB(ok, 1)[0]
    x2 = np.zeros((1, 1, N), np.int32); x2[0, 0, :4] = [4, 5, 6, 7]
    r2 = s.small(1)
    c2, _ = eng.commit(x2, r2)
    assert o.commitment_verify(c2[0], x2[0], r2[0])
    assert not o.commitment_verify(c2[0], x[0], r[0])          # commit.rs:169
    assert not o.commitment_verify(c[0], x2[0], r2[0])         # commit.rs:170


def test_commit_edge_values(setup):
    """extreme residues, all-zero and non-canonical representatives"""
    eng, o, s = setup
    half = (3515337053 - 1) // 2
    x = np.zeros((4, 1, N), np.int32)
    x[0] = half; x[1] = -half; x[3, 0, ::2] = half
    r = s.small(4); r[2] = 0
    c, _ = eng.commit(x, r)
    c_o, _ = o.commit_batch(x, r)
    assert (c == c_o).all()
    # a non-canonical i32 representative is canonicalised like ZqI64::from
    xn = x.copy(); xn[2, 0, 0] = np.int32(2 ** 31 - 1)
    c2, _ = eng.commit(xn, r)
    xc = o.center(xn.astype(np.int64))
    c2_o, _ = o.commit_batch(xc, r)
    assert (c2 == c2_o).all()


def test_commit_constraint_flag(setup):
    """check_commit_constraint (params.rs:102-108) needs a wide r to fail: use the device API with i8 limits"""
    eng, o, s = setup
    B = 9
    x, r = s.message(B), s.small(B)
    r[3, 1, :] = 127          # norm = 127*sqrt(512) = 2873 << bound: still ok
    c, ok = eng.commit(x, r)
    c_o, ok_o = o.commit_batch(x, r)
    assert (c == c_o).all() and (UB(ok, B) == ok_o.astype(bool)).all()


@pytest.mark.parametrize("B", [1, 65])
def test_open_proof(setup, B):
    eng, o, s = setup
    x, r, y, d = s.message(B, ragged=True), s.small(B), s.gaussian(B), s.challenge(B)
    c, t, ok = eng.open_commit(x, r, y)
    c_o, t_o, ok_o = o.open_commit_batch(x, r, y)
    assert (c == c_o).all() and (t == t_o).all() and UB(ok, B).all()
    z = eng.open_respond(y, r, d)
    assert (z == o.open_respond_batch(y, r, d)).all()
    c1 = np.ascontiguousarray(c[:, :1])
    v = UB(eng.open_verify(z, t, c1, d), B)
    assert v.all() and o.open_verify_batch(z, t, c1, d).all()
    # tampering: every single-coefficient change flips the bit, in the engine and in the oracle
    for name in ("z", "t", "c", "d"):
        zz, tt, cc, dd = z.copy(), t.copy(), c1.copy(), d.copy()
        arr = {"z": zz, "t": tt, "c": cc, "d": dd}[name]
        arr[:, ..., 5] += 1
        v = UB(eng.open_verify(zz, tt, cc, dd), B)
        v_o = o.open_verify_batch(zz, tt, cc, dd).astype(bool)
        assert (v == v_o).all() and not v.any(), name
    # oversized z: norm check (params.rs:112-118) fails
    zb = z.copy(); zb[::2, 2, 9] = eng.verify_bound() + 1
    v = UB(eng.open_verify(zb, t, c1, d), B)
    assert (v == o.open_verify_batch(zb, t, c1, d).astype(bool)).all()
    assert not v[::2].any()


def test_open_verify_norm_boundary(setup):
    """norm_2 exactly at / just above the bound (floor sqrt semantics, polynomial.rs:60-73)"""
    eng, o, s = setup
    B = 4
    vb = eng.verify_bound()
    z = np.zeros((B, 3, N), np.int32)
    z[0, 0, 0] = vb            # norm == bound -> passes the norm check
    z[1, 0, 0] = vb + 1        # fails
    z[2, 1, 0] = vb; z[2, 1, 1] = 1165   # sqrt(vb^2 + 1165^2) = vb + 0.998.. -> floor == vb passes
    z[3, 1, 0] = vb; z[3, 1, 1] = 1167   # just over -> fails
    d = s.challenge(B)
    t = np.zeros((B, 1, N), np.int32)
    c1 = np.zeros((B, 1, N), np.int32)
    # make the equation hold: t = A1.z - c1*d with c1 = 0
    t_o = o.mat_dot(o.a1, z[0].astype(np.int64)[:, None, :])
    for i in range(B):
        t[i] = o.mat_dot(o.a1, z[i].astype(np.int64)[:, None, :])[:, 0, :]
    v = UB(eng.open_verify(z, t, c1, d), B)
    v_o = o.open_verify_batch(z, t, c1, d).astype(bool)
    assert (v == v_o).all()
    assert list(v) == [True, False, True, False]
    assert t_o is not None


def test_range_error_reported(setup):
    eng, o, s = setup
    x, r, y = s.message(2), s.small(2), s.gaussian(2)
    y[1, 2, 100] = eng.small_limit() + 1
    with pytest.raises(engine.RzkError) as ei:
        eng.open_commit(x, r, y)
    assert ei.value.code == engine.RZK_ERR_RANGE
    y[1, 2, 100] = eng.small_limit()          # at the limit: exact
    c, t, _ = eng.open_commit(x, r, y)
    c_o, t_o, _ = o.open_commit_batch(x, r, y)
    assert (t == t_o).all()
    y[1, 0, 100] = 2 ** 30                     # y0 is never multiplied: any size is fine
    c, t, _ = eng.open_commit(x, r, y)
    c_o, t_o, _ = o.open_commit_batch(x, r, y)
    assert (t == t_o).all()


@pytest.mark.parametrize("B", [1, 33])
def test_linear_proof(setup, B):
    eng, o, s = setup
    x, g = s.message(B, ragged=True), s.scalar(B)
    r, rp, y, yp, d = s.small(B), s.small(B), s.gaussian(B), s.gaussian(B), s.challenge(B)
    lc = eng.linear_commit(g, x, rp, r, y, yp)
    lo = o.linear_commit_batch(g, x, rp, r, y, yp)
    for kname in ("gx", "cp", "c", "t", "tp", "u"):
        assert (lc[kname] == lo[kname]).all(), kname
    assert UB(lc["ok"], B).all()
    z, zp = eng.linear_respond(y, yp, r, rp, d)
    z_o, zp_o = o.linear_respond_batch(y, yp, r, rp, d)
    assert (z == z_o).all() and (zp == zp_o).all()
    v = UB(eng.linear_verify(z, zp, lc["c"], lc["cp"], g, lc["t"], lc["tp"], lc["u"], d), B)
    assert v.all() and o.linear_verify_batch(z, zp, lc["c"], lc["cp"], g, lc["t"], lc["tp"], lc["u"], d).all()
    for name in ("z", "zp", "c", "cp", "g", "t", "tp", "u", "d"):
        a = dict(z=z.copy(), zp=zp.copy(), c=lc["c"].copy(), cp=lc["cp"].copy(), g=g.copy(), t=lc["t"].copy(),
                 tp=lc["tp"].copy(), u=lc["u"].copy(), d=d.copy())
        a[name][:, ..., 17] += 1
        v = UB(eng.linear_verify(a["z"], a["zp"], a["c"], a["cp"], a["g"], a["t"], a["tp"], a["u"], a["d"]), B)
        v_o = o.linear_verify_batch(a["z"], a["zp"], a["c"], a["cp"], a["g"], a["t"], a["tp"], a["u"], a["d"])
        assert (v == v_o.astype(bool)).all() and not v.any(), name


@pytest.mark.parametrize("B,T", [(3, 1), (5, 4), (2, 64)])
def test_sum_proof(setup, B, T):
    eng, o, s = setup
    gs, xs = s.scalar(B, T), s.uniform_q(B, T, 1)
    rs, ys = s.small(B, T), s.gaussian(B, T)
    rp, yp, d = s.small(B), s.gaussian(B), s.challenge(B)
    sc = eng.sum_commit(gs, xs, rp, rs, ys, yp)
    so = o.sum_commit_batch(gs, xs, rp, rs, ys, yp)
    for kname in ("xp", "cp", "cs", "ts", "tp", "u"):
        assert (sc[kname] == so[kname]).all(), kname
    assert UB(sc["ok"], B).all()
    zs, zp = eng.sum_respond(ys, yp, rs, rp, d)
    zs_o, zp_o = o.sum_respond_batch(ys, yp, rs, rp, d)
    assert (zs == zs_o).all() and (zp == zp_o).all()
    v = UB(eng.sum_verify(zs, zp, sc["cs"], sc["cp"], gs, sc["ts"], sc["tp"], sc["u"], d), B)
    assert v.all() and o.sum_verify_batch(zs, zp, sc["cs"], sc["cp"], gs, sc["ts"], sc["tp"], sc["u"], d).all()
    # tamper a single term of a single instance: only that instance fails
    for name in ("zs", "cs", "gs", "ts", "u", "zp"):
        a = dict(zs=zs.copy(), zp=zp.copy(), cs=sc["cs"].copy(), cp=sc["cp"].copy(), gs=gs.copy(), ts=sc["ts"].copy(),
                 tp=sc["tp"].copy(), u=sc["u"].copy(), d=d.copy())
        arr = a[name]
        if arr.ndim >= 4 or name == "gs":
            arr[B - 1, T - 1, ..., 3] += 1
        else:
            arr[B - 1, ..., 3] += 1
        v = UB(eng.sum_verify(a["zs"], a["zp"], a["cs"], a["cp"], a["gs"], a["ts"], a["tp"], a["u"], a["d"]), B)
        v_o = o.sum_verify_batch(a["zs"], a["zp"], a["cs"], a["cp"], a["gs"], a["ts"], a["tp"], a["u"], a["d"])
        assert (v == v_o.astype(bool)).all(), name
        assert not v[B - 1] and v[:B - 1].all(), name


def test_worst_case_magnitudes(setup):
    """all operands at +-(q-1)/2: exercises the full range of the 2- and 3-prime CRT"""
    eng, o, s = setup
    half = (3515337053 - 1) // 2
    B = 2
    g = np.full((B, N), half, np.int32); g[1, ::2] = -half
    x = np.full((B, 1, N), half, np.int32); x[1, 0, 1::3] = -half
    r, rp, y, yp = s.small(B), s.small(B), s.gaussian(B), s.gaussian(B)
    lim = eng.small_limit()
    y[0] = lim; y[1, :, ::2] = -lim
    lc = eng.linear_commit(g, x, rp, r, y, yp)
    lo = o.linear_commit_batch(g, x, rp, r, y, yp)
    for kname in ("gx", "cp", "c", "t", "tp", "u"):
        assert (lc[kname] == lo[kname]).all(), kname


def test_i64_staging(setup):
    eng, o, s = setup
    rng = np.random.default_rng(5)
    a = rng.integers(-2 ** 62, 2 ** 62, size=3 * N + 17, dtype=np.int64)
    p = eng.pack_i64(a)
    assert (p == o.center(a)).all()
    assert (eng.unpack_i64(p) == p.astype(np.int64)).all()


def test_full_size_properties(setup):
    """BASELINE configs 2 and 3 at full size (2^16): EVERY item bit-exact against the oracle (commitment c, t = A1.y,
    response z, verify verdicts of a batch with a tampered subset), plus the size-independent properties
    - linearity of the commitment: com(x1; r1) + com(x2; r2) == com(x1 + x2; r1 + r2)  (mod q)
    - honest Open transcripts verify; a tampered subset does not."""
    eng, o, s = setup
    B = 1 << 16
    x1, x2 = s.message(B), s.message(B)
    r1, r2 = s.small(B), s.small(B)
    c1, ok1 = eng.commit(x1, r1)
    c2, ok2 = eng.commit(x2, r2)
    x12 = o.center(x1.astype(np.int64) + x2).astype(np.int32)
    c12, _ = eng.commit(x12, (r1 + r2).astype(np.int8))
    assert (o.center(c1.astype(np.int64) + c2) == c12).all()
    assert UB(ok1, B).all() and UB(ok2, B).all()
    y, d = s.gaussian(B), s.challenge(B)
    c, t, _ = eng.open_commit(x1, r1, y)
    assert (c == c1).all()
    c_o, t_o, ok_o = o.open_commit_batch(x1, r1, y)          # the whole batch, every coefficient
    assert (c == c_o).all() and (t == t_o).all() and ok_o.all()
    z = eng.open_respond(y, r1, d)
    assert (z == o.open_respond_batch(y, r1, d)).all()
    cc1 = np.ascontiguousarray(c[:, :1])
    assert UB(eng.open_verify(z, t, cc1, d), B).all()
    z[::1000, 1, 77] ^= 1
    t2 = t.copy(); t2[7::1777, 0, 300] += 1
    v = UB(eng.open_verify(z, t2, cc1, d), B)
    bad = np.zeros(B, bool); bad[::1000] = True; bad[7::1777] = True
    assert (v == ~bad).all()
    assert (v == o.open_verify_batch(z, t2, cc1, d).astype(bool)).all()


def test_full_size_linear(setup):
    """BASELINE config 4 at full size (2^14 instances) through the host entry points (several pipeline chunks on
    several streams): EVERY instance bit-exact against the oracle in every output of the prover's commit phase and in
    the responses; honest transcripts verify, a tampered subset fails and nothing else does, verdicts equal the oracle's."""
    eng, o, s = setup
    B = 1 << 14
    x, g = s.message(B), s.scalar(B)
    r, rp, y, yp, d = s.small(B), s.small(B), s.gaussian(B), s.gaussian(B), s.challenge(B)
    lc = eng.linear_commit(g, x, rp, r, y, yp)
    assert UB(lc["ok"], B).all()
    lo = o.linear_commit_batch(g, x, rp, r, y, yp)
    for kname in ("gx", "cp", "c", "t", "tp", "u"):
        assert (lc[kname] == lo[kname]).all(), kname
    z, zp = eng.linear_respond(y, yp, r, rp, d)
    z_o, zp_o = o.linear_respond_batch(y, yp, r, rp, d)
    assert (z == z_o).all() and (zp == zp_o).all()
    assert UB(eng.linear_verify(z, zp, lc["c"], lc["cp"], g, lc["t"], lc["tp"], lc["u"], d), B).all()
    u = lc["u"].copy()
    u[::997, ..., 5] += 1                                   # third equation only
    zp2 = zp.copy()
    zp2[3::1999, 2, 500] -= 1                               # second first-type equation and the third
    v = UB(eng.linear_verify(z, zp2, lc["c"], lc["cp"], g, lc["t"], lc["tp"], u, d), B)
    bad = np.zeros(B, bool); bad[::997] = True; bad[3::1999] = True
    assert (v == ~bad).all()
    assert (v == o.linear_verify_batch(z, zp2, lc["c"], lc["cp"], g, lc["t"], lc["tp"], u, d).astype(bool)).all()


def test_full_size_sum(setup):
    """BASELINE config 5 at full size (2^12 instances of 64 terms) through the host entry points: EVERY instance
    bit-exact against the oracle (x', c', c_i, t_i, t', u, z_i, z'); honest transcripts verify, single tampered terms
    fail exactly their instance, verdicts equal the oracle's."""
    eng, o, s = setup
    B, T = 1 << 12, 64
    gs, xs = s.scalar(B, T), s.uniform_q(B, T, 1)
    rs, ys = s.small(B, T), s.gaussian(B, T)
    rp, yp, d = s.small(B), s.gaussian(B), s.challenge(B)
    sc = eng.sum_commit(gs, xs, rp, rs, ys, yp)
    assert UB(sc["ok"], B).all()
    so = o.sum_commit_batch(gs, xs, rp, rs, ys, yp)
    for kname in ("xp", "cp", "cs", "ts", "tp", "u"):
        assert (sc[kname] == so[kname]).all(), kname
    zs, zp = eng.sum_respond(ys, yp, rs, rp, d)
    zs_o, zp_o = o.sum_respond_batch(ys, yp, rs, rp, d)
    assert (zs == zs_o).all() and (zp == zp_o).all()
    del so, zs_o, zp_o
    assert UB(eng.sum_verify(zs, zp, sc["cs"], sc["cp"], gs, sc["ts"], sc["tp"], sc["u"], d), B).all()
    ts = sc["ts"].copy()
    ts[5::501, 63, ..., 0] += 1                             # first equation of the last term
    gs2 = gs.copy()
    gs2[7::1013, 31, 9] ^= 1                                # third equation through one scalar
    v = UB(eng.sum_verify(zs, zp, sc["cs"], sc["cp"], gs2, ts, sc["tp"], sc["u"], d), B)
    bad = np.zeros(B, bool); bad[5::501] = True; bad[7::1013] = True
    assert (v == ~bad).all()
    assert (v == o.sum_verify_batch(zs, zp, sc["cs"], sc["cp"], gs2, ts, sc["tp"], sc["u"], d).astype(bool)).all()


@pytest.mark.parametrize("B,T", [(3, 4), (700, 4), (40, 64), (2500, 1)])
def test_device_resident_linear_sum_match_host_entry_points(setup, B, T):
    """The `_dev` Linear / Sum entry points (device pointers, caller's stream; what bench.py times) against the host
    entry points on the same inputs: identical outputs and verdicts, at instance counts where small product sums are
    cut into segments (3, 40, 700) and where they are not (2500)."""
    import torch
    eng, o, s = setup
    dev = torch.device("cuda:0")
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    E = lambda *sh: torch.empty(sh, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    gs, xs = s.scalar(B, T), s.uniform_q(B, T, 1)
    rs, ys = s.small(B, T), s.gaussian(B, T)
    rp, yp, d = s.small(B), s.gaussian(B), s.challenge(B)
    h = eng.sum_commit(gs, xs, rp, rs, ys, yp)
    xp, cp, cs, ts, tp, u = E(B, 1, N), E(B, 2, N), E(B, T, 2, N), E(B, T, 1, N), E(B, 1, N), E(B, 1, N)
    fl = torch.zeros(B, dtype=torch.int32, device=dev)
    eng.dev("sum_commit_batch", B, T, up(gs), up(xs), up(rp), up(rs), up(ys), up(yp), xp, cp, cs, ts, tp, u, fl, stream=st)
    torch.cuda.synchronize()
    for name, t_dev in (("xp", xp), ("cp", cp), ("cs", cs), ("ts", ts), ("tp", tp), ("u", u)):
        assert (t_dev.cpu().numpy() == h[name]).all(), name
    assert not fl.any()
    zs_h, zp_h = eng.sum_respond(ys, yp, rs, rp, d)
    zs, zp = E(B, T, 3, N), E(B, 3, N)
    eng.dev("sum_respond_batch", B, T, up(ys), up(yp), up(rs), up(rp), up(d), zs, zp, stream=st)
    torch.cuda.synchronize()
    assert (zs.cpu().numpy() == zs_h).all() and (zp.cpu().numpy() == zp_h).all()
    u_bad = u.clone(); u_bad[::3, ..., 9] += 1
    for u_t, expect in ((u, np.ones(B, bool)), (u_bad, np.arange(B) % 3 != 0)):
        fl.zero_()
        eng.dev("sum_verify_batch", B, T, zs, zp, cs, cp, up(gs), ts, tp, u_t, up(d), fl, stream=st)
        torch.cuda.synchronize()
        assert ((fl.cpu().numpy() == 0) == expect).all()
        v_h = UB(eng.sum_verify(zs_h, zp_h, h["cs"], h["cp"], gs, h["ts"], h["tp"], u_t.cpu().numpy(), d), B)
        assert (v_h == expect).all()
    if T == 1:
        g, x, r, y = gs[:, 0], xs[:, 0], rs[:, 0], ys[:, 0]
        hl = eng.linear_commit(g, x, rp, r, y, yp)
        gx, cpl, cl, tl, tpl, ul = E(B, 1, N), E(B, 2, N), E(B, 2, N), E(B, 1, N), E(B, 1, N), E(B, 1, N)
        fl.zero_()
        eng.dev("linear_commit_batch", B, up(g), up(x), up(rp), up(r), up(y), up(yp), gx, cpl, cl, tl, tpl, ul, fl, stream=st)
        torch.cuda.synchronize()
        for name, t_dev in (("gx", gx), ("cp", cpl), ("c", cl), ("t", tl), ("tp", tpl), ("u", ul)):
            assert (t_dev.cpu().numpy() == hl[name]).all(), name
        zl, zpl = E(B, 3, N), E(B, 3, N)
        eng.dev("linear_respond_batch", B, up(y), up(yp), up(r), up(rp), up(d), zl, zpl, stream=st)
        fl.zero_()
        eng.dev("linear_verify_batch", B, zl, zpl, cl, cpl, up(g), tl, tpl, ul, up(d), fl, stream=st)
        torch.cuda.synchronize()
        assert not fl.any()
        assert UB(eng.linear_verify(zl.cpu().numpy(), zpl.cpu().numpy(), hl["c"], hl["cp"], g, hl["t"], hl["tp"], hl["u"], d), B).all()


@pytest.mark.parametrize("B", [300, 5000])
def test_commit_split_key_range_and_masked_redo(setup, B):
    """The split-key commitment program is bit-exact against the oracle up to the |r| = 15 edge of its range and for arbitrary
    int32 representatives of x; items with |r| = 16 or more are redone by the two-prime program in a masked launch on the
    same stream (no host round trip, no second pass over the batch)."""
    eng, o, s = setup
    rng = np.random.default_rng(5)
    x, r = s.message(B, ragged=True), s.small(B)
    r[1] = rng.integers(-15, 16, size=r[1].shape)
    r[2, 1:] = 15; r[3, 1:] = -15
    x[4, 0, ::3] = np.int32(2 ** 31 - 1); x[4, 0, 1::3] = np.int32(-2 ** 31)
    launches0 = eng.kernel_launches()
    c, ok = eng.commit(x, r)
    clean = eng.kernel_launches() - launches0
    c_o, ok_o = o.commit_batch(o.center(x.astype(np.int64)).astype(np.int32), r)
    assert (c == c_o).all() and UB(ok, B).all()
    r[7, 2, 100] = 16                       # outside the one-word range
    r[B - 1, 1] = rng.integers(-127, 128, size=N)
    launches0 = eng.kernel_launches()
    c2, ok2 = eng.commit(x, r)
    assert eng.kernel_launches() - launches0 == clean       # same launches as a clean batch: the redo is the masked launch
    c2_o, _ = o.commit_batch(o.center(x.astype(np.int64)).astype(np.int32), r)
    assert (c2 == c2_o).all() and UB(ok2, B).all()


def test_masked_redo_full_batch_one_bad_item(setup):
    """2^16 commitments with ONE |r| = 16 item: every commitment equals the oracle's, and the call costs the launches
    and the PCIe bytes of a clean batch (no whole-batch redo on a range error)."""
    eng, o, s = setup
    B = 1 << 16
    x, r = s.message(B), s.small(B)
    launches0 = eng.kernel_launches()
    c0, _ = eng.commit(x, r)
    clean = eng.kernel_launches() - launches0
    r[40000, 2, 9] = 16
    launches0 = eng.kernel_launches()
    c, ok = eng.commit(x, r)
    assert eng.kernel_launches() - launches0 == clean
    assert UB(ok, B).all()
    same = np.ones(B, bool); same[40000] = False
    assert (c[same] == c0[same]).all()
    c_o, _ = o.commit_batch(x[39990:40010], r[39990:40010])
    assert (c[39990:40010] == c_o).all()
    # the Open / Linear / Sum provers share the path: one bad item among 300 instances of 3 terms
    Bs, T = 300, 3
    gs, xs = s.scalar(Bs, T), s.uniform_q(Bs, T, 1)
    rs, ys = s.small(Bs, T), s.gaussian(Bs, T)
    rp, yp = s.small(Bs), s.gaussian(Bs)
    rs[123, 1, 2, 500] = -77; rp[7, 1, 0] = 16
    sc = eng.sum_commit(gs, xs, rp, rs, ys, yp)
    so = o.sum_commit_batch(gs, xs, rp, rs, ys, yp)
    for kname in ("xp", "cp", "cs", "ts", "tp", "u"):
        assert (sc[kname] == so[kname]).all(), kname
    assert UB(sc["ok"], Bs).all()


@pytest.mark.parametrize("B", [1, 37, 9000])
def test_commit_packed_randomness(setup, B):
    """rzk_commit_batch_r2: the randomness at 2 bits per coefficient (what Params::default() needs) gives the commitments of
    rzk_commit_batch and of the oracle bit for bit, across chunk boundaries of the host pipeline (8192 items), and an entry
    -2 (representable, outside the small-prime program's range) is redone by the masked launch like any other."""
    eng, o, s = setup
    x, r = s.message(B, ragged=True), s.small(B)
    if B > 5:
        r[3, 1, 17] = -2; r[B - 1, 2, 511] = -2; r[4, 0, :] = -2
    r2 = engine.pack_r2(r)
    assert r2.shape == (B, 3, N // 4) and r2.dtype == np.uint8
    c, ok = eng.commit_r2(x, r2)
    c8, ok8 = eng.commit(x, r)
    assert (c == c8).all() and (ok == ok8).all()
    nb = min(B, 64)
    c_o, ok_o = o.commit_batch(x[:nb], r[:nb])
    assert (c[:nb] == c_o).all() and (UB(ok, B)[:nb] == ok_o.astype(bool)).all()
    c_o, _ = o.commit_batch(x[B - 1:], r[B - 1:])
    assert (c[B - 1:] == c_o).all()
    with pytest.raises(engine.RzkError):
        bad = r.copy(); bad[0, 0, 0] = 2
        engine.pack_r2(bad)


def test_thirty_bit_split_key_program_still_exact():
    """Engines with b = 1 commit modulo the small prime (signed lazy arithmetic); the 30-bit split-key program serves
    2 <= b <= 15 and stays selectable for b = 1 (RZK_TUNE=commit_small=0): both give the oracle's commitments."""
    import os
    s = synth.Synth(17, N=N)
    a1p, a2p = s.key()
    o = orc.Oracle(orc.Params(N=N), a1p, a2p)
    B = 130
    x, r = s.message(B, ragged=True), s.small(B)
    c_o, _ = o.commit_batch(x, r)
    old = os.environ.get("RZK_TUNE")
    os.environ["RZK_TUNE"] = "commit_small=0"
    try:
        eng = engine.Engine(N=N, device=0)
    finally:
        if old is None:
            os.environ.pop("RZK_TUNE")
        else:
            os.environ["RZK_TUNE"] = old
    try:
        eng.set_key_blocks(a1p, a2p)
        c, ok = eng.commit(x, r)
        assert (c == c_o).all() and UB(ok, B).all()
    finally:
        eng.close()


def test_large_b_runs_generic_commit():
    """b > 15 (accepted while b * kappa <= 74): every commitment runs the two-prime program, exact for any int8 r, on the host
    AND on the `_dev` entry points (no garbage c behind a FLAG_RANGE bit)."""
    import torch
    s = synth.Synth(91, N=N)
    a1p, a2p = s.key()
    P = engine.lib().rzk_default_params(N)
    P.b, P.kappa = 24, 3
    eng = engine.Engine(N=N, device=0, params=P)
    try:
        eng.set_key_blocks(a1p, a2p)
        o = orc.Oracle(orc.Params(N=N, b=24, kappa=3), a1p, a2p)
        B = 70
        rng = np.random.default_rng(3)
        x = s.message(B)
        r = rng.integers(-24, 25, size=(B, 3, N)).astype(np.int8)
        c, ok = eng.commit(x, r)
        c_o, ok_o = o.commit_batch(x, r)
        assert (c == c_o).all() and (UB(ok, B) == ok_o.astype(bool)).all()
        dev = torch.device("cuda:0")
        cd = torch.empty((B, 2, N), dtype=torch.int32, device=dev)
        fl = torch.zeros(B, dtype=torch.int32, device=dev)
        eng.dev("commit_batch", B, torch.from_numpy(x).to(dev), torch.from_numpy(r).to(dev), cd, fl,
                stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert (cd.cpu().numpy() == c_o).all() and not fl.any()
    finally:
        eng.close()


def test_commitment_verify_batch(setup):
    """Commitment::verify (commit.rs:173-210), both branches, against the oracle item by item:
    honest openings, swapped openings (commit.rs:169-170), tampered c / r, a randomised opening
    r' = f*r with f in the challenge space (commit.rs:203-207), and an r that fails the commit constraint."""
    eng, o, s = setup
    B = 24
    x, r = s.message(B, ragged=True), s.small(B)
    c, _ = eng.commit(x, r)
    v = UB(eng.commitment_verify(c, x, r), B)
    assert v.all() and all(o.commitment_verify(c[i], x[i], r[i]) for i in range(B))
    xs = np.roll(x, 1, axis=0)
    assert not UB(eng.commitment_verify(c, xs, r), B).any()
    ct = c.copy(); ct[::2, 1, 100] += 1
    rt = r.copy(); rt[1::4, 0, 3] += 1
    v = UB(eng.commitment_verify(ct, x, rt), B)
    expect = np.array([o.commitment_verify(ct[i], x[i], rt[i]) for i in range(B)])
    assert (v == expect).all() and not expect[::2].any() and not expect[1::4].any() and expect.sum() == B - 12 - 6
    # Some(f)
    f = s.challenge(B)
    rf = np.stack([np.stack([o.poly_mul(r[i, j], f[i]) for j in range(3)]) for i in range(B)]).astype(np.int8)
    v = UB(eng.commitment_verify(c, x, rf, f), B)
    assert v.all() and all(o.commitment_verify(c[i], x[i], rf[i], f[i]) for i in range(B))
    assert not UB(eng.commitment_verify(c, x, r, f), B).any()              # plain r does not open f*c
    ft = f.copy(); ft[::3, 0] += 1
    v = UB(eng.commitment_verify(c, x, rf, ft), B)
    expect = np.array([o.commitment_verify(c[i], x[i], rf[i], ft[i]) for i in range(B)])
    assert (v == expect).all() and not expect[::3].any()
    # check_commit_constraint(r) is part of the verdict (commit.rs:182); unreachable with int8 rows at the
    # default bound (1,359,072 > 127*sqrt(512)), so the verdicts above are the equation's alone
    assert eng.commit_bound() > 127 * 23


def test_api_commitment_verify_with_f():
    api = importlib.import_module("ring-zk_b200.api")
    rng = np.random.default_rng(3)
    params = api.Params.default()
    ck = params.generate_commitment_key(rng, N)
    x = params.prepare_value([[1, 2, 3, 4]], N)
    opening, com = ck.commit(rng, x, params)
    assert com.verify(opening, ck, params)
    f = np.zeros(N, np.int8); f[7] = 1                                     # the unit X^7
    rf = np.zeros_like(opening.r)
    for j in range(3):
        rf[j, 7:] = opening.r[j, :N - 7]
        rf[j, :7] = -opening.r[j, N - 7:]
    assert com.verify(api.Opening(opening.x, rf, f), ck, params)
    assert not com.verify(api.Opening(opening.x, opening.r, f), ck, params)
    assert api.Commitment.verify_batch(com.c[None], opening.x[None], rf[None], ck, f[None]).all()


def test_respond_rotation_kernel_and_fallback(setup):
    """z = y + d*r: the rotation kernel (byte accumulators) answers the honest items, the NTT program redoes the
    items it declines (|r| beyond the byte range, dense or non-{-1,0,1} d); any int32 representative of y."""
    eng, o, s = setup
    B = 40
    y, r, d = s.gaussian(B), s.small(B), s.challenge(B)
    rng = np.random.default_rng(11)
    r[1] = rng.integers(-3, 4, size=r[1].shape)                 # still inside the byte range (36 * 6 < 256)
    r[2] = rng.integers(-127, 128, size=r[2].shape)             # declined: |r| > 3
    d[3] = rng.integers(-1, 2, size=d[3].shape)                 # declined: ~340 non-zeros
    d[4, 5] = 2                                                 # declined: entry outside {-1, 0, 1}
    d[5] = 0                                                    # d = 0: z = y
    d[6] = 0; d[6, :127] = 1                                    # 127 terms: bias 1, sums up to 254
    half = (3515337053 - 1) // 2
    y[7, 0, ::2] = np.int32(2 ** 31 - 1); y[7, 0, 1::2] = np.int32(-2 ** 31)
    y[7, 1, :] = half - (np.arange(N) % 40); y[7, 2, :] = -half + (np.arange(N) % 40)
    z = eng.open_respond(y, r, d)
    z_o = o.open_respond_batch(o.center(y.astype(np.int64)), r, d)
    assert (z == z_o).all()
    # Sum-proof shape: one challenge per instance shared by its T terms
    T = 3
    ys, rs = s.gaussian(B, T), s.small(B, T)
    yp, rp = s.gaussian(B), s.small(B)
    rs[2, 1] = 100
    zs, zp = eng.sum_respond(ys, yp, rs, rp, d)
    zs_o, zp_o = o.sum_respond_batch(ys, yp, rs, rp, d)
    assert (zs == zs_o).all() and (zp == zp_o).all()


def test_device_outputs_stay_in_bounds(setup):
    """Outputs of the device-resident entry points are carved out of larger buffers with canary words on both
    sides (compute-sanitizer is not available on this pool): odd batch sizes, the rotation-kernel response with its
    masked fallback and Open verify by rotations must leave every canary intact."""
    import torch
    eng, o, s = setup
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    CAN = 4096

    def carve(shape, dtype=torch.int32):
        n = int(np.prod(shape))
        buf = torch.full((n + 2 * CAN,), -1234567, dtype=dtype, device=dev)
        return buf, buf[CAN:CAN + n].view(*shape)

    def intact(buf, n):
        return bool((buf[:CAN] == -1234567).all()) and bool((buf[CAN + n:] == -1234567).all())

    for B in (4099, 333):
        x, r, y, d = (torch.from_numpy(a).to(dev) for a in (s.message(B), s.small(B), s.gaussian(B), s.challenge(B)))
        r[5, 1, 7] = 100                                       # one item for the masked NTT fallback of the response
        cb, c = carve((B, 2, N)); tb, t = carve((B, 1, N)); zb, z = carve((B, 3, N)); fb, flags = carve((B,))
        flags.zero_()
        eng.dev("open_commit_batch", B, x, r, y, c, t, flags, stream=st)
        eng.dev("open_respond_batch", B, y, r, d, z, stream=st)
        eng.dev("open_verify_batch", B, z, t, c, 2, d, flags, stream=st)
        eng.dev("commitment_verify_batch", B, c, x, r, None, flags, stream=st)
        torch.cuda.synchronize()
        assert intact(cb, B * 2 * N) and intact(tb, B * N) and intact(zb, B * 3 * N) and intact(fb, B)
        fl = flags.cpu().numpy()
        assert (fl[np.arange(B) != 5] == 0).all() and (fl[5] & 2)     # item 5: range flag from the one-word commitment kernels
        idx = [0, 5, B - 1]
        xs, rs, ys, ds = (a[idx].cpu().numpy() for a in (x, r, y, d))
        c_o, t_o, _ = o.open_commit_batch(xs, rs, ys)
        z_o = o.open_respond_batch(ys, rs, ds)
        assert (z[idx].cpu().numpy() == z_o).all() and (t[idx].cpu().numpy() == t_o).all()
        assert (c[[0, B - 1]].cpu().numpy() == c_o[[0, 2]]).all()


def test_device_group_matches_single_engine(setup):
    """rzk_group_*: one engine + one host thread per listed device, the batch split in 8-aligned contiguous ranges.
    Results equal the single-engine ones bit for bit.  With one GPU the group lists device 0 three times (three
    engines, three threads); with more GPUs it uses them."""
    import torch
    eng, o, s = setup
    ndev = torch.cuda.device_count()
    ids = list(range(ndev)) if ndev >= 2 else [0, 0, 0]
    grp = engine.Group(ids, N=N)
    try:
        a1p, a2p = synth.Synth(42, N=N).key()
        grp.set_key_blocks(a1p, a2p)
        assert grp.size() == len(ids)
        B = 5003                                           # ragged: ranges of 1672, 1672, 1659 items on three engines
        x, r, y, d = s.message(B, ragged=True), s.small(B), s.gaussian(B), s.challenge(B)
        c, t, ok = eng.open_commit(x, r, y)
        cg, tg, okg = grp.open_commit(x, r, y)
        assert (c == cg).all() and (t == tg).all() and (ok == okg).all()
        z = eng.open_respond(y, r, d)
        assert (z == grp.open_respond(y, r, d)).all()
        zt = z.copy(); zt[::7, 2, 3] += 1
        c1 = np.ascontiguousarray(c[:, :1])
        v, vg = eng.open_verify(zt, t, c1, d), grp.open_verify(zt, t, c1, d)
        assert (v == vg).all() and UB(vg, B).sum() == B - len(range(0, B, 7))
        assert (grp.commitment_verify(c, x, r) == eng.commitment_verify(c, x, r)).all()
        cc, okc = grp.commit(x, r)
        assert (cc == c).all() and UB(okc, B).all()
        # Linear and Sum through the group (small batches: more engines than 8-item ranges is fine)
        Bl, T = 21, 3
        g, rp, yp = s.scalar(Bl), s.small(Bl), s.gaussian(Bl)
        L1 = eng.linear_commit(g, x[:Bl], rp, r[:Bl], y[:Bl], yp)
        L2 = grp.linear_commit(g, x[:Bl], rp, r[:Bl], y[:Bl], yp)
        assert all((L1[k] == L2[k]).all() for k in L1)
        gs, xs, rs, ys = s.scalar(Bl, T), s.uniform_q(Bl, T, 1), s.small(Bl, T), s.gaussian(Bl, T)
        S1 = eng.sum_commit(gs, xs, rp, rs, ys, yp)
        S2 = grp.sum_commit(gs, xs, rp, rs, ys, yp)
        assert all((S1[k] == S2[k]).all() for k in S1)
        zs, zp = grp.sum_respond(ys, yp, rs, rp, d[:Bl])
        assert UB(grp.sum_verify(zs, zp, S2["cs"], S2["cp"], gs, S2["ts"], S2["tp"], S2["u"], d[:Bl]), Bl).all()
        assert grp.kernel_launches() > 0
    finally:
        grp.close()


@pytest.mark.parametrize("b,kappa,B", [(4, 18, 33), (2, 37, 1 << 16)])
def test_other_admissible_parameters(b, kappa, B):
    """Parameter sets at the edge of what rzk_create admits (b * kappa <= 74; sigma, both norm bounds and the small
    operands scale with b * kappa): (4, 18) makes the rotation kernel decline the responses (|r| = 4 > its bias) so the
    NTT program answers them; (2, 37) is the accepted extreme -- rzk_small_limit() = 10.09 sigma -- run at the full 2^16
    batch: honest N(0, sigma) masking vectors never raise RZK_ERR_RANGE and every output is bit-exact against the oracle
    built with the same parameters (test the accepted extremes at full size)."""
    s = synth.Synth(31, N=N, b=b)
    a1p, a2p = s.key()
    P = engine.lib().rzk_default_params(N)
    P.b, P.kappa = b, kappa
    eng = engine.Engine(N=N, device=0, params=P)
    try:
        eng.set_key_blocks(a1p, a2p)
        o = orc.Oracle(orc.Params(N=N, b=b, kappa=kappa), a1p, a2p)
        sigma = b * 11 * kappa * 39
        assert eng.sigma() == o.sigma() == sigma and eng.verify_bound() == o.verify_bound() == 2 * sigma * 22
        assert eng.small_limit() >= 10 * sigma
        rng = np.random.default_rng(b * 100 + kappa)
        x, r = s.message(B, ragged=True), s.small(B)
        y = np.trunc(rng.normal(0.0, float(sigma), size=(B, 3, N))).astype(np.int32)
        d = np.zeros((B, N), np.int8)
        pos = np.argsort(rng.random((B, N)), axis=1)[:, :kappa]
        np.put_along_axis(d, pos, (rng.integers(0, 2, size=(B, kappa)) * 2 - 1).astype(np.int8), axis=1)
        assert np.abs(r).max() == b
        c, t, ok = eng.open_commit(x, r, y)
        c_o, t_o, ok_o = o.open_commit_batch(x, r, y)
        assert (c == c_o).all() and (t == t_o).all() and UB(ok, B).all() and ok_o.all()
        z = eng.open_respond(y, r, d)
        assert (z == o.open_respond_batch(y, r, d)).all()
        c1 = np.ascontiguousarray(c[:, :1])
        assert UB(eng.open_verify(z, t, c1, d), B).all()
        zb = z.copy(); zb[0, 0, 0] = eng.verify_bound() + 1; zb[1, 2, 9] += 1
        v = UB(eng.open_verify(zb, t, c1, d), B)
        v_o = o.open_verify_batch(zb, t, c1, d).astype(bool)
        assert (v == v_o).all() and not v[0] and not v[1] and v[2:].all()
    finally:
        eng.close()
