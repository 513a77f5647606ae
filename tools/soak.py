"""Differential soak test on the GPU: random batch sizes; the split-key commitment program with / without phase mixing and the
two-prime program run by the generic interpreter against each other; the rotation-kernel response against the NTT response;
Open verify with c1*d as signed rotations against the NTT-domain product (honest and tampered transcripts); Sum / Linear proofs
in their default lowering against the plain lowering run by the generic interpreter.
usage: python tools/soak.py [rounds] [seed]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("ring-zk_b200")
engine = importlib.import_module("ring-zk_b200.engine")
N = 512


def make(env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        e = engine.Engine(N=N, device=0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return e


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    s = pkg.synth.Synth(seed, N=N)
    a1p, a2p = s.key()
    engines = {"int+pp": make({}), "int": make({"RZK_TUNE": "commit_pp=9"}), "generic": make({"RZK_TEST_LOWERING": "generic"}),
               "ntt": make({"RZK_TEST_LOWERING": "nosparse,norot"})}
    for e in engines.values():
        e.set_key_blocks(a1p, a2p)
    UB = engine.unpack_bitmap
    for it in range(rounds):
        B = int(rng.choice([1, 2, 7, 8, 9, 63, 148, 149, 1000, 2367, 2368, 2369, 4095, 4096, 4097, 8191, 8193, int(rng.integers(1, 20000))]))
        x, r, y, d = s.message(B, ragged=bool(it & 1)), s.small(B), s.gaussian(B), s.challenge(B)
        if it % 3 == 0:
            r = rng.integers(-3, 4, size=r.shape).astype(np.int8)
        ref = None
        for name in ("int+pp", "int", "generic"):
            c, ok = engines[name].commit(x, r)
            assert UB(ok, B).all(), (name, B)
            if ref is None:
                ref = c
            else:
                assert (c == ref).all(), (name, B, it)
        e0 = engines["int+pp"]
        c, t, _ = e0.open_commit(x, r, y)
        assert (c == ref).all()
        z = e0.open_respond(y, r, d)
        z2 = engines["ntt"].open_respond(y, r, d)
        assert (z == z2).all(), ("respond", B, it)
        c1 = np.ascontiguousarray(c[:, :1])
        v = UB(e0.open_verify(z, t, c1, d), B)
        assert v.all(), ("verify", B, it)
        bad = rng.random(B) < 0.25
        zt = z.copy(); zt[bad, int(rng.integers(0, 3)), int(rng.integers(0, N))] += 1
        dt = d.copy(); dt[bad, int(rng.integers(0, N))] ^= 1                   # one rotation more / fewer / a sign flipped
        for args in ((zt, t, c1, d), (z, t, c1, dt)):
            v1, v2 = UB(e0.open_verify(*args), B), UB(engines["ntt"].open_verify(*args), B)
            assert (v1 == ~bad).all() and (v2 == ~bad).all(), ("verify tampered", B, it)
        assert UB(e0.commitment_verify(c, x, r), B).all()
        print(f"round {it:3d} B={B:6d} ok", flush=True)
    # Linear / Sum proofs: the default lowering (three-prime kernels with the chunked epilogue, shared challenge image,
    # one-launch product sums) against the plain one (every item transforms its own challenge, one launch per product
    # sum, no cutting of small batches into segments) run through the generic interpreter, at instance counts around the launch-shape boundaries of the
    # half-warp-per-item kernels (148 SMs x 2 x warps)
    plain = make({"RZK_TEST_LOWERING": "generic,nofuse,nodimg,nosegments,norot,nosparse"})
    plain.set_key_blocks(a1p, a2p)
    e0 = engines["int+pp"]
    for it in range(max(4, rounds // 3)):
        B = int(rng.choice([1, 2, 3, 147, 149, 295, 297, 591, 593, 1185, int(rng.integers(1, 1500))]))
        T = int(rng.choice([1, 2, 3, 5, 8]))
        gs, xs = s.scalar(B, T), s.uniform_q(B, T, 1)
        rs, ys = s.small(B, T), s.gaussian(B, T)
        rp, yp, d = s.small(B), s.gaussian(B), s.challenge(B)
        a, b = e0.sum_commit(gs, xs, rp, rs, ys, yp), plain.sum_commit(gs, xs, rp, rs, ys, yp)
        for k in ("xp", "cp", "cs", "ts", "tp", "u"):
            assert (a[k] == b[k]).all(), ("sum_commit", k, B, T)
        zs, zp = e0.sum_respond(ys, yp, rs, rp, d)
        args = [zs, zp, a["cs"], a["cp"], gs, a["ts"], a["tp"], a["u"], d]
        assert UB(e0.sum_verify(*args), B).all() and UB(plain.sum_verify(*args), B).all(), ("sum_verify", B, T)
        bad = rng.random(B) < 0.3
        u2 = a["u"].copy(); u2[bad, ..., 11] += 1
        ts2 = a["ts"].copy(); ts2[bad, T - 1, ..., 3] -= 1
        # (a changed c1 / c2 row and a changed challenge go through the rotation sums of the first equations)
        cs2 = a["cs"].copy(); cs2[bad, int(rng.integers(0, T)), int(rng.integers(0, 2)), int(rng.integers(0, N))] += 1
        cp2 = a["cp"].copy(); cp2[bad, 1, int(rng.integers(0, N))] -= 1
        d2 = d.copy(); d2[bad, int(rng.integers(0, N))] ^= 1
        for args2 in ([zs, zp, a["cs"], a["cp"], gs, a["ts"], a["tp"], u2, d], [zs, zp, a["cs"], a["cp"], gs, ts2, a["tp"], a["u"], d],
                      [zs, zp, cs2, a["cp"], gs, a["ts"], a["tp"], a["u"], d], [zs, zp, a["cs"], cp2, gs, a["ts"], a["tp"], a["u"], d],
                      [zs, zp, a["cs"], a["cp"], gs, a["ts"], a["tp"], a["u"], d2]):
            v1, v2 = UB(e0.sum_verify(*args2), B), UB(plain.sum_verify(*args2), B)
            assert (v1 == ~bad).all() and (v2 == ~bad).all(), ("sum_verify tampered", B, T)
        if T == 1:
            g, x = gs[:, 0], xs[:, 0]
            la, lb = e0.linear_commit(g, x, rp, rs[:, 0], ys[:, 0], yp), plain.linear_commit(g, x, rp, rs[:, 0], ys[:, 0], yp)
            for k in ("gx", "cp", "c", "t", "tp", "u"):
                assert (la[k] == lb[k]).all(), ("linear_commit", k, B)
            z, zq = e0.linear_respond(ys[:, 0], yp, rs[:, 0], rp, d)
            c2t = la["c"].copy(); c2t[bad, 1, int(rng.integers(0, N))] += 1
            for cc, dd, want in ((la["c"], d, np.ones(B, bool)), (c2t, d, ~bad), (la["c"], d2, ~bad)):
                largs = [z, zq, cc, la["cp"], g, la["t"], la["tp"], la["u"], dd]
                v1, v2 = UB(e0.linear_verify(*largs), B), UB(plain.linear_verify(*largs), B)
                assert (v1 == want).all() and (v2 == want).all(), ("linear_verify", B)
        print(f"sum round {it:3d} B={B:6d} T={T} ok", flush=True)
    plain.close()
    for e in engines.values():
        e.close()
    print("SOAK PASSED")


if __name__ == "__main__":
    main()
