"""Differential soak test on the GPU: random batch sizes, the commitment kernels (integer with / without phase mixing,
FP64, hybrid) against each other, the rotation-kernel response against the NTT response, honest proofs verify.
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
    engines = {"int+pp": make({"RZK_COMMIT_MODE": "0", "RZK_COMMIT_PP": "2"}), "int": make({"RZK_COMMIT_MODE": "0", "RZK_COMMIT_PP": "9"}),
               "f64": make({"RZK_COMMIT_MODE": "1"}), "hybrid": make({"RZK_COMMIT_MODE": "2"}), "ntt-respond": make({"RZK_NO_SPARSE": "1"})}
    for e in engines.values():
        e.set_key_blocks(a1p, a2p)
    UB = engine.unpack_bitmap
    for it in range(rounds):
        B = int(rng.choice([1, 2, 7, 8, 9, 63, 148, 149, 1000, 2367, 2368, 2369, 4095, 4096, 4097, 8191, 8193, int(rng.integers(1, 20000))]))
        x, r, y, d = s.message(B, ragged=bool(it & 1)), s.small(B), s.gaussian(B), s.challenge(B)
        if it % 3 == 0:
            r = rng.integers(-3, 4, size=r.shape).astype(np.int8)
        ref = None
        for name in ("int+pp", "int", "f64", "hybrid"):
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
        z2 = engines["ntt-respond"].open_respond(y, r, d)
        assert (z == z2).all(), ("respond", B, it)
        v = UB(e0.open_verify(z, t, np.ascontiguousarray(c[:, :1]), d), B)
        assert v.all(), ("verify", B, it)
        assert UB(e0.commitment_verify(c, x, r), B).all()
        print(f"round {it:3d} B={B:6d} ok", flush=True)
    for e in engines.values():
        e.close()
    print("SOAK PASSED")


if __name__ == "__main__":
    main()
