"""Host-path (pinned host buffers -> C ABI -> host buffers) throughput for several pipeline chunk sizes.
Development helper; bench.py's `e2e` is the reported number.  Usage: python tools/e2e_sweep.py [B]"""
import importlib
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N = 512


def one(chunk, B):
    import numpy as np
    import torch
    pkg = importlib.import_module("ring-zk_b200")
    engine = importlib.import_module("ring-zk_b200.engine")
    s = pkg.synth.Synth(1, N=N)
    eng = engine.Engine(N=N, device=0)
    eng.set_key_blocks(*s.key())

    def pin(a):
        t = torch.from_numpy(a).pin_memory()
        return t, t.numpy()
    xt, x = pin(s.message(B)); rt, r = pin(s.small(B)); yt, y = pin(s.gaussian(B)); dt_, d = pin(s.challenge(B))
    ct, c = pin(np.zeros((B, 2, N), np.int32)); tt, t = pin(np.zeros((B, 1, N), np.int32))
    zt, z = pin(np.zeros((B, 3, N), np.int32)); okt, ok = pin(np.zeros((B + 7) // 8, np.uint8))
    c1t, c1 = pin(np.zeros((B, 1, N), np.int32))

    def tm(fn, it=5):
        fn(); fn()
        t0 = time.perf_counter()
        for _ in range(it):
            fn()
        return (time.perf_counter() - t0) / it
    P = lambda a: a.ctypes.data
    dtc = tm(lambda: eng._call("rzk_commit_batch", B, P(x), P(r), P(c), P(ok)))
    eng._call("rzk_open_commit_batch", B, P(x), P(r), P(y), P(c), P(t), P(ok))
    eng._call("rzk_open_respond_batch", B, P(y), P(r), P(d), P(z))
    c1[:] = c[:, :1]
    dtv = tm(lambda: eng._call("rzk_open_verify_batch", B, P(z), P(t), P(c1), P(d), P(ok)))
    assert np.unpackbits(ok, bitorder="little")[:B].all()
    print(f"chunk={chunk:6d} commit {B / dtc / 1e6:7.3f} M/s ({dtc * 1e3:6.2f} ms, {B * (3584 + 4096) / dtc / 1e9:5.1f} GB/s both ways)  "
          f"open_verify {B / dtv / 1e6:7.3f} M/s ({dtv * 1e3:6.2f} ms, {B * 10752 / dtv / 1e9:5.1f} GB/s h2d)", flush=True)
    eng.close()


if __name__ == "__main__":
    if len(sys.argv) > 2:
        one(int(sys.argv[1]), int(sys.argv[2]))
    else:
        B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
        for chunk in (1024, 2048, 4096, 8192, 16384):
            env = dict(os.environ, RZK_CHUNK_ITEMS=str(chunk))
            subprocess.run([sys.executable, __file__, str(chunk), str(B)], env=env, check=False)
