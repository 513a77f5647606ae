"""Quick device-resident timing of every phase (CUDA events on torch's current stream).
Development helper; bench.py is the contract the driver runs."""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("ring-zk_b200")
engine = importlib.import_module("ring-zk_b200.engine")

N = 512


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    return min(ts), sum(ts) / len(ts)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
    dev = torch.device("cuda:0")
    s = pkg.synth.Synth(1, N=N)
    eng = engine.Engine(N=N, device=0)
    eng.set_key_blocks(*s.key())
    st = torch.cuda.current_stream().cuda_stream
    T = lambda a: torch.from_numpy(a).to(dev)
    x, r, y, d = T(s.message(B)), T(s.small(B)), T(s.gaussian(B)), T(s.challenge(B))
    c = torch.empty((B, 2, N), dtype=torch.int32, device=dev)
    t = torch.empty((B, 1, N), dtype=torch.int32, device=dev)
    z = torch.empty((B, 3, N), dtype=torch.int32, device=dev)
    flags = torch.zeros(B, dtype=torch.int32, device=dev)
    res = {}
    mn, av = timeit(lambda: eng.dev("commit_batch", B, x, r, c, flags, stream=st))
    res["commit"] = (mn, B / mn * 1e3)
    mn, av = timeit(lambda: eng.dev("open_commit_batch", B, x, r, y, c, t, flags, stream=st))
    res["open_commit"] = (mn, B / mn * 1e3)
    mn, av = timeit(lambda: eng.dev("open_respond_batch", B, y, r, d, z, stream=st))
    res["open_respond"] = (mn, B / mn * 1e3)
    mn, av = timeit(lambda: eng.dev("open_verify_batch", B, z, t, c, 2, d, flags, stream=st))
    res["open_verify"] = (mn, B / mn * 1e3)
    print("flags any:", int(flags.any()))
    for k, (ms, rate) in res.items():
        print(f"{k:14s} B={B}  {ms:8.3f} ms   {rate / 1e6:8.3f} M items/s")
    # linear at 2^14, sum T=64 at 2^10
    BL = min(B, 1 << 14)
    g = T(s.scalar(BL)); xl = x[:BL].contiguous(); rl = r[:BL].contiguous(); rpl = T(s.small(BL))
    yl = y[:BL].contiguous(); ypl = T(s.gaussian(BL)); dl = d[:BL].contiguous()
    E = lambda *sh: torch.empty(sh, dtype=torch.int32, device=dev)
    gx, cp, cl, tl, tpl, u = E(BL, 1, N), E(BL, 2, N), E(BL, 2, N), E(BL, 1, N), E(BL, 1, N), E(BL, 1, N)
    zl, zpl = E(BL, 3, N), E(BL, 3, N)
    fl = torch.zeros(BL, dtype=torch.int32, device=dev)
    mn, _ = timeit(lambda: eng.dev("linear_commit_batch", BL, g, xl, rpl, rl, yl, ypl, gx, cp, cl, tl, tpl, u, fl, stream=st))
    print(f"linear_commit  B={BL} {mn:8.3f} ms {BL / mn * 1e3 / 1e6:8.3f} M/s")
    mn, _ = timeit(lambda: eng.dev("linear_respond_batch", BL, yl, ypl, rl, rpl, dl, zl, zpl, stream=st))
    print(f"linear_respond B={BL} {mn:8.3f} ms {BL / mn * 1e3 / 1e6:8.3f} M/s")
    mn, _ = timeit(lambda: eng.dev("linear_verify_batch", BL, zl, zpl, cl, cp, g, tl, tpl, u, dl, fl, stream=st))
    print(f"linear_verify  B={BL} {mn:8.3f} ms {BL / mn * 1e3 / 1e6:8.3f} M/s   flags any: {int(fl.any())}")
    BS, TT = min(B, 1 << 12), 64
    gs, xs = T(s.scalar(BS, TT)), T(s.uniform_q(BS, TT, 1))
    rs, ys = T(s.small(BS, TT)), T(s.gaussian(BS, TT))
    rps, yps, ds = T(s.small(BS)), T(s.gaussian(BS)), T(s.challenge(BS))
    xp, cps, css, tss, tps, us = E(BS, 1, N), E(BS, 2, N), E(BS, TT, 2, N), E(BS, TT, 1, N), E(BS, 1, N), E(BS, 1, N)
    zs, zps = E(BS, TT, 3, N), E(BS, 3, N)
    fs = torch.zeros(BS, dtype=torch.int32, device=dev)
    mn, _ = timeit(lambda: eng.dev("sum_commit_batch", BS, TT, gs, xs, rps, rs, ys, yps, xp, cps, css, tss, tps, us, fs, stream=st), iters=3, warm=1)
    print(f"sum_commit  T={TT} B={BS} {mn:8.3f} ms {BS / mn * 1e3 / 1e3:8.3f} K/s")
    mn, _ = timeit(lambda: eng.dev("sum_respond_batch", BS, TT, ys, yps, rs, rps, ds, zs, zps, stream=st), iters=3, warm=1)
    print(f"sum_respond T={TT} B={BS} {mn:8.3f} ms {BS / mn * 1e3 / 1e3:8.3f} K/s")
    mn, _ = timeit(lambda: eng.dev("sum_verify_batch", BS, TT, zs, zps, css, cps, gs, tss, tps, us, ds, fs, stream=st), iters=3, warm=1)
    print(f"sum_verify  T={TT} B={BS} {mn:8.3f} ms {BS / mn * 1e3 / 1e3:8.3f} K/s   flags any: {int(fs.any())}")
    # host API end to end (pinned)
    xh, rh = s.message(B), s.small(B)
    t0 = time.perf_counter(); eng.commit(xh, rh); t1 = time.perf_counter()
    t0 = time.perf_counter(); eng.commit(xh, rh); t1 = time.perf_counter()
    print(f"host commit (pageable) B={B}: {(t1 - t0) * 1e3:.2f} ms  {B / (t1 - t0) / 1e6:.3f} M/s")


if __name__ == "__main__":
    main()
