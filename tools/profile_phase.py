"""Short program for ncu / launch lists: `reps` launches of the named phases on device-resident data.
usage: python tools/profile_phase.py <phase[,phase...]> [reps] [B]
phases: commit  open_commit  open_respond  open_verify  linear (2^14 commit+respond+verify)  sum (T = 64, commit+respond+verify)
(replaces the per-phase profile_*.py scripts of round 1)"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("ring-zk_b200")
engine = importlib.import_module("ring-zk_b200.engine")
N = 512


def main():
    phases = sys.argv[1].split(",")
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 16
    dev = torch.device("cuda:0")
    s = pkg.synth.Synth(3, N=N)
    eng = engine.Engine(N=N, device=0)
    eng.set_key_blocks(*s.key())
    T = lambda a: torch.from_numpy(a).to(dev)
    E = lambda *sh: torch.empty(sh, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    x, r, y, d = T(s.message(B)), T(s.small(B)), T(s.gaussian(B)), T(s.challenge(B))
    c, t, z = E(B, 2, N), E(B, 1, N), E(B, 3, N)
    flags = torch.zeros(B, dtype=torch.int32, device=dev)
    eng.dev("open_commit_batch", B, x, r, y, c, t, flags, stream=st)
    eng.dev("open_respond_batch", B, y, r, d, z, stream=st)
    torch.cuda.synchronize()
    for ph in phases:
        if ph == "commit":
            fn = lambda: eng.dev("commit_batch", B, x, r, c, flags, stream=st)
        elif ph == "open_commit":
            fn = lambda: eng.dev("open_commit_batch", B, x, r, y, c, t, flags, stream=st)
        elif ph == "open_respond":
            fn = lambda: eng.dev("open_respond_batch", B, y, r, d, z, stream=st)
        elif ph == "open_verify":
            fn = lambda: eng.dev("open_verify_batch", B, z, t, c, 2, d, flags, stream=st)
        elif ph == "linear":
            BL = min(B, 1 << 14)
            g, rpl, ypl = T(s.scalar(BL)), T(s.small(BL)), T(s.gaussian(BL))
            xl, rl, yl, dl = x[:BL].contiguous(), r[:BL].contiguous(), y[:BL].contiguous(), d[:BL].contiguous()
            gx, cp, cl, tl, tpl, u = E(BL, 1, N), E(BL, 2, N), E(BL, 2, N), E(BL, 1, N), E(BL, 1, N), E(BL, 1, N)
            zl, zpl = E(BL, 3, N), E(BL, 3, N)
            fl = torch.zeros(BL, dtype=torch.int32, device=dev)

            def fn():
                eng.dev("linear_commit_batch", BL, g, xl, rpl, rl, yl, ypl, gx, cp, cl, tl, tpl, u, fl, stream=st)
                eng.dev("linear_respond_batch", BL, yl, ypl, rl, rpl, dl, zl, zpl, stream=st)
                eng.dev("linear_verify_batch", BL, zl, zpl, cl, cp, g, tl, tpl, u, dl, fl, stream=st)
        elif ph == "sum":
            BS, TT = min(B, 1 << 12), 64
            gs, xs = T(s.scalar(BS, TT)), T(s.uniform_q(BS, TT, 1))
            rs, ys = T(s.small(BS, TT)), T(s.gaussian(BS, TT))
            rps, yps, ds = T(s.small(BS)), T(s.gaussian(BS)), T(s.challenge(BS))
            xp, cps, css, tss, tps, us = E(BS, 1, N), E(BS, 2, N), E(BS, TT, 2, N), E(BS, TT, 1, N), E(BS, 1, N), E(BS, 1, N)
            zs, zps = E(BS, TT, 3, N), E(BS, 3, N)
            fs = torch.zeros(BS, dtype=torch.int32, device=dev)

            def fn():
                eng.dev("sum_commit_batch", BS, TT, gs, xs, rps, rs, ys, yps, xp, cps, css, tss, tps, us, fs, stream=st)
                eng.dev("sum_respond_batch", BS, TT, ys, yps, rs, rps, ds, zs, zps, stream=st)
                eng.dev("sum_verify_batch", BS, TT, zs, zps, css, cps, gs, tss, tps, us, ds, fs, stream=st)
        else:
            raise SystemExit(f"unknown phase {ph}")
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
    print("flags", int(flags.any()))


if __name__ == "__main__":
    main()
