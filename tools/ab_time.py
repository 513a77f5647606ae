"""A/B timing of one phase under several engine settings (RZK_TEST_LOWERING / RZK_TUNE), device resident, CUDA events.
usage: python tools/ab_time.py <phase> [B] -- "<env assignments>" "<env assignments>" ...
   e.g. python tools/ab_time.py open_verify 65536 -- "" "RZK_TEST_LOWERING=norot" "RZK_TUNE=pp=2"
Development helper (keeps consolidated what used to be one script per experiment)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("ring-zk_b200")
engine = importlib.import_module("ring-zk_b200.engine")
N = 512


def make(assign):
    env = dict(a.split("=", 1) for a in assign.split()) if assign.strip() else {}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return engine.Engine(N=N, device=0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    phase = sys.argv[1]
    sep = sys.argv.index("--")
    B = int(sys.argv[2]) if sep > 2 else 1 << 16
    settings = sys.argv[sep + 1:] or [""]
    dev = torch.device("cuda:0")
    s = pkg.synth.Synth(1, N=N)
    key = s.key()
    st = torch.cuda.current_stream().cuda_stream
    T = lambda a: torch.from_numpy(a).to(dev)
    x, r, y, d = T(s.message(B)), T(s.small(B)), T(s.gaussian(B)), T(s.challenge(B))
    c = torch.empty((B, 2, N), dtype=torch.int32, device=dev)
    t = torch.empty((B, 1, N), dtype=torch.int32, device=dev)
    z = torch.empty((B, 3, N), dtype=torch.int32, device=dev)
    flags = torch.zeros(B, dtype=torch.int32, device=dev)
    e0 = engine.Engine(N=N, device=0)
    e0.set_key_blocks(*key)
    e0.dev("open_commit_batch", B, x, r, y, c, t, flags, stream=st)
    e0.dev("open_respond_batch", B, y, r, d, z, stream=st)
    torch.cuda.synchronize()
    E = lambda *sh: torch.empty(sh, dtype=torch.int32, device=dev)
    big = phase.startswith(("linear", "sum"))
    if big:
        BL = min(B, 1 << 14)
        g, rpl, ypl = T(s.scalar(BL)), T(s.small(BL)), T(s.gaussian(BL))
        xl, rl, yl, dl = x[:BL].contiguous(), r[:BL].contiguous(), y[:BL].contiguous(), d[:BL].contiguous()
        gx, cp, cl, tl, tpl, u = E(BL, 1, N), E(BL, 2, N), E(BL, 2, N), E(BL, 1, N), E(BL, 1, N), E(BL, 1, N)
        zl, zpl = E(BL, 3, N), E(BL, 3, N)
        fl = torch.zeros(BL, dtype=torch.int32, device=dev)
        e0.dev("linear_commit_batch", BL, g, xl, rpl, rl, yl, ypl, gx, cp, cl, tl, tpl, u, fl, stream=st)
        e0.dev("linear_respond_batch", BL, yl, ypl, rl, rpl, dl, zl, zpl, stream=st)
        BS, TT = min(B, 1 << 12), 64
        gs, xs = T(s.scalar(BS, TT)), T(s.uniform_q(BS, TT, 1))
        rs, ys = T(s.small(BS, TT)), T(s.gaussian(BS, TT))
        rps, yps, ds = T(s.small(BS)), T(s.gaussian(BS)), T(s.challenge(BS))
        xp, cps, css, tss, tps, us = E(BS, 1, N), E(BS, 2, N), E(BS, TT, 2, N), E(BS, TT, 1, N), E(BS, 1, N), E(BS, 1, N)
        zs, zps = E(BS, TT, 3, N), E(BS, 3, N)
        fs = torch.zeros(BS, dtype=torch.int32, device=dev)
        e0.dev("sum_commit_batch", BS, TT, gs, xs, rps, rs, ys, yps, xp, cps, css, tss, tps, us, fs, stream=st)
        e0.dev("sum_respond_batch", BS, TT, ys, yps, rs, rps, ds, zs, zps, stream=st)
        torch.cuda.synchronize()
    for assign in settings:
        eng = make(assign)
        eng.set_key_blocks(*key)
        if big:
            fns = {
                "linear_commit": (BL, lambda: eng.dev("linear_commit_batch", BL, g, xl, rpl, rl, yl, ypl, gx, cp, cl, tl, tpl, u, fl, stream=st)),
                "linear_verify": (BL, lambda: eng.dev("linear_verify_batch", BL, zl, zpl, cl, cp, g, tl, tpl, u, dl, fl, stream=st)),
                "sum_commit": (BS, lambda: eng.dev("sum_commit_batch", BS, TT, gs, xs, rps, rs, ys, yps, xp, cps, css, tss, tps, us, fs, stream=st)),
                "sum_verify": (BS, lambda: eng.dev("sum_verify_batch", BS, TT, zs, zps, css, cps, gs, tss, tps, us, ds, fs, stream=st)),
            }
            nB, fn = fns[phase]
            flags = fl if phase.startswith("linear") else fs
            run(phase, nB, assign, fn, flags)
            eng.close()
            continue
        fns = {
            "commit": lambda: eng.dev("commit_batch", B, x, r, c, flags, stream=st),
            "open_commit": lambda: eng.dev("open_commit_batch", B, x, r, y, c, t, flags, stream=st),
            "open_respond": lambda: eng.dev("open_respond_batch", B, y, r, d, z, stream=st),
            "open_verify": lambda: eng.dev("open_verify_batch", B, z, t, c, 2, d, flags, stream=st),
        }
        run(phase, B, assign, fns[phase], flags)
        eng.close()


def run(phase, B, assign, fn, flags):
    flags.zero_()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    iters = 10
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    print(f"{phase:12s} B={B} [{assign or 'default':40s}] min {ts[0]:.4f} ms  med {ts[iters // 2]:.4f} ms  {B / ts[iters // 2] / 1e3:8.3f} M/s  flags_any={int(flags.any())}", flush=True)


if __name__ == "__main__":
    main()
