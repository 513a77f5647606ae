"""Device-resident timing of one Sum-proof pass (T = 64) at a small instance count (development helper)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("ring-zk_b200")
engine = importlib.import_module("ring-zk_b200.engine")
N, TT = 512, 64
BS = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
s = pkg.synth.Synth(5, N=N)
eng = engine.Engine(N=N, device=0)
eng.set_key_blocks(*s.key())
T = lambda a: torch.from_numpy(a).to(dev)
E = lambda *sh: torch.empty(sh, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
gs, xs = T(s.scalar(BS, TT)), T(s.uniform_q(BS, TT, 1))
rs, ys = T(s.small(BS, TT)), T(s.gaussian(BS, TT))
rps, yps, ds = T(s.small(BS)), T(s.gaussian(BS)), T(s.challenge(BS))
xp, cps, css, tss, tps, us = E(BS, 1, N), E(BS, 2, N), E(BS, TT, 2, N), E(BS, TT, 1, N), E(BS, 1, N), E(BS, 1, N)
zs, zps = E(BS, TT, 3, N), E(BS, 3, N)
fs = torch.zeros(BS, dtype=torch.int32, device=dev)


def tm(fn, it=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


c = tm(lambda: eng.dev("sum_commit_batch", BS, TT, gs, xs, rps, rs, ys, yps, xp, cps, css, tss, tps, us, fs, stream=st))
eng.dev("sum_respond_batch", BS, TT, ys, yps, rs, rps, ds, zs, zps, stream=st)
v = tm(lambda: eng.dev("sum_verify_batch", BS, TT, zs, zps, css, cps, gs, tss, tps, us, ds, fs, stream=st))
print(f"sum T=64 B={BS}: commit {c:.3f} ms, verify {v:.3f} ms, flags any {int(fs.any())}, segments {'off' if os.environ.get('RZK_NO_SEGMENTS') else 'on'}")
