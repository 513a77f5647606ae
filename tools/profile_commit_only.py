"""Short program for ncu / clock checks: commit launches only, device-resident data.
usage: profile_commit_only.py [launches] [seconds-of-looping for a clock/power sample]"""
import importlib, os, subprocess, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("ring-zk_b200")
engine = importlib.import_module("ring-zk_b200.engine")
N, B = 512, 1 << 16
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
dev = torch.device("cuda:0")
s = pkg.synth.Synth(3, N=N)
eng = engine.Engine(N=N, device=0)
eng.set_key_blocks(*s.key())
T = lambda a: torch.from_numpy(a).to(dev)
x, r = T(s.message(B)), T(s.small(B))
c = torch.empty((B, 2, N), dtype=torch.int32, device=dev)
flags = torch.zeros(B, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(n):
    eng.dev("commit_batch", B, x, r, c, flags, stream=st)
torch.cuda.synchronize()
if secs > 0:
    t0 = time.time(); k = 0
    samples = []
    while time.time() - t0 < secs:
        for _ in range(50):
            eng.dev("commit_batch", B, x, r, c, flags, stream=st)
        k += 50
        torch.cuda.synchronize()
        if k % 500 == 0:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader", "-i", "0"],
                                 capture_output=True, text=True).stdout.strip()
            samples.append(out)
    dt = time.time() - t0
    print(f"mode={os.environ.get('RZK_COMMIT_MODE','0')} sustained {k * B / dt / 1e6:.1f} M/s over {dt:.1f}s; smi samples: {samples[:2]} ... {samples[-2:]}")
print("flags", int(flags.any()))
