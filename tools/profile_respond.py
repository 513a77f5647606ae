"""Short program for ncu: response launches only (device-resident data)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("ring-zk_b200")
engine = importlib.import_module("ring-zk_b200.engine")
N, B = 512, 1 << 16
dev = torch.device("cuda:0")
s = pkg.synth.Synth(3, N=N)
eng = engine.Engine(N=N, device=0)
eng.set_key_blocks(*s.key())
T = lambda a: torch.from_numpy(a).to(dev)
r, y, d = T(s.small(B)), T(s.gaussian(B)), T(s.challenge(B))
z = torch.empty((B, 3, N), dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    eng.dev("open_respond_batch", B, y, r, d, z, stream=st)
torch.cuda.synchronize()
print("done", int(z[0, 0, 0]))
