"""Short program for ncu: Open verify launches only (device-resident data)."""
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
x, r, y, d = T(s.message(B)), T(s.small(B)), T(s.gaussian(B)), T(s.challenge(B))
c = torch.empty((B, 2, N), dtype=torch.int32, device=dev)
t = torch.empty((B, 1, N), dtype=torch.int32, device=dev)
z = torch.empty((B, 3, N), dtype=torch.int32, device=dev)
flags = torch.zeros(B, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
eng.dev("open_commit_batch", B, x, r, y, c, t, flags, stream=st)
eng.dev("open_respond_batch", B, y, r, d, z, stream=st)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    eng.dev("open_verify_batch", B, z, t, c, 2, d, flags, stream=st)
torch.cuda.synchronize()
print("flags", int(flags.any()))
