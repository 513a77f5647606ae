"""GPU: the optional on-device samplers (rzk_sample_*_dev) against their numpy restatement, and a
prover flow whose r and y never leave the device."""
import importlib

import numpy as np
import pytest

import philox_ref as pr

pytestmark = pytest.mark.gpu
engine = importlib.import_module("ring-zk_b200.engine")
synth = importlib.import_module("ring-zk_b200.synth")
from oracle import oracle as orc  # noqa: E402  (checker only)

N = 512


def test_samplers_match_the_numpy_restatement():
    import torch
    dev = torch.device("cuda:0")
    eng = engine.Engine(N=N, device=0)
    try:
        st = torch.cuda.current_stream().cuda_stream
        for b in (1, 3, 127):
            out = torch.empty((50, N), dtype=torch.int8, device=dev)
            eng.dev("sample_small", 50, b, 0x1234567890ABCDEF, 7, out, stream=st)
            assert (out.cpu().numpy() == pr.sample_small(50, b, 0x1234567890ABCDEF, 7)).all()
        d = torch.empty((300, N), dtype=torch.int8, device=dev)
        eng.dev("sample_challenge", 300, 36, 99, 3, d, stream=st)
        assert (d.cpu().numpy() == pr.sample_challenge(300, 36, 99, 3)).all()
        y = torch.empty((400, N), dtype=torch.int32, device=dev)
        eng.dev("sample_gaussian", 400, 15444.0, 2024, 5, y, stream=st)
        yn, yr = y.cpu().numpy().astype(np.int64), pr.sample_gaussian(400, 15444.0, 2024, 5).astype(np.int64)
        diff = np.abs(yn - yr)
        assert diff.max() <= 1 and (diff != 0).mean() < 1e-4        # last-ulp differences of log / cos only
        v = yn.astype(np.float64).ravel()
        assert abs(v.std() / 15444.0 - 1) < 0.01 and abs(v.mean()) < 5 * 15444.0 / np.sqrt(v.size)
    finally:
        eng.close()


def test_prover_with_device_resident_randomness():
    """commit -> challenge -> response -> verify with r, y and d sampled on the device: only x goes up, c / t / z come
    down; the transcript verifies on the engine and on the oracle, and r and y are what the seed says they are."""
    import torch
    dev = torch.device("cuda:0")
    s = synth.Synth(3, N=N)
    a1p, a2p = s.key()
    eng = engine.Engine(N=N, device=0)
    try:
        eng.set_key_blocks(a1p, a2p)
        o = orc.Oracle(orc.Params(N=N), a1p, a2p)
        B, seed = 4500, 0xC0FFEE
        st = torch.cuda.current_stream().cuda_stream
        x = torch.from_numpy(s.message(B)).to(dev)
        r = torch.empty((B, 3, N), dtype=torch.int8, device=dev)
        y = torch.empty((B, 3, N), dtype=torch.int32, device=dev)
        d = torch.empty((B, N), dtype=torch.int8, device=dev)
        c = torch.empty((B, 2, N), dtype=torch.int32, device=dev)
        t = torch.empty((B, 1, N), dtype=torch.int32, device=dev)
        z = torch.empty((B, 3, N), dtype=torch.int32, device=dev)
        flags = torch.zeros(B, dtype=torch.int32, device=dev)
        eng.dev("sample_small", 3 * B, 1, seed, 1, r, stream=st)
        eng.dev("sample_gaussian", 3 * B, float(eng.sigma()), seed, 2, y, stream=st)
        eng.dev("open_commit_batch", B, x, r, y, c, t, flags, stream=st)
        eng.dev("sample_challenge", B, 36, seed + 1, 3, d, stream=st)             # the verifier's draw
        eng.dev("open_respond_batch", B, y, r, d, z, stream=st)
        eng.dev("open_verify_batch", B, z, t, c, 2, d, flags, stream=st)
        torch.cuda.synchronize()
        assert int(flags.any()) == 0
        idx = [0, 1, B - 1]
        rn = pr.sample_small(3 * B, 1, seed, 1).reshape(B, 3, N)
        assert (r.cpu().numpy() == rn).all()
        xs, ys, ds = x[idx].cpu().numpy(), y[idx].cpu().numpy(), d[idx].cpu().numpy()
        c_o, t_o, _ = o.open_commit_batch(xs, rn[idx], ys)
        z_o = o.open_respond_batch(ys, rn[idx], ds)
        assert (c[idx].cpu().numpy() == c_o).all() and (t[idx].cpu().numpy() == t_o).all() and (z[idx].cpu().numpy() == z_o).all()
        assert o.open_verify_batch(z_o, t_o, c_o[:, :1], ds).all()
    finally:
        eng.close()
