"""Single-call latency (B = 1) of every protocol phase through the host C ABI, next to the CPU oracle on one thread.
These are the twelve calls the reference's own Criterion benches time (benches/bench.rs:35-305, N = 512, Sum with 4 terms)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("ring-zk_b200")
engine = importlib.import_module("ring-zk_b200.engine")
from oracle import oracle as orc
N = 512


def med(fn, n):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e6


def main():
    s = pkg.synth.Synth(5, N=N)
    a1p, a2p = s.key()
    eng = engine.Engine(N=N, device=0)
    eng.set_key_blocks(a1p, a2p)
    o = orc.Oracle(orc.Params(N=N), a1p, a2p)
    B, T = 1, 4
    x, r, y, d = s.message(B), s.small(B), s.gaussian(B), s.challenge(B)
    g, rp, yp = s.scalar(B), s.small(B), s.gaussian(B)
    gs, xs, rs, ys = s.scalar(B, T), s.uniform_q(B, T, 1), s.small(B, T), s.gaussian(B, T)
    c, t, _ = eng.open_commit(x, r, y); z = eng.open_respond(y, r, d); c1 = np.ascontiguousarray(c[:, :1])
    L = eng.linear_commit(g, x, rp, r, y, yp); lz, lzp = eng.linear_respond(y, yp, r, rp, d)
    S = eng.sum_commit(gs, xs, rp, rs, ys, yp); zs, zp = eng.sum_respond(ys, yp, rs, rp, d)
    calls = {
        "commit": (lambda: eng.commit(x, r), lambda: o.commit_batch(x, r, 1)),
        "open_proof_commit": (lambda: eng.open_commit(x, r, y), lambda: o.open_commit_batch(x, r, y, 1)),
        "open_proof_create_response": (lambda: eng.open_respond(y, r, d), lambda: o.open_respond_batch(y, r, d, 1)),
        "open_proof_verify": (lambda: eng.open_verify(z, t, c1, d), lambda: o.open_verify_batch(z, t, c1, d, 1)),
        "linear_proof_commit": (lambda: eng.linear_commit(g, x, rp, r, y, yp), lambda: o.linear_commit_batch(g, x, rp, r, y, yp, 1)),
        "linear_proof_create_response": (lambda: eng.linear_respond(y, yp, r, rp, d), lambda: o.linear_respond_batch(y, yp, r, rp, d, 1)),
        "linear_proof_verify": (lambda: eng.linear_verify(lz, lzp, L["c"], L["cp"], g, L["t"], L["tp"], L["u"], d),
                                lambda: o.linear_verify_batch(lz, lzp, L["c"], L["cp"], g, L["t"], L["tp"], L["u"], d, 1)),
        "sum_proof_commit": (lambda: eng.sum_commit(gs, xs, rp, rs, ys, yp), lambda: o.sum_commit_batch(gs, xs, rp, rs, ys, yp, 1)),
        "sum_proof_create_response": (lambda: eng.sum_respond(ys, yp, rs, rp, d), lambda: o.sum_respond_batch(ys, yp, rs, rp, d, 1)),
        "sum_proof_verify": (lambda: eng.sum_verify(zs, zp, S["cs"], S["cp"], gs, S["ts"], S["tp"], S["u"], d),
                             lambda: o.sum_verify_batch(zs, zp, S["cs"], S["cp"], gs, S["ts"], S["tp"], S["u"], d, 1)),
    }
    out = {}
    for name, (gpu, cpu) in calls.items():
        for _ in range(5):
            gpu()
        out[name] = {"gpu_us": round(med(gpu, 40), 1), "cpu_oracle_1thread_us": round(med(cpu, 5), 1)}
        print(f"{name:30s} gpu {out[name]['gpu_us']:9.1f} us   cpu (C restatement, 1 thread) {out[name]['cpu_oracle_1thread_us']:10.1f} us", flush=True)
    eng.close()
    return out


if __name__ == "__main__":
    main()
