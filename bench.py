#!/usr/bin/env python
"""bench.py -- headline benchmark of the R_q hot path (BASELINE.json: commitments/s and
open-proof verifies/s at N=512, batch 2^16, on 1/2/4/8 B200).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA engine)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU restatement
                                                           # of the reference (oracle/), all host cores

One "step" = one pass of the hot path over one batch of 2^16 synthetic items per GPU
(BASELINE.json configs[1], "Batched commitment generation: 2^16 random messages at N=512,
one shared commitment key"), inputs resident in HBM.  Batches are sharded by item across
ranks (weak scaling: 2^16 items per rank); the only inter-GPU traffic is the NCCL all-gather
of the per-shard ok/verify bitmaps.  `value` is whole-job commitments/s; the same line carries
open-proof verifies/s (configs[2]), the roofline of the dominant kernel, the CPU baseline
timed on this box, and the end-to-end number through the host C ABI.
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N = 512
BATCH = 1 << 16
# algorithmic bytes (4 B per coefficient of every polynomial that must cross HBM) and modular multiplies
# per unit of work
ALG = {
    "commit": {"bytes": 12288, "mulmods": 21504},            # 4 polys in + 2 out
    "open_verify": {"bytes": 12288, "mulmods": 15872},       # 6 polys in
    "open_instance": {"bytes": 53248, "mulmods": 53248},     # commit + respond + verify
    "linear": {"bytes": 112640, "mulmods": 204288},
    "sum64": {"bytes": 3596288, "mulmods": 7268352},
}
ALG_BYTES_COMMIT = ALG["commit"]["bytes"]
ALG_MULMODS_COMMIT = ALG["commit"]["mulmods"]
# Integer denominators, measured by the builder on this pool's B200s with tools/imad_bench.cu (MEASURED_PEAKS.json carries
# no integer peak): the plain 32-bit multiply rate R_IMAD of SURVEY 8(d)'s formula frac = 3 * mulmods * rate / R_IMAD,
# and the rate of a complete Shoup modular multiplication (IMAD.HI + 2 IMAD: IMAD.HI issues at 24 /clk/SM, 2.6 x an IMAD)
MEASURED_IMAD_TPS = 18.1        # T IMAD/s, profiles/r1_imad_bench.jsonl
MEASURED_MULMOD_TPS = 4.617     # T Shoup-mulmods/s, same file
METRIC = "commitments/s"
UNIT = "commitments/s"
COMMIT_KERNEL = ("rzk_vm_kernel<1, MODE_SPLITKEY_S, SPCommitSplitKeyS> (split-key program modulo one 26-bit prime, signed lazy arithmetic, "
                 "CTA halves phase-mixed)", "commit_splitkey_s")
WORKLOAD = "configs[1]: batched commitment generation, 2^16 messages/GPU at N=512, Params::default(), one shared key"


def bench_config():
    """`config` of the JSON line: the same object for both arms (ours and --impl reference)."""
    return {"workload": WORKLOAD, "items_per_gpu_per_step": BATCH, "N": N, "q": 3515337053, "params": "Params::default()",
            "seed": 1000}


def int_roofline(name, rate_per_s):
    """Fraction of the integer-multiply roofline for `name` at `rate_per_s` units/s on ONE GPU, both ways:
    frac_contract = 3 * mulmods * rate / R_IMAD (R_IMAD builder-measured, 18.1 T/s) and
    frac_shoup = mulmods * rate / (measured rate of a whole Shoup mulmod, 4.617 T/s)."""
    mm = ALG[name]["mulmods"]
    return {"mulmods_per_unit": mm, "units_per_s_per_gpu": rate_per_s,
            "frac_contract": 3.0 * mm * rate_per_s / (MEASURED_IMAD_TPS * 1e12),
            "frac_shoup": mm * rate_per_s / (MEASURED_MULMOD_TPS * 1e12),
            "hbm_frac": ALG[name]["bytes"] * rate_per_s / 1e9 / measured_peaks()[0]}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel_key):
    """dram bytes per launch of the dominant kernel from the committed ncu summary, or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel_key)
        except Exception:
            return None
    return None


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._th = threading.Thread(target=self._run, daemon=True)
            self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._th:
            self._th.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


CPU_LABEL = ("C restatement of the reference: schoolbook O(N^2) negacyclic products in the reference's operation order incl. the "
             "identity / zero key blocks, zero coefficients of the left operand skipped (so r in {-1,0,1} costs 2/3 of the dense "
             "MAC count of BASELINE.md section 3 -- the CPU figure is the faster, conservative one), OpenMP over items")


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def bind_to_gpu_numa(index):
    """Runs this rank's host threads and places its pinned buffers on the NUMA node the GPU hangs off (end-to-end leg:
    with 8 ranks on one socket every copy crosses the inter-socket link).  Best effort: reports what it could do."""
    info = {"gpu": index, "node": None, "cpus_bound": False, "mem_bound": False}
    try:
        import ctypes
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = int(open(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node").read())
        info["node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus_bound"] = True
        # set_mempolicy(MPOL_PREFERRED, {node}): later page faults (pinned allocations included) prefer the GPU's node
        mask = ctypes.c_ulong(1 << node)
        rc = ctypes.CDLL(None, use_errno=True).syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))
        info["mem_bound"] = (rc == 0)
    except Exception as ex:      # no NVML / sysfs / permission: keep the default placement
        info["error"] = type(ex).__name__
    return info


def cpu_commit_rate(target_s=12.0, seed=7):
    """Times the CPU restatement of the reference (oracle/, schoolbook products, the reference's own
    operation order incl. the identity/zero key blocks) on a bounded sample with all host threads."""
    from oracle import oracle as orc
    pkg = importlib.import_module("ring-zk_b200")
    s = pkg.synth.Synth(seed, N=N)
    o = orc.Oracle(orc.Params(N=N), *s.key())
    cores = orc.max_threads()
    probe = 4 * cores
    x, r = s.message(probe), s.small(probe)
    t0 = time.perf_counter()
    o.commit_batch(x, r)
    dt = time.perf_counter() - t0
    per = dt / probe
    n = int(max(probe, min(BATCH, target_s / per)))
    n = (n // cores) * cores or cores
    x, r = s.message(n), s.small(n)
    t0 = time.perf_counter()
    o.commit_batch(x, r)
    dt = time.perf_counter() - t0
    return n / dt, cores, n, dt


def cpu_other_configs(seed=11):
    """The CPU restatement on the other configurations, all host threads, on stated sub-batches (Linear and
    Sum are timed on a sub-batch and scale linearly in the instance count): Open verify, one Linear instance
    (commit + respond + verify) and one Sum-64 instance."""
    from oracle import oracle as orc
    pkg = importlib.import_module("ring-zk_b200")
    s = pkg.synth.Synth(seed, N=N)
    o = orc.Oracle(orc.Params(N=N), *s.key())
    cores = orc.max_threads()
    out = {}
    B = 2048 * cores
    x, r, y, d = s.message(B), s.small(B), s.gaussian(B), s.challenge(B)
    c, t, _ = o.open_commit_batch(x, r, y)
    z = o.open_respond_batch(y, r, d)
    c1 = np.ascontiguousarray(c[:, :1])
    t0 = time.perf_counter(); ok = o.open_verify_batch(z, t, c1, d); dt = time.perf_counter() - t0
    assert ok.all()
    out["open_verify"] = {"value": B / dt, "unit": "open-proof verifies/s", "sample": f"{B} verifies in {dt:.2f} s"}
    B = 512 * cores
    g, x, r, rp, y, yp, d = s.scalar(B), s.message(B), s.small(B), s.small(B), s.gaussian(B), s.gaussian(B), s.challenge(B)
    t0 = time.perf_counter()
    lc = o.linear_commit_batch(g, x, rp, r, y, yp)
    z, zp = o.linear_respond_batch(y, yp, r, rp, d)
    ok = o.linear_verify_batch(z, zp, lc["c"], lc["cp"], g, lc["t"], lc["tp"], lc["u"], d)
    dt = time.perf_counter() - t0
    assert ok.all()
    out["linear"] = {"value": B / dt, "unit": "Linear proofs (commit + respond + verify)/s", "sample": f"{B} instances in {dt:.2f} s"}
    B, T = 16 * cores, 64
    gs, xs, rs, ys = s.scalar(B, T), s.uniform_q(B, T, 1), s.small(B, T), s.gaussian(B, T)
    rp, yp, d = s.small(B), s.gaussian(B), s.challenge(B)
    t0 = time.perf_counter()
    sc = o.sum_commit_batch(gs, xs, rp, rs, ys, yp)
    zs, zp = o.sum_respond_batch(ys, yp, rs, rp, d)
    ok = o.sum_verify_batch(zs, zp, sc["cs"], sc["cp"], gs, sc["ts"], sc["tp"], sc["u"], d)
    dt = time.perf_counter() - t0
    assert ok.all()
    out["sum64"] = {"value": B / dt, "unit": "Sum proofs with 64 terms (commit + respond + verify)/s",
                    "sample": f"{B} instances in {dt:.2f} s"}
    return out


def single_call_latency(eng, oracle_obj=None, reps=30):
    """Median latency in microseconds of one call on ONE item through the host C ABI (pageable numpy buffers) for the
    phases the reference's own Criterion benches time (benches/bench.rs:35-305: N = 512, Sum with 4 terms).
    With oracle_obj the same calls are timed on the CPU restatement, one thread (cpu_baseline leg only)."""
    pkg = importlib.import_module("ring-zk_b200")
    s = pkg.synth.Synth(5, N=N)
    B, T = 1, 4
    x, r, y, d = s.message(B), s.small(B), s.gaussian(B), s.challenge(B)
    g, rp, yp = s.scalar(B), s.small(B), s.gaussian(B)
    gs, xs, rs, ys = s.scalar(B, T), s.uniform_q(B, T, 1), s.small(B, T), s.gaussian(B, T)
    c, t, _ = eng.open_commit(x, r, y)
    z = eng.open_respond(y, r, d)
    c1 = np.ascontiguousarray(c[:, :1])
    L = eng.linear_commit(g, x, rp, r, y, yp)
    lz, lzp = eng.linear_respond(y, yp, r, rp, d)
    S = eng.sum_commit(gs, xs, rp, rs, ys, yp)
    zs, zp = eng.sum_respond(ys, yp, rs, rp, d)
    o = oracle_obj
    calls = {
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

    def med(fn, n):
        ts = []
        for _ in range(n):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return round(float(np.median(ts)) * 1e6, 1)
    out = {}
    for name, (gpu, cpu) in calls.items():
        if o is None:
            for _ in range(3):
                gpu()
            out[name] = med(gpu, reps)
        else:
            out[name] = med(cpu, 3)
    return out


def run_reference(args):
    """--impl reference: the reference's CPU path (C restatement; the Rust crate cannot be built here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; this arm is the one CPU job of the node and is
    # meant to use every host core, so the variable is reset before the OpenMP runtime of the oracle library starts
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    from oracle import oracle as orc
    pkg = importlib.import_module("ring-zk_b200")
    s = pkg.synth.Synth(999, N=N)
    o = orc.Oracle(orc.Params(N=N), *s.key())
    cores = orc.max_threads()
    # one step = the same 2^16 commitments our arm's step computes (bounded sample: a few seconds of CPU per step)
    n = BATCH
    x, r = pkg.synth.Synth(1000, N=N).message(n), pkg.synth.Synth(1000, N=N).small(n)
    t_all0 = time.perf_counter()
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        o.commit_batch(x, r)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
        if time.perf_counter() - t_all0 > 240 and len(times) >= 3:       # keep the whole run within a few minutes
            break
    dt = float(np.mean(times))
    value = n / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": bench_config(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "cpu_model": cpu_model(),
                         "sample": f"{n} commitments per step ({CPU_LABEL})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all0,
        "notes": "C restatement of the reference's CPU path (oracle/), not the Rust binary: poly-ring-xnp1 is not in the tree and there is no Rust toolchain",
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--ref-seconds", type=float, default=8.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the Linear / Sum lines (configs[3], configs[4])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("ring-zk_b200")
    engine = importlib.import_module("ring-zk_b200.engine")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    s = pkg.synth.Synth(1000 + rank, N=N)          # every rank draws its own shard of the batch
    key = pkg.synth.Synth(999, N=N).key()          # one shared key, replicated per GPU
    eng = engine.Engine(N=N, device=local)
    eng.set_key_blocks(*key)
    stream = torch.cuda.current_stream().cuda_stream

    def up(a):
        return torch.from_numpy(a).to(dev)

    # ---- device-resident inputs (config 2 and config 3) ----
    x, r, y, d = up(s.message(B)), up(s.small(B)), up(s.gaussian(B)), up(s.challenge(B))
    c = torch.empty((B, 2, N), dtype=torch.int32, device=dev)
    t = torch.empty((B, 1, N), dtype=torch.int32, device=dev)
    z = torch.empty((B, 3, N), dtype=torch.int32, device=dev)
    flags = torch.zeros(B, dtype=torch.int32, device=dev)
    rng_word = torch.zeros(1, dtype=torch.int32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The only collective of the path: the all-gather of the per-shard ok / verify bitmaps (8 KiB per rank and step).
    # The bitmaps of a job's steps are packed on the device into one [steps][bytes] buffer by the bitmap kernel (same
    # stream as the commit / verify kernel) and gathered with ONE NCCL all-gather per job, inside the timed region:
    # the per-step gather of round 1 cost 17 us of a 0.45 ms step at 8 GPUs (latency-bound plain NCCL, 4 % of the step).
    nbm = (B + 7) // 8
    max_steps = max(args.steps, args.warmup, 8) + 1
    bitmaps = torch.zeros((max_steps, nbm), dtype=torch.uint8, device=dev)
    gathered = torch.zeros((world, max_steps, nbm), dtype=torch.uint8, device=dev) if world > 1 else None
    step_no = [0]

    def pack_bitmap():
        k = step_no[0] % max_steps
        step_no[0] += 1
        eng.dev("flags_to_bitmap", B, flags, bitmaps[k], rng_word, stream=stream)

    def gather_job():
        if world > 1:
            dist.all_gather_into_tensor(gathered, bitmaps)

    def commit_step():
        flags.zero_()
        eng.dev("commit_batch", B, x, r, c, flags, stream=stream)
        pack_bitmap()

    def timed(step_fn, steps, warmup, gather=True):
        for _ in range(warmup):
            step_fn()
        if gather:
            gather_job()                      # warms the communicator up
        barrier()
        step_no[0] = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_fn()
        if gather:
            gather_job()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    # ---- headline: commitments/s, whole job ----
    launches0 = eng.kernel_launches()
    with ClockSampler(local) as clk:
        ms_total = timed(commit_step, args.steps, args.warmup)
    launches = eng.kernel_launches() - launches0 - 2 * args.warmup   # 2 launches per step
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    assert bool((flags == 0).all()), "commit constraint flags set on honest inputs"

    if world > 1:
        # every rank's ok bitmap of every timed step arrived on every rank
        assert bool((gathered[:, :args.steps] == 0xFF).all()), "gathered ok bitmaps are not all-ones"

    # ---- dominant kernel alone (roofline) ----
    def kern_only():
        eng.dev("commit_batch", B, x, r, c, flags, stream=stream)
    ms_k = timed(kern_only, args.steps, args.warmup, gather=False) / args.steps
    peak, peak_src = measured_peaks()
    achieved = ALG_BYTES_COMMIT * B / (ms_k * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(COMMIT_KERNEL[1]),
                "traffic_source": "ncu --set full capture of this kernel (dram__bytes_read.sum + dram__bytes_write.sum per launch), "
                                  "committed as profiles/traffic.json -- a constant of the committed build, not an in-run counter",
                "kernel": COMMIT_KERNEL[0],
                "kernel_ms": ms_k, "algorithmic_bytes_per_launch": ALG_BYTES_COMMIT * B, "peak_source": peak_src,
                "note": "integer-multiply bound path (21504 modular multiplies per commitment by SURVEY.md 8(d)): the binding "
                        "roofline is int_roofline.commit; HBM is second"}

    # ---- second half of the metric: open-proof verifies/s (config 3) ----
    eng.dev("open_commit_batch", B, x, r, y, c, t, flags, stream=stream)
    eng.dev("open_respond_batch", B, y, r, d, z, stream=stream)
    torch.cuda.synchronize()

    def verify_step():
        flags.zero_()
        eng.dev("open_verify_batch", B, z, t, c, 2, d, flags, stream=stream)
        pack_bitmap()
    ms_v = timed(verify_step, args.steps, args.warmup)
    verifies = world * B * args.steps / (ms_v * 1e-3)
    assert bool((flags == 0).all()), "honest Open proofs failed to verify"

    def verify_only():
        eng.dev("open_verify_batch", B, z, t, c, 2, d, flags, stream=stream)
    ms_vk = timed(verify_only, args.steps, args.warmup, gather=False) / args.steps
    roofline["open_verify"] = {"kernel": "rzk_vm_kernel<2, MODE_SPLIT, SPVerifyFirstRot> (two-prime program; c1*d as signed rotations, OP_ROT)",
                               "kernel_ms": ms_vk, "achieved": ALG["open_verify"]["bytes"] * B / (ms_vk * 1e-3) / 1e9, "unit": "GB/s",
                               "frac": ALG["open_verify"]["bytes"] * B / (ms_vk * 1e-3) / 1e9 / peak,
                               "verifies_per_s_per_gpu": B / (ms_vk * 1e-3), "traffic": ncu_traffic("open_verify_rot")}

    def prove_step():
        eng.dev("open_commit_batch", B, x, r, y, c, t, flags, stream=stream)
        eng.dev("open_respond_batch", B, y, r, d, z, stream=stream)
    ms_p = timed(prove_step, args.steps, args.warmup, gather=False)
    proves = world * B * args.steps / (ms_p * 1e-3)

    # ---- configs[3] and configs[4]: Linear proofs (2^14 instances) and Sum proofs with 64 terms (2^12 instances),
    # device-resident, commit + respond + verify per instance.  Inputs are drawn on the device (same distributions
    # as ring-zk_b200/synth.py); correctness inside the bench = every honest instance verifies (flags == 0).
    extras = None
    if not args.no_extras:
        extras = {}
        half = (3515337053 - 1) // 2
        gen = torch.Generator(device=dev); gen.manual_seed(77 + rank)
        U = lambda *sh: torch.randint(-half, half + 1, sh, device=dev, generator=gen, dtype=torch.int64).to(torch.int32)
        S3 = lambda *sh: torch.randint(-1, 2, sh, device=dev, generator=gen, dtype=torch.int32).to(torch.int8)
        G = lambda *sh: (torch.randn(sh, device=dev, generator=gen, dtype=torch.float64) * 15444.0).trunc().to(torch.int32)
        E = lambda *sh: torch.empty(sh, dtype=torch.int32, device=dev)
        ksteps = max(3, args.steps // 4)
        # Linear
        BL = 1 << 14
        g_, x_, r_, rp_, y_, yp_, d_ = U(BL, N), U(BL, 1, N), S3(BL, 3, N), S3(BL, 3, N), G(BL, 3, N), G(BL, 3, N), d[:BL].contiguous()
        gx, cp, cl, tl, tpl, u = E(BL, 1, N), E(BL, 2, N), E(BL, 2, N), E(BL, 1, N), E(BL, 1, N), E(BL, 1, N)
        zl, zpl = E(BL, 3, N), E(BL, 3, N)
        fl = torch.zeros(BL, dtype=torch.int32, device=dev)

        def linear_step():
            eng.dev("linear_commit_batch", BL, g_, x_, rp_, r_, y_, yp_, gx, cp, cl, tl, tpl, u, fl, stream=stream)
            eng.dev("linear_respond_batch", BL, y_, yp_, r_, rp_, d_, zl, zpl, stream=stream)
            eng.dev("linear_verify_batch", BL, zl, zpl, cl, cp, g_, tl, tpl, u, d_, fl, stream=stream)
        ms_l = timed(linear_step, ksteps, 2, gather=False) / ksteps
        assert bool((fl == 0).all()), "honest Linear proofs failed to verify"
        extras["linear"] = {"instances_per_gpu": BL, "instances_per_s": world * BL / (ms_l * 1e-3), "ms_per_step": ms_l,
                            "int_roofline": int_roofline("linear", BL / (ms_l * 1e-3))}
        del g_, x_, r_, rp_, y_, yp_, gx, cp, cl, tl, tpl, u, zl, zpl
        # Sum, T = 64
        BS, TT = 1 << 12, 64
        gs, xs, rs, ys = U(BS, TT, N), U(BS, TT, 1, N), S3(BS, TT, 3, N), G(BS, TT, 3, N)
        rps, yps, ds = S3(BS, 3, N), G(BS, 3, N), d[:BS].contiguous()
        xp, cps, css, tss, tps, us = E(BS, 1, N), E(BS, 2, N), E(BS, TT, 2, N), E(BS, TT, 1, N), E(BS, 1, N), E(BS, 1, N)
        zs, zps = E(BS, TT, 3, N), E(BS, 3, N)
        fs = torch.zeros(BS, dtype=torch.int32, device=dev)

        def sum_step():
            eng.dev("sum_commit_batch", BS, TT, gs, xs, rps, rs, ys, yps, xp, cps, css, tss, tps, us, fs, stream=stream)
            eng.dev("sum_respond_batch", BS, TT, ys, yps, rs, rps, ds, zs, zps, stream=stream)
            eng.dev("sum_verify_batch", BS, TT, zs, zps, css, cps, gs, tss, tps, us, ds, fs, stream=stream)
        ms_s = timed(sum_step, ksteps, 1, gather=False) / ksteps
        assert bool((fs == 0).all()), "honest Sum proofs failed to verify"
        extras["sum64"] = {"instances_per_gpu": BS, "terms": TT, "instances_per_s": world * BS / (ms_s * 1e-3), "ms_per_step": ms_s,
                           "int_roofline": int_roofline("sum64", BS / (ms_s * 1e-3))}
        del gs, xs, rs, ys, xp, cps, css, tss, tps, us, zs, zps
        torch.cuda.empty_cache()
        if world == 1:
            extras["single_call_latency_us"] = single_call_latency(eng)
        # ---- row f1 (optional on-device samplers): Open prover end to end, host -> c, t, z on the host, with the
        # randomness drawn on the host and copied up (the reference's flow) or sampled on the device (r, y never move).
        # One stream, no chunk pipelining in either variant: the difference is the bytes that cross PCIe.
        Bp = B
        pin = lambda *sh, dt=torch.int32: torch.empty(sh, dtype=dt).pin_memory()
        xh, rh, yh, dh = pin(Bp, 1, N), pin(Bp, 3, N, dt=torch.int8), pin(Bp, 3, N), pin(Bp, N, dt=torch.int8)
        ch2, th2, zh2 = pin(Bp, 2, N), pin(Bp, 1, N), pin(Bp, 3, N)
        xh.copy_(x.cpu()); rh.copy_(r.cpu()); yh.copy_(y.cpu()); dh.copy_(d.cpu())
        xd, rd, yd, dd = torch.empty_like(x), torch.empty_like(r), torch.empty_like(y), torch.empty_like(d)
        cE, tE, zE, fE = torch.empty_like(c), torch.empty_like(t), torch.empty_like(z), torch.zeros_like(flags)
        sig = float(eng.sigma())

        def prove_e2e(device_randomness, seed):
            xd.copy_(xh, non_blocking=True)
            if device_randomness:
                eng.dev("sample_small", 3 * Bp, 1, seed, 1, rd, stream=stream)
                eng.dev("sample_gaussian", 3 * Bp, sig, seed, 2, yd, stream=stream)
            else:
                rd.copy_(rh, non_blocking=True); yd.copy_(yh, non_blocking=True)
            eng.dev("open_commit_batch", Bp, xd, rd, yd, cE, tE, fE, stream=stream)
            ch2.copy_(cE, non_blocking=True); th2.copy_(tE, non_blocking=True)
            dd.copy_(dh, non_blocking=True)                                   # the verifier's challenge arrives
            eng.dev("open_respond_batch", Bp, yd, rd, dd, zE, stream=stream)
            zh2.copy_(zE, non_blocking=True)
            torch.cuda.synchronize()
        res = {}
        for name, devr in (("host_randomness", False), ("device_randomness", True)):
            prove_e2e(devr, 1)
            t0 = time.perf_counter()
            for i in range(ksteps):
                prove_e2e(devr, 2 + i)
            res[name] = world * Bp * ksteps / (time.perf_counter() - t0)
        assert int(fE.any()) == 0
        # ---- rows f1 + f2 chained: messages arrive from the host (2 KB per item), r and y are drawn on the device, the proof
        # (commit -> Fiat-Shamir challenge -> response, docs/FIAT_SHAMIR.md) and its verification run back to back on one
        # stream, and only the verdict bitmap (1 bit per item) goes back: nothing but x and the bitmap crosses PCIe
        api = importlib.import_module("ring-zk_b200.api")
        pre = b"ring-zk/fs/open/v1".ljust(32, b"\0") + bytes(32) + bytes(24)          # tag || (benchmark: zero key digest) || shape words
        d8, d8v = torch.empty((Bp, N), dtype=torch.int8, device=dev), torch.empty((Bp, N), dtype=torch.int8, device=dev)
        fV = torch.zeros_like(flags)
        bmd = torch.zeros((Bp + 7) // 8, dtype=torch.uint8, device=dev)
        bmh = torch.zeros((Bp + 7) // 8, dtype=torch.uint8).pin_memory()

        def chain_step(seed):
            xd.copy_(xh, non_blocking=True)
            eng.dev("sample_small", 3 * Bp, 1, seed, 1, rd, stream=stream)
            eng.dev("sample_gaussian", 3 * Bp, sig, seed, 2, yd, stream=stream)
            fE.zero_(); fV.zero_()
            eng.open_prove_fs(xd, rd, yd, pre, cE, tE, d8, zE, fE, stream=stream)
            eng.open_verify_fs(cE, tE, zE, pre, d8v, fV, stream=stream)
            eng.dev("flags_to_bitmap", Bp, fV, bmd, rng_word, stream=stream)
            bmh.copy_(bmd, non_blocking=True)
            torch.cuda.synchronize()
        chain_step(1)
        l0 = eng.kernel_launches()
        t0 = time.perf_counter()
        for i in range(ksteps):
            chain_step(100 + i)
        dt_chain = time.perf_counter() - t0
        assert int(fE.any()) == 0 and bool((bmh == 0xFF).all()) and bool((d8 == d8v).all()), "chained non-interactive proofs must verify"
        extras["open_fs_chain_e2e"] = {
            "instances_per_s": world * Bp * ksteps / dt_chain, "unit": "non-interactive open proofs proved AND verified/s, host messages in, verdict bitmap out",
            "h2d_bytes_per_item": 4 * N, "d2h_bytes_per_item": 0.125, "kernel_launches_per_step": (eng.kernel_launches() - l0) / ksteps,
            "note": "rzk_sample_*_dev (test-only samplers) + rzk_open_prove_fs_batch_dev + rzk_open_verify_fs_batch_dev on one stream"}
        del xd, rd, yd, dd, cE, tE, zE, xh, rh, yh, dh, ch2, th2, zh2, d8, d8v
        extras["open_prove_e2e"] = {"instances_per_s": res, "unit": "open proofs (commit + response)/s, host buffers in and out",
                                    "h2d_bytes_per_item": {"host_randomness": 4 * N + 3 * N + 12 * N + N, "device_randomness": 4 * N + N},
                                    "d2h_bytes_per_item": 8 * N + 4 * N + 12 * N}

    # ---- end to end through the host C ABI (pinned host buffers, H2D + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        xh = torch.empty((B, 1, N), dtype=torch.int32).pin_memory()
        rh = torch.empty((B, 3, N), dtype=torch.int8).pin_memory()
        ch = torch.empty((B, 2, N), dtype=torch.int32).pin_memory()
        okh = torch.zeros((B + 7) // 8, dtype=torch.uint8).pin_memory()
        xh.copy_(x.cpu()); rh.copy_(r.cpu())
        xn, rn, cn, okn = xh.numpy(), rh.numpy(), ch.numpy(), okh.numpy()

        # two forms of the same call: randomness as int8 (any |r| <= 127) and packed at 2 bits per coefficient
        # (rzk_commit_batch_r2: what the Rust shim sends for Params::default(), b = 1, r in {-1, 0, 1}); the packing is host
        # marshaling like the i64 -> int8 narrowing either form needs, and stays outside the timed region like it
        r2h = torch.from_numpy(engine.pack_r2(rn)).pin_memory()
        r2n = r2h.numpy()

        def time_host(fn):
            for _ in range(max(1, args.warmup)):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(ksteps):
                fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                tt = torch.tensor([dt], device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            return dt
        ksteps = max(3, args.steps // 2)
        dt_i8 = time_host(lambda: eng._call("rzk_commit_batch", B, xn.ctypes.data, rn.ctypes.data, cn.ctypes.data, okn.ctypes.data))
        assert bool((ch.to(dev) == c).all()), "host-path and device-path commitments differ"
        ch.zero_()
        dt = time_host(lambda: eng._call("rzk_commit_batch_r2", B, xn.ctypes.data, r2n.ctypes.data, cn.ctypes.data, okn.ctypes.data))
        assert bool((ch.to(dev) == c).all()), "host-path (packed randomness) and device-path commitments differ"
        e2e = {"value": world * B * ksteps / dt, "unit": UNIT,
               "h2d_bytes_per_step": B * (N * 4 + 3 * N // 4), "d2h_bytes_per_step": B * 2 * N * 4 + (B + 7) // 8,
               "steps": ksteps, "api": "rzk_commit_batch_r2 (host pointers, pinned; r at 2 bits per coefficient), chunked 4-stream pipeline (8192 items per chunk)",
               "int8_r": {"value": world * B * ksteps / dt_i8, "unit": UNIT, "h2d_bytes_per_step": B * (N * 4 + 3 * N),
                          "d2h_bytes_per_step": B * 2 * N * 4 + (B + 7) // 8, "api": "rzk_commit_batch (r as int8)"}}

        # the second half of BASELINE.json's metric, same way: Open-proof verifies/s through rzk_open_verify_batch
        pin = lambda tdev: tdev.cpu().pin_memory()
        zh, th, c1h, dh = pin(z), pin(t), pin(c[:, :1].contiguous()), pin(d)
        vbm = torch.zeros((B + 7) // 8, dtype=torch.uint8).pin_memory()
        zn, tn, c1n, dn, vn = zh.numpy(), th.numpy(), c1h.numpy(), dh.numpy(), vbm.numpy()

        def host_verify():
            eng._call("rzk_open_verify_batch", B, zn.ctypes.data, tn.ctypes.data, c1n.ctypes.data, dn.ctypes.data, vn.ctypes.data)
        for _ in range(max(1, args.warmup)):
            host_verify()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            host_verify()
        torch.cuda.synchronize()
        dtv = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dtv], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dtv = float(tt.item())
        assert bool((vbm == 0xFF).all()), "honest Open proofs must verify through the host path"
        e2e["open_verify"] = {"value": world * B * ksteps / dtv, "unit": "open-proof verifies/s",
                              "h2d_bytes_per_step": B * (3 * N * 4 + N * 4 + N * 4 + N), "d2h_bytes_per_step": (B + 7) // 8,
                              "api": "rzk_open_verify_batch (host pointers, pinned)"}

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, cores, n, dt = cpu_commit_rate(target_s=12.0)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "cpu_model": cpu_model(),
               "sample": f"{n} commitments in {dt:.1f} s ({CPU_LABEL})"}
        if not args.no_extras:
            from oracle import oracle as orc
            cpu["single_call_latency_us_1thread"] = single_call_latency(eng, orc.Oracle(orc.Params(N=N), *key))
            cpu["other_configs"] = cpu_other_configs()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": bench_config(),
            "numa": numa,
            "notes": {"l2": "inputs_larger_than_l2 (480 MB touched per step vs 126 MB L2)",
                      "collective": "one NCCL all_gather of the job's ok bitmaps (steps x 8 KiB per rank), inside the timed region" if world > 1 else "none",
                      "items_per_gpu_per_step": B},
            "open_verifies_per_s": verifies, "open_proves_per_s": proves,
            "ms_per_step_open_verify": ms_v / args.steps, "ms_per_step_open_prove": ms_p / args.steps,
            "other_configs": extras, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clk.summary(),
            # integer-multiply roofline of every configuration, per GPU (step rates divided by the number of ranks)
            "int_roofline": {
                "R_IMAD_T_per_s": MEASURED_IMAD_TPS, "shoup_mulmod_T_per_s": MEASURED_MULMOD_TPS,
                "peak_source": "builder-measured on this pool's B200s (tools/imad_bench.cu, profiles/r1_imad_bench.jsonl); "
                               "MEASURED_PEAKS.json has no integer peak.  frac_contract = 3 * mulmods * rate / R_IMAD is "
                               "SURVEY.md 8(d)'s formula; frac_shoup divides by the measured rate of a whole Shoup mulmod",
                "commit": int_roofline("commit", B / (ms_k * 1e-3)),
                "open_verify": int_roofline("open_verify", B / (ms_vk * 1e-3)),
                "open_instance": int_roofline("open_instance", 1.0 / (ms_p / args.steps * 1e-3 / B + ms_vk * 1e-3 / B)),
                "linear": extras["linear"]["int_roofline"] if extras else None,
                "sum64": extras["sum64"]["int_roofline"] if extras else None,
                # kept for continuity with round 1's line
                "mulmods_per_item": ALG_MULMODS_COMMIT, "frac": ALG_MULMODS_COMMIT * B / (ms_k * 1e-3) / (MEASURED_MULMOD_TPS * 1e12),
                "frac_contract": 3.0 * ALG_MULMODS_COMMIT * B / (ms_k * 1e-3) / (MEASURED_IMAD_TPS * 1e12)},
        }
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def emit(line):
    """The one JSON line of this run, on the process's real stdout."""
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


if __name__ == "__main__":
    # stdout carries exactly one JSON line: whatever a library prints to file descriptor 1 (NCCL's "NCCL version ..."
    # line under torchrun) is routed to stderr, and the JSON line is written to a duplicate of the original stdout
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    main()
