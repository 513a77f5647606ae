"""How fast can this host move the end-to-end commit traffic (235 MB up, 268 MB down per 2^16 commitments and GPU) with nothing
else going on?  Pinned buffers, one stream per direction per GPU: the ceiling of bench.py's e2e line behind the host C ABI.

    python tools/pcie_duplex.py                                            # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_duplex.py
                                                                            # all GPUs copying at the same time
    RZK_BIND_NUMA=1 ... tools/pcie_duplex.py                               # ranks bound to their GPU's NUMA node (as bench.py)

Prints one JSON line (rank 0): per-GPU and aggregate GB/s, and the commitments/s ceiling they imply."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
numa = None
if os.environ.get("RZK_BIND_NUMA"):
    import bench
    numa = bench.bind_to_gpu_numa(local)
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

up_b, down_b = 234881024, 268443648
hu = torch.empty(up_b, dtype=torch.uint8).pin_memory(); du = torch.empty(up_b, dtype=torch.uint8, device="cuda")
hd = torch.empty(down_b, dtype=torch.uint8).pin_memory(); dd = torch.empty(down_b, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(up, down, chunks=1, it=10):
    barrier(); t0 = time.perf_counter()
    for _ in range(it):
        for k in range(chunks):
            if up:
                a, b = k * up_b // chunks, (k + 1) * up_b // chunks
                with torch.cuda.stream(s1):
                    du[a:b].copy_(hu[a:b], non_blocking=True)
            if down:
                a, b = k * down_b // chunks, (k + 1) * down_b // chunks
                with torch.cuda.stream(s2):
                    hd[a:b].copy_(dd[a:b], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / it
    if world > 1:                       # the slowest rank decides, as in bench.py
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return dt


run(True, True)
t_up, t_down, t_both, t_both8 = run(True, False), run(False, True), run(True, True), run(True, True, 8)
if rank == 0:
    print(json.dumps({
        "n_gpus": world, "numa": numa, "bytes_up_per_gpu": up_b, "bytes_down_per_gpu": down_b,
        "h2d_alone_GBps_per_gpu": up_b / t_up / 1e9, "d2h_alone_GBps_per_gpu": down_b / t_down / 1e9,
        "duplex_GBps_per_gpu": (up_b + down_b) / t_both / 1e9, "duplex_GBps_aggregate": world * (up_b + down_b) / t_both / 1e9,
        "duplex_ms": t_both * 1e3, "duplex_8_chunks_ms": t_both8 * 1e3,
        "commitments_per_s_ceiling": world * 65536 / t_both,
        "note": "all ranks copy at the same time; times are the maximum over ranks"}))
if world > 1:
    dist.destroy_process_group()
