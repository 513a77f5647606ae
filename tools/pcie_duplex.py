"""How fast can this host move the end-to-end commit traffic (235 MB up, 268 MB down per 2^16 commitments) with nothing else
going on?  Pinned buffers, one stream per direction; upper bound for bench.py's e2e line (development helper)."""
import time
import torch
up_b, down_b = 234881024, 268443648
hu = torch.empty(up_b, dtype=torch.uint8).pin_memory(); du = torch.empty(up_b, dtype=torch.uint8, device="cuda")
hd = torch.empty(down_b, dtype=torch.uint8).pin_memory(); dd = torch.empty(down_b, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(both, chunks=1, it=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it):
        for k in range(chunks):
            a, b = k * up_b // chunks, (k + 1) * up_b // chunks
            with torch.cuda.stream(s1):
                du[a:b].copy_(hu[a:b], non_blocking=True)
            if both:
                a, b = k * down_b // chunks, (k + 1) * down_b // chunks
                with torch.cuda.stream(s2):
                    hd[a:b].copy_(dd[a:b], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / it


run(True)
t_up = run(False)
t_both = run(True)
t_both8 = run(True, 8)
print(f"H2D alone {up_b / t_up / 1e9:.1f} GB/s ({t_up * 1e3:.2f} ms); both directions together {(up_b + down_b) / t_both / 1e9:.1f} GB/s "
      f"({t_both * 1e3:.2f} ms -> at most {65536 / t_both / 1e6:.2f} M commitments/s end to end); in 8 chunks {t_both8 * 1e3:.2f} ms")
