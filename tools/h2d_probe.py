#!/usr/bin/env python
"""Host-to-device bandwidth of pinned copies on this box (what bounds bench.py's e2e number): one stream vs two,
105 MB chunks (a 256-frame batch) vs one 1 GiB copy."""
import torch

dev = torch.device("cuda", 0)
n = 1 << 28                                   # 1 GiB of float32
host = torch.empty(n, dtype=torch.float32).pin_memory()
devb = torch.empty(n, dtype=torch.float32, device=dev)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


ms = timed(lambda: devb.copy_(host, non_blocking=True))
print(f"1 GiB, one copy:            {n * 4 / ms / 1e6:7.1f} GB/s")
chunk = 256 * 320 * 320
def chunks():
    for i in range(0, n, chunk):
        devb[i:i + chunk].copy_(host[i:i + chunk], non_blocking=True)
ms = timed(chunks)
print(f"1 GiB, 105 MB chunks:       {n * 4 / ms / 1e6:7.1f} GB/s")
s2 = torch.cuda.Stream()
def two():
    half = n // 2
    devb[:half].copy_(host[:half], non_blocking=True)
    with torch.cuda.stream(s2):
        devb[half:].copy_(host[half:], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)
ms = timed(two)
print(f"1 GiB, two streams:         {n * 4 / ms / 1e6:7.1f} GB/s")
ms = timed(lambda: host.copy_(devb, non_blocking=True))
print(f"1 GiB D2H, one copy:        {n * 4 / ms / 1e6:7.1f} GB/s")
