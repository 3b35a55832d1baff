"""Host-to-device copy rate of pinned memory on this box (one-shot commits upload 17 - 22 MB): sizes, repeats, and
whether the source was just written by the CPU."""
import time

import torch

dev = torch.device("cuda:0")
for mb in (1, 4, 22, 64, 256):
    n = mb << 20
    src = torch.empty(n, dtype=torch.uint8).pin_memory()
    dst = torch.empty(n, dtype=torch.uint8, device=dev)
    back = torch.empty(n, dtype=torch.uint8).pin_memory()
    for label, touch in (("clean", False), ("just written", True)):
        ts = []
        for i in range(6):
            if touch:
                src.fill_(i)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        best = min(ts[1:])
        print(f"H2D {mb:4d} MiB {label:13s}: {best * 1e3:7.3f} ms  {n / best / 1e9:6.1f} GB/s", flush=True)
    ts = []
    for i in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        back.copy_(dst, non_blocking=True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print(f"D2H {mb:4d} MiB              : {min(ts[1:]) * 1e3:7.3f} ms  {n / min(ts[1:]) / 1e9:6.1f} GB/s", flush=True)
