#!/usr/bin/env python3
"""Write-only HBM bandwidth of the box: cudaMemset / torch fill of 4..16 GB, CUDA-event timed.  The ceiling of a pure store
stream (K1 writes 16 B/voxel and reads nothing) can differ from the copy bandwidth in MEASURED_PEAKS.json."""
import torch
for gb in (4, 8, 16):
    n = gb * (1 << 30) // 4
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    for name, fn in (("fill_", lambda: x.fill_(1.5)), ("zero_ (memset)", lambda: x.zero_())):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print("%2d GB %-15s %.3f ms = %.0f GB/s" % (gb, name, ms, n * 4 / ms / 1e6), flush=True)
    y = torch.empty(n // 2, dtype=torch.float32, device="cuda")
    z = torch.empty(n // 2, dtype=torch.float32, device="cuda")
    for _ in range(3):
        z.copy_(y)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        z.copy_(y)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print("%2d GB copy (r+w)       %.3f ms = %.0f GB/s" % (gb, ms, n * 4 / ms / 1e6), flush=True)
    del x, y, z
