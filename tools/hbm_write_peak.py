#!/usr/bin/env python3
"""Pure-write HBM bandwidth (the enumerator writes and never reads): fill of a 2 GiB buffer."""
import torch
x = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
for _ in range(3):
    x.fill_(1)
torch.cuda.synchronize()
best = 0.0
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); x.fill_(2); b.record(); torch.cuda.synchronize()
    best = max(best, x.numel() / (a.elapsed_time(b) * 1e-3) / 1e9)
print(f"fill (write only): {best:.0f} GB/s")
y = torch.empty_like(x)
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); y.copy_(x); b.record(); torch.cuda.synchronize()
    best2 = 2 * x.numel() / (a.elapsed_time(b) * 1e-3) / 1e9
print(f"copy (read + write): {best2:.0f} GB/s")
