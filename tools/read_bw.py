#!/usr/bin/env python
"""What read bandwidth does this B200 give a plain streaming reduction?  (context for roofline numbers)"""
import torch
n = 65536
x = torch.empty((n, n), dtype=torch.float32, device="cuda").normal_()
for name, fn in (("sum", lambda: x.sum()), ("max", lambda: x.max()), ("sum_dim0", lambda: x.sum(0)), ("sum_dim1", lambda: x.sum(1))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"torch {name}: {ms:.3f} ms  {x.numel() * 4 / ms / 1e6:.0f} GB/s")
y = torch.empty(1 << 30, dtype=torch.float32, device="cuda")
z = torch.empty_like(y)
for _ in range(2):
    z.copy_(y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    z.copy_(y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"torch copy 4 GiB: {ms:.3f} ms  {2 * y.numel() * 4 / ms / 1e6:.0f} GB/s (read+write)")
