#!/usr/bin/env python
"""Achievable read+write HBM bandwidth for tensors of the step's sizes (torch copy / add as the yardstick)."""
import torch
dev = torch.device("cuda", 0)
for rows, C in ((81920, 320), (20480, 640), (5120, 1280), (1280, 1280)):
    x = torch.randn(rows, C, device=dev).bfloat16()
    y = torch.empty_like(x)
    big = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    for name, fn in (("copy", lambda: y.copy_(x)), ("silu", lambda: torch.nn.functional.silu(x, inplace=False))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            big.zero_()  # flush L2
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        ts.sort()
        by = 2 * rows * C * 2
        print(f"{name} [{rows} x {C}] bf16: median {ts[5]:.1f} us  -> {by / ts[5] / 1e6:.2f} TB/s (read + write, L2 flushed)")
