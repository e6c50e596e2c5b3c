#!/usr/bin/env python
"""Measures pinned host<->device copy bandwidth of this box (the ceiling of the e2e number)."""
import time, torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    print(name, "GiB/s", 5 * n / 2**30 / (time.perf_counter() - t0))
