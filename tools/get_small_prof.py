#!/usr/bin/env python
"""One cmb200_get_small launch over T-class pages for ncu (tools: ncu -k regex:k_get_small ...)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import edge_fuse_b200 as E
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cls = "RTZM".find(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = E.Engine(pshift=16, accel=12, capacity=1 << 14, arena_bytes=1 << 30, max_batch=1024)
allc = np.arange(16 * n, dtype=np.uint64)
cids = allc[((allc + (allc >> np.uint64(3))) & np.uint64(3)) == cls][:n]
pages = np.stack([E.gen_chunk_host(42, int(c), 65536) for c in cids])
u = np.full(n, 9, dtype=np.uint64); l = np.arange(n, dtype=np.uint64)
eng.put(u, l, pages)
for _ in range(3):
    out, st = eng.get_small(u, l)
assert (st == E.HIT).all() and (out == pages).all()
print("ok")
