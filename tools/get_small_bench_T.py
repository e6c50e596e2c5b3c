import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import edge_fuse_b200 as E
n = 64
eng = E.Engine(pshift=16, accel=12, capacity=1 << 14, arena_bytes=1 << 30, max_batch=1024)
hp = E.lib().cmb200_host_alloc(n * 65536)
row = {}
for k, cls in [(k, c) for k, c in enumerate("RTZM") if c in os.environ.get("CLASSES", "TM")]:
    allc = np.arange(16 * n, dtype=np.uint64)
    cids = allc[((allc + (allc >> np.uint64(3))) & np.uint64(3)) == k][:n]
    pages = np.stack([E.gen_chunk_host(42, int(c), 65536) for c in cids])
    u = np.full(n, 100 + k, dtype=np.uint64); l = np.arange(n, dtype=np.uint64)
    eng.put(u, l, pages)
    out, st = eng.get_small(u, l)
    assert (st == E.HIT).all() and ((out == pages).all() or os.environ.get('NOCHECK'))
    for m in (1, 32):
        f = lambda: eng.get_small(u[:m], l[:m], out=hp)
        f(); f()
        t0 = time.perf_counter()
        for _ in range(20): f()
        row[f"{cls}_n{m}_us"] = round((time.perf_counter() - t0) / 20 * 1e6, 1)
print(json.dumps(row))
