#!/usr/bin/env python
"""cmb200_get_small (one page per call) from T caller threads, straight at the engine (no C-layer
combining): how the launch path and the kernel behave with many single-CTA kernels in flight.
Classes: R incompressible, T text-like, Z zero, M mixed, B blend."""
import ctypes as C
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import edge_fuse_b200 as E

CH = 65536
N = 256
L = E.lib()
eng = E.Engine(pshift=16, accel=12, capacity=1 << 14, arena_bytes=1 << 30, max_batch=1024)
sets = {}
allc = np.arange(16 * N, dtype=np.uint64)
cls_of = (allc + (allc >> np.uint64(3))) & np.uint64(3)
for k, name in ((0, "R"), (1, "T"), (2, "Z"), (3, "M")):
    cids = allc[cls_of == k][:N]
    pages = np.stack([E.gen_chunk_host(42, int(c), CH) for c in cids])
    u = np.full(N, 100 + k, dtype=np.uint64)
    l = np.arange(N, dtype=np.uint64)
    eng.put(u, l, pages)
    sets[name] = np.stack([u, l], axis=1).copy()          # cmb200_addr = {u, l}
sets["B"] = np.concatenate([sets[c][:N // 4] for c in "RTZM"])
np.random.default_rng(1).shuffle(sets["B"])


def run(addrs, threads, per):
    bufs = [L.cmb200_host_alloc(CH) for _ in range(threads)]
    bad = []

    def worker(t):
        st = np.zeros(1, dtype=np.int32)
        for i in range(per):
            a = addrs[(t * per + i) % len(addrs)]
            rc = L.cmb200_get_small(eng.h, 1, a.ctypes.data, bufs[t], st.ctypes.data)
            if rc or st[0] != E.HIT:
                bad.append((rc, int(st[0])))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    [x.start() for x in th]
    [x.join() for x in th]
    dt = time.perf_counter() - t0
    for b in bufs:
        L.cmb200_host_free(b)
    assert not bad, bad[:3]
    return threads * per / dt


out = {}
for cls in os.environ.get("CLASSES", "RZTB"):
    for T in (1, 4, 8, 16, 32):
        run(sets[cls], T, 20)
        k = run(sets[cls], T, 300)
        out[f"{cls}_T{T}"] = {"kops": round(k / 1e3, 1), "gibs": round(k * CH / 2**30, 2), "us_per_get": round(T / k * 1e6, 1)}
        print(cls, T, out[f"{cls}_T{T}"], flush=True)
print(json.dumps(out))
