#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel on a
handful of chunks of each class and both table modes, through the C ABI.

    compute-sanitizer --tool memcheck python tools/sanitize_small.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import datagen
import edge_fuse_b200 as E

for n in (65536, 131072, 4096, 5000):
    pages = [datagen.make_page(k, n, 3 + i) for i, k in enumerate("RTZMPAXS" if n != 5000 else "RTZP")]
    blocks, fps = E.lz4_encode_batch(datagen.pad_rows(pages), nbytes=n, accel=12, fingerprints=True)
    out, used = E.lz4_decode_batch(blocks, n)
    assert all((out[i][:n] == pages[i]).all() for i in range(len(pages))) and (used == [len(b) for b in blocks]).all()
eng = E.Engine(pshift=16, accel=12, capacity=2048, arena_bytes=64 << 20, max_batch=32, flags=E.FINGERPRINT)
pages = np.stack([E.gen_chunk_host(42, c, 65536) for c in range(40)])
u = np.full(40, 9, dtype=np.uint64); l = (np.arange(40) % 30).astype(np.uint64)
eng.put(u, l, pages)
out, st = eng.get(u, l)
assert (st == E.HIT).all()
eng.unset(u[:5], l[:5])
eng.sample(np.arange(6, dtype=np.uint64))
eng.import_remote(u[:4], l[:4] + np.uint64(100), np.ones(4, np.uint32), np.arange(4, dtype=np.uint64) + np.uint64(10**6))
eng.locate(u, l)
eng.read_records(u[5:9], l[5:9]); eng.read_fingerprints(u[5:9], l[5:9])
eng.close()
print("sanitize_small ok")
