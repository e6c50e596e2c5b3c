#!/usr/bin/env python
"""Per-class device timing of the encode (put) and decode (get) kernels: CUDA-event durations the
engine records around k_encode / k_decode, pages resident in HBM.  Tuning aid, not the bench line.

    python tools/kernel_bench.py [--chunks 4096] [--classes RTZMB] [--reps 3]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import edge_fuse_b200 as E  # noqa: E402

CHUNK = 65536


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=4096)
    ap.add_argument("--classes", default="RTZMB")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--pshift", type=int, default=16)
    ap.add_argument("--fingerprint", type=int, default=1)
    a = ap.parse_args()
    n, bs = a.chunks, 1 << a.pshift
    eng = E.Engine(pshift=a.pshift, accel=12, capacity=8 * n * (a.reps + 1) * len(a.classes),
                   arena_bytes=int(n * bs * 1.1 * (a.reps + 1) * len(a.classes)) + (1 << 28), max_batch=n,
                   flags=E.FINGERPRINT if a.fingerprint else 0)
    d = eng.dev_alloc(n * bs)
    d_out = eng.dev_alloc(n * bs)
    res = {}
    gen = 0
    for cls in a.classes:
        k = "RTZM".find(cls)
        allc = np.arange(8 * n, dtype=np.uint64)
        cids = allc[((allc + (allc >> np.uint64(3))) & np.uint64(3)) == k][:n] if k >= 0 else allc[:n]
        eng.gen_chunks_dev(42, cids, d)
        u = np.full(n, 7, dtype=np.uint64)
        best_e, best_d, lens = 1e30, 1e30, None
        for r in range(a.reps + 1):
            gen += 1
            l = np.arange(n, dtype=np.uint64) | (np.uint64(gen) << np.uint64(44))
            s0 = eng.stats()
            lens = eng.put(u, l, d, on_dev=True)
            s1 = eng.stats()
            _, status = eng.get(u, l, out=d_out, on_dev=True)
            s2 = eng.stats()
            assert (status == E.HIT).all()
            if r:
                best_e = min(best_e, (s1["encode_kernel_ns"] - s0["encode_kernel_ns"]) * 1e-9)
                best_d = min(best_d, (s2["decode_kernel_ns"] - s1["decode_kernel_ns"]) * 1e-9)
        f0 = eng.stats()["fingerprint_kernel_ns"]
        eng.fingerprint_dev(n, d)
        eng.fingerprint_dev(n, d)
        fp_s = (eng.stats()["fingerprint_kernel_ns"] - f0) * 0.5e-9
        stored = float(lens.sum())
        res[cls] = {"encode_gibs": n * bs / 2**30 / best_e, "decode_gibs": n * bs / 2**30 / best_d,
                    "ratio": stored / (n * bs),
                    "encode_alg_gbs": (n * (bs + 88) + stored) / best_e / 1e9,
                    "decode_alg_gbs": (n * (bs + 56) + stored) / best_d / 1e9,
                    "fingerprint_gbs": n * (bs + 16) / fp_s / 1e9}
        print(cls, json.dumps(res[cls]), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
