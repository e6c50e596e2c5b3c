#!/usr/bin/env python
"""BASELINE.json configs 2 and 3 on one GPU, with their parity gates:

  C2  stream with 50 % same-address duplicates -> put path incl. the key table (overwrites in place)
  C3  read-hit path: everything resident compressed in HBM, get all (lookup + LZ4 decode)

Pages are generated on the device slice by slice (a 16 GiB stream does not need 16 GiB of host
RAM); gates: entries == distinct keys, every get hits, and every decoded page equals the
regenerated input (compared on the device).  Device-timed with the engine's own CUDA events.

    python tools/configs_bench.py [--gib 16] [--dup 0.5]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import edge_fuse_b200 as E

CH = 65536


class DevView:
    """torch view of a raw device allocation (for on-device comparisons only)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=16.0)
    ap.add_argument("--dup", type=float, default=0.5)
    ap.add_argument("--slice", type=int, default=32768, help="chunks generated / put per call")
    a = ap.parse_args()
    n = int(a.gib * (1 << 30) / CH)
    cids, distinct = E.gen_stream_ids(n, a.dup)
    off, nh = E.gen_addr(42, cids, 16)
    page = off >> np.uint64(16)
    eng = E.Engine(pshift=16, accel=12, capacity=2 * distinct, arena_bytes=int(distinct * CH * 0.72) + (1 << 30),
                   max_batch=16384, flags=E.FINGERPRINT)
    S = a.slice
    d_in = eng.dev_alloc(S * CH)
    d_out = eng.dev_alloc(S * CH)
    t_in = torch.as_tensor(DevView(d_in, S * CH), device="cuda")
    t_out = torch.as_tensor(DevView(d_out, S * CH), device="cuda")
    s0 = eng.stats()
    w0 = time.perf_counter()
    for at in range(0, n, S):
        m = min(S, n - at)
        eng.gen_chunks_dev(42, cids[at:at + m], d_in)
        eng.put(nh[at:at + m], page[at:at + m], d_in, on_dev=True)
    s1 = eng.stats()
    put_s = (s1["encode_kernel_ns"] - s0["encode_kernel_ns"]) * 1e-9
    put_wall = time.perf_counter() - w0
    assert s1["entries"] == distinct, (s1["entries"], distinct)
    assert s1["dropped_puts"] == 0
    # C3: get every distinct key (last content written under it = the chunk itself: repeats carry the same id)
    qc = np.arange(distinct, dtype=np.uint64)
    qo, qn = E.gen_addr(42, qc, 16)
    qp = qo >> np.uint64(16)
    bad = 0
    for at in range(0, distinct, S):
        m = min(S, distinct - at)
        _, status = eng.get(qn[at:at + m], qp[at:at + m], out=d_out, on_dev=True)
        assert (status == E.HIT).all()
        eng.gen_chunks_dev(42, qc[at:at + m], d_in)
        torch.cuda.synchronize()
        bad += int((t_in[: m * CH] != t_out[: m * CH]).any().item())
    s2 = eng.stats()
    get_s = (s2["decode_kernel_ns"] - s1["decode_kernel_ns"]) * 1e-9
    assert bad == 0, "decoded pages differ from the regenerated stream"
    res = {
        "C2": {"stream_gib": n * CH / 2**30, "chunks": n, "distinct": distinct, "dup_frac": 1 - distinct / n,
               "put_gibs_kernel": n * CH / 2**30 / put_s, "put_gibs_wall_incl_generation": n * CH / 2**30 / put_wall,
               "entries": s1["entries"], "arena_used_gib": s1["arena_used"] / 2**30,
               "arena_garbage_gib": s1["arena_garbage"] / 2**30, "gate": "entries == distinct, no dropped puts"},
        "C3": {"resident_pages_gib": distinct * CH / 2**30, "get_gibs_kernel": distinct * CH / 2**30 / get_s,
               "hits": int(s2["get_hits"] - s1["get_hits"]), "gate": "all hit, decoded pages == regenerated input"},
    }
    print(json.dumps(res))
    eng.close()


if __name__ == "__main__":
    main()
