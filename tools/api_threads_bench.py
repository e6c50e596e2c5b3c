#!/usr/bin/env python
"""Throughput of the DROP-IN single-page API (cachemap_put / cachemap_get, one 64 KiB page per
call, as edgefs_read/write issue them) from T caller threads — the flat-combining path — next
to the reference's own library under the same calls (LMDB on tmpfs).  ctypes releases the GIL
around the C calls, so Python threads are real concurrent callers."""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import edge_fuse_b200 as E

CH = 65536


def run(lib, cm, pages, threads, per_thread, do_get):
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    errs = []

    def worker(t):
        try:
            for i in range(per_thread):
                p = pages[(t * per_thread + i) % len(pages)]
                off = (t * per_thread + i) << 16
                if do_get:
                    r = lib.cachemap_get(cm, off, 0x77, 0)
                    if not r:
                        errs.append("miss")
                    else:
                        libc.free(r)
                else:
                    lib.cachemap_put(cm, off, 0x77, 0, p.ctypes.data)
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    [x.start() for x in th]
    [x.join() for x in th]
    dt = time.perf_counter() - t0
    assert not errs, errs[:3]
    return threads * per_thread * CH / 2**30 / dt, threads * per_thread / dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-thread", type=int, default=256)
    ap.add_argument("--threads", default="1,8,32,64")
    ap.add_argument("--no-ref", action="store_true")
    a = ap.parse_args()
    pages = [E.gen_chunk_host(42, c, CH) for c in range(64)]
    os.environ.setdefault("CMB200_ARENA_MB", "8192")
    L = E.lib()
    out = {}
    tlist = [int(x) for x in a.threads.split(',')]
    for threads in tlist:
        with tempfile.TemporaryDirectory() as d:
            cm = L.cachemap_create(d.encode(), 1 << 16, 12, 16)
            L.cachemap_put(cm, 1 << 40, 1, 0, pages[0].ctypes.data)      # engine start outside the clock
            put = run(L, cm, pages, threads, a.per_thread, False)
            st0 = E.engine_stats(L.cachemap_engine(cm)) if hasattr(E, "engine_stats") else None
            get = run(L, cm, pages, threads, a.per_thread, True)
            st1 = E.engine_stats(L.cachemap_engine(cm)) if hasattr(E, "engine_stats") else None
            L.cachemap_free(cm)
        out[f"ours_T{threads}"] = {"put_gibs": put[0], "put_kops": put[1] / 1e3, "get_gibs": get[0], "get_kops": get[1] / 1e3}
        if st0 and st1:    # gets that reached the GPU and the launches that carried them
            out[f"ours_T{threads}"]["gpu_gets"] = st1["get_requests"] - st0["get_requests"]
            out[f"ours_T{threads}"]["get_launches"] = st1["kernel_launches"] - st0["kernel_launches"]
        print(threads, out[f"ours_T{threads}"], flush=True)
    from oracle import ef_oracle as O
    R = O.ref()
    if R is not None and not a.no_ref:
        for threads in tlist:
            with tempfile.TemporaryDirectory(dir="/dev/shm") as d:
                cm = R.cachemap_create(d.encode(), 1 << 16, 12, 16)
                put = run(R, cm, pages, threads, a.per_thread, False)
                get = run(R, cm, pages, threads, a.per_thread, True)
            out[f"ref_T{threads}"] = {"put_gibs": put[0], "get_gibs": get[0]}
            print("ref", threads, out[f"ref_T{threads}"], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
