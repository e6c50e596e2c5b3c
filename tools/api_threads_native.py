#!/usr/bin/env python
"""Builds tools/api_threads.c and runs it for the drop-in and for the reference's own library
(oracle/_ref, LMDB on tmpfs): GiB/s of 64 KiB pages through cachemap_put / cachemap_get, one page
per call, from T native threads.  One JSON object on the last line."""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import edge_fuse_b200 as E


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", default="1,8,32,64")
    ap.add_argument("--per-thread", type=int, default=512)
    ap.add_argument("--no-ref", action="store_true")
    a = ap.parse_args()
    os.environ.setdefault("CMB200_ARENA_MB", "8192")
    work = tempfile.mkdtemp(prefix="api_threads_")
    exe = os.path.join(work, "api_threads")
    subprocess.run(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "tools", "api_threads.c"), "-ldl", "-lpthread"], check=True)
    pages = np.stack([E.gen_chunk_host(42, c, 65536) for c in range(64)])
    pbin = os.path.join(work, "pages.bin")
    pages.tofile(pbin)
    libs = {"ours": os.environ.get("CMB200_LIB") or E.library_path()}
    ref = os.path.join(ROOT, "oracle", "_ref", "libcachemap_ref.so")
    if os.path.exists(ref) and not a.no_ref:
        libs["reference_cpu"] = ref
    out = {}
    for name, lib in libs.items():
        for t in [int(x) for x in a.threads.split(",")]:
            per = min(a.per_thread, 60000 // t)
            with tempfile.TemporaryDirectory(dir="/dev/shm" if name != "ours" else None) as d:
                r = subprocess.run([exe, lib, d, pbin, str(t), str(per)], capture_output=True, text=True, timeout=120)
            line = [x for x in r.stdout.splitlines() if x.startswith("{")]
            if not line:
                print(name, t, "failed", r.stdout[-300:], r.stderr[-300:], flush=True)
                continue
            out[f"{name}_T{t}"] = json.loads(line[-1])
            print(name, t, out[f"{name}_T{t}"], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
