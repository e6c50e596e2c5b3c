"""TEST INFRASTRUCTURE ONLY — independent writer / reader of the cache-directory snapshot format.

The product writes and reads `<cachedir>/cachemap_b200.snap` in CUDA/C++ (engine.cu, cmb200_save /
cmb200_load).  This module restates the format from its description (include/cachemap_b200.h,
DESIGN.md §0 f3) in plain Python over the oracle's store model, so that the tests can hand the
engine a file it did not write, and read back one it did:

    header  "CMB200S1" | u32 version = 1 | u32 pshift | u64 records | u64 record bytes | u32 flags | pad to 64
    record  u64 ts | u64 fp_hi | u64 fp_lo | u32 len | u32 0 | len bytes | pad to 16
            (the len bytes are the reference's LMDB value: {u64 u, u64 l, i32 compressed_length, 4 pad}
             + LZ4 block or raw page, cachemap/filemap.c:9-12,140-147)
    flags   bit 0: the fingerprints are meaningful
"""
from __future__ import annotations

import struct

MAGIC = b"CMB200S1"


def write_snapshot(path: str, pshift: int, records, with_fingerprints: bool = False) -> None:
    """records: iterable of (ts, fp_hi, fp_lo, record_bytes) with record_bytes = prefix + payload."""
    records = list(records)
    total = sum(len(r[3]) for r in records)
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<IIQQI", 1, pshift, len(records), total, 1 if with_fingerprints else 0) + b"\0" * 28)
        for ts, hi, lo, rec in records:
            f.write(struct.pack("<QQQII", ts, hi, lo, len(rec), 0))
            f.write(rec)
            f.write(b"\0" * ((16 - len(rec) % 16) % 16))


def read_snapshot(path: str):
    """-> (pshift, flags, [(ts, fp_hi, fp_lo, record_bytes)])"""
    with open(path, "rb") as f:
        head = f.read(64)
        assert head[:8] == MAGIC, "not a snapshot"
        version, pshift, count, total, flags = struct.unpack("<IIQQI", head[8:36])
        assert version == 1
        out = []
        for _ in range(count):
            ts, hi, lo, n, zero = struct.unpack("<QQQII", f.read(32))
            rec = f.read(n)
            f.read((16 - n % 16) % 16)
            out.append((ts, hi, lo, rec))
        assert sum(len(r[3]) for r in out) == total and f.read(1) == b""
    return pshift, flags, out
