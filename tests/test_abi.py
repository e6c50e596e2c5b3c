"""The C-ABI library loads without a GPU and exports every symbol the headers in include/
declare.  No compute calls here (CPU box)."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for h in ("cachemap.h", "filemap.h", "cachemap_b200.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"\b([a-z_0-9]+)\s*\(", text):
            n = m.group(1)
            if n.startswith(("cachemap_", "filemap_", "cmb200_")):
                names.add(n)
    return names


def test_exports_match_headers(E):
    decl = declared_functions()
    assert decl == set(E.EXPORTED_SYMBOLS), decl ^ set(E.EXPORTED_SYMBOLS)
    L = ctypes.CDLL(E.library_path())
    for name in sorted(decl):
        assert hasattr(L, name), name


def test_reference_surface_present(E):
    # the thirteen functions of the reference's cachemap.h:33-47 and filemap.h:19-29
    ref = ["cachemap_create", "cachemap_free", "cachemap_get", "cachemap_put", "cachemap_put_async",
           "cachemap_print_stats", "filemap_create", "filemap_free", "filemap_set", "filemap_unset",
           "filemap_get", "filemap_get_rand", "filemap_entries"]
    L = ctypes.CDLL(E.library_path())
    for n in ref:
        assert hasattr(L, n)


def test_create_argument_checks_need_no_gpu(E, tmp_path):
    # cachemap.c:113-114 / filemap.c:51: NULL for a missing directory or capacity < 1024; the
    # device is not touched at create time (lazy init, fork safety).
    assert not E.Cachemap(str(tmp_path / "missing"), 4096).ok
    assert not E.Cachemap(str(tmp_path), 1023).ok
    cm = E.Cachemap(str(tmp_path), 1024)
    assert cm.ok
    cm.free()


def test_fnv_hash_header_inline(tmp_path):
    # include/uint128.h's FNV_hash must give the reference's values (edgefs.c:209,1911 call it).
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "uint128.h"\n'
                   'int main(void){uint64_t h; FNV_hash("/bk1",4,&h); printf("%016lx ",(unsigned long)h);'
                   'uint128_t a={0x1122334455667788ULL,(7ULL<<44)|3}; FNV_hash(&a,sizeof a,&h);'
                   'printf("%016lx %zu\\n",(unsigned long)h,sizeof a);return 0;}\n')
    exe = tmp_path / "t"
    import subprocess
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert out == ["1e400c9ca688f534", "1e041ed74a444846", "16"]


def test_host_stream_generator(E):
    # class layout of the synthetic stream and determinism
    r, t, z, m = (E.gen_chunk_host(42, c, 65536) for c in range(4))
    assert (z[2:] == 0).all() and z[0] == 2 and z[1] == 0
    assert (m[:32768] == m[32768:]).all()
    assert ((t >= 97) & (t <= 100)).mean() > 0.7
    assert len(np.unique(r)) == 256
    assert (E.gen_chunk_host(42, 5, 4096) == E.gen_chunk_host(42, 5, 4096)).all()
    cids, distinct = E.gen_stream_ids(10000, 0.5)
    assert distinct == len(np.unique(cids)) and 0.4 < 1 - distinct / 10000 < 0.6
    assert cids.max() == distinct - 1
    cids0, d0 = E.gen_stream_ids(1000, 0.0)
    assert d0 == 1000 and (cids0 == np.arange(1000)).all()


def test_edgefs_glue_header_matches_oracle(tmp_path, oracle):
    """include/edgefs_glue.h (the gate and the object id that edgefs.c computes
    before it calls the cache, edgefs.c:192-212,1911) against the oracle restatement, plus known
    answers taken with the reference's own FNV_hash."""
    import subprocess
    so = tmp_path / "glue_shim.so"
    subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "glue_shim.c"), "-o", str(so)])
    L = ctypes.CDLL(str(so))
    L.shim_cache_check.restype = ctypes.c_int
    L.shim_cache_check.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64,
                                   ctypes.c_void_p, ctypes.c_void_p]
    L.shim_build_nhid.restype = ctypes.c_uint64
    L.shim_build_nhid.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    rng = np.random.default_rng(5)
    for pshift in range(12, 18):
        page = 1 << pshift
        cases = [(0, 0), (0, page), (page, 2 * page), (1, page), (page, page - 1), (page - 1, 1),
                 (3 * page, 131072), ((1 << 60) + page, page), (2**64 - page, page)]
        cases += [(int(rng.integers(0, 1 << 40)) & ~(page - 1 if rng.random() < 0.5 else 0),
                   int(rng.integers(0, 1 << 20)) & ~(page - 1 if rng.random() < 0.5 else 0)) for _ in range(200)]
        for have in (0, 1):
            for off, size in cases:
                ps, ao = ctypes.c_uint64(), ctypes.c_uint64()
                got = L.shim_cache_check(have, pshift, off, size, ctypes.byref(ps), ctypes.byref(ao))
                assert (bool(got), ps.value, ao.value) == oracle.cache_check(bool(have), pshift, off, size)
    # FNV_hash("/bk1") = 0x1e400c9ca688f534, FNV_hash("") = offset basis (reference objects)
    assert L.shim_build_nhid(b"", b"/bk1") == 0xcbf29ce484222325 ^ 0x1e400c9ca688f534
    assert L.shim_build_nhid(b"a", b"") == 0xaf63dc4c8601ec8c ^ 0xcbf29ce484222325
    for _ in range(100):
        name = bytes(rng.integers(1, 256, int(rng.integers(0, 64)), dtype=np.uint8))
        path = b"/" + bytes(rng.integers(1, 256, int(rng.integers(0, 200)), dtype=np.uint8))
        assert L.shim_build_nhid(name, path) == oracle.build_nhid(name, path) \
            == oracle.fnv1a64(name) ^ oracle.fnv1a64(path)


def test_snapshot_restatement_roundtrip(tmp_path, oracle):
    """oracle/snapshot.py (the independent statement of the cache-directory file format used by the
    GPU tests) writes what it reads: header fields, 16-byte padding, timestamps, record bytes."""
    import datagen
    from oracle import snapshot
    model = oracle.StoreModel(12, 12)
    recs = []
    for i, kind in enumerate("RTZMPAXS"):
        page = datagen.make_page(kind, 4096, 50 + i)
        model.put(i << 12, 9, 0, page)
        u, l = oracle.addr_compose(i << 12, 9, 0, 12)
        recs.append((100 + i, i, ~i & 0xFFFFFFFFFFFFFFFF, model.record_bytes(u, l)))
    path = str(tmp_path / "s.snap")
    snapshot.write_snapshot(path, 12, recs, with_fingerprints=True)
    assert os.path.getsize(path) == 64 + sum(32 + ((len(r[3]) + 15) & ~15) for r in recs)
    pshift, flags, got = snapshot.read_snapshot(path)
    assert (pshift, flags, got) == (12, 1, recs)
