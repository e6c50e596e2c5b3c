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
