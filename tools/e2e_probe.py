import sys, os, time
sys.path.insert(0, '.')
import numpy as np, edge_fuse_b200 as E
n=16384; CH=65536
for mb in (4096,):
    eng = E.Engine(pshift=16, accel=12, capacity=64*n, arena_bytes=24<<30, max_batch=mb, flags=E.FINGERPRINT)
    cids=np.arange(n,dtype=np.uint64); off,nh=E.gen_addr(42,cids,16)
    d=eng.dev_alloc(n*CH); eng.gen_chunks_dev(42,cids,d)
    hp=E.lib().cmb200_host_alloc(n*CH)
    h=np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8*(n*CH)).from_address(hp)); eng.d2h(h,d)
    page=off>>np.uint64(16)
    for it in range(3):
        l=page|(np.uint64(it+1)<<np.uint64(44))
        t0=time.perf_counter(); eng.put(nh,l,hp,on_dev=False); t1=time.perf_counter()
        print("host ms", round((t1-t0)*1e3,1), flush=True)
