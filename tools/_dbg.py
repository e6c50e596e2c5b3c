import sys, os, json, hashlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, datagen
import edge_fuse_b200 as E
from oracle import ef_oracle as O
g = json.load(open('tests/golden/lz4_blocks.json'))['cases']
bad = 0
for r in g:
    if r['n'] == 0: continue
    p = datagen.make_page(r['kind'], r['n'], r['seed'])
    b, _ = E.lz4_encode_batch(datagen.pad_rows([p]), nbytes=r['n'], accel=r['accel'])
    exp = O.lz4_encode(p, r['accel'])
    if b[0] != exp:
        bad += 1
        d = next((i for i in range(min(len(b[0]), len(exp))) if b[0][i] != exp[i]), -1)
        print('MISMATCH', r['kind'], r['n'], r['accel'], r['seed'], len(b[0]), len(exp), 'first diff at', d)
print('bad', bad, 'of', len(g))
