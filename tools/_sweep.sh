for ring in 0 2048 4096; do echo "ring=$ring"; CMB200_ENC_RING=$ring timeout 120 python tools/kernel_bench.py --classes RTZMB --chunks 4096 --reps 2 2>&1 | grep -E '^[RTZMB] ' | cut -c1-60; done
echo accel13; timeout 100 python - <<'PY'
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, datagen, edge_fuse_b200 as E
from oracle import ef_oracle as O
import os
for accel in (13, 40):
    pages=[datagen.make_page(k,65536,5+i) for i,k in enumerate("RTZMPAXS")]
    b,_=E.lz4_encode_batch(np.stack(pages),accel=accel)
    print(accel, all(x==O.lz4_encode(p,accel) for x,p in zip(b,pages)))
PY
CMB200_ENC_RING=0 timeout 300 python -m pytest tests -m gpu -q --timeout 240 -x 2>&1 | tail -2
