"""Cycle stamps of the fused bottleneck kernel (leader CTA of pair 0, first 64 chunks of the last k_bneck_tc launch of a forward).
usage: YOLACT_B200_BNECK_TRACE=1 python tools/bneck_trace.py [batch]
events per chunk q:  0 MMA warp reaches A(q)   1 A(q) issued + committed   2 MMA warp reaches B(q)   3 xs_full(q) observed
                     4 epilogue warp 2 observes accA_full(q)   5 its TMEM loads are back   6 arrive + TMA store issued"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('YOLACT_B200_BNECK_TRACE', '1')
from oracle import forward_torch as ft
from yolact_minimal_b200 import _lib
from yolact_minimal_b200.config import make_config
from yolact_minimal_b200.modules.yolact import Yolact

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = make_config('res101_coco', 550)
cfg.precision, cfg.max_batch = 'fp16', B
net = Yolact(cfg)
net.load_state_dict(ft.synth_state_dict('res101', seed=0), strict=True)
net = net.cuda().eval()
img = torch.randn(B, 3, 550, 550, device='cuda')
with torch.no_grad():
    for _ in range(3):
        net(img)
torch.cuda.synchronize()
L = _lib.load() if hasattr(_lib, 'load') else net.engine(B).L
buf = (ctypes.c_ulonglong * 512)()
L.yb_debug_bneck_trace.restype = ctypes.c_int
rc = L.yb_debug_bneck_trace(buf, 512)
assert rc == 0, rc
t = np.array(buf, dtype=np.int64).reshape(8, 64)
t0 = t[0, 0]
names = ['A reach', 'A issued', 'B reach', 'xs_full', 'accA_full', 'ld back', 'arrived']
print('q   ' + '  '.join(f'{n:>9s}' for n in names) + '   | E1 latency (accA_full -> xs_full seen by MMA)')
for q in range(40):
    row = [(t[e, q] - t0) if t[e, q] else -1 for e in range(7)]
    print(f'{q:2d}  ' + '  '.join(f'{v:9d}' for v in row) + f'   | {row[3] - row[4]:6d}  ld {row[5] - row[4]:5d}  pack+arrive {row[6] - row[5]:5d}')
