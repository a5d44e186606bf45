"""Probe: does a UMMA descriptor that starts 128 B (one row) into a 128B-swizzled tile read rows 1..128 correctly,
and does it need base_offset = 1?  Compares a 1x1 conv run normally against the same conv with the A tile loaded one row
early and consumed through the shifted descriptor (rows whose slab row falls outside the 128-row box are ignored)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from yolact_minimal_b200 import _lib
L = _lib.lib(); dev = torch.device('cuda:0')
B, Cin, H, Cout = 2, 128, 35, 128
torch.manual_seed(0); np.random.seed(0)
x = torch.randn(B, Cin, H, H, device=dev)
w = (np.random.randn(Cout, Cin, 1, 1) / np.sqrt(Cin)).astype(np.float32); b = np.zeros(Cout, np.float32)
def run():
    out = torch.empty(B, Cout, H, H, device=dev)
    _lib.check(L.yb_conv2d(x.data_ptr(), B, Cin, H, w.ctypes.data, b.ctypes.data, Cout, 1, 1, 0, None, 2, 1, out.data_ptr()), 'yb_conv2d')
    return out.cpu().numpy()
os.environ['YOLACT_B200_PAIR'] = '0'
ref = run()
Hp = H + 2
m = ((np.arange(B)[:, None, None] * Hp + (np.arange(H)[None, :, None] + 1)) * Hp + (np.arange(H)[None, None, :] + 1))   # haloed row index
ok_rows = (m % 128) != 127
for mode in (1, 2):
    os.environ['YOLACT_B200_DBG_SLAB'] = str(mode)
    y = run()
    d = np.abs(y - ref).max(axis=1)
    print('mode', mode, '(base_offset %d)' % (1 if mode == 1 else 0), 'max err on covered rows', float(d[ok_rows].max()), 'on uncovered', float(d[~ok_rows].max()))
