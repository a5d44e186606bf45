"""Time single engine conv layers (tcgen05 path) at bench shapes: python tools/prof_layer.py [name ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from yolact_minimal_b200 import _lib
LAYERS = {  # name: (B, Cin, H, Cout, k, stride, relu, residual)
    'expand35': (64, 256, 35, 1024, 1, 1, 1, 1),
    'expand35_nores': (64, 256, 35, 1024, 1, 1, 1, 0),
    'expand138': (64, 64, 138, 256, 1, 1, 1, 1),
    'expand69': (64, 128, 69, 512, 1, 1, 1, 1),
    'expand18': (64, 512, 18, 2048, 1, 1, 1, 1),
    'reduce35': (64, 1024, 35, 256, 1, 1, 1, 0),
    'c3x3_35': (64, 256, 35, 256, 3, 1, 1, 0),
    'c3x3_64_138': (64, 64, 138, 64, 3, 1, 1, 0),
}
os.environ.setdefault('YOLACT_B200_CONV_REPS', '20')
L = _lib.lib()
dev = torch.device('cuda:0')
for name in (sys.argv[1:] or list(LAYERS)):
    B, Cin, H, Cout, k, stride, relu, res = LAYERS[name]
    Ho = (H - 1) // 2 + 1 if stride == 2 else H
    x = torch.randn(B, Cin, H, H, device=dev)
    r = torch.randn(B, Cout, Ho, Ho, device=dev) if res else None
    w = (np.random.randn(Cout, Cin, k, k) / np.sqrt(Cin * k * k)).astype(np.float32)
    b = np.zeros(Cout, np.float32)
    out = torch.empty(B, Cout, Ho, Ho, device=dev)
    print(name, flush=True)
    _lib.check(L.yb_conv2d(x.data_ptr(), B, Cin, H, w.ctypes.data, b.ctypes.data, Cout, k, stride, relu,
                           r.data_ptr() if r is not None else None, 2, 1, out.data_ptr()), 'yb_conv2d')
