"""Fixed cost per launch of the tcgen05 conv kernel: time layer shapes at 1, 2, 3, 4 full rounds of tiles
(batch chosen so that B*(H+2)^2/128 ~ k*148) and print per-round increments and the intercept.
python tools/prof_t0.py   (on the GPU box; set YOLACT_B200_NO_PDL=1 for the A/B run)"""
import os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SHAPES = {  # name: (Cin, H, Cout, k, residual)
    'reduce35': (1024, 35, 256, 1, 0), 'c3x3_35': (256, 35, 256, 3, 0), 'expand35': (256, 35, 1024, 1, 1),
    'c3x3_18': (512, 18, 512, 3, 0), 'expand18': (512, 18, 2048, 1, 1),
}
if len(sys.argv) > 1 and sys.argv[1] == '--one':
    import numpy as np, torch
    from yolact_minimal_b200 import _lib
    name, B = sys.argv[2], int(sys.argv[3])
    Cin, H, Cout, k, res = SHAPES[name]
    os.environ['YOLACT_B200_CONV_REPS'] = '50'
    L = _lib.lib(); dev = torch.device('cuda:0')
    x = torch.randn(B, Cin, H, H, device=dev)
    r = torch.randn(B, Cout, H, H, device=dev) if res else None
    w = (np.random.randn(Cout, Cin, k, k) / np.sqrt(Cin * k * k)).astype(np.float32)
    b = np.zeros(Cout, np.float32)
    out = torch.empty(B, Cout, H, H, device=dev)
    _lib.check(L.yb_conv2d(x.data_ptr(), B, Cin, H, w.ctypes.data, b.ctypes.data, Cout, k, 1, 1,
                           r.data_ptr() if r is not None else None, 2, 1, out.data_ptr()), 'yb_conv2d')
    sys.exit(0)
for name, (Cin, H, Cout, k, res) in SHAPES.items():
    plane = (H + 2) ** 2
    n_tiles = max(1, Cout // 256)
    ts = []
    for rounds in (1, 2, 3, 4):
        B = (rounds * 148 // n_tiles * 128) // plane
        p = subprocess.run([sys.executable, __file__, '--one', name, str(B)], capture_output=True, text=True)
        m = re.search(r'([\d.]+) us/launch', p.stderr)
        ts.append((B, -(-B * plane // 128) * n_tiles, float(m.group(1)) if m else float('nan')))
    inc = (ts[3][2] - ts[0][2]) / 3
    print(name, ' '.join('B=%d tiles=%d %.1fus' % t for t in ts), '| per round %.1f us, intercept %.1f us' % (inc, ts[0][2] - inc), flush=True)
