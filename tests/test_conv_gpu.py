"""GPU: one engine convolution layer (yb_conv2d through the C ABI) vs a plain PyTorch fp32
reference of the same op, at shapes taken from SURVEY.md App. A (incl. the odd 550 extents,
stride 2, residual, multi-N-tile and multi-wave persistent cases).  fp32 = CUDA-core kernel;
bf16 / fp16 = tcgen05 kernel (operands pre-rounded in the reference, fp32 accumulation)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import synth

pytestmark = pytest.mark.gpu

# (B, Cin, H, Cout, k, stride, relu, residual)
SHAPES = [
    (1, 64, 8, 64, 1, 1, 0, 0),
    (2, 64, 17, 64, 3, 1, 1, 0),
    (2, 256, 35, 256, 3, 1, 1, 0),
    (1, 128, 69, 128, 3, 2, 1, 0),
    (2, 512, 18, 2048, 1, 1, 1, 1),
    (1, 256, 35, 512, 1, 2, 0, 0),
    (1, 256, 69, 32, 1, 1, 1, 0),
    (2, 256, 9, 352, 3, 1, 0, 0),
    (2, 256, 5, 256, 3, 2, 1, 0),
    (3, 1024, 35, 256, 1, 1, 1, 0),
    (4, 256, 138, 256, 3, 1, 1, 0),
    (2, 96, 35, 288, 1, 1, 0, 0),      # Swin qkv, Cin not a multiple of 64 (TMA zero-fills the K tail)
    (2, 96, 35, 384, 1, 1, 2, 0),      # Swin fc1 + exact GELU
    (2, 384, 35, 96, 1, 1, 0, 1),      # Swin fc2 + residual
    (1, 768, 18, 2304, 1, 1, 0, 0),    # Swin stage-3 qkv
]


def conv_case(seed, B, Cin, H, Cout, k, stride, res):
    x = synth.normal(seed, 1, (B, Cin, H, H)).astype(np.float32)
    w = (synth.normal(seed, 2, (Cout, Cin, k, k)) / np.sqrt(Cin * k * k)).astype(np.float32)
    b = (0.1 * synth.normal(seed, 3, (Cout,))).astype(np.float32)
    Ho = (H - 1) // 2 + 1 if stride == 2 else H
    r = synth.normal(seed, 4, (B, Cout, Ho, Ho)).astype(np.float32) if res else None
    return x, w, b, r


def run_conv(cuda, x, w, b, r, k, stride, relu, precision, use_tc):
    from yolact_minimal_b200 import _lib
    L = _lib.lib()
    B, Cin, H, _ = x.shape
    Cout = w.shape[0]
    Ho = (H - 1) // 2 + 1 if stride == 2 else H
    xd = torch.from_numpy(x).to(cuda)
    rd = torch.from_numpy(r).to(cuda) if r is not None else None
    out = torch.empty(B, Cout, Ho, Ho, dtype=torch.float32, device=cuda)
    wc, bc = np.ascontiguousarray(w), np.ascontiguousarray(b)
    _lib.check(L.yb_conv2d(xd.data_ptr(), B, Cin, H, wc.ctypes.data, bc.ctypes.data, Cout, k, stride, relu,
                           rd.data_ptr() if rd is not None else None, precision, use_tc, out.data_ptr()), 'yb_conv2d')
    return out


def reference(cuda, x, w, b, r, k, stride, relu, rnd):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    q = (lambda t: t) if rnd is None else (lambda t: t.to(rnd).float())
    y = F.conv2d(q(torch.from_numpy(x).to(cuda)).double(), q(torch.from_numpy(w).to(cuda)).double(),
                 torch.from_numpy(b).to(cuda).double(), stride=stride, padding=k // 2)
    if r is not None:
        y = y + q(torch.from_numpy(r).to(cuda)).double()
    if relu == 1:
        y = F.relu(y)
    elif relu == 2:
        y = F.gelu(y)
    return y.float()


@pytest.mark.parametrize('shape', SHAPES, ids=lambda s: 'B%d_Cin%d_H%d_Cout%d_k%d_s%d_relu%d_res%d' % s)
def test_conv_fp32_simt(cuda, shape):
    B, Cin, H, Cout, k, stride, relu, res = shape
    if B * H * H * Cout * Cin * k * k > 3e11:
        pytest.skip('too slow for the CUDA-core kernel in a unit test')
    x, w, b, r = conv_case(5, B, Cin, H, Cout, k, stride, res)
    y = run_conv(cuda, x, w, b, r, k, stride, relu, 0, 0)
    ref = reference(cuda, x, w, b, r, k, stride, relu, None)
    assert float((y - ref).abs().max()) < 1e-4 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize('pair', ['auto', 'single', 'pair', 'resmma', 'pair+resmma'])
@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
@pytest.mark.parametrize('shape', SHAPES, ids=lambda s: 'B%d_Cin%d_H%d_Cout%d_k%d_s%d_relu%d_res%d' % s)
def test_conv_tcgen05(cuda, shape, precision, pair, monkeypatch):
    """`pair` forces the launch form of the tcgen05 kernel: one CTA per 128-row tile, or a CTA pair (cluster of 2,
    cta_group::2 MMAs over both CTAs' operands) per 256-row tile; 'resmma' forces the residual-through-the-tensor-core
    form (identity k-blocks) wherever the tile width allows; 'auto' is the per-layer choice the engine makes."""
    if pair != 'auto':
        monkeypatch.setenv('YOLACT_B200_PAIR', '1' if pair.startswith('pair') else '0')
        monkeypatch.setenv('YOLACT_B200_RESMMA', '1' if pair.endswith('resmma') else '0')
    if pair.endswith('resmma') and not shape[7]:
        pytest.skip('no residual')
    B, Cin, H, Cout, k, stride, relu, res = shape
    x, w, b, r = conv_case(6, B, Cin, H, Cout, k, stride, res)
    prec, rnd, eps = (1, torch.bfloat16, 2.0 ** -8) if precision == 'bf16' else (2, torch.float16, 2.0 ** -11)
    ref = reference(cuda, x, w, b, r, k, stride, relu, rnd)
    y_tc = run_conv(cuda, x, w, b, r, k, stride, relu, prec, 1)
    scale = max(1.0, float(ref.abs().max()))
    err = float((y_tc - ref).abs().max())
    assert err < 1.5 * eps * scale, (err, scale)          # only the final 16-bit rounding of the output
    if pair == 'auto':
        y_simt = run_conv(cuda, x, w, b, r, k, stride, relu, prec, 0)
        assert float((y_tc - y_simt).abs().max()) < 1.5 * eps * scale       # same math on CUDA cores
