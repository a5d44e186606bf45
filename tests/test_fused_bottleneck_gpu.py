"""GPU: the fused bottleneck kernel k_bneck_tc (conv3 + residual + ReLU of one ResNet block chained into conv1 of the next through
shared memory, modules/resnet.py:20-40) computes exactly what the two separate k_conv_tc launches compute: same operand rounding
points, same k order, residual added by the tensor core after the main k-blocks -- so the network outputs and the stage output c4
must be BIT-IDENTICAL with the fusion on and off (YOLACT_B200_NO_FUSE=1), at tile counts below / at / above one wave and with ragged
last tiles.  (The 1e-2 / 1e-3 parity against the fp32 oracle is tests/test_forward_gpu.py and tests/test_parity_gaps_gpu.py: they
run fused.)"""
import os

import numpy as np
import pytest
import torch

from oracle import synth, forward_torch as ft

pytestmark = pytest.mark.gpu


def _outputs(arch, S, B, precision, cuda, fuse, env=None):
    from yolact_minimal_b200.config import make_config
    from yolact_minimal_b200.modules.yolact import Yolact
    old = os.environ.pop('YOLACT_B200_NO_FUSE', None)
    if not fuse:
        os.environ['YOLACT_B200_NO_FUSE'] = '1'
    for k, v in (env or {}).items():
        os.environ[k] = v
    try:
        cfg = make_config(arch + '_coco', S)
        cfg.precision, cfg.max_batch = precision, B
        net = Yolact(cfg)
        net.load_state_dict(ft.synth_state_dict(arch, seed=0), strict=True)
        net = net.to(cuda).eval()
        img = torch.from_numpy(synth.image_batch(21, B, S)).to(cuda)
        with torch.no_grad():
            out = [o.clone() for o in net(img)]
            again = [o.clone() for o in net(img)]
        c4 = net.engine(B).read_activation('c4', B).clone()
        launches = net.engine(B).launches_per_forward() if hasattr(net.engine(B), 'launches_per_forward') else None
        torch.cuda.synchronize()
    finally:
        for k in (env or {}):
            os.environ.pop(k, None)
        os.environ.pop('YOLACT_B200_NO_FUSE', None)
        if old is not None:
            os.environ['YOLACT_B200_NO_FUSE'] = old
    return out, again, c4, launches


@pytest.mark.parametrize('arch,S,B,precision', [('res101', 64, 1, 'fp16'), ('res101', 128, 3, 'fp16'), ('res50', 256, 2, 'fp16'),
                                                ('res101', 550, 2, 'fp16'), ('res101', 550, 9, 'fp16'), ('res101', 320, 5, 'bf16')])
def test_fused_equals_unfused_bitwise(cuda, arch, S, B, precision):
    fused, fused2, c4f, _ = _outputs(arch, S, B, precision, cuda, True, env={'YOLACT_B200_NO_FUSE_DOWN': '1'})
    plain, _, c4p, _ = _outputs(arch, S, B, precision, cuda, False)
    assert torch.equal(c4f, c4p), float((c4f - c4p).abs().max())
    for name, a, b, c in zip(('cls', 'box', 'coef', 'proto'), fused, plain, fused2):
        assert torch.equal(a, c), name                                   # deterministic run to run
        assert torch.equal(a, b), (name, float((a - b).abs().max()))
    assert np.isfinite(fused[0].cpu().numpy()).all()


@pytest.mark.parametrize('arch,S,B', [('res50', 128, 2), ('res101', 550, 3)])
def test_stem_on_pixel_rows_equals_overlapping_rows_bitwise(cuda, arch, S, B):
    """The stem convolution on 32-byte pixel rows (one slab per tap row, the dx taps as row-shifted K = 16 MMAs) issues the same MMAs in the
    same order as the overlapping-rows K = 64 form it replaced (YOLACT_B200_NO_STEM16=1): the network output must not change by a bit."""
    new, _, c4n, _ = _outputs(arch, S, B, 'fp16', cuda, True, env={'YOLACT_B200_NO_FUSE_DOWN': '1'})
    old, _, c4o, _ = _outputs(arch, S, B, 'fp16', cuda, True, env={'YOLACT_B200_NO_STEM16': '1', 'YOLACT_B200_NO_FUSE_DOWN': '1'})
    assert torch.equal(c4n, c4o), float((c4n - c4o).abs().max())
    for name, a, b in zip(('cls', 'box', 'coef', 'proto'), new, old):
        assert torch.equal(a, b), (name, float((a - b).abs().max()))


@pytest.mark.parametrize('arch,S,B,precision', [('res50', 128, 2, 'fp16'), ('res101', 550, 3, 'fp16'), ('res101', 256, 5, 'bf16')])
def test_folded_downsample_branch_matches_unfused(cuda, arch, S, B, precision):
    """First block of layer1: the 1x1 downsample convolution of the residual branch is folded into the fused kernel's first GEMM
    (conv3 and downsample weights side by side along K, biases summed), so the residual is never rounded to 16 bits / written / re-read.
    One rounding point fewer than the reference-ordered path: outputs agree with the unfused network to a few 16-bit ulps, and the
    fused path is deterministic.  (Parity against the fp32 oracle: tests/test_forward_gpu.py, which runs this form.)"""
    fused, fused2, c4f, _ = _outputs(arch, S, B, precision, cuda, True)
    plain, _, c4p, _ = _outputs(arch, S, B, precision, cuda, False)
    ulp = 2.0 ** -10 if precision == 'fp16' else 2.0 ** -7
    dmax, dmean = float((c4f - c4p).abs().max()), float((c4f - c4p).abs().mean())
    assert dmax <= 8 * ulp * float(c4p.abs().max()), (dmax, float(c4p.abs().max()))
    assert dmean <= 2 * ulp * float(c4p.abs().mean()) + 1e-12, (dmean, float(c4p.abs().mean()))
    for name, a, b, c in zip(('cls', 'box', 'coef', 'proto'), fused, plain, fused2):
        assert torch.equal(a, c), name
        assert float((a - b).abs().max()) <= 8 * ulp * max(1.0, float(b.abs().max())), (name, float((a - b).abs().max()))
