"""GPU: the mask output stage and the numpy post-process twins (8(f) rank 4) vs the oracle and goldens
minted from the reference:
  * bit-packed masks written by after_nms == the packed uint8 masks; pack / unpack round trip
  * mask_iou (AND + popcount over packed words) == the reference's float matmul formulation (golden from utils/box_utils.py:189-200)
    and the numpy oracle, bit for bit; box_iou == the oracle's separately rounded restatement
  * RLE run lengths == the plain restatement of pycocotools' rleEncode on Fortran-ordered masks; ASCII compression round trip
  * nms_numpy / after_nms_numpy (no clip, bool masks) vs what the reference's own numpy functions returned."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import synth, postprocess_np as pp, train_np as tn

pytestmark = pytest.mark.gpu


def _after_nms_case(cuda, seed=5, S=128, h=80, w=120):
    from yolact_minimal_b200.utils.output_utils import after_nms
    anchors = pp.make_anchors(S)
    cls, box, coef = synth.head_outputs(seed, anchors.shape[0], 81, 'realistic')
    proto = synth.proto(seed, S // 4)
    ids, scores, boxes, aidx = pp.nms(cls, box, anchors)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    args = (t(ids), t(scores), t(boxes), t(coef[aidx]), t(proto), h, w)
    return after_nms(*args, mask_dtype=torch.uint8)[3], after_nms(*args, mask_dtype='bits')[3]


def test_bitpacked_masks_equal_byte_masks(cuda):
    from yolact_minimal_b200.utils import mask_utils as mu
    for h, w in ((80, 120), (97, 64), (33, 31), (60, 250)):
        u8, bits = _after_nms_case(cuda, h=h, w=w)
        assert bits.shape == (u8.shape[0], h, (w + 31) // 32) and u8.sum() > 0
        assert torch.equal(mu.pack_masks(u8), bits)                 # the assembly kernel's packed output == packing its byte output
        assert torch.equal(mu.unpack_masks(bits, w), u8)
        assert torch.equal(mu.pack_masks(u8.float()), bits)         # float32 {0,1} masks (the reference's dtype) pack identically


def test_mask_iou_and_box_iou(cuda):
    from yolact_minimal_b200.utils.box_utils import mask_iou, box_iou
    g = load_golden('train_stages.npz')
    m1 = (synth.uniform(31, 1, (5, 400)) > 0.5).astype(np.float32)
    m2 = (synth.uniform(31, 2, (7, 400)) > 0.6).astype(np.float32)
    got = mask_iou(torch.from_numpy(m1).to(cuda), torch.from_numpy(m2).to(cuda))
    assert not got.is_cuda                                          # the reference returns .cpu() (box_utils.py:200)
    assert np.array_equal(got.numpy(), g['mask_iou'])               # bit-equal to what the reference's matmul formulation returned
    assert np.array_equal(got.numpy(), tn.mask_iou(m1, m2))
    # an all-zero pair -> 0/0 = NaN, like the reference's float division
    z = torch.zeros(1, 400, device=cuda)
    assert torch.isnan(mask_iou(z, z)).all()
    a = np.sort(synth.uniform(32, 1, (9, 2, 2)), axis=1).transpose(0, 2, 1).reshape(9, 4).astype(np.float32)[:, [0, 2, 1, 3]]
    b = np.sort(synth.uniform(32, 2, (6, 2, 2)), axis=1).transpose(0, 2, 1).reshape(6, 4).astype(np.float32)[:, [0, 2, 1, 3]]
    got = box_iou(torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda)).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), pp.box_iou(a[None], b[None])[0].view(np.uint32))
    gb = box_iou(torch.from_numpy(a).to(cuda)[None], torch.from_numpy(b).to(cuda)[None])
    assert gb.shape == (1, 9, 6) and np.array_equal(gb[0].cpu().numpy(), got)


def test_rle_matches_pycocotools_semantics(cuda):
    from yolact_minimal_b200.utils import mask_utils as mu
    u8, bits = _after_nms_case(cuda, h=97, w=64)
    rles = mu.encode_rle(bits, 97, 64)
    counts = mu.rle_counts(bits, 97, 64)
    host = u8.cpu().numpy()
    for i in range(host.shape[0]):
        ref = pp.rle_counts(host[i])
        assert np.array_equal(counts[i], ref), i
        assert int(counts[i].sum()) == 97 * 64
        assert rles[i]['size'] == [97, 64] and mu.rle_string_to_counts(rles[i]['counts']) == [int(c) for c in ref]
        assert np.array_equal(pp.rle_decode(counts[i], 97, 64), host[i])
    # edge cases: empty, full, first pixel set, a single row / column, a run buffer that is too small at first
    h, w = 7, 37
    cases = np.zeros((6, h, w), np.uint8)
    cases[1] = 1
    cases[2, 0, 0] = 1
    cases[3, :, 5] = 1
    cases[4, 3, :] = 1
    cases[5] = (synth.uniform(8, 1, (h, w)) > 0.5)
    b = mu.pack_masks(torch.from_numpy(cases).to(cuda))
    for max_runs in (4096, 2):
        got = mu.rle_counts(b, h, w, max_runs=max_runs)
        for i in range(len(cases)):
            assert np.array_equal(got[i], pp.rle_counts(cases[i])), (i, max_runs)
    assert list(got[0]) == [h * w] and list(got[1]) == [0, h * w] and list(got[2])[:2] == [0, 1]


def test_numpy_twins_vs_reference_goldens(cuda):
    from yolact_minimal_b200.config import make_config
    from yolact_minimal_b200.utils.output_utils import nms_numpy, after_nms_numpy
    g = load_golden('numpy_twins.npz')
    cfg = make_config('res101_coco', 544)
    for name, S, regime, seed, h, w in (('stress_S128', 128, 'stress', 1, 80, 120), ('realistic_S256', 256, 'realistic', 2, 97, 64),
                                        ('wild_S128', 128, 'stress', 4, 60, 60)):
        anchors = pp.make_anchors(S)
        cls, box, coef = synth.head_outputs(seed, anchors.shape[0], 81, regime)
        if name.startswith('wild'):
            box = (box * 3).astype(np.float32)
        proto = synth.proto(seed, S // 4)
        ids, scores, boxes, coefs, pr = nms_numpy(cls[None], box[None], coef[None], proto[None], anchors.reshape(-1).tolist(), cfg)
        assert isinstance(ids, np.ndarray) and np.array_equal(ids, g[name + '/class']) and np.array_equal(scores, g[name + '/score'])
        assert np.allclose(boxes, g[name + '/box'], rtol=0, atol=1e-5) and np.array_equal(coefs, g[name + '/coef'])
        if name.startswith('wild'):
            assert boxes.min() < 0 or boxes.max() > 1                 # no clip: boxes leave [0,1] exactly as the reference's twin lets them
        o = pp.nms_numpy(cls, box, anchors)
        assert np.array_equal(boxes.view(np.uint32), o[2].view(np.uint32))
        rid, rsc, rbx, rmask = after_nms_numpy(ids, scores, boxes.copy(), coefs, proto, h, w, cfg)
        assert rmask.dtype == bool and rmask.shape == tuple(g[name + '/mask_shape'])
        assert np.array_equal(rbx, g[name + '/boxes_px'])
        ref = np.unpackbits(g[name + '/mask_bits'])[:rmask.size].reshape(rmask.shape).astype(bool)
        assert float((rmask != ref).mean()) < 2e-3                    # cv2.INTER_LINEAR vs the half-pixel bilinear kernel: contour pixels only
