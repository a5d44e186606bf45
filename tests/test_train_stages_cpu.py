"""CPU: the training branch stage by stage.  tests/golden/train_stages.npz holds what the reference's match(), OHEM mining,
four loss functions and mask_iou return on synthetic head outputs (tests/golden/make_golden.py stages); checked here are
(1) the numpy oracle oracle/train_np.py -- the checker native training kernels will be held to -- and
(2) the torch-autograd checker oracle/train_torch.py, function by function (the native kernels are held to both on the GPU:
tests/test_train_gpu.py)."""
import numpy as np
import torch

from conftest import load_golden
from oracle import synth, train_np as tn

S, B, NCLS = 128, 3, 81


def inputs(g):
    A, P = g['anchors'].shape[0], S // 4
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    tg, mk = synth.train_targets(9, B, S, n=4)
    class_p = f32(synth.normal(21, 1, (B, A, NCLS)) * 2)
    box_p = f32(synth.normal(21, 2, (B, A, 4)) * 0.5)
    coef_p = np.tanh(f32(synth.normal(21, 3, (B, A, 32))))
    proto_p = np.maximum(f32(synth.normal(21, 4, (B, P, P, 32))), 0)
    seg_p = f32(synth.normal(21, 5, (B, NCLS - 1, S // 8, S // 8)))
    return tg, mk, class_p, box_p, coef_p, proto_p, seg_p


def test_oracle_matches_reference_stage_by_stage():
    g = load_golden('train_stages.npz')
    tg, mk, class_p, box_p, coef_p, proto_p, seg_p = inputs(g)
    res = [tn.match(tg[i][:, :4], g['anchors'], tg[i][:, 4]) for i in range(B)]
    offsets, labels, matched, idx = (np.stack(x) for x in zip(*res))
    assert np.array_equal(labels, g['labels']) and np.array_equal(idx, g['matched_idx'])
    assert np.array_equal(matched, g['matched'])
    assert np.allclose(offsets, g['offsets'], rtol=1e-5, atol=1e-6, equal_nan=True)
    assert np.array_equal(tn.ohem_negatives(class_p, labels), g['ohem_neg'])
    assert np.isclose(tn.category_loss(class_p, labels), g['loss_c'], rtol=2e-6)
    assert np.isclose(tn.box_loss(box_p, offsets, labels), g['loss_b'], rtol=2e-6)
    assert np.isclose(tn.mask_loss(labels, idx, coef_p, proto_p, mk, matched), g['loss_m'], rtol=2e-5)
    assert np.isclose(tn.semantic_loss(seg_p, mk, [t[:, 4] for t in tg]), g['loss_s'], rtol=2e-6)
    m1 = (synth.uniform(31, 1, (5, 400)) > 0.5).astype(np.float32)
    m2 = (synth.uniform(31, 2, (7, 400)) > 0.6).astype(np.float32)
    assert np.allclose(tn.mask_iou(m1, m2), g['mask_iou'], rtol=1e-6)


def test_torch_checker_training_functions_match_reference():
    from oracle import train_torch as tt
    from yolact_minimal_b200.config import make_config
    g = load_golden('train_stages.npz')
    tg, mk, class_p, box_p, coef_p, proto_p, seg_p = inputs(g)
    cfg = make_config('res50_coco', S, mode='train', train_bs=B)
    anchors = torch.from_numpy(g['anchors'])
    out = [tt.assign_targets(cfg, torch.from_numpy(tg[i][:, :4]), anchors, torch.from_numpy(tg[i][:, 4]).long()) for i in range(B)]
    offsets, labels, matched, idx = (torch.stack(x) for x in zip(*out))
    assert np.array_equal(labels.numpy(), g['labels']) and np.array_equal(idx.numpy(), g['matched_idx'])
    pos = labels > 0
    t = torch.from_numpy
    assert np.isclose(float(tt.category_loss(cfg, t(class_p), labels, pos)), g['loss_c'], rtol=1e-5)
    assert np.isclose(float(tt.box_loss(cfg, t(box_p), offsets, pos)), g['loss_b'], rtol=1e-5)
    assert np.isclose(float(tt.mask_loss(cfg, pos, idx, t(coef_p), t(proto_p), [t(m) for m in mk], matched)), g['loss_m'], rtol=1e-5)
    assert np.isclose(float(tt.semantic_loss(cfg, t(seg_p), [t(m) for m in mk], [t(x[:, 4]).long() for x in tg])), g['loss_s'], rtol=1e-5)
