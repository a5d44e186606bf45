"""oracle/train_np.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

numpy fp32 restatement of the reference's training targets and losses, function by function, so that
the native kernels for SURVEY.md 8 row a12 / f3 (yolact_minimal_b200/csrc/losses.cu) can be checked stage by stage:

  match            utils/box_utils.py:57-83    per image: IoU[gt, anchor] -> best gt per anchor, each gt claims
                                               its best anchor (IoU := 2, later gt wins a shared anchor),
                                               labels  >0 fg / 0 bg / -1 neutral
  encode           utils/box_utils.py:104-114  SSD offsets, variances 0.1 / 0.2
  ohem_negatives   modules/yolact.py:205-225   hard-negative mining, 3 negatives per positive, per image
  category_loss    modules/yolact.py:227-231   cross entropy (sum) over positives + mined negatives / #pos
  box_loss         modules/yolact.py:233-238   smooth-L1 (sum) over positives / #pos
  mask_loss        modules/yolact.py:240-290   sigmoid(proto @ coef^T), crop to the gt box, BCE / box area
                                               (the > masks_to_train random subset is NOT restated: keep n small)
  semantic_loss    modules/yolact.py:292-313   per-class max of the downsampled gt masks, BCE with logits
  mask_iou         utils/box_utils.py:189-200  pairwise IoU of flattened binary masks (prep_metrics)

Pinned by tests/golden/train_stages.npz (minted from the reference by tests/golden/make_golden.py).
"""
import numpy as np

from . import postprocess_np as pp

F32 = np.float32


def encode(matched, anchors):
    matched, anchors = matched.astype(F32), anchors.astype(F32)
    cxcy = (matched[:, :2] + matched[:, 2:]) / F32(2) - anchors[:, :2]
    cxcy = cxcy / (F32(0.1) * anchors[:, 2:])
    with np.errstate(divide='ignore', invalid='ignore'):
        wh = np.log((matched[:, 2:] - matched[:, :2]) / anchors[:, 2:]) / F32(0.2)
    return np.concatenate([cxcy, wh], 1).astype(F32)


def match(box_gt, anchors, class_gt, pos_thr=0.5, neg_thr=0.4):
    """-> offsets [A,4] f32, labels [A] i64, matched gt box [A,4] f32, matched gt index [A] i64."""
    box_gt, anchors = box_gt.astype(F32), anchors.astype(F32)
    corners = np.concatenate([anchors[:, :2] - anchors[:, 2:] / F32(2), anchors[:, :2] + anchors[:, 2:] / F32(2)], 1)
    iou = pp.box_iou(box_gt[None], corners[None])[0]               # [num_gt, A], same op order as box_utils.py:8-37
    best_anchor_of_gt = iou.argmax(1)
    best_iou, best_gt = iou.max(0).copy(), iou.argmax(0).astype(np.int64)
    best_iou[best_anchor_of_gt] = F32(2)
    for j in range(len(best_anchor_of_gt)):                        # sequential: a later gt wins a shared anchor
        best_gt[best_anchor_of_gt[j]] = j
    matched = box_gt[best_gt]
    labels = class_gt.astype(np.int64)[best_gt] + 1
    labels[best_iou < F32(pos_thr)] = -1
    labels[best_iou < F32(neg_thr)] = 0
    return encode(matched, anchors), labels, matched, best_gt


def _hardness(class_p):
    flat = class_p.reshape(-1, class_p.shape[-1]).astype(F32)
    mx = flat.max()
    return (np.log(np.sum(np.exp(flat - mx), 1, dtype=F32)) + mx - flat[:, 0]).astype(F32).reshape(class_p.shape[0], -1)


def ohem_negatives(class_p, labels, ratio=3):
    """-> bool [B, A]: the negatives that enter the classification loss."""
    pos = labels > 0
    mark = _hardness(class_p)
    mark[pos] = 0
    mark[labels < 0] = 0
    rank = np.argsort(np.argsort(-mark, axis=1, kind='stable'), axis=1, kind='stable')
    num_neg = np.minimum(ratio * pos.sum(1, keepdims=True), pos.shape[1] - 1)
    neg = rank < num_neg
    neg[pos] = False
    neg[labels < 0] = False
    return neg


def _log_softmax(x):
    x = x.astype(np.float64)
    m = x.max(1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(1, keepdims=True))


def category_loss(class_p, labels, conf_alpha=1.0, ratio=3):
    pos = labels > 0
    chosen = pos | ohem_negatives(class_p, labels, ratio)
    ls = _log_softmax(class_p[chosen])
    nll = -ls[np.arange(ls.shape[0]), labels[chosen]]
    return conf_alpha * nll.sum() / pos.sum()


def box_loss(box_p, offsets, labels, bbox_alpha=1.5):
    pos = labels > 0
    d = np.abs(box_p[pos].astype(np.float64) - offsets[pos].astype(np.float64))
    return bbox_alpha * np.where(d < 1, 0.5 * d * d, d - 0.5).sum() / pos.sum()


def _downsample_binarise(masks, h, w):
    """F.interpolate(..., mode='bilinear', align_corners=False) then > 0.5 (yolact.py:246-250,:303-305)."""
    return (pp.bilinear_resize(masks.astype(F32), h, w) > F32(0.5)).astype(F32)


def mask_loss(labels, best_gt, coef_p, proto_p, masks_gt, matched, mask_alpha=6.125):
    ph, pw = proto_p.shape[1:3]
    total = 0.0
    for i in range(coef_p.shape[0]):
        gt = _downsample_binarise(masks_gt[i], ph, pw).transpose(1, 2, 0)            # [ph, pw, n_gt]
        sel = labels[i] > 0
        if not sel.any():
            continue
        idx, boxes, coef = best_gt[i][sel], matched[i][sel].astype(F32), coef_p[i][sel].astype(F32)
        pred = pp.sigmoid(proto_p[i].astype(F32) @ coef.T)
        pred = np.clip(pp.crop(pred, boxes), 0, 1).astype(np.float64)
        tgt = gt[:, :, idx].astype(np.float64)
        with np.errstate(divide='ignore'):
            bce = -(tgt * np.maximum(np.log(pred), -100) + (1 - tgt) * np.maximum(np.log(1 - pred), -100))   # torch clamps log at -100
        area = ((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])).astype(np.float64)
        total += (bce.sum((0, 1)) / area).sum()
    return mask_alpha * total / ph / pw / (labels > 0).sum()


def semantic_loss(seg_p, masks_gt, class_gt, semantic_alpha=1.0):
    B, _, mh, mw = seg_p.shape
    total = 0.0
    for i in range(B):
        gt = _downsample_binarise(masks_gt[i], mh, mw)
        target = np.zeros(seg_p[i].shape, np.float64)
        for j in range(gt.shape[0]):
            c = int(class_gt[i][j])
            target[c] = np.maximum(target[c], gt[j])
        x = seg_p[i].astype(np.float64)
        total += (np.maximum(x, 0) - x * target + np.log1p(np.exp(-np.abs(x)))).sum()
    return semantic_alpha * total / mh / mw / B


def mask_iou(m1, m2):
    m1, m2 = m1.astype(F32), m2.astype(F32)
    inter = m1 @ m2.T
    a1, a2 = m1.sum(1).reshape(1, -1), m2.sum(1).reshape(1, -1)
    with np.errstate(divide='ignore', invalid='ignore'):
        return inter / (a1.T + a2 - inter)
