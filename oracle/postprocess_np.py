"""numpy restatement of the reference post-process chain (TEST INFRASTRUCTURE -- see
oracle/__init__.py; never imported by the product package).

Follows, op by op and in float32:
  make_anchors        utils/box_utils.py:86-101  (+ modules/yolact.py:111-114)
  box_iou             utils/box_utils.py:8-37
  sanitize/crop       utils/box_utils.py:117-132, :147-168
  nms                 utils/output_utils.py:126-163
  fast_nms            utils/output_utils.py:11-43
  traditional_nms     utils/output_utils.py:84-123
  hard_nms (cnms)     cython_nms.pyx:24-74
  after_nms           utils/output_utils.py:200-233
Differences from the reference that are deliberate and documented in DESIGN.md:
  * exp() in the box decode is the correctly-rounded float32 exponential (computed in
    float64 and rounded once) instead of torch-CPU's SLEEF expf (<=1 ulp): decoded boxes may
    differ from the reference by 1 ulp; selected indices are unaffected on every golden.
  * sort ties are broken by ascending candidate index (what torch's CPU stable sort does,
    SURVEY.md App. B.4); the reference leaves this implementation-defined.
  * the functions return the selected (class, anchor) indices as well, which the reference
    only exposes through the coef-carries-index trick.
"""
from itertools import product
from math import sqrt, ceil

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------- anchors
def make_anchors_level(img_size, aspect_ratios, conv_h, conv_w, scale):
    """utils/box_utils.py:86-101 -- float64 python arithmetic, list of [cx, cy, w, h]."""
    out = []
    for j, i in product(range(conv_h), range(conv_w)):
        x = (i + 0.5) / conv_w
        y = (j + 0.5) / conv_h
        for ar in aspect_ratios:
            ar = sqrt(ar)
            out += [x, y, scale * ar / img_size, scale / ar / img_size]
    return out


def default_scales(img_size):
    """config.py:80"""
    return [int(img_size / 544 * aa) for aa in (24, 48, 96, 192, 384)]


def make_anchors(img_size, scales=None, aspect_ratios=(1, 1 / 2, 2)):
    """modules/yolact.py:111-114 -> float32 [A,4] (rounded once from float64 as
    utils/output_utils.py:133 does)."""
    scales = default_scales(img_size) if scales is None else scales
    flat = []
    for lvl, stride in enumerate((8, 16, 32, 64, 128)):
        size = ceil(img_size / stride)
        flat += make_anchors_level(img_size, aspect_ratios, size, size, scales[lvl])
    return np.asarray(flat, dtype=np.float64).astype(F32).reshape(-1, 4)


# ----------------------------------------------------------------------------- decode / IoU
def exp_f32(x):
    """Correctly-rounded float32 exp (see module docstring)."""
    return np.exp(x.astype(np.float64)).astype(F32)


def decode(box, anchor):
    """utils/output_utils.py:148-153 on float32 [n,4] arrays (separately rounded ops)."""
    box = box.astype(F32, copy=False)
    anchor = anchor.astype(F32, copy=False)
    xy = anchor[:, :2] + box[:, :2] * F32(0.1) * anchor[:, 2:]
    wh = anchor[:, 2:] * exp_f32(box[:, 2:] * F32(0.2))
    out = np.concatenate((xy, wh), axis=1).astype(F32)
    out[:, :2] -= out[:, 2:] / F32(2)
    out[:, 2:] += out[:, :2]
    # torch.clip propagates NaN; np.clip does too.
    return np.clip(out, F32(0.0), F32(1.0)).astype(F32)


def box_iou(box_a, box_b):
    """utils/box_utils.py:8-37 for batched [n,A,4] x [n,B,4] float32 (NaN for 0/0)."""
    a = box_a[:, :, None, :]
    b = box_b[:, None, :, :]
    max_xy = np.minimum(a[..., 2:], b[..., 2:])
    min_xy = np.maximum(a[..., :2], b[..., :2])
    inter = np.maximum(max_xy - min_xy, F32(0))          # clamp(min=0), NaN-propagating
    inter_area = inter[..., 0] * inter[..., 1]
    area_a = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    area_b = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    with np.errstate(invalid='ignore', divide='ignore'):
        return (inter_area / (area_a + area_b - inter_area)).astype(F32)


def _argsort_desc_stable(x, axis=-1):
    return np.argsort(-x, axis=axis, kind='stable')


# ----------------------------------------------------------------------------- Fast NMS
def fast_nms(box_thre, class_thre, cand_anchor, top_k=200, iou_thre=0.5, max_det=100):
    """utils/output_utils.py:11-43.
    box_thre [n,4] decoded boxes, class_thre [C-1,n] scores, cand_anchor [n] original anchor
    index of every candidate.  Returns (class_ids int64 [d], scores f32 [d], boxes f32 [d,4],
    anchor_idx int64 [d]) in the reference's output order."""
    num_classes, n = class_thre.shape
    idx = _argsort_desc_stable(class_thre, axis=1)[:, :top_k]              # :12-14
    scores = np.take_along_axis(class_thre, idx, axis=1)                   # :15
    k = idx.shape[1]
    boxes = box_thre[idx.reshape(-1)].reshape(num_classes, k, 4)           # :18
    iou = box_iou(boxes, boxes)                                            # :21
    iou = np.triu(iou, k=1)                                                # :22 (keeps NaN above diag)
    iou_max = iou.max(axis=1)                                              # :23 (NaN-propagating)
    with np.errstate(invalid='ignore'):
        keep = iou_max <= F32(iou_thre)                                    # :26
    class_ids = np.broadcast_to(np.arange(num_classes)[:, None], keep.shape)
    class_ids, box_nms, score_nms = class_ids[keep], boxes[keep], scores[keep]   # :31 class-major
    anchor_nms = cand_anchor[idx][keep]
    order = _argsort_desc_stable(score_nms, axis=0)[:max_det]              # :34-37
    return (class_ids[order].astype(np.int64), score_nms[order], box_nms[order],
            anchor_nms[order].astype(np.int64))


# ----------------------------------------------------------------------------- hard NMS
def hard_nms(dets, thresh):
    """cython_nms.pyx:24-74 -- greedy NMS, '+1' areas, suppress on ovr >= thresh, returns the
    surviving indices in ascending original order.  dets float32 [n,5] (x1,y1,x2,y2,score).
    `scores.argsort()[::-1]` (numpy default quicksort, reversed) is restated as: descending
    score, ties by DESCENDING index (what reversing a stable ascending sort yields)."""
    dets = np.asarray(dets, dtype=F32)
    n = dets.shape[0]
    x1, y1, x2, y2, sc = (dets[:, i] for i in range(5))
    areas = (x2 - x1 + F32(1)) * (y2 - y1 + F32(1))
    order = np.argsort(sc, kind='stable')[::-1]
    suppressed = np.zeros(n, dtype=bool)
    thresh = F32(thresh)
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(F32(0), xx2 - xx1 + F32(1))
        h = np.maximum(F32(0), yy2 - yy1 + F32(1))
        inter = w * h
        with np.errstate(invalid='ignore', divide='ignore'):
            ovr = inter / (areas[i] + areas[rest] - inter)
            hit = ovr >= thresh
        suppressed[rest[hit & ~suppressed[rest]]] = True
    return np.where(~suppressed)[0].astype(np.int64)


def traditional_nms(box_thre, class_thre, cand_anchor, img_size, score_thre=0.05, iou_thre=0.5,
                    max_det=100):
    """utils/output_utils.py:84-123."""
    boxes = (box_thre * F32(img_size)).astype(F32)                          # :90
    idx_l, cls_l, scr_l = [], [], []
    for c in range(class_thre.shape[0]):
        s = class_thre[c]
        m = s > F32(score_thre)                                             # :94
        if not m.any():
            continue
        ids = np.nonzero(m)[0]
        keep = hard_nms(np.concatenate([boxes[m], s[m][:, None]], axis=1), iou_thre)
        idx_l.append(ids[keep]); cls_l.append(np.full(len(keep), c)); scr_l.append(s[m][keep])
    if not idx_l:
        # the reference would raise on torch.cat([]) here; report "no detections"
        return (np.zeros(0, np.int64), np.zeros(0, F32), np.zeros((0, 4), F32), np.zeros(0, np.int64))
    idx = np.concatenate(idx_l); cls = np.concatenate(cls_l); scr = np.concatenate(scr_l)
    order = _argsort_desc_stable(scr, axis=0)[:max_det]                      # :115-117
    idx = idx[order]
    return (cls[order].astype(np.int64), scr[order], (boxes[idx] / F32(img_size)).astype(F32),
            cand_anchor[idx].astype(np.int64))


# ----------------------------------------------------------------------------- nms()
def nms(class_pred, box_pred, anchors, score_thre=0.05, iou_thre=0.5, top_k=200, max_det=100,
        traditional=False, img_size=None):
    """utils/output_utils.py:126-163 for ONE image: class_pred [A,C] (post-softmax),
    box_pred [A,4], anchors [A,4] float32.  Returns None when nothing passes the score filter
    (:155), else (class_ids, scores, boxes, anchor_idx)."""
    class_p = np.ascontiguousarray(class_pred.astype(F32, copy=False).T)[1:, :]   # :135-138
    class_p_max = class_p.max(axis=0)                                              # :140
    with np.errstate(invalid='ignore'):
        keep = class_p_max > F32(score_thre)                                       # :143
    if not keep.any():
        return None
    cand_anchor = np.nonzero(keep)[0]
    class_thre = class_p[:, keep]
    box_thre = decode(box_pred[keep], anchors[keep])
    if traditional:
        return traditional_nms(box_thre, class_thre, cand_anchor, img_size, score_thre, iou_thre, max_det)
    return fast_nms(box_thre, class_thre, cand_anchor, top_k, iou_thre, max_det)


# ----------------------------------------------------------------------------- after_nms()
def sanitize_coordinates(x1, x2, size, padding):
    """utils/box_utils.py:117-132 (float32, no rounding of the results)."""
    x1 = x1 * F32(size)
    x2 = x2 * F32(size)
    lo = np.minimum(x1, x2)
    hi = np.maximum(x1, x2)
    lo = np.maximum(lo - F32(padding), F32(0))
    hi = np.minimum(hi + F32(padding), F32(size))
    return lo.astype(F32), hi.astype(F32)


def crop(masks, boxes, padding=1):
    """utils/box_utils.py:147-168: masks [h,w,n], boxes [n,4]."""
    h, w, n = masks.shape
    x1, x2 = sanitize_coordinates(boxes[:, 0], boxes[:, 2], w, padding)
    y1, y2 = sanitize_coordinates(boxes[:, 1], boxes[:, 3], h, padding)
    cols = np.arange(w, dtype=F32)[None, :, None]
    rows = np.arange(h, dtype=F32)[:, None, None]
    m = (cols >= x1[None, None, :]) & (cols < x2[None, None, :]) & \
        (rows >= y1[None, None, :]) & (rows < y2[None, None, :])
    return masks * m.astype(F32)


def _bilinear_axis(in_size, out_size):
    """torch upsample_bilinear2d, align_corners=False, float32 index arithmetic
    (ATen UpSample.h area_pixel_compute_source_index)."""
    scale = F32(in_size) / F32(out_size)
    dst = np.arange(out_size, dtype=F32)
    src = scale * (dst + F32(0.5)) - F32(0.5)
    src = np.maximum(src, F32(0))
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    i1 = np.minimum(i0 + 1, in_size - 1)
    l1 = (src - i0.astype(F32)).astype(F32)
    l0 = (F32(1) - l1).astype(F32)
    return i0, i1, l0, l1


def bilinear_resize(masks, out_h, out_w):
    """masks [n,h,w] float32 -> [n,out_h,out_w], F.interpolate(mode='bilinear',
    align_corners=False)."""
    n, h, w = masks.shape
    y0, y1, ly0, ly1 = _bilinear_axis(h, out_h)
    x0, x1, lx0, lx1 = _bilinear_axis(w, out_w)
    top = masks[:, y0][:, :, x0] * lx0[None, None, :] + masks[:, y0][:, :, x1] * lx1[None, None, :]
    bot = masks[:, y1][:, :, x0] * lx0[None, None, :] + masks[:, y1][:, :, x1] * lx1[None, None, :]
    return (top * ly0[None, :, None] + bot * ly1[None, :, None]).astype(F32)


def sigmoid(x):
    return (F32(1) / (F32(1) + np.exp(-x.astype(np.float64)).astype(F32))).astype(F32)


def mask_logits(proto, coef):
    """proto [P,P,K] @ coef[d,K]^T with float64 accumulation (reference: torch.matmul fp32;
    summation order is library-defined, so the oracle uses the exact sum rounded once)."""
    return (proto.astype(np.float64) @ coef.astype(np.float64).T).astype(F32)


def after_nms(class_ids, scores, boxes, coefs, proto, img_h, img_w, visual_thre=0.0, no_crop=False,
              return_soft=False):
    """utils/output_utils.py:200-233 for one image.  Returns (ids, scores, boxes_int32 [d,4],
    masks float32 {0,1} [d,h,w]) or None."""
    if class_ids is None:
        return None
    if visual_thre > 0:                                                     # :204-212
        keep = scores >= F32(visual_thre)
        if not keep.any():
            return None
        class_ids, scores, boxes, coefs = class_ids[keep], scores[keep], boxes[keep], coefs[keep]
    masks = sigmoid(mask_logits(proto, coefs))                              # :217
    if not no_crop:
        masks = crop(masks, boxes)                                          # :219-220
    masks = np.ascontiguousarray(masks.transpose(2, 0, 1))                  # :222
    ori = max(img_h, img_w)
    soft = bilinear_resize(masks, ori, ori)                                 # :226
    hard = (soft > F32(0.5)).astype(F32)                                    # :227
    hard = hard[:, :img_h, :] if img_h < img_w else hard[:, :, :img_w]      # :228
    soft = soft[:, :img_h, :] if img_h < img_w else soft[:, :, :img_w]
    boxes_px = (boxes * F32(ori)).astype(F32).astype(np.int32)              # :230-231 (trunc)
    if return_soft:
        return class_ids, scores, boxes_px, hard, soft
    return class_ids, scores, boxes_px, hard


# ----------------------------------------------------------------------------- val_aug (pre-process)
NORM_MEAN = np.array([103.94, 116.78, 123.68], dtype=F32)     # BGR, config.py:66
NORM_STD = np.array([57.38, 57.12, 58.40], dtype=F32)         # config.py:67


def _cv_linear_axis(src_size, dst_size):
    """OpenCV resize (INTER_LINEAR, float path) coordinate tables: fx = (dx+0.5)*scale-0.5 in double
    rounded to float, sx = floor(fx); clamped at both borders with weight 0 on the outside tap."""
    scale = float(src_size) / float(dst_size)
    d = np.arange(dst_size, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(F32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(F32)).astype(F32)
    lo = s < 0
    s[lo] = 0; f[lo] = 0
    hi = s >= src_size - 1
    s[hi] = src_size - 1; f[hi] = 0
    s1 = np.minimum(s + 1, src_size - 1)
    return s, s1, (F32(1) - f).astype(F32), f


def val_aug(img_bgr_u8, val_size):
    """utils/augmentations.py:219-227: float32 -> pad to square (top-left, mean fill, :138-165) ->
    cv2.resize bilinear (:168-189) -> (x - mean) / std, BGR->RGB, HWC->CHW (:212-216)."""
    img = img_bgr_u8.astype(F32)
    h, w = img.shape[:2]
    P = max(h, w)
    if h != w:
        pad = np.empty((P, P, 3), dtype=F32)
        pad[:] = NORM_MEAN
        pad[:h, :w] = img
        img = pad
    x0, x1, ax0, ax1 = _cv_linear_axis(P, val_size)
    y0, y1, ay0, ay1 = _cv_linear_axis(P, val_size)
    rows0 = img[y0][:, x0] * ax0[None, :, None] + img[y0][:, x1] * ax1[None, :, None]
    rows1 = img[y1][:, x0] * ax0[None, :, None] + img[y1][:, x1] * ax1[None, :, None]
    out = (rows0 * ay0[:, None, None] + rows1 * ay1[:, None, None]).astype(F32)
    out = (out - NORM_MEAN) / NORM_STD
    return np.ascontiguousarray(out[:, :, ::-1].transpose(2, 0, 1)).astype(F32)


# ----------------------------------------------------------------------------- numpy twins / mask output stage
def nms_numpy(class_pred, box_pred, anchors, score_thre=0.05, iou_thre=0.5, top_k=200, max_det=100):
    """utils/output_utils.py:166-197 for ONE image: nms() WITHOUT the clip of the decoded boxes to [0,1] (compare :186-190 with
    :153).  fp32 restatement (the reference's twin decodes in float64 because its anchors are a float64 array: agreement ~1e-7)."""
    class_p = np.ascontiguousarray(class_pred.astype(F32, copy=False).T)[1:, :]
    keep = class_p.max(axis=0) > F32(score_thre)
    if not keep.any():
        return None
    cand_anchor = np.nonzero(keep)[0]
    b, a = box_pred[keep].astype(F32), anchors[keep].astype(F32)
    xy = a[:, :2] + b[:, :2] * F32(0.1) * a[:, 2:]
    wh = a[:, 2:] * exp_f32(b[:, 2:] * F32(0.2))
    out = np.concatenate((xy, wh), axis=1).astype(F32)
    out[:, :2] -= out[:, 2:] / F32(2)
    out[:, 2:] += out[:, :2]
    return fast_nms(out, class_p[:, keep], cand_anchor, top_k, iou_thre, max_det)


def rle_counts(mask):
    """Run lengths of pycocotools.mask.encode(np.asfortranarray(mask)) for ONE {0,1} mask [h,w]: column-major scan, alternating
    runs starting with a (possibly empty) run of zeros (pycocotools/common/maskApi.c rleEncode)."""
    v = np.asarray(mask).astype(np.uint8).reshape(-1, order='F')
    counts, prev, run = [], 0, 0
    for x in v:
        if x != prev:
            counts.append(run); run = 0; prev = x
        run += 1
    counts.append(run)
    return np.asarray(counts, dtype=np.uint32)


def rle_decode(counts, h, w):
    v = np.zeros(h * w, np.uint8)
    p, val = 0, 0
    for c in counts:
        v[p:p + int(c)] = val
        p += int(c); val ^= 1
    return v.reshape(h, w, order='F')
