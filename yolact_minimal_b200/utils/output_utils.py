"""Post-process API kept from the reference's utils/output_utils.py -- `nms` (:126-163) and
`after_nms` (:200-233) -- with the work done by the CUDA kernels behind yb_detect /
yb_mask_assemble (include/yolact_b200.h).  Same arguments, same return tuples, same
"(None, ...)" convention for "no detections".  `detect_batched` is the batched form the
reference lacks (its nms() squeezes a batch of 1, output_utils.py:127-130).

All inputs must be CUDA float32 tensors; nothing here falls back to the CPU.
"""
import ctypes

import torch

from .. import _lib
from .box_utils import anchors_tensor

_ws_cache = {}
_anchor_cache = {}


def _params(cfg, num_classes, coef_dim, no_clip=False):
    return _lib.DetectParams(float(cfg.nms_score_thre), float(cfg.nms_iou_thre), int(cfg.top_k),
                             int(cfg.max_detections), int(num_classes), int(coef_dim),
                             1 if getattr(cfg, 'traditional_nms', False) else 0, float(getattr(cfg, 'img_size', 0)), 1 if no_clip else 0)


def _workspace(nbytes, device):
    key = (device.index, 'detect')
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _device_anchors(anchors, device):
    if isinstance(anchors, torch.Tensor) and anchors.is_cuda and anchors.dtype == torch.float32:
        return anchors.reshape(-1, 4).contiguous()
    key = (id(anchors), len(anchors), device.index)
    hit = _anchor_cache.get(key)
    if hit is None:
        hit = anchors_tensor(anchors, device)
        _anchor_cache.clear()
        _anchor_cache[key] = hit
    return hit


def _require_cuda(*tensors):
    for t in tensors:
        if not (isinstance(t, torch.Tensor) and t.is_cuda):
            raise _lib.YolactB200Error('yolact_minimal_b200 post-process needs CUDA tensors (no CPU fallback)')


def record_fields(B, D, K):
    """(name, int32 offset, element count, shape, dtype) of every field of the flat detection record.  Field offsets are
    16-byte aligned: the post-process kernels store boxes / coefficients with 128-bit vector writes."""
    fields, o = [], 0
    for name, n, shape, dt in (('count', B, (B,), torch.int32), ('cls', B * D, (B, D), torch.int32), ('anchor', B * D, (B, D), torch.int32),
                               ('score', B * D, (B, D), torch.float32), ('box', B * D * 4, (B, D, 4), torch.float32),
                               ('coef', B * D * K, (B, D, K), torch.float32)):
        fields.append((name, o, n, shape, dt))
        o += (n + 3) // 4 * 4
    return fields, o


def record_numel(B, D, K):
    return record_fields(B, D, K)[1]


def record_views(flat, B, D, K):
    """Typed views of a flat int32 detection record -> the detect_batched dict; '_flat' is the buffer itself."""
    out = {'_flat': flat}
    for name, o, n, shape, dt in record_fields(B, D, K)[0]:
        out[name] = flat[o:o + n].view(dt).view(shape)
    return out


def detect_batched(class_pred, box_pred, coef_pred, anchors, cfg, no_clip=False):
    """Batched decode + (Fast|traditional) NMS + top-k.
    class_pred [B,A,C] (post-softmax), box_pred [B,A,4], coef_pred [B,A,K].
    Returns dict of padded tensors: count [B] int32, class [B,D] int32, anchor [B,D] int32,
    score [B,D], box [B,D,4], coef [B,D,K]  (D = cfg.max_detections)."""
    _require_cuda(class_pred, box_pred, coef_pred)
    dev = class_pred.device
    cls = class_pred.detach().to(torch.float32).contiguous()
    box = box_pred.detach().to(torch.float32).contiguous()
    coef = coef_pred.detach().to(torch.float32).contiguous()
    if cls.dim() == 2:
        cls, box, coef = cls[None], box[None], coef[None]
    B, A, C = cls.shape
    K = coef.shape[-1]
    anc = _device_anchors(anchors, dev)
    if anc.shape[0] != A:
        raise ValueError(f'anchors has {anc.shape[0]} rows, predictions have {A}')
    p = _params(cfg, C, K, no_clip)
    D = p.max_det
    L = _lib.lib()
    # ONE flat int32 buffer [count B | class B*D | anchor B*D | score B*D | box B*D*4 | coef B*D*K]; the dict entries are typed views
    # of it, so the kernels write the record that dist.gather_detections all-gathers directly (no pack pass)
    out = record_views(torch.empty(record_numel(B, D, K), dtype=torch.int32, device=dev), B, D, K)
    with torch.cuda.device(dev):
        nbytes = L.yb_detect_workspace_bytes(B, A, ctypes.byref(p))
        ws = _workspace(nbytes, dev)
        _lib.check(L.yb_detect(cls.data_ptr(), box.data_ptr(), coef.data_ptr(), anc.data_ptr(), B, A, ctypes.byref(p),
                               ws.data_ptr(), ws.numel(), out['count'].data_ptr(), out['cls'].data_ptr(),
                               out['anchor'].data_ptr(), out['score'].data_ptr(), out['box'].data_ptr(),
                               out['coef'].data_ptr(), torch.cuda.current_stream().cuda_stream), 'yb_detect')
    return out


def nms(class_pred, box_pred, coef_pred, proto_out, anchors, cfg):
    """Reference signature (utils/output_utils.py:126).  Batch of one image.
    Returns (class_ids int64 [d], scores [d], boxes [d,4], coefs [d,K], proto [P,P,K]) or
    (None,)*5 when nothing passes the score filter (:155)."""
    cls = class_pred.squeeze(0) if class_pred.dim() == 3 else class_pred
    box = box_pred.squeeze(0) if box_pred.dim() == 3 else box_pred
    coef = coef_pred.squeeze(0) if coef_pred.dim() == 3 else coef_pred
    proto = proto_out.squeeze(0) if proto_out.dim() == 4 else proto_out
    r = detect_batched(cls[None], box[None], coef[None], anchors, cfg)
    d = int(r['count'][0].item())          # the reference syncs here too (boolean-mask compaction)
    if d == 0:
        return None, None, None, None, None
    return (r['cls'][0, :d].long(), r['score'][0, :d].clone(), r['box'][0, :d].clone(), r['coef'][0, :d].clone(), proto)


def after_nms(ids_p, class_p, box_p, coef_p, proto_p, img_h, img_w, cfg=None, img_name=None, mask_dtype=torch.float32):
    """Reference signature (utils/output_utils.py:200).  Returns (ids, scores, boxes int32 [d,4]
    in pixels of max(img_h,img_w), masks [d,img_h,img_w] with values {0,1}) or (None,)*4.
    `mask_dtype=torch.uint8` writes byte masks instead of the reference's float32 (4x less HBM
    traffic), `mask_dtype='bits'` bit-packed masks [d,img_h,ceil(img_w/32)] (32x less; utils/mask_utils.py has
    the IoU / RLE stage that consumes them); unlike the reference, box_p is not modified in place."""
    if ids_p is None:
        return None, None, None, None
    _require_cuda(box_p, coef_p, proto_p)
    if cfg is not None and getattr(cfg, 'visual_thre', 0) > 0:              # :204-212
        keep = class_p >= cfg.visual_thre
        if not bool(keep.any()):
            return None, None, None, None
        ids_p, class_p, box_p, coef_p = ids_p[keep], class_p[keep], box_p[keep], coef_p[keep]
    if cfg is not None and getattr(cfg, 'save_lincomb', False):
        raise NotImplementedError('save_lincomb is a visualisation option (draw_lincomb); out of scope')
    dev = proto_p.device
    proto = proto_p.detach().to(torch.float32).contiguous()
    coef = coef_p.detach().to(torch.float32).contiguous()
    box = box_p.detach().to(torch.float32).contiguous()
    d, K = coef.shape
    P = proto.shape[0]
    if proto.shape[1] != P or proto.shape[2] != K:
        raise ValueError(f'proto {tuple(proto.shape)} does not match coef {tuple(coef.shape)}')
    crop = 0 if (cfg is not None and getattr(cfg, 'no_crop', False)) else 1
    f32 = mask_dtype == torch.float32
    fmt = 2 if isinstance(mask_dtype, str) and mask_dtype == 'bits' else (1 if f32 else 0)
    if fmt == 2:
        masks = torch.empty(d, int(img_h), (int(img_w) + 31) // 32, dtype=torch.int32, device=dev)
    else:
        masks = torch.empty(d, int(img_h), int(img_w), dtype=torch.float32 if f32 else torch.uint8, device=dev)
    boxes_px = torch.empty(d, 4, dtype=torch.int32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        ws = torch.empty(int(L.yb_mask_workspace_bytes(d, P)), dtype=torch.uint8, device=dev)
        _lib.check(L.yb_mask_assemble(proto.data_ptr(), coef.data_ptr(), box.data_ptr(), d, P, K, int(img_h), int(img_w),
                                      crop, fmt, ws.data_ptr(), ws.numel(), masks.data_ptr(),
                                      boxes_px.data_ptr(), torch.cuda.current_stream().cuda_stream), 'yb_mask_assemble')
    return ids_p, class_p, boxes_px, masks


# ---------------------------------------------------------------------------------------------------------------------------
# The ONNX / TensorRT callers' numpy twins (utils/output_utils.py:46-81,:166-197,:236-273; detect_with_onnx.py, detect_with_trt.py):
# numpy arrays in, numpy arrays out, same return conventions -- computed by the same CUDA kernels.  Semantics kept from the
# reference: NO clip of the decoded boxes to [0,1] (compare :186-190 with :153) and boolean masks.  Two documented differences:
# the reference decodes in float64 (its anchors are a float64 array) and resizes the masks with cv2.INTER_LINEAR; here the decode
# is the fp32 kernel and the resize the same half-pixel bilinear kernel as after_nms (boxes agree to ~1e-6, masks to a few pixels
# on the contour).
# ---------------------------------------------------------------------------------------------------------------------------
def _np_dev():
    if not torch.cuda.is_available():
        raise _lib.YolactB200Error('the numpy post-process twins run on a CUDA device (no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


def nms_numpy(class_pred, box_pred, coef_pred, proto_out, anchors, cfg):
    """utils/output_utils.py:166-197.  Returns (class_ids, class_thre, box_thre, coef_thre, proto_p) as numpy arrays or (None,)*5."""
    import numpy as np
    assert not getattr(cfg, 'traditional_nms', False), 'Traditional nms is not supported with numpy.'          # :193
    dev = _np_dev()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).to(dev)
    cls, box, coef = t(np.squeeze(class_pred)), t(np.squeeze(box_pred)), t(np.squeeze(coef_pred))
    proto = np.squeeze(np.asarray(proto_out))
    anc = t(np.asarray(anchors, dtype=np.float64).reshape(-1, 4))
    r = detect_batched(cls[None], box[None], coef[None], anc, cfg, no_clip=True)
    d = int(r['count'][0].item())
    if d == 0:
        return None, None, None, None, None
    return (r['cls'][0, :d].cpu().numpy().astype(np.int64), r['score'][0, :d].cpu().numpy(), r['box'][0, :d].cpu().numpy(),
            r['coef'][0, :d].cpu().numpy(), proto)


def after_nms_numpy(ids_p, class_p, box_p, coef_p, proto_p, img_h, img_w, cfg=None):
    """utils/output_utils.py:236-273.  Returns (ids, scores, boxes int32 [d,4], masks bool [d,img_h,img_w]) or (None,)*4."""
    import numpy as np
    if ids_p is None:
        return None, None, None, None
    if cfg and getattr(cfg, 'visual_thre', 0) > 0:
        keep = class_p >= cfg.visual_thre
        if not keep.any():
            return None, None, None, None
        ids_p, class_p, box_p, coef_p = ids_p[keep], class_p[keep], box_p[keep], coef_p[keep]
    assert not (cfg and getattr(cfg, 'save_lincomb', False)), 'save_lincomb is not supported in onnx mode.'   # :253
    dev = _np_dev()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).to(dev)
    r = after_nms(ids_p, class_p, t(box_p), t(coef_p), t(proto_p), img_h, img_w, cfg=None if cfg is None else _NoVisual(cfg), mask_dtype=torch.uint8)
    return ids_p, class_p, r[2].cpu().numpy().astype(np.int32), r[3].cpu().numpy().astype(bool)


class _NoVisual:
    """cfg view for after_nms_numpy: the score filter was already applied on the numpy side."""

    def __init__(self, cfg):
        self._cfg = cfg

    def __getattr__(self, k):
        if k == 'visual_thre':
            return 0
        return getattr(self._cfg, k)
