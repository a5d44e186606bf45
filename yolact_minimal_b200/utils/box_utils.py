"""Host-side box helpers kept from the reference's utils/box_utils.py API.

make_anchors (box_utils.py:86-101) is pure float64 host arithmetic in the reference and stays
so here (anchors are computed once and cached on the device; the reference rebuilds the tensor
from a Python list for every image, utils/output_utils.py:132-133).
"""
from math import ceil, sqrt

import numpy as np
import torch


def make_anchors(cfg, conv_h, conv_w, scale):
    """Same contract as the reference: flat python list [cx, cy, w, h, ...] in float64, rows
    ordered (y, x, aspect_ratio) to line up with the head's NHWC output."""
    ys = (np.arange(conv_h, dtype=np.float64) + 0.5) / conv_h
    xs = (np.arange(conv_w, dtype=np.float64) + 0.5) / conv_w
    ws = [scale * sqrt(ar) / cfg.img_size for ar in cfg.aspect_ratios]
    hs = [scale / sqrt(ar) / cfg.img_size for ar in cfg.aspect_ratios]
    out = np.empty((conv_h, conv_w, len(ws), 4), dtype=np.float64)
    out[..., 0] = xs[None, :, None]
    out[..., 1] = ys[:, None, None]
    out[..., 2] = np.asarray(ws)[None, None, :]
    out[..., 3] = np.asarray(hs)[None, None, :]
    return out.reshape(-1).tolist()


def all_anchors(cfg):
    """modules/yolact.py:111-114: the five FPN levels concatenated."""
    flat = []
    for lvl, stride in enumerate((8, 16, 32, 64, 128)):
        size = ceil(cfg.img_size / stride)
        flat += make_anchors(cfg, size, size, cfg.scales[lvl])
    return flat


def anchors_tensor(anchors, device):
    """list (float64) or tensor -> float32 [A,4] on `device` (rounded once, as
    utils/output_utils.py:133 does)."""
    if isinstance(anchors, (list, tuple)):
        anchors = torch.tensor(anchors, dtype=torch.float64).reshape(-1, 4).to(torch.float32)
    return anchors.reshape(-1, 4).to(device=device, dtype=torch.float32).contiguous()


def box_iou(box_a, box_b):
    """utils/box_utils.py:8-37 on the GPU: IoU of [n,4] x [m,4] (-> [n,m]) or batched [b,n,4] x [b,m,4] (-> [b,n,m]) corner boxes."""
    from .. import _lib
    if not (isinstance(box_a, torch.Tensor) and box_a.is_cuda and box_b.is_cuda):
        raise _lib.YolactB200Error('box_iou needs CUDA tensors (no CPU fallback)')
    batched = box_a.dim() == 3
    a = box_a.to(torch.float32).contiguous()
    b = box_b.to(torch.float32).contiguous()
    if not batched:
        a, b = a[None], b[None]
    out = torch.empty(a.shape[0], a.shape[1], b.shape[1], dtype=torch.float32, device=a.device)
    L = _lib.lib()
    with torch.cuda.device(a.device):
        for i in range(a.shape[0]):
            _lib.check(L.yb_box_iou(a[i].data_ptr(), a.shape[1], b[i].data_ptr(), b.shape[1], out[i].data_ptr(),
                                    torch.cuda.current_stream().cuda_stream), 'yb_box_iou')
    return out if batched else out[0]


def mask_iou(mask1, mask2):
    """utils/box_utils.py:189-200: pairwise IoU of flattened {0,1} masks [n,N] x [m,N] -> [n,m], returned on the CPU like the
    reference (`ret.cpu()`).  Computed by AND + popcount over bit-packed words instead of a float matmul."""
    from .mask_utils import pack_masks, mask_iou_bits
    a = pack_masks(mask1.reshape(mask1.shape[0], 1, -1))
    b = pack_masks(mask2.reshape(mask2.shape[0], 1, -1))
    return mask_iou_bits(a, b).cpu()
