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
