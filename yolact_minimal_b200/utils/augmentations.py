"""`val_aug` kept from the reference's utils/augmentations.py:219-227, computed on the GPU by
yb_val_aug (one fused kernel: pad-to-square with the BGR mean, bilinear resize, normalise,
BGR->RGB, HWC->CHW).  Takes the uint8 BGR image (numpy HWC, as cv2.imread returns it, or a CUDA
uint8 tensor) and returns a CUDA float32 tensor [3, val_size, val_size]."""
import numpy as np
import torch

from .. import _lib


def val_aug(img, val_size, device='cuda'):
    if isinstance(img, np.ndarray):
        if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
            raise ValueError(f'expected a uint8 HxWx3 BGR image, got {img.dtype} {img.shape}')
        img = torch.from_numpy(np.ascontiguousarray(img)).to(device)
    if not (img.is_cuda and img.dtype == torch.uint8):
        raise _lib.YolactB200Error('val_aug needs a uint8 image on a CUDA device (no CPU fallback)')
    img = img.contiguous()
    h, w = int(img.shape[0]), int(img.shape[1])
    out = torch.empty(3, int(val_size), int(val_size), dtype=torch.float32, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(_lib.lib().yb_val_aug(img.data_ptr(), h, w, int(val_size), out.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), 'yb_val_aug')
    return out
