"""Drop-in for the reference's only native component, `cython_nms.nms(dets, thresh)`
(cython_nms.pyx:24-74): greedy NMS with '+1' pixel areas, suppress on ovr >= thresh, returns
the surviving indices in ascending original order -- computed by the CUDA kernel k_hard_nms
through the C ABI (yb_hard_nms_host / yb_hard_nms)."""
import numpy as np

from . import _lib


def nms(dets, thresh):
    """dets: np.float32 [n,5] (x1,y1,x2,y2,score) host array, or a CUDA torch tensor.
    Returns np.int64 [k] (host), like the reference."""
    try:
        import torch
        is_tensor = isinstance(dets, torch.Tensor)
    except ImportError:        # pragma: no cover
        is_tensor = False
    L = _lib.lib()
    if is_tensor and dets.is_cuda:
        d = dets.detach().to(torch.float32).contiguous()
        n = d.shape[0]
        keep = torch.empty(n, dtype=torch.uint8, device=d.device)
        if n:
            with torch.cuda.device(d.device):
                _lib.check(L.yb_hard_nms(d.data_ptr(), n, float(thresh), keep.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), 'yb_hard_nms')
        return np.nonzero(keep.cpu().numpy() == 1)[0].astype(np.int64)
    d = np.ascontiguousarray(np.asarray(dets), dtype=np.float32)
    if d.ndim != 2 or d.shape[1] != 5:
        raise ValueError(f'dets must be [n,5], got {d.shape}')
    n = d.shape[0]
    keep = np.zeros(n, dtype=np.uint8)
    if n:
        _lib.check(L.yb_hard_nms_host(d.ctypes.data, n, float(thresh), keep.ctypes.data), 'yb_hard_nms_host')
    return np.nonzero(keep == 1)[0].astype(np.int64)
