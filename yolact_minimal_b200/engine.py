"""Python handle on the C++/CUDA network (yb_net_* in include/yolact_b200.h): feeds it the
module's parameters by their state-dict names, keeps it in sync when they change, and runs
the forward on torch's current stream into torch-allocated outputs (torch owns device memory
and streams; every FLOP happens in libyolact_b200.so)."""
import ctypes

import numpy as np
import torch

from . import _lib

PRECISIONS = {'fp32': 0, 'bf16': 1, 'fp16': 2}
# fp16 operands / fp32 accumulate: the 16-bit mode that meets north_star's 1e-2 bound
DEFAULT_PRECISION = 'fp16'


class Engine:
    def __init__(self, depth, img_size, num_classes, num_ratios, coef_dim):
        self.L = _lib.lib()
        self.cfg = _lib.NetConfig(depth, img_size, num_classes, num_ratios, coef_dim)
        h = ctypes.c_void_p()
        _lib.check(self.L.yb_net_create(ctypes.byref(self.cfg), ctypes.byref(h)), 'yb_net_create')
        self.h = h
        self.num_anchors = self.L.yb_net_num_anchors(h)
        self.proto_size = self.L.yb_net_proto_size(h)
        self.names = []
        for i in range(self.L.yb_net_num_params(h)):
            name, cnt = ctypes.c_char_p(), ctypes.c_int64()
            _lib.check(self.L.yb_net_param_info(h, i, ctypes.byref(name), ctypes.byref(cnt)), 'yb_net_param_info')
            self.names.append((name.value.decode(), cnt.value))
        self._sig = None
        self.max_batch = 0
        self.precision = None
        self.device = None

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                self.L.yb_net_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def param_names(self):
        return [n for n, _ in self.names]

    def sync(self, module, precision=DEFAULT_PRECISION, min_batch=1):
        """(Re)load parameters when the module's tensors changed; (re)finalize when the batch
        grows or the precision changes."""
        if precision not in PRECISIONS:
            raise ValueError(f'precision must be one of {list(PRECISIONS)}, got {precision!r}')
        sd = dict(module.named_parameters())
        sd.update(dict(module.named_buffers()))
        sig = tuple((sd[n]._version, sd[n].data_ptr()) for n, _ in self.names)
        dev = next(module.parameters()).device
        if dev.type != 'cuda':
            raise _lib.YolactB200Error('model parameters must live on a CUDA device (no CPU fallback)')
        need = (sig != self._sig or precision != self.precision or min_batch > self.max_batch or dev != self.device)
        if not need:
            return
        with torch.cuda.device(dev):
            if sig != self._sig:
                self.load_state(sd)
            mb = max(min_batch, self.max_batch)
            _lib.check(self.L.yb_net_finalize(self.h, mb, PRECISIONS[precision]), 'yb_net_finalize')
        self._sig, self.max_batch, self.precision, self.device = sig, mb, precision, dev

    def invalidate(self):
        """Force the next forward to re-read every parameter.  sync() detects re-assigned or in-place
        modified tensors through (_version, data_ptr); writes through `.data` (p.data.copy_(), EMA / weight
        surgery) bump neither, so call this after them."""
        self._sig = None

    def load_state(self, sd):
        for name, cnt in self.names:
            if name not in sd:
                raise KeyError(f'parameter {name!r} missing from the state dict')
            a = np.ascontiguousarray(sd[name].detach().to('cpu', torch.float32).numpy())
            if a.size != cnt:
                raise ValueError(f'{name}: {a.size} elements, engine expects {cnt}')
            _lib.check(self.L.yb_net_set_param(self.h, name.encode(), a.ctypes.data, cnt), f'yb_net_set_param({name})')

    def finalize(self, max_batch, precision):
        _lib.check(self.L.yb_net_finalize(self.h, max_batch, PRECISIONS[precision]), 'yb_net_finalize')
        self.max_batch, self.precision = max_batch, precision

    # ------------------------------------------------------------------------------------------
    def forward(self, img):
        B = img.shape[0]
        S = self.cfg.img_size
        if tuple(img.shape[1:]) != (3, S, S):
            raise ValueError(f'expected input [B,3,{S},{S}], got {tuple(img.shape)}')
        img = img.detach().to(torch.float32).contiguous()
        dev = img.device
        A, P, C, K = self.num_anchors, self.proto_size, self.cfg.num_classes, self.cfg.coef_dim
        cls = torch.empty(B, A, C, dtype=torch.float32, device=dev)
        box = torch.empty(B, A, 4, dtype=torch.float32, device=dev)
        coef = torch.empty(B, A, K, dtype=torch.float32, device=dev)
        proto = torch.empty(B, P, P, K, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(self.L.yb_net_forward(self.h, img.data_ptr(), B, cls.data_ptr(), box.data_ptr(), coef.data_ptr(),
                                             proto.data_ptr(), torch.cuda.current_stream().cuda_stream), 'yb_net_forward')
        return cls, box, coef, proto

    def read_activation(self, name, batch):
        C, H, W = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(self.L.yb_net_read_activation(self.h, name.encode(), batch, None, 0, ctypes.byref(C), ctypes.byref(H),
                                                 ctypes.byref(W), None), 'yb_net_read_activation')
        out = torch.empty(batch, C.value, H.value, W.value, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.L.yb_net_read_activation(self.h, name.encode(), batch, out.data_ptr(), out.numel(), ctypes.byref(C),
                                                     ctypes.byref(H), ctypes.byref(W), torch.cuda.current_stream().cuda_stream),
                       'yb_net_read_activation')
        return out

    def set_profiling(self, enable):
        _lib.check(self.L.yb_net_set_profiling(self.h, 1 if enable else 0), 'yb_net_set_profiling')

    def profile(self):
        """Per-kernel device time / algorithmic work of the forwards since the last call."""
        buf = (_lib.ProfEntry * 16)()
        n = ctypes.c_int()
        _lib.check(self.L.yb_net_profile(self.h, buf, 16, ctypes.byref(n)), 'yb_net_profile')
        return {buf[i].name.decode(): dict(launches=buf[i].launches, forwards=buf[i].forwards, ms=buf[i].ms,
                                           flops=buf[i].flops, bytes=buf[i].bytes) for i in range(n.value)}

    def anchors(self):
        a = np.empty((self.num_anchors, 4), np.float32)
        _lib.check(self.L.yb_net_anchors_host(self.h, a.ctypes.data), 'yb_net_anchors_host')
        return a

    def set_anchors(self, anchors):
        """Use the caller's anchor table ([A,4] cx,cy,w,h) in the C pipelines (cfg.scales / cfg.aspect_ratios)."""
        a = np.ascontiguousarray(np.asarray(anchors, dtype=np.float64).reshape(-1, 4).astype(np.float32))
        if a.shape[0] != self.num_anchors:
            raise ValueError(f'{a.shape[0]} anchors, the engine has {self.num_anchors}')
        _lib.check(self.L.yb_net_set_anchors(self.h, a.ctypes.data, a.shape[0]), 'yb_net_set_anchors')

    def submit_host(self, img_host, params):
        """Pipelined end-to-end call (yb_net_submit_host): returns a ticket; the H2D copy of this batch
        overlaps the compute of the previous one.  img_host must be a contiguous float32 array that
        stays alive (ideally pinned) until collect_host(ticket)."""
        t = ctypes.c_int()
        _lib.check(self.L.yb_net_submit_host(self.h, img_host.ctypes.data, img_host.shape[0], ctypes.byref(params), ctypes.byref(t)),
                   'yb_net_submit_host')
        self._pending = getattr(self, '_pending', {})
        self._pending[t.value] = (img_host, img_host.shape[0], params.max_det)
        return t.value

    def collect_host(self, ticket):
        _, B, D = self._pending.pop(ticket)
        K = self.cfg.coef_dim
        out = dict(count=np.zeros(B, np.int32), cls=np.zeros((B, D), np.int32), anchor=np.zeros((B, D), np.int32),
                   score=np.zeros((B, D), np.float32), box=np.zeros((B, D, 4), np.float32), coef=np.zeros((B, D, K), np.float32))
        _lib.check(self.L.yb_net_collect_host(self.h, ticket, out['count'].ctypes.data, out['cls'].ctypes.data, out['anchor'].ctypes.data,
                                              out['score'].ctypes.data, out['box'].ctypes.data, out['coef'].ctypes.data),
                   'yb_net_collect_host')
        return out

    def detect_host(self, img_host, params):
        """End to end with host buffers (yb_net_detect_host): numpy in, numpy out."""
        img = np.ascontiguousarray(img_host, dtype=np.float32)
        B, D, K = img.shape[0], params.max_det, self.cfg.coef_dim
        out = dict(count=np.zeros(B, np.int32), cls=np.zeros((B, D), np.int32), anchor=np.zeros((B, D), np.int32),
                   score=np.zeros((B, D), np.float32), box=np.zeros((B, D, 4), np.float32), coef=np.zeros((B, D, K), np.float32))
        _lib.check(self.L.yb_net_detect_host(self.h, img.ctypes.data, B, ctypes.byref(params), out['count'].ctypes.data,
                                             out['cls'].ctypes.data, out['anchor'].ctypes.data, out['score'].ctypes.data,
                                             out['box'].ctypes.data, out['coef'].ctypes.data), 'yb_net_detect_host')
        return out
