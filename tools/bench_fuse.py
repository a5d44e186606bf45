"""A/B of the bottleneck fusion: img/s of the inference step (forward + post-process) with k_bneck_tc on / off, plus the per-op
profile of the fused layers.  usage: python tools/bench_fuse.py [batch]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import forward_torch as ft


def run(fuse, B, steps=10):
    from yolact_minimal_b200.config import make_config
    from yolact_minimal_b200.modules.yolact import Yolact
    from yolact_minimal_b200.utils.output_utils import detect_batched
    os.environ.pop('YOLACT_B200_NO_FUSE', None)
    if not fuse:
        os.environ['YOLACT_B200_NO_FUSE'] = '1'
    cfg = make_config('res101_coco', 550)
    cfg.precision, cfg.max_batch = 'fp16', B
    net = Yolact(cfg)
    net.load_state_dict(ft.synth_state_dict('res101', seed=0), strict=True)
    net = net.cuda().eval()
    eng = net.engine(B)
    anchors = torch.from_numpy(eng.anchors()).cuda()
    imgs = [torch.randn(B, 3, 550, 550, device='cuda') for _ in range(2)]

    def step(i):
        with torch.no_grad():
            cls, box, coef, proto = net(imgs[i & 1])
        return detect_batched(cls, box, coef, anchors, cfg)
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    eng.set_profiling(True)
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    prof = eng.profile()
    eng.set_profiling(False)
    print(f'fuse={int(fuse)} B={B}: {ms:.3f} ms/step  {B / ms * 1e3:.0f} img/s')
    for k, v in prof.items():
        if v['launches']:
            f = max(1, v['forwards'])
            print(f"   {k:16s} {v['ms'] / f:8.3f} ms/step  {v['launches'] / f:5.0f} launches  "
                  f"{(v['flops'] / (v['ms'] * 1e-3) / 1e12) if v['flops'] and v['ms'] else 0:7.1f} TF/s  {v['bytes'] / (v['ms'] * 1e-3) / 1e9 if v['ms'] else 0:7.0f} GB/s")
    del net, eng
    torch.cuda.empty_cache()


if __name__ == '__main__':
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    which = sys.argv[2] if len(sys.argv) > 2 else 'both'          # 'fused' | 'plain' | 'both'; other switches through the environment
    if which in ('plain', 'both'):
        run(False, B)
    if which in ('fused', 'both'):
        run(True, B)
