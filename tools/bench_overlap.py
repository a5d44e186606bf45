"""Does the post-process of batch i hide under the forward of batch i+1 when it runs on a second stream?
usage: python tools/bench_overlap.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import forward_torch as ft
from yolact_minimal_b200.config import make_config
from yolact_minimal_b200.modules.yolact import Yolact
from yolact_minimal_b200.utils.output_utils import detect_batched

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = make_config('res101_coco', 550)
cfg.precision, cfg.max_batch = 'fp16', B
net = Yolact(cfg)
net.load_state_dict(ft.synth_state_dict('res101', seed=0), strict=True)
net = net.cuda().eval()
anchors = torch.from_numpy(net.engine(B).anchors()).cuda()
imgs = [torch.randn(B, 3, 550, 550, device='cuda') for _ in range(2)]
main = torch.cuda.current_stream()


def timed(fn, steps=12):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    for s in streams:
        main.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


streams = []


def serial(i):
    with torch.no_grad():
        cls, box, coef, proto = net(imgs[i & 1])
    return detect_batched(cls, box, coef, anchors, cfg)


print(f'serial                : {timed(serial):.3f} ms/step')
for prio in (0, -1):
    side = torch.cuda.Stream(priority=prio)
    streams[:] = [side]
    keep = [None, None]

    def piped(i):
        with torch.no_grad():
            cls, box, coef, proto = net(imgs[i & 1])
        side.wait_stream(main)
        with torch.cuda.stream(side):
            det = detect_batched(cls, box, coef, anchors, cfg)
        for t in (cls, box, coef):
            t.record_stream(side)
        keep[i & 1] = (cls, box, coef, proto, det)
        return det
    print(f'post on side stream {prio:2d}: {timed(piped):.3f} ms/step')
