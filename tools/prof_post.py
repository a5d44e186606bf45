"""Time the post-process kernels at BASELINE size (A=19248, B=64) in the three regimes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth, postprocess_np as pp
from yolact_minimal_b200.config import make_config
from yolact_minimal_b200.utils.output_utils import detect_batched
dev = torch.device('cuda:0'); B = 64
cfg = make_config('res101_coco', 550)
anchors = torch.from_numpy(pp.make_anchors(550)).to(dev); A = anchors.shape[0]
for regime in ('stress', 'realistic', 'sparse'):
    c, b, k = synth.head_outputs(7, A, 81, regime)
    rep = lambda a: torch.from_numpy(a).to(dev)[None].expand(B, *a.shape).contiguous()
    cls, box, coef = rep(c), rep(b), rep(k)
    for _ in range(3): detect_batched(cls, box, coef, anchors, cfg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): detect_batched(cls, box, coef, anchors, cfg)
    e1.record(); torch.cuda.synchronize()
    print(f'{regime}: {1e3 * e0.elapsed_time(e1) / (10 * B):.2f} us/img', flush=True)
