"""One post-process call per regime (after a warm-up) for an ncu launch list."""
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
    for _ in range(2): detect_batched(cls, box, coef, anchors, cfg)
    torch.cuda.synchronize()
