"""One inference step (forward + post-process) at the bench configuration -- the target of ncu captures.
usage: python tools/one_forward.py [batch] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import forward_torch as ft
from yolact_minimal_b200.config import make_config
from yolact_minimal_b200.modules.yolact import Yolact
from yolact_minimal_b200.utils.output_utils import detect_batched

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = make_config('res101_coco', 550)
cfg.precision, cfg.max_batch = 'fp16', B
net = Yolact(cfg)
net.load_state_dict(ft.synth_state_dict('res101', seed=0), strict=True)
net = net.cuda().eval()
anchors = torch.from_numpy(net.engine(B).anchors()).cuda()
img = torch.randn(B, 3, 550, 550, device='cuda')
for _ in range(reps):
    with torch.no_grad():
        cls, box, coef, proto = net(img)
    det = detect_batched(cls, box, coef, anchors, cfg)
torch.cuda.synchronize()
print('ok', int(det['count'].sum()))
