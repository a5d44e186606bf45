import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import test_parity_gaps_gpu as T
from oracle import postprocess_np as pp
from yolact_minimal_b200.config import make_config
from yolact_minimal_b200.utils.output_utils import detect_batched
cuda = torch.device('cuda:0')
thr = 0.5
anchors, cls, box, coef, iou = T._near_threshold_case(thr)
cfg = make_config('res101_coco', 544)
cfg.nms_iou_thre, cfg.top_k, cfg.max_detections = thr, 200, 256
n = len(anchors)
owner = (np.arange(n) // 12) // 20
cls_b = np.repeat(cls[None], 4, 0)
for g in range(4):
    off = owner != g
    cls_b[g, off, 1:] = 0; cls_b[g, off, 0] = 1
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
for rep in range(3):
    r = detect_batched(t(cls_b), t(np.repeat(box[None], 4, 0)), t(np.repeat(coef[None], 4, 0)), t(anchors), cfg)
    for g in range(4):
        o = pp.nms(cls_b[g], box, anchors, iou_thre=thr, top_k=200, max_det=256)
        d = int(r['count'][g])
        rc, ra, rs = r['cls'][g, :d].cpu().numpy(), r['anchor'][g, :d].cpu().numpy(), r['score'][g, :d].cpu().numpy()
        same = d == len(o[0]) and np.array_equal(rc, o[0]) and np.array_equal(ra, o[3])
        print(f'rep {rep} image {g}: d={d} oracle={len(o[0])} same={same}')
        if not same:
            k = min(d, len(o[0]))
            bad = np.nonzero((rc[:k] != o[0][:k]) | (ra[:k] != o[3][:k]))[0]
            print('   first diffs at', bad[:10], 'ours (cls, anchor, score):', [(int(rc[i]), int(ra[i]), float(rs[i])) for i in bad[:6]],
                  'oracle:', [(int(o[0][i]), int(o[3][i]), float(o[1][i])) for i in bad[:6]])
            so, ss = set(zip(o[0].tolist(), o[3].tolist())), set(zip(rc.tolist(), ra.tolist()))
            print('   only ours', sorted(ss - so)[:10], 'only oracle', sorted(so - ss)[:10])
