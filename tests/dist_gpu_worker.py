"""torchrun worker for tests/test_parity_gaps_gpu.py::test_two_rank_gather_equals_single_gpu (and usable by hand:
`python -m torch.distributed.run --nproc-per-node N tests/dist_gpu_worker.py out.json`).

Every rank runs forward + decode/Fast-NMS/top-k on ITS shard of a seeded global batch and the detection records are
all-gathered over NCCL; rank 0 then runs the whole global batch on its own GPU and requires the gathered records to equal
the single-GPU ones bit-for-bit.  Imports oracle/ only for the seeded weights and images."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out_path):
    from oracle import synth, forward_torch as ft
    from yolact_minimal_b200 import dist as ydist
    from yolact_minimal_b200.config import make_config
    from yolact_minimal_b200.modules.yolact import Yolact
    from yolact_minimal_b200.utils.output_utils import detect_batched
    rank, world, local = ydist.init_from_env('nccl')
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    arch, S, per_rank = 'res50', 256, 4
    G = per_rank * world
    cfg = make_config(arch + '_coco', S)
    cfg.precision, cfg.max_batch = 'fp16', G
    cfg.nms_score_thre = 0.012                                     # random-init scores sit just above 1/81: keep plenty of candidates
    net = Yolact(cfg)
    net.load_state_dict(ft.synth_state_dict(arch, seed=0), strict=True)
    net = net.to(dev).eval()
    img = torch.from_numpy(synth.image_batch(77, G, S)).to(dev)
    lo, hi = ydist.shard_range(G, rank, world)
    with torch.no_grad():
        cls, box, coef, _ = net(img[lo:hi])
    det = detect_batched(cls, box, coef, net.anchors, cfg)
    gathered = ydist.gather_detections(det)
    torch.cuda.synchronize()
    if rank == 0:
        with torch.no_grad():
            cls, box, coef, _ = net(img)
        single = detect_batched(cls, box, coef, net.anchors, cfg)
        ok = all(torch.equal(gathered[k].view(torch.int32), single[k].view(torch.int32)) for k in ('count', 'cls', 'anchor', 'score', 'box', 'coef'))
        json.dump({'bit_exact': bool(ok), 'world': world, 'per_rank': per_rank, 'images': int(gathered['count'].shape[0]),
                   'detections': int(single['count'].sum())}, open(out_path, 'w'))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main(sys.argv[1])
