"""torchrun worker: BASELINE config 4 in miniature -- DistributedDataParallel (NCCL) around the Yolact module whose training
step is the native engine, SGD with the reference's warm-up-free settings, synthetic targets (seed = 1 + rank), a few steps.
Writes {loss_first, loss_last, weights_equal_across_ranks, finite} from rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main(out_path, arch='res50', S=128, per_rank=2, steps=6):
    from oracle import synth
    from yolact_minimal_b200 import dist as ydist
    import train_checks as tc
    rank, world, local = ydist.init_from_env('nccl')
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    net = tc.make_train_net(arch, S, per_rank, dev)
    ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], broadcast_buffers=True)      # train.py:76
    opt = torch.optim.SGD(ddp.parameters(), lr=3e-4, momentum=0.9, weight_decay=5e-4)
    img = torch.from_numpy(synth.image_batch(20 + rank, per_rank, S)).to(dev)
    tg, mk = synth.train_targets(1 + rank, per_rank, S)
    tgt = [torch.from_numpy(t).to(dev) for t in tg]
    mks = [torch.from_numpy(m).to(dev) for m in mk]
    hist = []
    for _ in range(steps):
        losses = ddp(img, tgt, mks)
        total = sum(losses)
        opt.zero_grad()
        total.backward()
        opt.step()
        t = total.detach().clone()
        dist.all_reduce(t)                                             # train.py:122 averages the losses for logging
        hist.append(float(t) / world)
    w = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    chk = torch.stack([w.double().sum(), (w.double() ** 2).sum()])
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    if rank == 0:
        json.dump({'loss_first': hist[0], 'loss_last': hist[-1], 'hist': hist, 'finite': bool(np.isfinite(hist).all()),
                   'weights_equal_across_ranks': bool(all(torch.equal(allc[0], c) for c in allc)), 'world': world}, open(out_path, 'w'))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main(sys.argv[1])
