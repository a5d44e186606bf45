"""CPU: the training branch (first cut on torch autograd, yolact_minimal_b200/train_torch.py) reproduces
the reference's four losses, gradients and BatchNorm statistics (goldens minted by the reference in train
mode), and runs under DDP with the gloo backend (world_size 2), like train.py does with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
from oracle import synth, forward_torch as ft


def make_train_net(arch, S, B):
    from yolact_minimal_b200.config import make_config
    from yolact_minimal_b200.modules.yolact import Yolact
    cfg = make_config(arch + '_coco', S, mode='train', train_bs=B)
    net = Yolact(cfg)
    net.load_state_dict(ft.synth_state_dict(arch, seed=0, train=True), strict=True)
    return net.train()


@pytest.mark.parametrize('arch,S,B', [('res50', 128, 2), ('res101', 96, 2)])
def test_training_losses_and_grads_match_reference(arch, S, B):
    g = load_golden('train.npz')
    key = f'{arch}_S{S}_B{B}'
    net = make_train_net(arch, S, B)
    img = torch.from_numpy(synth.image_batch(11, B, S))
    tg, mk = synth.train_targets(5, B, S)
    losses = net(img, [torch.from_numpy(t) for t in tg], [torch.from_numpy(m) for m in mk])
    assert len(losses) == 4
    got = np.asarray([float(l.detach()) for l in losses])
    assert np.allclose(got, g[key + '/losses'], rtol=2e-5, atol=1e-5), (got, g[key + '/losses'])
    sum(losses).backward()
    named = dict(net.named_parameters())
    for k in g.files:
        if k.startswith(key + '/grad/'):
            pn = k[len(key + '/grad/'):]
            assert np.isclose(float(named[pn].grad.double().norm()), float(g[k]), rtol=1e-4), pn
    assert np.allclose(net.backbone.bn1.running_mean.numpy(), g[key + '/bn1_mean'], atol=1e-6)      # BN is live in train mode


def _ddp_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    net = make_train_net('res50', 64, 2)
    ddp = torch.nn.parallel.DistributedDataParallel(net, broadcast_buffers=True)        # train.py:76
    opt = torch.optim.SGD(ddp.parameters(), lr=1e-4, momentum=0.9, weight_decay=5e-4)
    img = torch.from_numpy(synth.image_batch(20 + rank, 1, 64))
    tg, mk = synth.train_targets(7 + rank, 1, 64)
    losses = ddp(img, [torch.from_numpy(t) for t in tg], [torch.from_numpy(m) for m in mk])
    opt.zero_grad()
    sum(losses).backward()
    opt.step()
    w = net.prediction_layers.conf_layer.weight.detach()
    q.put((rank, float(w.double().sum()), [float(l.detach()) for l in losses]))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_training_step_gloo_world2():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict((r, (w, l)) for r, w, l in (q.get(timeout=300) for _ in range(2)))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert got[0][0] == got[1][0]                 # identical weights after the all-reduced step
    assert all(np.isfinite(got[r][1]).all() for r in (0, 1))
