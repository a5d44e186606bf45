"""CPU: (1) the torch-autograd CHECKER of the training branch (oracle/train_torch.py -- what the native engine's losses and
gradients are compared against on the GPU, tests/test_train_gpu.py) reproduces the reference's four losses, gradient norms and
BatchNorm statistics (goldens minted by the reference in train mode); (2) the product's autograd plumbing
(train_native._NativeTrainStep: ONE Function whose inputs are the parameters) delivers engine-made gradients to the
parameters and through DistributedDataParallel's all-reduce hooks (gloo, world_size 2), like train.py does with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
from oracle import synth, forward_torch as ft, train_torch as tt


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
    losses = tt.training_step_forward(net, img, [torch.from_numpy(t) for t in tg], [torch.from_numpy(m) for m in mk])
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


def test_training_on_cpu_is_refused():
    net = make_train_net('res50', 64, 1)
    img = torch.from_numpy(synth.image_batch(1, 1, 64))
    tg, mk = synth.train_targets(7, 1, 64)
    with pytest.raises(RuntimeError):                                  # no CPU fallback: the product path is the native engine
        net(img, [torch.from_numpy(t) for t in tg], [torch.from_numpy(m) for m in mk])


class _FakeEngine:
    """Stands in for train_native.TrainEngine on a box without a GPU: losses = sum of squares of the parameters, gradients made
    'by the engine' (2 * p * scale, rank-dependent) and handed to autograd through the same Function the product uses."""

    def __init__(self, params, scale):
        self._params, self.scale = params, scale
        self.sizes = [p.numel() for p in params]

    def forward(self, *a):
        with torch.no_grad():
            s = sum((p.double() ** 2).sum() for p in self._params).float()
        return torch.stack([s, s * 0, s * 0, s * 0])

    def backward(self, grad_losses):
        with torch.no_grad():
            return torch.cat([(2 * p * self.scale * grad_losses[0]).reshape(-1) for p in self._params])

    def grad_views(self, flat):
        return list(torch.split(flat, self.sizes))


def _ddp_worker(rank, world, port, q):
    from yolact_minimal_b200.train_native import _NativeTrainStep
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Conv2d(3, 4, 3)
            self.b = torch.nn.BatchNorm2d(4)

        def forward(self, x):
            params = [p for p in self.parameters()]
            eng = _FakeEngine(params, float(rank + 1))
            return _NativeTrainStep.apply(eng, x, None, None, None, 0, 0, 0, *params).unbind(0)
    net = Tiny()
    ddp = torch.nn.parallel.DistributedDataParallel(net, broadcast_buffers=True)        # train.py:76
    opt = torch.optim.SGD(ddp.parameters(), lr=1e-2)
    losses = ddp(torch.zeros(1, 3, 8, 8))
    opt.zero_grad()
    w0 = net.a.weight.detach().clone()
    sum(losses).backward()
    g = net.a.weight.grad.clone()
    opt.step()
    q.put((rank, float((g - 2 * w0 * 1.5).abs().max()), float(net.a.weight.detach().double().sum())))   # mean of scales 1 and 2
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_allreduces_engine_gradients_gloo_world2():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict((r, (e, w)) for r, e, w in (q.get(timeout=300) for _ in range(2)))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert got[0][0] < 1e-6 and got[1][0] < 1e-6       # every rank holds the rank-AVERAGED engine gradient
    assert got[0][1] == got[1][1]                      # identical weights after the all-reduced step
