"""Diagnostic: is the 16-bit tensor-core forward bit-wise batch-invariant and run-to-run deterministic?  Localises the first
differing tap."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import forward_torch as ft, synth
from yolact_minimal_b200.config import make_config
from yolact_minimal_b200.modules.yolact import Yolact

def run(arch, S, B, precision='fp16'):
    dev = torch.device('cuda:0')
    cfg = make_config(arch + '_coco', S); cfg.precision, cfg.max_batch = precision, B
    net = Yolact(cfg); net.load_state_dict(ft.synth_state_dict(arch, seed=0), strict=True); net = net.to(dev).eval()
    x = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(1234)).to(dev)
    taps = ('c3', 'c4', 'c5', 'p3', 'p4', 'p5', 'p6', 'p7')
    with torch.no_grad():
        a = [o.clone() for o in net(x)]
        ta = {t: net.engine(B).read_activation(t, B).clone() for t in taps}
        b = [o.clone() for o in net(x)]
    print(f'{arch}@{S} B={B} {precision}: run-to-run equal:', [bool(torch.equal(p, q)) for p, q in zip(a, b)])
    for bi in sorted({0, B // 2, B - 1}):
        with torch.no_grad():
            o = [t.clone() for t in net(x[bi:bi + 1])]
            to = {t: net.engine(B).read_activation(t, 1).clone() for t in taps}
        msg = []
        for n, f, g in zip(('cls', 'box', 'coef', 'proto'), a, o):
            d = (f[bi:bi + 1] - g).abs()
            msg.append(f'{n}: ndiff={int((d > 0).sum())} max={float(d.max()):.3e}')
        for t in taps:
            d = (ta[t][bi:bi + 1] - to[t]).abs()
            nz = (d > 0).nonzero()
            msg.append(f'{t}: ndiff={int((d > 0).sum())} max={float(d.max()):.3e}' + (f' first={nz[0].tolist()}' if len(nz) else ''))
        print(f'  image {bi}: ' + ' | '.join(msg))

if __name__ == '__main__':
    run('res50', 128, 8)
    run('res101', 550, 64)
    run('res101', 550, 64, 'bf16')
