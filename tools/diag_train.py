"""Diagnostic run of the native training path on a GPU: prints every comparison of tests/train_checks.py (no asserts)."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import train_checks as tc
from conftest import load_golden
from yolact_minimal_b200.config import make_config

dev = torch.device('cuda:0')

def stage():
    g = load_golden('train_stages.npz')
    tg, mk, class_p, box_p, coef_p, proto_p, seg_p = tc.stage_inputs(g)
    cfg = make_config('res50_coco', tc.S_ST, mode='train', train_bs=tc.B_ST)
    r = tc.run_native_losses(dev, cfg, g['anchors'], tg, mk, class_p, box_p, coef_p, proto_p, seg_p)
    print('labels equal', np.array_equal(r['labels'], g['labels']), 'n diff', int((r['labels'] != g['labels']).sum()))
    print('matched_idx equal', np.array_equal(r['matched_idx'], g['matched_idx']), int((r['matched_idx'] != g['matched_idx']).sum()))
    pos = g['labels'] > 0
    print('offsets max err on positives', float(np.abs(r['offsets'][pos] - g['offsets'][pos]).max()))
    print('ohem neg agreement', float((r['neg'].astype(bool) == g['ohem_neg']).mean()), 'n diff', int((r['neg'].astype(bool) != g['ohem_neg']).sum()),
          'count', int(r['neg'].sum()), int(g['ohem_neg'].sum()))
    print('losses', r['losses'].tolist(), 'golden', [float(g[k]) for k in ('loss_c', 'loss_b', 'loss_m', 'loss_s')])
    ref_l, ref_g = tc.torch_losses_and_grads(cfg, g['anchors'], tg, mk, class_p, box_p, coef_p, proto_p, seg_p)
    print('checker losses', ref_l)
    for k in ('d_cls', 'd_box', 'd_coef', 'd_proto', 'd_seg'):
        print(k, 'rel err', tc.rel(r[k], ref_g[k]), 'cos', tc.cos(r[k], ref_g[k]), 'max abs', float(np.abs(r[k] - ref_g[k]).max()), 'ref norm', float(np.linalg.norm(ref_g[k])))

def engine(arch, S, B, precision, mode='fp32'):
    o = tc.engine_vs_checker(arch, S, B, dev, precision, mode=mode)
    print(f'--- engine {arch}@{S} B={B} {precision} mode={mode}: launches/step {o["launches"]} substituted {o["substituted"]}')
    print('losses', o['losses'], 'checker', o['ref_losses'])
    big = lambda d: {k: (round(v, 5) if isinstance(v, float) else v) for k, v in sorted(d.items(), key=lambda kv: -(kv[1] if isinstance(kv[1], float) else 9))[:24]}
    print('activations (rel err), worst 24:', big(o['act']))
    print('activation gradients (rel err), worst 24:', big(o['gact']))
    worst = sorted(o['grads'].items(), key=lambda kv: -(kv[1][0] if kv[1][0] == kv[1][0] else 1e9))
    print('parameter gradients, worst 40 by rel err (rel, cos, ref norm):')
    for n, v in worst[:40]:
        print(f'   {n:60s} rel {v[0]:.4f} cos {v[1]:.5f} |ref| {v[2]:.3e}')
    rels = np.asarray([v[0] for v in o['grads'].values()])
    print('median rel', float(np.nanmedian(rels)), 'max', float(np.nanmax(rels)), 'nan', int(np.isnan(rels).sum()), 'of', len(rels))
    bn = sorted(o['bn'].items(), key=lambda kv: -kv[1])[:5]
    print('BN buffers worst abs err', bn)

import os
ORDER = os.environ.get('ORDER', 'bf16,fp16').split(',')
for fn, args in tuple((engine, ('res50', 128, 2, pr, 'subst')) for pr in ORDER):
    try:
        fn(*args)
    except Exception:
        traceback.print_exc()
    sys.stdout.flush()
