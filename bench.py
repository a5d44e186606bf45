#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on synthetic data.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp16|bf16|fp32] [--impl reference]

Metric (BASELINE.json): img/s of res101_coco 550x550, batch 64 per GPU, eval forward + fused
post-process (decode + Fast-NMS + top-k); Fast-NMS us/img reported alongside.  One "step" = one
pass of the hot path over one 64-image batch per GPU.  For N > 1 the driver launches one rank per
GPU with torch.distributed.run; images shard across ranks (no data-path collective) and the
detection records are all-gathered over NCCL once per step (weak scaling: 64 images per GPU).

Rank 0 prints ONE JSON line.  `value` is measured with inputs resident in HBM, `e2e` through the
C-ABI host-buffer entry point yb_net_detect_host (pinned host input, H2D + forward + post-process
+ D2H of the detection records inside the timed region).  `--impl reference` times the CPU port
of the reference path (oracle/) on the host cores for the same metric/config.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ARCH, IMG, BATCH = 'res101', 550, 64
# algorithmic 2*MAC of the reference forward per image
GFLOPS = {('res101', 550): 164.68, ('res101', 544): 157.16, ('res50', 550): 118.28, ('res50', 544): 113.38,
          ('swin_tiny', 550): 123.43, ('swin_tiny', 544): 119.19}
GFLOP_PER_IMG = GFLOPS[(ARCH, IMG)]


def host_cores():
    """CPU threads this process can actually use: min(affinity, cgroup cpu.max quota)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    try:
        q, per = open('/sys/fs/cgroup/cpu.max').read().split()
        if q != 'max':
            n = min(n, max(1, int(int(q) / int(per))))
    except (OSError, ValueError):
        pass
    return n


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf=d['bf16_tflops'], tf_sustained=d['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src='fallback')


# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100',
                                          '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self, t0=None, t1=None):
        """Summarise the samples taken between wall-clock times t0 and t1 (the timed region)."""
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(2)
        import datetime
        sm, pw, mx, reasons = [], [], None, set()
        for l in self.lines:
            f = [x.strip() for x in l.split(',')]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.05):
                    continue
                sm.append(float(f[1])); mx = float(f[2]); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'power_w_max': max(pw) if pw else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ----------------------------------------------------------------------------------------------------
def cpu_reference_sample(n_img, reps, threads):
    """The reference's CPU path restated (oracle/): eval forward + nms() per image, fp32, all
    host threads.  Returns (img_per_s, seconds, fast_nms_us_per_img)."""
    import torch
    from oracle import synth, forward_torch as ft, postprocess_np as pp
    torch.set_num_threads(threads)
    sd = ft.synth_state_dict(ARCH, seed=0)
    img = torch.from_numpy(synth.image_batch(1, n_img, IMG))
    anchors = pp.make_anchors(IMG)
    ft.forward(img[:1], sd, ARCH)                                  # warm-up
    t0 = time.perf_counter()
    t_nms = 0.0
    for _ in range(reps):
        cls, box, coef, proto = [t.numpy() for t in ft.forward(img, sd, ARCH)]
        t1 = time.perf_counter()
        for b in range(n_img):
            pp.nms(cls[b], box[b], anchors)
        t_nms += time.perf_counter() - t1
    dt = time.perf_counter() - t0
    return n_img * reps / dt, dt, 1e6 * t_nms / (n_img * reps)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = host_cores()
    n_img = 4
    for _ in range(args.warmup):
        pass                                                        # the sample's own warm-up forward is inside
    vals, t_all, nms_us = [], 0.0, 0.0
    steps = max(1, min(args.steps, 20))                             # bounded: each step is a 4-image sample
    for _ in range(steps):
        v, dt, nu = cpu_reference_sample(n_img, 1, cores)
        vals.append(v); t_all += dt; nms_us = nu
    value = float(np.mean(vals))
    line = {'impl': 'reference', 'metric': 'img/s', 'value': value, 'unit': 'img/s', 'n_gpus': args.gpus, 'steps': steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * t_all / steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'{ARCH}_coco {IMG}x{IMG} eval forward + Fast-NMS, CPU port of the reference path (oracle/), '
                                   f'bounded sample of {n_img} images per step (full workload: bs={BATCH}/GPU)'},
            'cpu_baseline': {'value': value, 'unit': 'img/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{steps} x {n_img} images, torch fp32 CPU forward + numpy nms()'},
            'fast_nms_us_per_img': nms_us,
            'e2e': {'value': value, 'unit': 'img/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from oracle import synth, forward_torch as ft, postprocess_np as pp        # cpu_baseline leg + synthetic inputs only
    from yolact_minimal_b200 import _lib, dist as ydist
    from yolact_minimal_b200.config import make_config
    from yolact_minimal_b200.modules.yolact import Yolact
    from yolact_minimal_b200.utils.output_utils import detect_batched

    rank, world, local = ydist.init_from_env()
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    K, W = args.steps, max(args.warmup, 3)
    pk = peaks()

    cfg = make_config(ARCH + '_coco', IMG)
    cfg.precision, cfg.max_batch = args.precision, BATCH
    net = Yolact(cfg)
    net.load_state_dict(ft.synth_state_dict(ARCH, seed=0), strict=True)
    net = net.to(dev).eval()
    eng = net.engine(BATCH)
    anchors = torch.from_numpy(eng.anchors()).to(dev)

    # two distinct resident input batches (232 MB each > 126 MB L2), alternated between steps
    gen = torch.Generator().manual_seed(1234 + rank)
    host = [torch.randn(BATCH, 3, IMG, IMG, generator=gen).pin_memory() for _ in range(2)]     # ~N(0,1) like normalised RGB
    imgs = [h.to(dev) for h in host]
    L = _lib.lib()

    def step(i):
        with torch.no_grad():
            cls, box, coef, proto = net(imgs[i & 1])
        det = detect_batched(cls, box, coef, anchors, cfg)
        if world > 1:
            det = ydist.gather_detections(det)
        return det, (cls, box, coef, proto)

    sampler = ClockSampler(local)
    sampler.start()                                                 # nvidia-smi needs ~0.5 s to start: begin before warm-up
    for i in range(W):
        det, outs = step(i)
    torch.cuda.synchronize()

    # ---- timed region: device-resident inputs --------------------------------------------------
    eng.set_profiling(True)
    eng.profile()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push('timed')                       # ncu --nvtx --nvtx-include "timed/" isolates these K steps
    ev0.record()
    for i in range(K):
        det, outs = step(i)
    ev1.record()
    torch.cuda.nvtx.range_pop()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(wall0, wall1)
    launches = _lib.launch_count() - l0
    prof = eng.profile()
    eng.set_profiling(False)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * BATCH * K / (ms / 1e3)

    # ---- Fast-NMS alone (decode + Fast-NMS + top-k), CUDA events -----------------------------------
    def time_detect(cls, box, coef, reps=10):
        for _ in range(3):
            detect_batched(cls, box, coef, anchors, cfg)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            detect_batched(cls, box, coef, anchors, cfg)
        b.record(); torch.cuda.synchronize()
        return 1e3 * a.elapsed_time(b) / (reps * cls.shape[0])            # us / img
    nms_us = {'network_output': time_detect(*outs[:3])}
    A = anchors.shape[0]
    for regime in ('stress', 'realistic'):
        c, b_, k_ = synth.head_outputs(7, A, 81, regime)
        rep = lambda a: torch.from_numpy(a).to(dev)[None].expand(BATCH, *a.shape).contiguous()
        nms_us[regime] = time_detect(rep(c), rep(b_), rep(k_))
    pp_bytes = A * 81 * 4 + A * 16 + A * 16 + 100 * (128 + 156)          # algorithmic bytes / img

    # ---- e2e: host buffers through the C ABI (yb_net_detect_host) --------------------------------------
    p = _lib.DetectParams(cfg.nms_score_thre, cfg.nms_iou_thre, cfg.top_k, cfg.max_detections, cfg.num_classes, 32, 0, float(IMG))
    host_np = [h.numpy() for h in host]
    for i in range(2):
        eng.detect_host(host_np[i & 1], p)
    for i in range(3):                                                   # both pipeline slots: buffers, graphs, events
        a_ = eng.submit_host(host_np[0], p); b_ = eng.submit_host(host_np[1], p)
        eng.collect_host(a_); eng.collect_host(b_)
    if world > 1:
        dist.barrier()
    # pipelined public API: submit batch i+1 (its H2D copy overlaps the compute of batch i), then collect batch i
    t0 = time.perf_counter()
    pending = eng.submit_host(host_np[0], p)
    for i in range(1, K):
        nxt = eng.submit_host(host_np[i & 1], p)
        r = eng.collect_host(pending)
        pending = nxt
    r = eng.collect_host(pending)
    e2e_s = time.perf_counter() - t0
    # the plain synchronous call, for reference
    t0 = time.perf_counter()
    for i in range(K):
        r = eng.detect_host(host_np[i & 1], p)
    e2e_sync_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s, e2e_sync_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s, e2e_sync_s = float(t[0].item()), float(t[1].item())
    D = cfg.max_detections
    e2e = {'value': world * BATCH * K / e2e_s, 'unit': 'img/s', 'h2d_bytes_per_step': BATCH * 3 * IMG * IMG * 4,
           'd2h_bytes_per_step': BATCH * (4 + D * (4 + 4 + 4 + 16 + 128)),
           'api': 'yb_net_submit_host / yb_net_collect_host (pinned host input, 2 batches in flight)',
           'synchronous_call_value': world * BATCH * K / e2e_sync_s}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---- parity gate printed with the number (oracle on a 1-image slice) --------------------------------
    with torch.no_grad():
        mine = [o[:1].cpu().numpy() for o in net(imgs[0][:1])]
    ref = [r.numpy() for r in ft.forward(host[0][:1], ft.synth_state_dict(ARCH, seed=0), ARCH)]
    parity = {n: float(np.abs(m - r).max()) for n, m, r in zip(('cls', 'box', 'coef', 'proto'), mine, ref)}
    o = pp.nms(mine[0][0], mine[1][0], eng.anchors())
    det1 = detect_batched(*[torch.from_numpy(m).to(dev) for m in mine[:3]], anchors, cfg)
    d = int(det1['count'][0])
    parity['fast_nms_indices_exact'] = bool(o is not None and d == len(o[0]) and
                                            np.array_equal(det1['cls'][0, :d].cpu().numpy(), o[0]) and
                                            np.array_equal(det1['anchor'][0, :d].cpu().numpy(), o[3]))

    # ---- the same step with bf16 operands (same kernels, same speed class; 8-bit mantissa) -----------------
    other = None
    if args.precision == 'fp16' and world == 1 and not os.environ.get('YB_BENCH_SKIP_BF16'):
        del net, eng
        torch.cuda.empty_cache()
        cfg2 = make_config(ARCH + '_coco', IMG)
        cfg2.precision, cfg2.max_batch = 'bf16', BATCH
        net2 = Yolact(cfg2)
        net2.load_state_dict(ft.synth_state_dict(ARCH, seed=0), strict=True)
        net2 = net2.to(dev).eval()

        def step2(i):
            with torch.no_grad():
                c_, b_, k_, _ = net2(imgs[i & 1])
            return detect_batched(c_, b_, k_, anchors, cfg2)
        for i in range(3):
            step2(i)
        torch.cuda.synchronize()
        a_, b2_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for i in range(10):
            step2(i)
        b2_.record(); torch.cuda.synchronize()
        with torch.no_grad():
            m2 = [o[:1].cpu().numpy() for o in net2(imgs[0][:1])]
        other = {'dtype': 'bf16', 'dtype_detail': 'bf16 operands, f32 accumulate', 'value_single_rank': BATCH * 10 / (a_.elapsed_time(b2_) / 1e3), 'steps': 10,
                 'parity_vs_fp32_oracle_max_abs_err': {n: float(np.abs(m - r).max()) for n, m, r in zip(('cls', 'box', 'coef', 'proto'), m2, ref)}}
        eng = net2.engine(BATCH)

    # ---- roofline of the dominant kernel ----
    tc = prof['conv_tc'] if prof['conv_tc']['launches'] else prof['conv_simt']
    dom = 'k_conv_tc' if prof['conv_tc']['launches'] else 'k_conv_simt'
    achieved = tc['flops'] / (tc['ms'] * 1e-3) / 1e12 if tc['ms'] else 0.0
    total_ms = sum(v['ms'] for v in prof.values())
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
    if os.path.exists(tpath) and dom == 'k_conv_tc':
        traffic = json.load(open(tpath)).get('dram_bytes_per_launch_avg')      # from the committed ncu --set full capture
    roofline = {'kernel': dom, 'bound': 'tensor', 'achieved': achieved, 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s',
                'frac': achieved / pk['tf_sustained'], 'traffic': traffic,
                'alg_bytes_per_launch': tc['bytes'] / max(1, tc['launches']), 'peak_source': pk['src'] + ' (sustained bf16 cuBLAS)',
                'launches_per_step': tc['launches'] / max(1, tc['forwards']),
                'share_of_forward': tc['ms'] / total_ms if total_ms else None,
                'alg_flops_per_launch': tc['flops'] / max(1, tc['launches']),
                'avg_launch_ms': tc['ms'] / max(1, tc['launches'])}
    breakdown = {k: {'ms_per_step': v['ms'] / max(1, v['forwards']), 'launches_per_step': v['launches'] / max(1, v['forwards']),
                     'alg_tflops': (v['flops'] / (v['ms'] * 1e-3) / 1e12) if v['ms'] and v['flops'] else None,
                     'alg_gbs': (v['bytes'] / (v['ms'] * 1e-3) / 1e9) if v['ms'] else None}
                 for k, v in prof.items() if v['launches']}
    pp_roof = {'kernel': 'k_filter_decode + k_class_fast_nms + k_final_topk', 'bound': 'hbm',
               'achieved': pp_bytes / (nms_us['stress'] * 1e-6) / 1e9, 'peak': pk['hbm'], 'unit': 'GB/s',
               'frac': pp_bytes / (nms_us['stress'] * 1e-6) / 1e9 / pk['hbm'], 'alg_bytes_per_img': pp_bytes, 'regime': 'stress'}

    cores = host_cores()
    skip_cpu = world > 1 or os.environ.get('YB_BENCH_SKIP_CPU')           # the CPU baseline is a rank-0, N=1 measurement
    cpu_v, cpu_s, cpu_nms_us = (0.0, 0.0, 0.0) if skip_cpu else cpu_reference_sample(4, 8, cores)
    line = {'metric': 'img/s', 'value': value, 'unit': 'img/s', 'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms / K,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'fp16': 'f16', 'bf16': 'bf16', 'fp32': 'f32'}[args.precision], 'dtype_detail': 'tensor-core operands in that type, f32 accumulation; fp32 inputs and outputs',
            'data': 'synthetic',
            'config': {'workload': f'{ARCH}_coco {IMG}x{IMG} bs={BATCH}/GPU eval forward + decode/Fast-NMS/top-k (BASELINE.json configs[2] '
                                   f'at one GPU per {BATCH} images), random-init weights', 'global_batch': BATCH * world,
                       'parallelism': f'dp{world} (image shards, NCCL all-gather of detection records)' if world > 1 else 'single GPU',
                       'l2_policy': 'inputs (232 MB/step, two alternating batches) larger than the 126 MB L2'},
            'tensor_frac_of_peak': value * GFLOP_PER_IMG * 1e9 / (world * pk['tf_sustained'] * 1e12),
            'gflop_per_img': GFLOP_PER_IMG,
            'fast_nms_us_per_img': nms_us, 'roofline': roofline, 'roofline_postprocess': pp_roof, 'kernel_breakdown': breakdown,
            'cpu_baseline': None if skip_cpu else {'value': cpu_v, 'unit': 'img/s', 'cores': cores, 'kind': 'port',
                                                   'sample': f'8 reps x 4 images ({cpu_s:.1f} s): torch fp32 CPU forward + numpy nms()',
                                                   'fast_nms_us_per_img': cpu_nms_us},
            'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks, 'parity_vs_fp32_oracle_max_abs_err': parity,
            'bf16_arm': other}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--precision', default=os.environ.get('YOLACT_B200_PRECISION', 'fp16'), choices=['fp16', 'bf16', 'fp32'])
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--arch', default=ARCH, choices=['res101', 'res50', 'swin_tiny'], help='default: the BASELINE metric config')
    ap.add_argument('--img', type=int, default=IMG)
    ap.add_argument('--batch', type=int, default=BATCH, help='images per GPU')
    args = ap.parse_args()
    ARCH, IMG, BATCH = args.arch, args.img, args.batch
    GFLOP_PER_IMG = GFLOPS.get((ARCH, IMG), float('nan'))
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)
