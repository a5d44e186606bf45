"""How fast is the reference CPU path on this box, and with how many threads? (bench sizing)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth, forward_torch as ft
print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)), 'torch default threads', torch.get_num_threads())
os.system("grep -m1 'model name' /proc/cpuinfo; cat /sys/fs/cgroup/cpu.max 2>/dev/null")
sd = ft.synth_state_dict('res101'); img = torch.from_numpy(synth.image_batch(1, 1, 550))
for th in (8, 16, 32, 64, 128):
    torch.set_num_threads(th)
    ft.forward(img, sd, 'res101')
    t0 = time.perf_counter(); ft.forward(img, sd, 'res101'); dt = time.perf_counter() - t0
    print(f'threads {th}: {dt:.2f} s/img', flush=True)
