"""after_nms (mask assembly) timing, replayed from a CUDA graph: python tools/bench_mask.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pk = bench.peaks()
print(bench.mask_stage_leg(torch.device('cuda:0'), pk))
