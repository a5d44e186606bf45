"""Time the native training step alone (bench.py's training leg): python tools/bench_train.py [arch] [size] [per_gpu] [steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
arch = sys.argv[1] if len(sys.argv) > 1 else 'res101'
S = int(sys.argv[2]) if len(sys.argv) > 2 else 550
per = int(sys.argv[3]) if len(sys.argv) > 3 else 2
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 50
print(json.dumps(bench.training_leg(arch, S, per, steps, torch.device('cuda:0'), 0, 1)))
