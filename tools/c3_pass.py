"""One C3 learner pass (PPO_RND on Atari-shaped flat frames) for profiling: python tools/c3_pass.py [n_passes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_kernels as BK
rows = []
BK.timed = lambda fn, iters, warm=3: ([fn() for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3)], torch.cuda.synchronize(), (1.0, 1.0))[2]
BK.bench_configs.__globals__["timed"] = BK.timed
BK.bench_configs(rows, 4)
