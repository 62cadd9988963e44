import sys, warnings
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
import ppo_exploration_b200 as ppx
dev = torch.device('cuda', 0)
for name in ('C2', 'C1', 'C3', 'C4'):
    p = B.PpxPass(name, torch, ppx, dev, 0, 1)
    out = []
    for i in range(10):
        p.step_resident()
        l = p.m.last_losses
        out.append((float(np.abs(l[:, :5]).max()), bool(np.isfinite(l).all())))
    print(name, out[0], out[4], out[9], 'param finite:', bool(torch.isfinite(p.m.policy.bank.flat).all()))
    del p
    torch.cuda.empty_cache()
