"""Small driver for profiling the fused tensor-core MLP kernels (mlp_tc.cu) at the C2 minibatch shape:
python tools/mlp_tc_prof.py [iters]   (ncu: -k regex:mlp3_tc)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_exploration_b200 as ppx  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4
M, D = 131072, 8
env = ppx.SyntheticVecEnv(4, D, ppx.Box((2,)), seed=0)
pol = ppx.models.Policy(env, 64, intrinsic_model=False, device="cuda")
assert pol.mlp._fused_args()["tc"]
x = torch.randn(M, D, device="cuda")
outs = pol.forward_raw(x)
d = [torch.randn_like(o) / M for o in outs]
for _ in range(iters):
    pol.forward_raw(x)
    pol.mlp.backward(d)
torch.cuda.synchronize()
print("ok")
