"""Measurement: the Nature-CNN trunk (Conv 8/4 -> 4/2 -> 3/1 -> FC 3136->512, ReLU) forward + backward on N frames of
4x84x84, CUDA events after warm-up; per-entry-point device time from one eager pass."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ppo_exploration_b200 import _lib as L
from ppo_exploration_b200 import models as PM

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda")
convs, fc = [(32, 8, 4, "relu"), (64, 4, 2, "relu"), (64, 3, 1, "relu")], (512, "relu")
bank = PM.ParamBank(PM.ConvTrunk.specs("trunk", (4, 84, 84), convs, fc), dev)
torch.manual_seed(0)
bank.flat.copy_(0.05 * torch.randn(bank.size, device=dev))
trunk = PM.ConvTrunk(bank, "trunk", (4, 84, 84), convs, fc, PM._Scratch(dev))
trunk.enable_tc()
x = torch.rand(N, 4 * 84 * 84, device=dev)
d = torch.randn(N, 512, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def step():
    feat, saved = trunk.forward(x)
    trunk.backward(saved, d, need_dx=False)


for _ in range(3):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    flush.zero_()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    feat, saved = trunk.forward(x)
    e1.record()
    trunk.backward(saved, d, need_dx=False)
    e2.record()
    torch.cuda.synchronize()
    ts.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
fwd, bwd = (sum(t[i] for t in ts) / len(ts) for i in (0, 1))
flops_fwd = 2.0 * N * (400 * 256 * 32 + 81 * 512 * 64 + 49 * 576 * 64 + 3136 * 512)
print(f"N={N}: forward {fwd:.3f} ms, backward (no dX of the frames) {bwd:.3f} ms -> {N / (fwd + bwd) * 1e3:.0f} frames/s; "
      f"forward {flops_fwd / fwd / 1e9:.1f} TFLOP/s fp32-equivalent, backward {2 * flops_fwd / bwd / 1e9:.1f}")
rec, orig = [], L.call


def timed(name, *a):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); rc = orig(name, *a); e.record(); rec.append((name, s, e)); return rc


L.call = PM.L.call = timed
step()
torch.cuda.synchronize()
L.call = PM.L.call = orig
agg = {}
for name, s, e in rec:
    a = agg.setdefault(name, [0.0, 0]); a[0] += s.elapsed_time(e); a[1] += 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"   {k:28s} {v[0]:8.3f} ms  x{v[1]}")
