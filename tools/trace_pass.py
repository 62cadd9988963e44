"""Diagnostic: host-side timeline of one learner pass (where does the wall time between kernels go?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
import ppo_exploration_b200 as ppx
from ppo_exploration_b200 import _lib as L
from ppo_exploration_b200 import buffer as BUF

dev = torch.device("cuda", 0)
np.random.seed(0); torch.manual_seed(0)
cfg = B.CONFIGS["C2"]
env = ppx.SyntheticVecEnv(cfg["N"], cfg["D"], ppx.Box((2,)), seed=0)
m = ppx.PPO(env=env, nstep=cfg["T"], batch_size=cfg["batch"], hidden_size=cfg["hidden"], sim_hash=True, hash_bits=cfg["hash_bits"], device=dev, **cfg["hp"])
ro = m.rollout
host = B.synth_rollout(cfg, 100)
ro.load_rollout(**{k: v for k, v in host.items() if k in B.ROLLOUT_FIELDS})
lv = torch.as_tensor(host["last_value"]).to(dev); dn = torch.as_tensor(host["masks"][-1].copy()).to(dev)
raw = ro.rewards.clone()
marks = []
orig_next = BUF.HostRngStream.next
def next_(self):
    t0 = time.perf_counter(); v = orig_next(self); marks.append(("rng.next", t0, time.perf_counter())); return v
BUF.HostRngStream.next = next_
def one():
    ro.rewards.copy_(raw)
    t0 = time.perf_counter()
    ro.sim_hash(ro.observations, ro.rewards)
    t1 = time.perf_counter()
    ro.compute_returns_and_advantages(lv, dn)
    t2 = time.perf_counter()
    m.train()
    t3 = time.perf_counter()
    return t0, t1, t2, t3
for _ in range(6): one()
torch.cuda.synchronize()
marks.clear()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); t = one(); e1.record(); torch.cuda.synchronize()
print("gpu ms %.2f | host: simhash %.2f gae %.2f train %.2f" % (e0.elapsed_time(e1), (t[1]-t[0])*1e3, (t[2]-t[1])*1e3, (t[3]-t[2])*1e3))
print("rng.next waits (ms):", ["%.2f@%.2f" % ((b-a)*1e3, (a-t[2])*1e3) for _, a, b in marks])
# GPU-only floor: kernels back to back without host gaps = replay one epoch's graphs 10x
torch.cuda.synchronize()
gs = [v[0] for k, v in m._graphs.items() if isinstance(v, tuple)]
print("graphs", len(gs))
e0.record()
for _ in range(max(1, m.n_epochs // max(1, len(gs)))):      # graphs are per epoch; stay inside one pass's staging buffers
    for g in gs: g.replay()
e1.record(); torch.cuda.synchronize()
print("epoch graphs of one pass back-to-back: %.2f ms" % (e0.elapsed_time(e1)))
