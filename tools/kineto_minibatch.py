"""Diagnostic (1 GPU, or torchrun N >= 2): CUPTI timeline of graph-mode passes; prints, for a late minibatch of the pass
(steady state: no staging traffic on the copy stream), every kernel of the main stream with its duration and the idle gap
before it, plus per-kernel medians over the whole trace."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity

import bench as B
import ppo_exploration_b200 as ppx

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
p = B.PpxPass(sys.argv[1] if len(sys.argv) > 1 else "C2", torch, ppx, dev, rank, world)
if len(sys.argv) > 2:
    p.m.shard_shuffle = sys.argv[2]
    if sys.argv[2] == "local":
        np.random.seed(1000 + rank)
for _ in range(6):
    p.step_resident()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        p.step_resident()
    torch.cuda.synchronize()
path = f"gpurun_out/kineto_mb_{rank}.json"
prof.export_chrome_trace(path)
if rank == 0:
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    streams = {}
    for e in ev:
        streams.setdefault(e["args"].get("stream"), []).append(e)
    main = max(streams.items(), key=lambda kv: len(kv[1]))[0]
    mk = streams[main]
    short = lambda n: n.replace("ppx::", "").replace("(anonymous namespace)::", "").split("(")[0][-46:]
    gi = [i for i, e in enumerate(mk) if "gather_kernel" in e["name"]]
    a, b = gi[35], gi[36]
    print(f"minibatch 35 of pass 1 on the main stream ({world} GPU(s), {p.m.shard_shuffle if world > 1 else 'single'}): kernel, dur us, gap before us")
    tot = 0.0
    for i in range(a, b):
        e = mk[i]
        gap = e["ts"] - (mk[i - 1]["ts"] + mk[i - 1]["dur"])
        tot += e["dur"] + gap
        print(f"  {short(e['name']):48s} {e['dur']:7.1f} {gap:6.1f}")
    print(f"  total {tot:.1f} us")
    med = {}
    for e in mk:
        med.setdefault(short(e["name"]), []).append(e["dur"])
    print("medians:", {k: (round(float(np.median(v)), 1), len(v)) for k, v in med.items() if len(v) >= 40})
os.remove(path)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
