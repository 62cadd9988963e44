"""Diagnostic: CUPTI (torch.profiler) kernel timeline of two graph-mode passes; prints the largest idle gaps / longest
kernels on the main stream and what the other streams ran meanwhile."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity

import bench as B
import ppo_exploration_b200 as ppx

dev = torch.device("cuda", 0)
p = B.PpxPass(sys.argv[1] if len(sys.argv) > 1 else "C2", torch, ppx, dev, 0, 1)
for _ in range(6):
    p.step_resident()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        p.step_resident()
    torch.cuda.synchronize()
path = "gpurun_out/kineto_pass.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
streams = {}
for e in ev:
    streams.setdefault(e["args"].get("stream"), []).append(e)
main = max(streams.items(), key=lambda kv: len(kv[1]))[0]
print("streams:", {k: len(v) for k, v in streams.items()}, "main:", main)
t0 = ev[0]["ts"]
mk = streams[main]
rows = []
for a, b in zip(mk, mk[1:]):
    gap = b["ts"] - (a["ts"] + a["dur"])
    rows.append((gap, a, b))
rows.sort(key=lambda r: -r[0])
print("largest gaps on the main stream (us): gap | after kernel (dur) -> next kernel | other-stream activity inside the gap")
for gap, a, b in rows[:8]:
    lo, hi = a["ts"] + a["dur"], b["ts"]
    other = [(e["name"][:40], round(e["ts"] - lo, 1), round(e["dur"], 1)) for s, v in streams.items() if s != main for e in v
             if e["ts"] < hi and e["ts"] + e["dur"] > lo]
    print(round(gap, 1), "| t=%.1f" % (lo - t0), a["name"][:40], round(a["dur"], 1), "->", b["name"][:40], "|", other[:8])
med = {}
for e in mk:
    med.setdefault(e["name"], []).append(e["dur"])
med = {k: float(np.median(v)) for k, v in med.items()}
slow = sorted(((e["dur"] - med[e["name"]], e) for e in mk), key=lambda r: -r[0])[:8]
print("kernels furthest above their median (us over | name dur median | other-stream activity meanwhile)")
for over, e in slow:
    lo, hi = e["ts"], e["ts"] + e["dur"]
    other = [(x["name"][:40], round(x["ts"] - lo, 1), round(x["dur"], 1)) for s, v in streams.items() if s != main for x in v
             if x["ts"] < hi and x["ts"] + x["dur"] > lo]
    print(round(over, 1), "| t=%.1f" % (lo - t0), e["name"][:50], round(e["dur"], 1), round(med[e["name"]], 1), "|", other[:8])
allev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X"]
cps = [e for e in ev if e.get("cat") == "gpu_memcpy" and "HtoD" in e["name"]]
print("H2D copies: t(us) dur(us) bytes")
print([(round(e["ts"] - t0), round(e["dur"], 1), e["args"].get("bytes")) for e in cps])
for c in [e for e in cps if e["dur"] > 150][:2]:
    lo, hi = c["ts"] - 300, c["ts"] + c["dur"] + 50
    print("--- host activity around the slow H2D at t=%.0f (name, t-rel, dur, tid)" % (c["ts"] - t0))
    for e in sorted(allev, key=lambda e: e["ts"]):
        if e.get("cat") in ("cuda_runtime", "cuda_driver", "cpu_op", "user_annotation") and e["ts"] < hi and e["ts"] + e["dur"] > lo and e["dur"] > 15:
            print("   ", e["name"][:50], round(e["ts"] - c["ts"], 1), round(e["dur"], 1), e.get("tid"))
os.remove(path)
