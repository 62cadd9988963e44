"""Diagnostic (torchrun, N >= 2): where does a sharded pass spend its time?  For each shard_shuffle mode: the phases of
one pass (L2 flush, bonus, GAE, train) bracketed with CUDA events in graph mode, then one eager pass with EVERY libppx
launch bracketed (device time per entry point, which for the p2p kernels includes the wait for the slowest peer).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/trace_sharded.py [C2]
"""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench as B
import ppo_exploration_b200 as ppx
from ppo_exploration_b200 import _lib as L


class AllOps:
    """CUDA-event brackets around every libppx call that takes a stream (last argument)."""

    def __init__(self):
        self.rec, self.orig = [], L.call

    def __enter__(self):
        def timed(name, *args):
            if name.startswith("ppx_np_") or "workspace" in name or threading.current_thread() is not threading.main_thread():
                return self.orig(name, *args)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = self.orig(name, *args)
            e.record()
            self.rec.append((name, s, e))
            return rc
        L.call = timed
        for mod in (ppx.models, ppx.algorithms, ppx.buffer, ppx.util, ppx.dist):
            if hasattr(mod, "L"):
                mod.L.call = timed
        return self

    def __exit__(self, *a):
        L.call = self.orig

    def table(self):
        torch.cuda.synchronize()
        agg = {}
        for name, s, e in self.rec:
            a = agg.setdefault(name, [0.0, 0])
            a[0] += s.elapsed_time(e)
            a[1] += 1
        return sorted(((k, round(v[0], 3), v[1]) for k, v in agg.items()), key=lambda x: -x[1])


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    barrier = dist.barrier if world > 1 else (lambda: None)
    out = {}
    for mode in (("local", "global") if world > 1 else ("single",)):
        p = B.PpxPass(name, torch, ppx, dev, rank, world)
        m = p.m
        if world > 1:
            m.shard_shuffle = mode
        for kv in os.environ.get("PPX_TRACE_SET", "").split(","):      # e.g. PPX_TRACE_SET=speculative_shuffle=0
            if kv:
                k, v = kv.split("=")
                setattr(m, k, type(getattr(m, k))(int(v)))
        np.random.seed(0 if mode == "global" else 1000 + rank)
        for _ in range(5):
            p.step_resident()
        torch.cuda.synchronize()
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        acc = np.zeros(5)
        w0 = m.rng_wait_s
        t0 = time.perf_counter()
        K = 5
        for _ in range(K):
            ev[0].record(); p.flush.zero_()
            ev[1].record(); p.ro.rewards.copy_(p.raw_rewards); p.ro.sim_hash_sharded(p.ro.observations, p.ro.rewards)
            ev[2].record(); p.ro.compute_returns_and_advantages(p.last_value_dev, p.dones_dev)
            ev[3].record(); m.train()
            ev[4].record()
            torch.cuda.synchronize()
            acc[:4] += [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
        wall = (time.perf_counter() - t0) / K * 1e3
        rec = {"flush_ms": acc[0] / K, "simhash_ms": acc[1] / K, "gae_ms": acc[2] / K, "train_ms": acc[3] / K,
               "wall_ms_per_pass_with_sync": wall, "rng_wait_ms": (m.rng_wait_s - w0) / K * 1e3}
        # per-minibatch device timeline of one pass: an event before every graph replay (+ one at the end of train())
        marks, host_t = [], []
        orig_gc = m._graph_call
        def gc(key, fn):
            e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e); host_t.append(time.perf_counter())
            return orig_gc(key, fn)
        m._graph_call = gc
        p.step_resident(); torch.cuda.synchronize(); barrier()
        rec["minibatch_gpu_us"], rec["minibatch_host_issue_us"] = [], []
        for _ in range(4):
            marks.clear(); host_t.clear()
            h0 = time.perf_counter()
            p.step_resident()
            e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e); torch.cuda.synchronize()
            rec["minibatch_gpu_us"].append([round(marks[i].elapsed_time(marks[i + 1]) * 1e3) for i in range(len(marks) - 1)])
            rec["minibatch_host_issue_us"].append([round((t - h0) * 1e6) for t in host_t][::4])
        m._graph_call = orig_gc
        if m._spec is not None:                                 # the stream that will serve the next pass + the one just used
            rec["spec_stream_profile_us"] = [[round(x * 1e6) for x in row] for row in m._spec[0].profile]
        rec["last_stream_profile_us"] = [[round(x * 1e6) for x in row] for row in getattr(m, "_last_rng", m._spec[0]).profile]
        # back-to-back replay of the captured minibatch graphs (no host work between them, same order on every rank)
        gs = [(k, v[0]) for k, v in m._graphs.items() if isinstance(v, tuple)]
        m._cursor.zero_()
        barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        # a graph covers one EPOCH (its minibatches find their index slices through the device step cursor): never replay more
        # epochs than the staging buffers of a pass hold
        reps = max(1, m.n_epochs // max(1, len(gs)))
        for _ in range(reps):
            for _, g in gs:
                g.replay()
        e1.record(); torch.cuda.synchronize()
        rec["graphs"] = [str(k) for k, _ in gs]
        rec["graph_replays_back_to_back_ms"] = e0.elapsed_time(e1)
        rec["graph_replays"] = reps * len(gs)
        m._cursor.zero_()
        # eager pass, every launch bracketed
        m.use_cuda_graph = False
        p.step_resident()
        barrier(); torch.cuda.synchronize()
        with AllOps() as ops:
            p.step_resident()
        rec["eager_ops_ms"] = ops.table()
        # anomalies of the eager timeline: launches that took > 1.6x their entry point's median, and idle gaps > 50 us
        med = {}
        for nm, s_, e_ in ops.rec:
            med.setdefault(nm, []).append(s_.elapsed_time(e_))
        med = {k: float(np.median(v)) for k, v in med.items()}
        first = ops.rec[0][1]
        odd = []
        for i, (nm, s_, e_) in enumerate(ops.rec):
            d = s_.elapsed_time(e_)
            gap = ops.rec[i - 1][2].elapsed_time(s_) if i else 0.0
            if d > 1.6 * med[nm] + 0.01 or gap > 0.05:
                odd.append([i, nm, round(first.elapsed_time(s_), 3), round(d, 3), round(med[nm], 3), round(gap, 3)])
        rec["eager_anomalies_idx_name_startms_durms_medianms_gapms"] = odd
        rec["eager_calls"] = len(ops.rec)
        m.use_cuda_graph = True
        out[mode] = rec
        del p, m
        torch.cuda.empty_cache()
    for r in range(world):
        barrier()
        if r == rank and r < 2:
            from ppo_exploration_b200.buffer import _partner_pool
            out["partner_pool"] = {str(n): len(v) for n, v in _partner_pool.slots.items()}
            print(json.dumps({"rank": rank, "world": world, "config": name, **out}))
            sys.stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
