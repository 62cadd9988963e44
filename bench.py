"""Benchmark of the learner hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

Workload (BASELINE.json configs[1], "C2"): PPO + SimHash count bonus, Swimmer-shaped obs (dim 8),
Box(2) actions, 64-bit codes, 2048 envs x 256 steps per GPU, reference hyper-parameters `swimmer_ppo`
(hyperparameters.py:7-8: hidden 64, lr 3e-4, gamma .999, lambda .95, 10 epochs, clip .2, vf 1,
max_grad_norm 5) with the minibatch count pinned to 4 per epoch.

One "step" = one learner pass over a rollout: SimHash bonus over all T*N transitions -> GAE ->
train() (10 epochs x 4 minibatches: shuffle-gather, MLP forward, fused PPO loss fwd+bwd, MLP backward,
clip+Adam).  value = transitions (T*N per GPU x N GPUs) per second.

  python bench.py [--gpus N] [--steps K] [--warmup W]           ppx arm (this repo's CUDA path)
  python bench.py --impl reference ...                          CPU arm: the oracle port of the reference
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T, N, D, A, K_BITS = 256, 2048, 8, 2, 64
HP = dict(lr=3e-4, gamma=0.999, gae_lam=0.95, vf_coef=1, max_grad_norm=5, n_epochs=10, clip_range=0.2, ent_coef=0.0)
HIDDEN = 64
N_MINIBATCH = 4
# DRAM bytes per launch of the fused MLP kernels at the C2 minibatch, from the committed ncu captures
NCU_TRAFFIC = {"ppx_mlp3_bwd": 144.0e6, "ppx_mlp3_fwd": 81.9e6, "ppx_mlp3_tc_bwd": 143.3e6, "ppx_mlp3_tc_fwd": 80.1e6}
NCU_TRAFFIC_SRC = ("ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch (profiles/ncu_mlp_tc_r01d.md; "
                   "SIMT pair: profiles/ncu_mlp3_r01b.md)")
WORKLOAD = ("C2: PPO+SimHash, obs 8, Box(2), k=64, 2048 envs x 256 steps per GPU, swimmer_ppo hparams, "
            "4 minibatches/epoch x 10 epochs")


def synth_rollout(seed, n_envs=N, t=T):
    rs = np.random.RandomState(seed)
    return dict(observations=rs.randn(t, n_envs, D).astype(np.float32), actions=rs.randn(t, n_envs, A),
                rewards=rs.randn(t, n_envs).astype(np.float32), values=rs.randn(t, n_envs).astype(np.float32),
                masks=(rs.rand(t, n_envs) < 0.02).astype(np.uint8),
                action_log_probs=(-1.4 + 0.3 * rs.randn(t, n_envs, A)).astype(np.float32),
                last_value=rs.randn(n_envs).astype(np.float32))


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference (numpy + torch CPU), bounded sample, scaled linearly
# ---------------------------------------------------------------------------------------------
def cpu_reference_pass(seed=0, hash_envs=128, hash_steps=64, train_minibatches=2):
    """Times the reference's algorithm (oracle port) on a bounded sample of C2 and scales to the full
    pass.  Returns (transitions_per_s, detail dict)."""
    import torch
    from oracle import rollout as OR
    from oracle import learner as OL
    np.random.seed(seed); torch.manual_seed(seed)
    ro = synth_rollout(seed)
    A_mat = np.random.randn(K_BITS, D)
    # SimHash: the reference walks every obs in python (buffer.py:195-199)
    tab = OR.CountTable(0.1)
    t0 = time.perf_counter()
    for t in range(hash_steps):
        tab.update(A_mat, ro["observations"][t, :hash_envs], ro["rewards"][t, :hash_envs])
    t_hash = (time.perf_counter() - t0) * (T * N) / (hash_steps * hash_envs)
    # GAE at full size
    t0 = time.perf_counter()
    adv, ret = OR.gae(ro["rewards"], ro["values"], ro["masks"].astype(np.int64), ro["last_value"], ro["masks"][-1],
                      HP["gamma"], HP["gae_lam"])
    t_gae = time.perf_counter() - t0
    # train(): `train_minibatches` optimiser steps at the full minibatch size, scaled to 40
    p = OL.make_policy_params(D, A, HIDDEN)
    opt = torch.optim.Adam(OL._policy_param_list(p), lr=HP["lr"])
    buf = dict(ro, advantages=adv, returns=ret)
    hp = dict(HP, batch_size=T * N // N_MINIBATCH)
    t0 = time.perf_counter()
    OL.ppo_train(p, opt, buf, hp, discrete=False, max_steps=train_minibatches)
    t_train = (time.perf_counter() - t0) * (HP["n_epochs"] * N_MINIBATCH) / train_minibatches
    total = t_hash + t_gae + t_train
    detail = dict(sim_hash_s=t_hash, gae_s=t_gae, train_s=t_train,
                  sample=(f"sim_hash on {hash_envs} envs x {hash_steps} steps, GAE full size, "
                          f"{train_minibatches} of 40 optimiser steps at B=131072; each scaled linearly to one pass"))
    return (T * N) / total, detail


def host_threads():
    import torch
    return dict(cpu_count=os.cpu_count(), affinity=len(os.sched_getaffinity(0)), torch_threads=torch.get_num_threads())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    vals, det = [], None
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_pass(hash_envs=32, hash_steps=8, train_minibatches=1)
    t_all = time.perf_counter()
    for s in range(args.steps):
        v, det = cpu_reference_pass(seed=s, hash_envs=512, hash_steps=64, train_minibatches=8)
        vals.append(v)
    wall = time.perf_counter() - t_all
    v = float(np.mean(vals))
    th = host_threads()
    line = {"impl": "reference", "metric": "transitions/s through GAE+bonus+PPO update", "value": v,
            "unit": "transitions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * (T * N) / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": v, "unit": "transitions/s", "cores": th["torch_threads"], "kind": "port",
                             "sample": det["sample"], "host": th, "measured_wall_s": wall},
            "e2e": {"value": v, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# ppx arm
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.index = None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class OpTimer:
    """CUDA-event brackets around selected C-ABI calls on the launching stream (roofline measurement)."""
    SHAPE_ARGS = {"ppx_linear_fwd": (4, 5, 6, 10), "ppx_linear_bwd_data": (3, 4, 5, 11), "ppx_linear_bwd_weight": (4, 5, 6, 10),
                  "ppx_mlp3_fwd": (2, 3, 4, 5), "ppx_mlp3_bwd": (2, 3, 4, 5),
                  "ppx_mlp3_tc_fwd": (2, 3, 4, 5), "ppx_mlp3_tc_bwd": (2, 3, 4, 5)}

    def __init__(self, L, torch):
        self.L, self.torch, self.rec, self.orig = L, torch, [], L.call

    def __enter__(self):
        def timed(name, *args):
            if name in self.SHAPE_ARGS:
                s, e = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
                s.record()
                rc = self.orig(name, *args)
                e.record()
                self.rec.append((name, tuple(args[i] for i in self.SHAPE_ARGS[name]), s, e))
                return rc
            return self.orig(name, *args)
        self.L.call = timed
        for mod in self._mods():
            mod.L.call = timed
        return self

    def _mods(self):
        import ppo_exploration_b200 as ppx
        return [ppx.models, ppx.algorithms, ppx.buffer, ppx.util]

    def __exit__(self, *a):
        self.L.call = self.orig

    def summary(self):
        self.torch.cuda.synchronize()
        agg = {}
        for name, shape, s, e in self.rec:
            k = (name, shape)
            ms = s.elapsed_time(e)
            a = agg.setdefault(k, [0.0, 0])
            a[0] += ms; a[1] += 1
        return agg


def es_step_bench(torch, ppx, dev, world, rank, steps=20, warmup=5):
    """ES-NSRA step at C5 (P=10 000 members, MLP 8-64-64-2 -> D=4736, noise table 2^28 f32, K=10, archive 10 000):
    sample offsets -> theta + sigma*eps for this rank's P/W members -> all-gather fitness -> novelty k-NN ->
    replicated update.  Strong scaling (the population is fixed).  Returns perturbations/s and ms/step."""
    import torch.distributed as dist
    P, M = 10000 - 10000 % world, 10000
    np.random.seed(0)
    es = ppx.EvolutionStrategy(obs_dim=8, n_actions=2, hidden_sizes=(64, 64), population_size=P, sigma=0.1,
                               learning_rate=0.01, decay=0.9995, novelty_param=0.5, device=dev,
                               noise_table_size=1 << 28, noise_seed=0)
    es.noise_table()
    g = torch.Generator(device=dev).manual_seed(1)
    archive = torch.randn(M, 2, dtype=torch.float64, device=dev, generator=g)
    queries = torch.randn(2, 2, dtype=torch.float64, device=dev, generator=g)
    fit_local = torch.randn(P // world, dtype=torch.float64, device=dev, generator=g)

    def step():
        pop = es._get_population()                          # identical offsets on every rank (shared seed)
        w = es.perturb_all(es.shard_population(pop))        # [P/W, D] f32: what the evaluators consume
        r_all = es.gather_fitness(fit_local)
        _, nov = es.novelty_batch(archive, queries)
        es._update_weights(r_all, pop, novelty=nov[0:1])      # novelty stays on the device
        return w

    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item()) / steps
    return {"metric": "ES perturbations/s", "value": P / (ms / 1e3), "unit": "perturbations/s", "ms_per_step": ms,
            "scaling": "strong", "config": {"workload": "C5: ES-NSRA, P=%d, MLP 8-64-64-2 (D=4736), noise table 2^28 f32, "
                                                        "K=10, archive 10000, z-score shaping" % P,
                                            "alg_bytes_per_perturbation": 56832}}


def run_ppx(args):
    import torch
    import torch.distributed as dist
    import ppo_exploration_b200 as ppx
    from ppo_exploration_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    np.random.seed(rank); torch.manual_seed(0)          # per-rank shuffle stream, replicated weights
    env = ppx.SyntheticVecEnv(N, D, ppx.Box((A,)), seed=rank)
    B = T * N // N_MINIBATCH
    m = ppx.PPO(env=env, nstep=T, batch_size=B, hidden_size=HIDDEN, sim_hash=True, hash_bits=K_BITS, device=dev, **HP)
    m.shard_shuffle = "local"                           # N>1: every rank shuffles its own rollout
    ro = m.rollout
    host = synth_rollout(100 + rank)
    pinned = {k: torch.as_tensor(v).pin_memory() for k, v in host.items()}
    dones = pinned["masks"][-1].clone().pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in pinned.values()) + dones.numel()
    # per epoch: the int32 Fisher-Yates partner list when the swaps run on the device (default), else the int64 permutation
    perm_bytes = HP["n_epochs"] * T * N * world * (4 if m._device_apply() else 8)

    def load():
        ro.load_rollout(**{k: v for k, v in pinned.items() if k != "last_value"})

    def bonus_and_gae():
        if world > 1:
            codes = ro.sim_hash_codes(ro.observations).view(T, N)
            allc = ppx.dist.interleave_env_shards(ppx.dist.all_gather_cat(codes)).reshape(-1)
            counts = ro.count_table.update_codes(allc).view(T, world, N)[:, rank].contiguous()
            L.call("ppx_simhash_bonus", counts.data_ptr(), T * N, 0.1, ro.rewards.data_ptr(), 0, L.stream())
        else:
            ro.sim_hash(ro.observations, ro.rewards)
        ro.compute_returns_and_advantages(last_value_dev, dones_dev)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    last_value_dev = pinned["last_value"].to(dev)
    dones_dev = dones.to(dev)
    load()
    raw_rewards = ro.rewards.clone()

    def step_resident():
        flush.zero_()                                   # L2 flush (256 MiB > 126 MB L2), inside the timed region
        ro.rewards.copy_(raw_rewards)                   # the bonus is applied in place; restore the raw rewards
        bonus_and_gae()
        m.train()                                       # ends with the D2H read of the loss log

    def step_e2e():
        flush.zero_()
        load()                                          # H2D of the whole rollout from pinned host memory
        last_value_dev.copy_(pinned["last_value"], non_blocking=True)
        dones_dev.copy_(dones, non_blocking=True)
        bonus_and_gae()
        m.train()

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item())

    if args.profile:
        for _ in range(args.warmup):
            step_resident()
        ms = timed(step_resident, args.steps)
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms / args.steps, "launches": L.launch_count()}))
        return
    clocks = ClockSampler(local_rank)                   # sampled over warm-up + timed region (same load; nvidia-smi
    clocks.start()                                      # needs a few hundred ms to produce its first line)
    for _ in range(max(args.warmup, 5)):
        step_resident()
    torch.cuda.synchronize()
    l0 = L.launch_count()
    ms = timed(step_resident, args.steps)
    launches = L.launch_count() - l0
    clk = clocks.stop()
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    # per-op device timing of the dense-layer calls over one more pass (CUDA events on the launching stream)
    m.use_cuda_graph = False                            # per-op events need the individual launches, not a graph replay
    step_resident()
    with OpTimer(L, torch) as ot:
        step_resident()
    agg = ot.summary()
    m.use_cuda_graph = True

    trans = T * N * world
    value = trans * args.steps / (ms / 1e3)
    e2e = trans * args.steps / (ms_e2e / 1e3)
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    top = max(agg.items(), key=lambda kv: kv[1][0])
    (op, (M_, K_, N_, b_)), (tot_ms, cnt) = top
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    if op in ("ppx_mlp3_fwd", "ppx_mlp3_tc_fwd"):   # shape = (M, D, H, G): 2 flops per MAC of the three layers of every net
        flops = 2.0 * M_ * (b_ * (K_ * N_ + N_ * N_) + N_ * (A + b_ - 1))
    elif op in ("ppx_mlp3_bwd", "ppx_mlp3_tc_bwd"):  # dgrad (layers 3,2) + wgrad (layers 3,2,1)
        flops = 2.0 * M_ * (b_ * (K_ * N_ + 2 * N_ * N_) + 2 * N_ * (A + b_ - 1))
    else:
        flops = 2.0 * M_ * K_ * N_ * b_
    # algorithmic bytes of one fused-MLP launch: X read + the two saved activations of every net
    # written (forward) or read (backward) + the head outputs / their gradients
    mlp_bytes = 4.0 * M_ * (K_ + 2 * b_ * N_ + (A + b_ - 1))
    achieved = flops / (tot_ms / cnt / 1e3) / 1e12
    ops = sorted(((f"{k[0]}{list(k[1])}", round(v[0], 3), v[1]) for k, v in agg.items()), key=lambda x: -x[1])[:8]
    ms_launch = tot_ms / cnt
    if op.startswith("ppx_mlp3_tc"):
        # the tensor-core pair: the 64x64 GEMMs are off the FMA pipe, what remains is elementwise work + HBM streams
        # of the saved activations -> HBM roofline (the tensor pipe needs ~3 x flops / 1.1 PFLOP/s tf32 = a few us)
        gbs = mlp_bytes / (ms_launch / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": f"{op} M={M_} D={K_} H={N_} G={b_}", "achieved": gbs, "peak": hbm_peak,
                    "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": NCU_TRAFFIC.get(op),
                    "traffic_source": NCU_TRAFFIC_SRC if op in NCU_TRAFFIC else None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback",
                    "alg_bytes_per_launch": mlp_bytes, "fp32_equiv_tflops": achieved,
                    "tf32_mma_tflops": 3.0 * achieved, "tf32_frac_of_bf16_peak": 3.0 * achieved / tf_peak,
                    "note": "fused policy-MLP kernel with its two 64x64 GEMMs on tcgen05 (3xTF32, fp32-equivalent); "
                            "algorithmic bytes = 4 M (D + 2 G H + sum o); tf32_mma_tflops counts the three tensor passes",
                    "ms_per_launch": ms_launch, "launches_per_step": cnt, "top_ops_ms_per_step": ops}
    else:
        roofline = {"bound": "tensor", "kernel": f"{op} M={M_} K/D={K_} N/H={N_} batch/G={b_}",
                    "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                    "traffic": NCU_TRAFFIC.get(op), "traffic_source": NCU_TRAFFIC_SRC if op in NCU_TRAFFIC else None,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
                    "note": "exact-fp32 SIMT kernel (FFMA-bound; 1e-5 parity path), fraction quoted against the bf16 tensor peak; fp32_frac = achieved / 74.4 TFLOP/s (148 SMs x 128 FMA/clk x 1.965 GHz)",
                    "fp32_frac": achieved / 74.4,
                    "ms_per_launch": ms_launch, "launches_per_step": cnt, "top_ops_ms_per_step": ops}
    line = {"metric": "transitions/s through GAE+bonus+PPO update", "value": value, "unit": "transitions/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 5), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "flushed every step (256 MiB memset inside the timed region)",
                       "shuffle": ("np.random.permutation each epoch inside the timed region (bit-exact reference stream): draws on the host, "
                                   + ("swaps on the GPU (copy stream)" if m._device_apply() else "swaps on host worker threads")),
                       "global_minibatch": B * world, "cuda_graph": "per-minibatch launch sequence (incl. the NCCL collectives when N>1) replayed as a CUDA graph",
                       "shard_shuffle": "local (per-rank shuffle stream)" if world > 1 else "n/a (1 GPU)"},
            "e2e": {"value": e2e, "unit": "transitions/s", "h2d_bytes_per_step": int(h2d + perm_bytes // world),
                    "d2h_bytes_per_step": int(HP["n_epochs"] * N_MINIBATCH * 64), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roofline}
    if os.environ.get("PPX_BENCH_TRACE") == "1":        # diagnosis: per-rank host timeline of one more pass
        from ppo_exploration_b200 import buffer as BUF
        marks, orig_next = [], BUF.HostRngStream.next
        def next_(self):
            t0 = time.perf_counter(); v = orig_next(self); marks.append((t0, time.perf_counter())); return v
        BUF.HostRngStream.next = next_
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter(); flush.zero_(); ro.rewards.copy_(raw_rewards); bonus_and_gae(); torch.cuda.synchronize()
        t1 = time.perf_counter(); m.train(); torch.cuda.synchronize(); t2 = time.perf_counter()
        BUF.HostRngStream.next = orig_next
        gs = [v[0] for k, v in m._graphs.items() if isinstance(v, tuple)]
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for _ in range(5):
            for gph in gs:
                gph.replay()
        eb.record(); torch.cuda.synchronize()
        print(f"[trace rank {rank}] bonus+gae {1e3 * (t1 - t0):.2f} ms, train {1e3 * (t2 - t1):.2f} ms, 40 graphs back-to-back "
              f"{ea.elapsed_time(eb):.2f} ms, rng waits " + " ".join(f"{1e3 * (b - a):.2f}@{1e3 * (a - t1):.1f}" for a, b in marks),
              file=sys.stderr, flush=True)
    es = es_step_bench(torch, ppx, dev, world, rank)
    line["es"] = es
    if world == 1 and rank == 0:
        v, det = cpu_reference_pass(hash_envs=512, hash_steps=64, train_minibatches=8)
        th = host_threads()
        line["cpu_baseline"] = {"value": v, "unit": "transitions/s", "cores": th["torch_threads"], "kind": "port",
                                "sample": det["sample"], "host": th,
                                "split_s_per_pass": {k: det[k] for k in ("sim_hash_s", "gae_s", "train_s")}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL communicators referenced by captured CUDA graphs do not tear down cleanly (destroy_process_group
        # blocks); everything is flushed and synchronised here, so leave without the destructor chain.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def main():
    # watchdog: a wedged collective must not burn the box -- dump every thread's stack and exit
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("PPX_BENCH_WATCHDOG_S", "600")), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ppx", choices=["ppx", "reference"])
    ap.add_argument("--profile", action="store_true", help="short run for ncu: W warm-up + K steps only, no e2e/CPU legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ppx(args)


if __name__ == "__main__":
    main()
