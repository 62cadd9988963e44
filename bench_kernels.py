"""Per-kernel roofline microbenchmark of libppx (not the driver's bench: see bench.py for that contract).

For every hot kernel: the shape the headline configs use AND a shape large enough to be bandwidth-bound,
timed with CUDA events on the launching stream after warm-up, with an L2 flush (256 MiB memset) before
every timed launch.  achieved = ALGORITHMIC bytes per launch (SURVEY §8d / DESIGN.md) / time;
frac = achieved / MEASURED_PEAKS.json hbm_gbs.  Writes one JSON line per case and, with --md, a table.

  python bench_kernels.py [--only gae,simhash,...] [--md profiles/kernels_rNN.md] [--iters 20]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import ppo_exploration_b200 as ppx  # noqa: E402
from ppo_exploration_b200 import _lib as L  # noqa: E402
from ppo_exploration_b200 import models as PM  # noqa: E402

DEV = "cuda"
LARGE_ONLY = os.environ.get("PPX_KERNELS_LARGE_ONLY") == "1"        # profiling runs: only the bandwidth-sized shapes
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = PEAKS.get("hbm_gbs", 6650.0)
TF = PEAKS.get("bf16_tflops", 1590.0)
_flush = None


def timed(fn, iters, warm=3):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        _flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts = np.array(ts)
    return float(np.median(ts)), float(ts.min())


def report(rows, name, shape, bytes_, flops, ms, ms_min, launches=1, note=""):
    r = {"kernel": name, "shape": shape, "ms_median": ms, "ms_min": ms_min, "launches": launches, "note": note}
    if bytes_:
        r.update(alg_bytes=int(bytes_), gbs=bytes_ / ms / 1e6, hbm_frac=bytes_ / ms / 1e6 / HBM)
    if flops:
        r.update(flops=float(flops), tflops=flops / ms / 1e9, tensor_frac=flops / ms / 1e9 / TF)
    rows.append(r)
    print(json.dumps(r), flush=True)


def bench_gae(rows, iters):
    for T, N, dual in ((256, 2048, False), (256, 131072, False), (128, 128, True), (128, 131072, True))[(1 if LARGE_ONLY else 0)::(2 if LARGE_ONLY else 1)]:
        o, a = ppx.Box((4,)), ppx.Box((1,))
        cls = ppx.IntrinsicStorage if dual else ppx.RolloutStorage
        buf = cls(T, N, o, a)
        g = torch.Generator(device=DEV).manual_seed(0)
        for nm in ("rewards", "values") + (("int_rewards", "int_values") if dual else ()):
            getattr(buf, nm).copy_(torch.randn(T, N, device=DEV, generator=g))
        buf.masks.copy_((torch.rand(T, N, device=DEV, generator=g) < 0.02).to(torch.uint8))
        lv = torch.randn(N, device=DEV); d = buf.masks[-1].clone()
        fn = (lambda: buf.compute_returns_and_advantages(lv, lv, d)) if dual else (lambda: buf.compute_returns_and_advantages(lv, d))
        ms, mn = timed(fn, iters)
        report(rows, "gae_dual" if dual else "gae", f"T={T} N={N}", T * N * (33 if dual else 17), 0, ms, mn)
        del buf


def bench_simhash(rows, iters):
    for n, D, k in ((2048, 8, 64), (524288, 8, 64), (8 << 20, 8, 64))[(1 if LARGE_ONLY else 0):(2 if LARGE_ONLY else 3)]:
        np.random.seed(0)
        buf = ppx.RolloutStorage(1, 1, ppx.Box((D,)), ppx.Box((1,)), sim_hash=True, hash_bits=k, table_capacity=1 << 25)
        obs = torch.randn(n, D, device=DEV)
        r = torch.zeros(n, device=DEV)
        fn = lambda: buf.sim_hash(obs, r)
        ms, mn = timed(fn, min(iters, 10))
        report(rows, "simhash_update", f"n={n} D={D} k={k}", n * (4 * D + 8 + 8 + 16), 0, ms, mn,
               note="codes + ordered count update + bonus; includes the 8-byte ctrl memset")
        codes = buf.sim_hash_codes(obs)
        ms, mn = timed(lambda: buf.sim_hash_codes(obs), iters)
        report(rows, "simhash_codes", f"n={n} D={D} k={k}", n * (4 * D + 8), 0, ms, mn)
        del buf


def bench_gather(rows, iters):
    for T, N, D, A, B in ((256, 2048, 8, 2, 131072), (256, 2048, 8, 2, 524288), (128, 128, 28224, 1, 4096))[(1 if LARGE_ONLY else 0):]:
        buf = ppx.RolloutStorage(T, N, ppx.Box((D,)), ppx.Box((A,)))
        buf.observations.normal_()
        idx = torch.randperm(T * N, device=DEV)[:B].contiguous()
        bufs = buf._minibatch_buffers(B)
        ms, mn = timed(lambda: buf.gather_into(idx, bufs), iters)
        per = 2 * (4 * D + 8 * A + 4 * A + 12) + 8
        report(rows, "gather_minibatch", f"T={T} N={N} D={D} A={A} B={B}", B * per, 0, ms, mn)
        del buf, bufs


def bench_loss(rows, iters):
    for B, A, disc in ((131072, 2, 0), (4 << 20, 2, 0), (131072, 18, 1), (4 << 20, 1, 0))[(1 if LARGE_ONLY else 0):(2 if LARGE_ONLY else 4)]:
        f = lambda *s: torch.randn(*s, device=DEV)
        actor, lstd = f(B, A), torch.zeros(A, device=DEV)
        actions = (torch.randint(0, A, (B,), device=DEV).double() if disc else f(B, A).double())
        oldlp = -1.0 + 0.1 * f(B, A if not disc else 1)
        adv, val, oval, ret = f(B), f(B), f(B), f(B)
        stats = torch.tensor([0.0, 1.0], dtype=torch.float64, device=DEV)
        d_actor, d_lstd, d_val = torch.empty(B, A, device=DEV), torch.empty(A, device=DEV), torch.empty(B, device=DEV)
        losses = torch.zeros(8, dtype=torch.float64, device=DEV)
        ws = torch.empty(L.call("ppx_ppo_loss_workspace", B, A) // 8 + 1, dtype=torch.float64, device=DEV)
        cfg = L.PpoCfg(B, 0, A, disc, 0, 0.2, 0.01, 1.0, 0.0, 1.0)
        fn = lambda: L.call("ppx_ppo_loss_fwd_bwd", C.byref(cfg), actor.data_ptr(), lstd.data_ptr(), actions.data_ptr(),
                            oldlp.data_ptr(), adv.data_ptr(), stats.data_ptr(), val.data_ptr(), oval.data_ptr(),
                            ret.data_ptr(), None, None, None, None, None, d_actor.data_ptr(), d_lstd.data_ptr(),
                            d_val.data_ptr(), None, losses.data_ptr(), ws.data_ptr(), L.stream())
        ms, mn = timed(fn, iters)
        # actor_out 4A + actions 8A(Box)/8 + old_lp 4A(/4) + adv 4 + v,ov,R 12 (twice: head + dvalue) ; writes d_actor 4A + d_v 4
        byt = B * ((4 * A + 8 * A + 4 * A + 4 * A) if not disc else (4 * A + 8 + 4 + 4 * A)) + B * (4 + 12 + 12 + 4)
        report(rows, "ppo_loss_fwd_bwd", f"B={B} A={A} discrete={disc}", byt, 0, ms, mn, launches=4,
               note="head + sum + finalize + dvalue")


def bench_adam(rows, iters):
    for n in (9732, 16 << 20)[(1 if LARGE_ONLY else 0):]:
        bank = PM.ParamBank([("w", (n,))], torch.device(DEV))
        bank.flat.normal_(); bank.grad.normal_()
        ms, mn = timed(lambda: bank.adam_step(3e-4, 5.0), iters)
        report(rows, "clip_adam", f"n={n}", n * 32, 0, ms, mn, launches=3, note="step bump + sumsq + adam")


def bench_es(rows, iters):
    P, D = 10000, 4736
    np.random.seed(0)
    es = ppx.EvolutionStrategy(hidden_sizes=[64, 64], obs_dim=8, n_actions=2, population_size=P, noise_table_size=1 << 28)
    ms, mn = timed(lambda: L.call("ppx_noise_fill", es.noise_table().data_ptr(), 1 << 28, 1, L.stream()), 5)
    report(rows, "noise_fill", "n=2^28 f32", (1 << 28) * 4, 0, ms, mn)
    off = es._get_population()
    r = torch.randn(P, dtype=torch.float64, device=DEV)
    out = torch.empty(P, D, device=DEV)
    fnp = lambda: L.call("ppx_es_perturb", es.theta.data_ptr(), es.noise_table().data_ptr(), off.data_ptr(), 0.1, P, D,
                         out.data_ptr(), 0, L.stream())
    ms, mn = timed(fnp, iters)
    report(rows, "es_perturb", f"P={P} D={D} f32 out", P * D * 8, 0, ms, mn)
    obs = torch.randn(P, 8, dtype=torch.float64, device=DEV)
    ms, mn = timed(lambda: es.predict_population(off, obs), iters)
    report(rows, "es_forward", f"P={P} D={D} (8-64-64-2)", P * (D * 4 + 8 * 8 + 2 * 8), 2.0 * P * D, ms, mn,
           note=f"{P / ms * 1e3:.0f} member-steps/s; weights theta + sigma*eps formed on the fly (alg bytes = one read of eps)")
    ms, mn = timed(lambda: es._update_weights(r, off, 0.5), iters)
    report(rows, "es_update", f"P={P} D={D}", P * D * 4, 0, ms, mn, launches=3, note="stats+coef, gemv, apply")
    es.fitness_shaping = "centered_rank"
    ms, mn = timed(lambda: es._update_weights(r, off, 0.5), iters)
    report(rows, "es_update(rank)", f"P={P} D={D}", P * D * 4, 0, ms, mn, launches=5)
    for M, Q in ((10000, 2), (10000, 1024)):
        arch = torch.randn(M, 2, dtype=torch.float64, device=DEV)
        q = torch.randn(Q, 2, dtype=torch.float64, device=DEV)
        ms, mn = timed(lambda: es.novelty_batch(arch, q), iters)
        report(rows, "knn_novelty", f"M={M} Q={Q} K=10", 0, 0, ms, mn, note=f"{Q / ms * 1e3:.0f} queries/s (latency-bound, archive is L2-resident)")


def bench_linear(rows, iters):
    sc = PM._Scratch(torch.device(DEV))
    for M, K, N, batch in ((131072, 8, 128, 1), (131072, 64, 64, 2), (131072, 64, 2, 1), (4096, 28224, 384, 1), (16384, 28224, 256, 1)):
        ld = N * batch
        x = torch.randn(M, K * batch, device=DEV)
        w = torch.randn(batch, K, N, device=DEV) / np.sqrt(K)
        b = torch.randn(batch, N, device=DEV)
        y = torch.empty(M, ld, device=DEV)
        dy = torch.randn(M, ld, device=DEV)
        dx = torch.empty(M, K * batch, device=DEV)
        dw, db = torch.empty_like(w), torch.empty_like(b)
        fl = 2.0 * M * K * N * batch
        ms, mn = timed(lambda: PM.linear_fwd(x.data_ptr(), K * batch, w.data_ptr(), b.data_ptr(), M, K, N, 1, y.data_ptr(), ld,
                                             batch, K, K * N, N, N), iters)
        report(rows, "linear_fwd", f"M={M} K={K} N={N} b={batch}", 4 * M * (K + N) * batch, fl, ms, mn)
        ms, mn = timed(lambda: PM.linear_bwd_data(dy.data_ptr(), ld, w.data_ptr(), M, K, N, x.data_ptr(), K * batch, 1, dx.data_ptr(),
                                                  K * batch, batch, N, K * N, K, K), iters)
        report(rows, "linear_bwd_data", f"M={M} K={K} N={N} b={batch}", 4 * M * (2 * K + N) * batch, fl, ms, mn)
        ms, mn = timed(lambda: PM.linear_bwd_weight(sc, x.data_ptr(), K * batch, dy.data_ptr(), ld, M, K, N, dw.data_ptr(), db.data_ptr(),
                                                    batch, K, N, K * N, N), iters)
        report(rows, "linear_bwd_weight", f"M={M} K={K} N={N} b={batch}", 4 * M * (K + N) * batch, fl, ms, mn, launches=2)
        del x, y, dy, dx


def bench_tc(rows, iters):
    for M, R, N in ((131072, 64, 64), (131072, 128, 128), (16384, 28224, 256), (16384, 28224, 128), (32768, 3136, 512)):
        A = torch.randn(M, R, device=DEV)
        W = torch.randn(N, R, device=DEV) / np.sqrt(R)
        hi, lo = torch.empty_like(W), torch.empty_like(W)
        L.call("ppx_tc_split", W.data_ptr(), N, R, hi.data_ptr(), lo.data_ptr(), None, None, L.stream())
        b = torch.randn(N, device=DEV)
        Cc = torch.empty(M, N, device=DEV)
        fn = lambda: L.call("ppx_tc_linear", A.data_ptr(), R, hi.data_ptr(), lo.data_ptr(), R, M, R, N, b.data_ptr(), None, 0, 1, 0,
                            None, None, 0.0, Cc.data_ptr(), N, L.stream())
        ms, mn = timed(fn, iters)
        report(rows, "tc_linear(3xTF32)", f"M={M} R={R} N={N}", 4 * M * (R + N), 2.0 * M * R * N, ms, mn)


def bench_mlp3(rows, iters):
    """Fused policy-MLP forward / backward (mlp_fused.cu) at the C2 minibatch and at the RND/C1 width."""
    for M, D, h, space, intr in ((131072, 8, 64, ppx.Box((2,)), False), (131072, 8, 64, ppx.Box((2,)), True),
                                 (16384, 4, 128, ppx.Discrete(2), False)):
        env = ppx.SyntheticVecEnv(4, D, space, seed=0)
        pol = ppx.models.Policy(env, h, intrinsic_model=intr, device=DEV)
        assert pol.mlp._fused_args()["ok"]
        G, so = len(pol.outs), sum(pol.outs)
        x = torch.randn(M, D, device=DEV)
        fa = pol.mlp._fused_args()
        for tc in ((True, False) if fa["tc"] else (False,)):
            fa["tc"] = tc
            tag = "_tc" if tc else ""
            kind = "tcgen05 3xTF32 (fp32-equivalent flops)" if tc else "fp32 SIMT"
            outs = pol.forward_raw(x)
            d = [torch.randn_like(o) / M for o in outs]
            ms, mn = timed(lambda: pol.forward_raw(x), iters)
            fl = 2.0 * M * (G * (D * h + h * h) + h * so)
            report(rows, "mlp3_fwd" + tag, f"M={M} D={D} H={h} G={G}", 4 * M * (D + 2 * G * h + so), fl, ms, mn,
                   note=f"{kind}: {fl / ms / 1e9 / 74.4:.3f} of 74.4 TFLOP/s FFMA peak")
            ms, mn = timed(lambda: pol.mlp.backward(d), iters)
            fl = 2.0 * M * (G * (D * h + 2 * h * h) + 2 * h * so)
            report(rows, "mlp3_bwd" + tag, f"M={M} D={D} H={h} G={G}", 4 * M * (D + 2 * G * h + so), fl, ms, mn, launches=2,
                   note=f"bwd + partial reduce; {kind}: {fl / ms / 1e9 / 74.4:.3f} of 74.4 TFLOP/s FFMA peak")


def bench_bonus(rows, iters):
    """Bonus nets at the Atari-shaped configs: C3 RND (16384 obs x 28224, h=128) and C4 ICM (32768 pairs x 3136, h=f=512)."""
    np.random.seed(0); torch.manual_seed(0)
    M, D, h = 16384, 28224, 128
    env = ppx.SyntheticVecEnv(4, 8, ppx.Discrete(18), seed=0)
    rnd = ppx.RndNetwork(D, hidden_size=h, device=DEV)
    obs = torch.rand(M, D, device=DEV)
    rms = ppx.RunningMeanStd(shape=(D,), device=DEV)
    fl = 2.0 * M * (2 * D * h + 3 * h * h + 2 * h)
    for tag, tc in (("tcgen05 3xTF32 first layers", True), ("SIMT fp32", False)):
        if not tc:
            rnd.predictor.tc, rnd.target.tc = {}, {}
        fn = lambda: rnd.int_reward(obs, rms=rms)
        ms, mn = timed(fn, max(3, iters // 4))
        report(rows, "rnd_bonus", f"C3: M={M} D={D} h={h}", 4 * M * D, fl, ms, mn, launches=9,
               note=f"{tag}; normalize_obs (fused into the tensor-core kernels / separate pass for SIMT) + target/predictor forward + (p-t)^2; {M / ms * 1e3:.0f} obs/s; alg bytes = one read of the observations")
    del rnd, obs, rms
    torch.cuda.empty_cache()
    M, D, h, nA = 32768, 3136, 512, 18
    icm = ppx.IntrinsicCuriosityModule(D, ppx.ActionConverter(ppx.Discrete(nA)), hidden_size=h, device=DEV)
    s0, s1 = torch.rand(M, D, device=DEV), torch.rand(M, D, device=DEV)
    act = torch.randint(0, nA, (M,), device=DEV)
    fl = 2.0 * (2 * M * (D * h + h * h) + M * ((nA + h) * h + h * h))
    for tag, tc in (("tcgen05 3xTF32 first layer", True), ("SIMT fp32", False)):
        if not tc:
            for st in (icm.enc, icm.fwd, icm.inv):
                st.tc = {}
        fn = lambda: icm.int_reward(s0, s1, act)
        ms, mn = timed(fn, max(3, iters // 4))
        report(rows, "icm_bonus", f"C4: M={M} D={D} h=f={h}", 4 * 2 * M * D, fl, ms, mn, launches=8,
               note=f"{tag}; encoder(s), encoder(s'), forward model, clamp(mse); {M / ms * 1e3:.0f} transitions/s")


def bench_configs(rows, iters):
    """Whole learner passes at the Atari-shaped configs (BASELINE configs[2], [3]) with synthetic rollouts:
    C3 = PPO_RND, 128 envs x 128 steps, flattened 84x84x4 frames (D=28224), dual-head GAE (gamma .999 / .99), h=128;
    C4 = PPO_ICM, 32 envs x 128 steps per GPU, D=3136 features, h=f=512, Discrete(18)."""
    for name in ("C3", "C4"):
        np.random.seed(0); torch.manual_seed(0)
        if name == "C3":
            T, N, D, space = 128, 128, 28224, ppx.Discrete(18)
            env = ppx.SyntheticVecEnv(N, D, space, seed=0)
            m = ppx.PPO_RND(env=env, nstep=T, batch_size=4096, n_epochs=4, gamma=0.999, int_gamma=0.99, hidden_size=128,
                            int_hidden_size=128, device=DEV)
        else:
            T, N, D, space = 128, 32, 3136, ppx.Discrete(18)
            env = ppx.SyntheticVecEnv(N, D, space, seed=0)
            m = ppx.PPO_ICM(env=env, nstep=T, batch_size=1024, n_epochs=4, hidden_size=128, int_hidden_size=512, device=DEV)
        ro = m.rollout
        g = torch.Generator(device=DEV).manual_seed(0)
        ro.observations.copy_(torch.rand(T, N, D, device=DEV, generator=g))
        ro.actions.copy_(torch.randint(0, 18, (T, N, 1), device=DEV, generator=g).double())
        for nm in ("rewards", "values", "action_log_probs") + (("int_values",) if name == "C3" else ()):
            getattr(ro, nm).copy_(torch.randn(getattr(ro, nm).shape, device=DEV, generator=g) * (0.1 if nm == "action_log_probs" else 1.0) - (2.9 if nm == "action_log_probs" else 0.0))
        ro.masks.copy_((torch.rand(T, N, device=DEV, generator=g) < 0.02).to(torch.uint8))
        ro.pos, ro.full = T, True
        lv = torch.randn(N, device=DEV)
        next_obs = torch.rand(T, N, D, device=DEV, generator=g)

        def one_pass():
            if name == "C3":
                ro.int_rewards.copy_(m.rnd_bonus_rollout(next_obs))
                ro.compute_returns_and_advantages(lv, lv, ro.masks[-1])
            else:
                ro.compute_returns_and_advantages(lv, ro.masks[-1])
            m.train()
        ms, mn = timed(one_pass, max(3, iters // 4), warm=3)
        report(rows, f"{name}_learner_pass", f"T={T} N={N} D={D}", 0, 0, ms, mn, launches=0,
               note=f"{T * N / ms * 1e3:.0f} transitions/s ({'RND bonus over the rollout + dual GAE + ' if name == 'C3' else 'GAE + '}"
                    f"{m.n_epochs} epochs x {-(-T * N // m.batch_size)} minibatches of {m.batch_size})")
        del m, ro, next_obs
        torch.cuda.empty_cache()


ALL = {"configs": bench_configs, "bonus": bench_bonus, "mlp3": bench_mlp3, "tc": bench_tc, "gae": bench_gae, "simhash": bench_simhash, "gather": bench_gather, "loss": bench_loss, "adam": bench_adam,
       "es": bench_es, "linear": bench_linear}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=",".join(ALL))
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--md", default=None)
    args = ap.parse_args()
    rows = []
    for k in args.only.split(","):
        ALL[k](rows, args.iters)
        torch.cuda.empty_cache()
    if args.md:
        with open(args.md, "w") as f:
            f.write(f"| kernel | shape | ms (median) | alg. bytes | GB/s | frac of measured HBM ({HBM:.0f} GB/s) | TFLOP/s | note |\n|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                f.write(f"| {r['kernel']} | {r['shape']} | {r['ms_median']:.4f} | {r.get('alg_bytes', '')} | "
                        f"{r.get('gbs', 0):.0f} | {r.get('hbm_frac', 0):.3f} | {r.get('tflops', 0):.2f} | {r['note']} |\n")


if __name__ == "__main__":
    main()
