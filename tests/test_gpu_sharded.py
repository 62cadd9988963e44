"""Multi-GPU parity (needs >= 2 GPUs, skipped otherwise): a W-rank sharded run against ONE GPU holding
all envs, inside the same processes (the 1-GPU answer is computed by every rank with the dist layer masked out).

  * PPO.train, shard_shuffle="global": the reference's global np.random.permutation(T*N_total) on every rank, rollout
    replicated once per pass, rank r takes rows [r*B, (r+1)*B) of every global minibatch -> first-minibatch losses at
    1e-5, end-of-train weights at 2e-4 (Adam amplifies the summation-order noise of a different gradient partition);
  * shard_shuffle="local": replicas bit-identical;
  * RolloutStorage.sim_hash_sharded: rewards and the count table equal the 1-GPU ones bit for bit;
  * sharded RunningMeanStd / RND rollout bonus: moments and bonuses equal the 1-GPU ones;
  * PPO_RND.train and PPO_ICM.train (halo row for the shuffled-consecutive pairing), "global";
  * ES ask / tell: identical populations on every rank, the sharded update (partial GEMV + peer-memory all-reduce) equals
    the single-GPU update to 1e-12 and leaves bit-identical replicas.
Run on the box with:  gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu"""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rollout(T, n_envs, D_, space, seed, dual=False):
    rs = np.random.RandomState(seed)
    kind, n = space
    A = n if kind == "Box" else 1
    out = dict(observations=rs.randn(T, n_envs, D_).astype(np.float32),
               actions=(rs.randn(T, n_envs, A) if kind == "Box" else rs.randint(0, n, size=(T, n_envs, 1)).astype(np.float64)),
               rewards=rs.randn(T, n_envs).astype(np.float32), values=rs.randn(T, n_envs).astype(np.float32),
               masks=(rs.rand(T, n_envs) < 0.05).astype(np.uint8),
               action_log_probs=(-1.4 + 0.3 * rs.randn(T, n_envs, A)).astype(np.float32))
    if dual:
        out["int_values"] = rs.randn(T, n_envs).astype(np.float32)
        out["int_rewards"] = np.abs(rs.randn(T, n_envs)).astype(np.float32)
    return out


class _single:
    """Mask the dist layer: the code inside runs as on one GPU."""

    def __enter__(self):
        import ppo_exploration_b200.dist as PD
        self.PD, self.ws, self.rk = PD, PD.world_size, PD.rank
        PD.world_size, PD.rank = (lambda: 1), (lambda: 0)

    def __exit__(self, *a):
        self.PD.world_size, self.PD.rank = self.ws, self.rk


def _worker(rank, world, port):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      LOCAL_WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev, rank=rank, world_size=world)
    import ppo_exploration_b200 as ppx
    T, N, D_ = 64, 32, 8                 # per rank
    sl = slice(rank * N, (rank + 1) * N)
    hp = dict(lr=3e-4, gae_lam=0.95, vf_coef=1, max_grad_norm=0.5, n_epochs=3, clip_range=0.2, ent_coef=0.01)

    def shard(d):
        return {k: v[:, sl] for k, v in d.items()}

    def finish(m, data, lvv, dual, livv=None):
        m.rollout.load_rollout(**data)
        if dual:
            m.rollout.compute_returns_and_advantages(torch.tensor(lvv), torch.tensor(livv), data["masks"][-1])
        else:
            m.rollout.compute_returns_and_advantages(torch.tensor(lvv), data["masks"][-1])
        m.train(); m.train()
        torch.cuda.synchronize()

    def check(tag, w, w_ref, l, l_ref, wtol=2e-4):
        err = float((w - w_ref).abs().max())
        l0 = float(np.abs(l[0] - l_ref[0]).max() / max(1.0, np.abs(l_ref[0]).max()))       # before any Adam step
        lerr = float(np.abs(l - l_ref).max())
        assert l0 < 1e-5 and err < wtol and lerr < 1e-3, f"{tag} rank {rank}: first-minibatch loss rel diff {l0:.2e}, max|w - w_1gpu| {err:.2e}, loss diff {lerr:.2e}"

    # ---------------- PPO: global == 1 GPU; local replicas identical ----------------
    full = _rollout(T, N * world, D_, ("Box", 2), 7)
    lv = np.random.RandomState(8).randn(N * world).astype(np.float32)

    def ppo(n_envs, data, lvv, mode):
        np.random.seed(3); torch.manual_seed(3)
        env = ppx.SyntheticVecEnv(n_envs, D_, ppx.Box((2,)), seed=0)
        m = ppx.PPO(env=env, nstep=T, batch_size=T * n_envs // 2, hidden_size=64, device=dev, gamma=0.99, **hp)
        m.shard_shuffle = mode
        finish(m, data, lvv, False)
        return m.policy.bank.flat.clone(), m.last_losses.copy()

    with _single():
        w_ref, l_ref = ppo(N * world, full, lv, "global")
    w_g, l_g = ppo(N, shard(full), lv[sl], "global")
    check("PPO global", w_g, w_ref, l_g[:, :4], l_ref[:, :4])
    w_l, l_l = ppo(N, shard(full), lv[sl], "local")
    allw = [torch.empty_like(w_l) for _ in range(world)]
    dist.all_gather(allw, w_l)
    assert all(torch.equal(allw[0], a) for a in allw) and np.isfinite(l_l).all(), "local-mode replicas diverged"

    # ---------------- SimHash: sharded == 1 GPU, bit for bit ----------------
    def hashed(n_envs, data):
        np.random.seed(5)
        ro = ppx.RolloutStorage(T, n_envs, ppx.Box((D_,)), ppx.Box((2,)), sim_hash=True, device=dev, hash_bits=16)
        ro.load_rollout(**data)
        # few distinct codes: coarse observations make repeated keys (the order-dependent part of the update)
        ro.observations.copy_(torch.round(ro.observations * 0.6))
        for _ in range(2):                                     # the table persists across rollouts
            ro.sim_hash_sharded(ro.observations, ro.rewards)
        torch.cuda.synchronize()
        return ro.rewards.cpu().numpy(), ro.count_table.items()

    with _single():
        r_ref, tab_ref = hashed(N * world, full)
    r_sh, tab_sh = hashed(N, shard(full))
    assert np.array_equal(r_sh, r_ref[:, sl]), "sharded SimHash bonuses differ from the 1-GPU ones"
    assert tab_sh == tab_ref, "sharded SimHash count table differs from the 1-GPU one"

    # ---------------- RND: running moments, rollout bonus, train ----------------
    fulld = _rollout(T, N * world, D_, ("Discrete", 3), 11, dual=True)

    liv = np.random.RandomState(9).randn(N * world).astype(np.float32)

    def rnd(n_envs, data, lvv, livv, mode):
        np.random.seed(4); torch.manual_seed(4)
        env = ppx.SyntheticVecEnv(n_envs, D_, ppx.Discrete(3), seed=0)
        m = ppx.PPO_RND(env=env, nstep=T, batch_size=T * n_envs // 2, hidden_size=64, int_hidden_size=16, device=dev,
                        gamma=0.999, int_gamma=0.99, int_vf_coef=0.5, **hp)
        m.shard_shuffle = mode
        for t in range(3):                                     # warm-up statistics over env shards
            m.obs_rms.update(data["observations"][t])
        nxt = np.concatenate([data["observations"][1:], data["observations"][:1]], 0)
        bonus = m.rnd_bonus_rollout(nxt).clone()
        finish(m, data, lvv, True, livv)
        return m.policy.bank.flat.clone(), m.last_losses.copy(), bonus.cpu().numpy(), (m.obs_rms.mean, m.obs_rms.var, m.obs_rms.count,
                                                                                     m.int_rew_rms.var, m.int_rew_rms.count)

    with _single():
        w_ref, l_ref, b_ref, st_ref = rnd(N * world, fulld, lv, liv, "global")
    w_g, l_g, b_g, st_g = rnd(N, shard(fulld), lv[sl], liv[sl], "global")
    for a, b in zip(st_g, st_ref):
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-15, err_msg="sharded running moments")
    np.testing.assert_allclose(b_g, b_ref[:, sl], rtol=1e-6, atol=1e-9, err_msg="sharded RND rollout bonus")
    check("RND global", w_g, w_ref, l_g[:, :5], l_ref[:, :5])

    # ---------------- ICM: global, halo row for the shuffled-consecutive pairing ----------------
    def icm(n_envs, data, lvv):
        np.random.seed(6); torch.manual_seed(6)
        env = ppx.SyntheticVecEnv(n_envs, D_, ppx.Discrete(3), seed=0)
        m = ppx.PPO_ICM(env=env, nstep=T, batch_size=T * n_envs // 2, hidden_size=64, int_hidden_size=16, device=dev, **hp)
        finish(m, {k: v for k, v in data.items() if not k.startswith("int_")}, lvv, False)
        return torch.cat([m.policy.bank.flat, m.intrinsic_module.bank.flat]).clone(), m.last_losses.copy()

    with _single():
        w_ref, l_ref = icm(N * world, fulld, lv)
    w_g, l_g = icm(N, shard(fulld), lv[sl])
    check("ICM global", w_g, w_ref, l_g[:, :6], l_ref[:, :6], wtol=5e-4)
    # ---------------- ES: population sharded over the ranks, fused peer-memory all-reduce of the update ----------------
    def es_run(sharded):
        np.random.seed(21)
        es = ppx.EvolutionStrategy(obs_dim=8, n_actions=2, hidden_sizes=(64, 64), population_size=1000, sigma=0.1, learning_rate=0.01,
                                   decay=0.9995, novelty_param=0.5, device=dev, noise_table_size=1 << 22, noise_seed=3)
        fit = torch.as_tensor(np.random.RandomState(22).randn(1000)).to(dev)
        arch = torch.as_tensor(np.random.RandomState(23).randn(500, 2)).to(dev)
        qs = torch.as_tensor(np.random.RandomState(24).randn(2, 2)).to(dev)
        mine = fit[rank * 500:(rank + 1) * 500].contiguous() if sharded else fit
        pops = []
        for _ in range(4):                                     # eager, capture, two graph replays
            pop, w, _ = es.ask(arch, qs)
            pops.append(pop.clone())
            es.tell(mine)
        torch.cuda.synchronize()
        return es.theta.clone(), es.learning_rate, torch.stack(pops), w.clone(), es.update_mode

    with _single():
        th_ref, lr_ref, pops_ref, w_ref1, _ = es_run(False)
    th_s, lr_s, pops_s, w_s, mode = es_run(True)
    assert mode.startswith("sharded"), mode
    assert torch.equal(pops_s, pops_ref), "ranks must draw the population of the 1-GPU run"
    assert torch.equal(w_s, w_ref1[rank * 500:(rank + 1) * 500]), "perturbed weights of this rank's members"
    np.testing.assert_allclose(th_s.cpu().numpy(), th_ref.cpu().numpy(), rtol=1e-12, atol=1e-14, err_msg="sharded ES update")
    assert abs(lr_s - lr_ref) < 1e-15
    allt = [torch.empty_like(th_s) for _ in range(world)]
    dist.all_gather(allt, th_s)
    assert all(torch.equal(allt[0], a) for a in allt), "ES replicas diverged"
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)                                                # NCCL communicators referenced by captured graphs do not tear down cleanly


def test_sharded_equals_single_gpu():
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    ctx = mp.spawn(_worker, args=(world, _free_port()), nprocs=world, join=False)
    ok = ctx.join(timeout=600)
    while not ok:
        ok = ctx.join(timeout=600)
