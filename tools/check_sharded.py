"""torchrun --nproc-per-node W tools/check_sharded.py : sharded PPO.train() against a single-GPU run holding all envs.

shard_shuffle="global": every rank must end with the weights of the 1-GPU run over the concatenated rollout
(same global permutation, owner-computes).  shard_shuffle="local": ranks must agree with each other bit-for-bit
(replicated weights after all-reduced gradients) and the loss log must be finite."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import ppo_exploration_b200 as ppx

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
T, N, D_, A = 64, 32, 8, 2          # per rank
hp = dict(lr=3e-4, gamma=0.99, gae_lam=0.95, vf_coef=1, max_grad_norm=0.5, n_epochs=3, clip_range=0.2, ent_coef=0.01)

def rollout(n_envs, seed):
    rs = np.random.RandomState(seed)
    return dict(observations=rs.randn(T, n_envs, D_).astype(np.float32), actions=rs.randn(T, n_envs, A),
                rewards=rs.randn(T, n_envs).astype(np.float32), values=rs.randn(T, n_envs).astype(np.float32),
                masks=(rs.rand(T, n_envs) < 0.05).astype(np.uint8),
                action_log_probs=(-1.4 + 0.3 * rs.randn(T, n_envs, A)).astype(np.float32))

full = rollout(N * world, 7)
mine = {k: v[:, rank * N:(rank + 1) * N] for k, v in full.items()}
lv = np.random.RandomState(8).randn(N * world).astype(np.float32)

def run(n_envs, data, lvv, mode, sharded):
    np.random.seed(3); torch.manual_seed(3)
    env = ppx.SyntheticVecEnv(n_envs, D_, ppx.Box((A,)), seed=0)
    m = ppx.PPO(env=env, nstep=T, batch_size=T * n_envs // 2, hidden_size=64, device=dev, **hp)
    m.shard_shuffle = mode
    m.rollout.load_rollout(**data)
    m.rollout.compute_returns_and_advantages(torch.tensor(lvv), data["masks"][-1])
    if not sharded:
        import ppo_exploration_b200.dist as PD
        ws, rk = PD.world_size, PD.rank
        PD.world_size, PD.rank = (lambda: 1), (lambda: 0)
        try:
            m.train(); m.train()
        finally:
            PD.world_size, PD.rank = ws, rk
    else:
        m.train(); m.train()
    return m.policy.bank.flat.clone(), m.last_losses.copy()

w_ref, l_ref = run(N * world, full, lv, "global", sharded=False)            # every rank computes the 1-GPU answer
w_g, l_g = run(N, mine, lv[rank * N:(rank + 1) * N], "global", sharded=True)
err = float((w_g - w_ref).abs().max())
lerr = float(np.abs(l_g[:, :4] - l_ref[:, :4]).max())
l0 = float(np.abs(l_g[0, :4] - l_ref[0, :4]).max() / max(1.0, np.abs(l_ref[0, :4]).max()))   # before any Adam step
w_l, l_l = run(N, mine, lv[rank * N:(rank + 1) * N], "local", sharded=True)
allw = [torch.empty_like(w_l) for _ in range(world)]
dist.all_gather(allw, w_l)
same = all(torch.equal(allw[0], a) for a in allw)
# first minibatch: pure arithmetic parity (1e-5); later ones: Adam amplifies summation-order noise
ok = l0 < 1e-5 and err < 2e-4 and lerr < 1e-3 and same and np.isfinite(l_l).all()
print(f"rank {rank}: global-mode first-minibatch loss rel diff {l0:.2e}, max|w - w_1gpu| = {err:.2e}, loss diff {lerr:.2e}; local-mode replicas identical: {same}; ok={ok}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
