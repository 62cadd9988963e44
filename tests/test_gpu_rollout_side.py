"""GPU tests of the rollout-side widening: device-side action sampling in Policy.act (f1) and the
device-side VecNormalize (f3).  Parity anchors: the log-probability of the drawn action against torch.distributions on
the same actor output (1e-6), the sample statistics against the distribution (the Philox stream itself is ppx's own:
torch's generator cannot be replayed from a kernel -- stated in DESIGN.md §4), VecNormalize against the numpy restatement
of stable_baselines3's arithmetic (oracle.rollout.VecNormalizeRef; parity unpinned upstream)."""
import numpy as np
import pytest
import torch

from oracle import rollout as OR

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("space", ["box", "discrete"])
def test_act_device_sampling(space):
    import ppo_exploration_b200 as ppx
    torch.manual_seed(5)
    N, D = 4096, 8
    sp = ppx.Box((2,)) if space == "box" else ppx.Discrete(5)
    env = ppx.SyntheticVecEnv(N, D, sp, seed=0)
    pol = ppx.Policy(env, 64)
    if space == "box":
        pol.bank.view("action_log_std").copy_(torch.tensor([-0.3, 0.2]))
    obs = torch.randn(N, D, device="cuda")
    actions, values, lp = pol.act(obs)
    a2, _, _ = pol.act(obs)
    assert not torch.equal(actions, a2), "a new draw number must give new actions"
    outs = pol.forward_raw(obs)
    dist = pol._dist(outs[0])
    if space == "box":
        assert actions.dtype == torch.float64 and tuple(actions.shape) == (N, 2) and tuple(lp.shape) == (N, 2)
        want = dist.log_prob(actions.float())
        torch.testing.assert_close(lp, want, rtol=1e-5, atol=1e-5)
        z = ((actions.float() - dist.mean) / dist.stddev).cpu().numpy()
        assert abs(z.mean()) < 0.05 and abs(z.std() - 1.0) < 0.05
    else:
        assert actions.dtype == torch.int64 and tuple(actions.shape) == (N,) and tuple(lp.shape) == (N,)
        want = dist.log_prob(actions)
        torch.testing.assert_close(lp, want, rtol=1e-5, atol=1e-6)
        freq = torch.bincount(actions, minlength=5).float().cpu().numpy() / N
        np.testing.assert_allclose(freq, dist.probs.mean(0).cpu().numpy(), atol=0.03)
    torch.testing.assert_close(values, outs[1].squeeze(-1))


def test_vecnormalize_matches_restatement():
    import ppo_exploration_b200 as ppx

    class Env:
        def __init__(self, n, d, seed):
            self.num_envs, self.rs = n, np.random.RandomState(seed)
            self.observation_space, self.action_space = ppx.Box((d,)), ppx.Box((2,))

        def reset(self):
            return (self.rs.randn(self.num_envs, 6) * 3 + 1).astype(np.float32)

        def step(self, a):
            return ((self.rs.randn(self.num_envs, 6) * 3 + 1).astype(np.float32), (self.rs.randn(self.num_envs) * 5).astype(np.float32),
                    self.rs.rand(self.num_envs) < 0.2, [{} for _ in range(self.num_envs)])

    n = 16
    vn = ppx.VecNormalize(Env(n, 6, 3))
    ref_env, ref = Env(n, 6, 3), OR.VecNormalizeRef(n, (6,))
    got = vn.reset().cpu().numpy()
    want = ref.reset(ref_env.reset())
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
    for _ in range(25):
        o, r, d, _ = vn.step(np.zeros((n, 2)))
        ro, rr, rd, _ = ref_env.step(None)
        wo, wr = ref.step(ro, rr, rd)
        assert np.array_equal(d, rd)
        np.testing.assert_allclose(o.cpu().numpy(), wo, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(r.cpu().numpy(), wr, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(vn.ret.cpu().numpy(), ref.ret, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(vn.ret_rms.var, ref.ret_rms.var, rtol=1e-12)
    # (numpy averages the f32 observations in f32, the device in f64: 1e-6 as in the RunningMeanStd fixture test)
    np.testing.assert_allclose(vn.obs_rms.mean, ref.obs_rms.mean, rtol=1e-6, atol=1e-7)
    raw = vn.get_original_obs()
    back = vn.unnormalize_obs(vn.normalize_obs(raw))
    inside = (vn.normalize_obs(raw).abs() < 9.99)
    torch.testing.assert_close(back[inside], raw[inside], rtol=1e-4, atol=1e-4)


def test_collect_samples_with_vecnormalize_and_csv_logger(tmp_path):
    """The widened rollout side end to end: VecNormalize'd env -> act (device sampling) -> add -> train; learn() logs
    through ppx's logger in the reference's CSV schema."""
    import csv
    import glob
    import ppo_exploration_b200 as ppx
    np.random.seed(0); torch.manual_seed(0)
    env = ppx.VecNormalize(ppx.SyntheticVecEnv(8, 6, ppx.Box((2,)), seed=1))
    m = ppx.PPO(env=env, env_id="synthetic", nstep=16, batch_size=64, n_epochs=1, hidden_size=64)
    import os
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        m.learn(total_timesteps=8 * 16 * 2, log_interval=1, log_to_file=True)
    finally:
        os.chdir(cwd)
    assert np.isfinite(m.last_losses).all() and m.num_timesteps == 256
    files = glob.glob(str(tmp_path / "logs" / "PPO" / "synthetic" / "run-*.csv"))
    assert len(files) == 1
    rows = list(csv.DictReader(open(files[0])))
    assert len(rows) == 2 and "total timesteps" in rows[0] and rows[1]["total timesteps"] == "256" and "value_loss" in rows[1]


def test_overlapped_rollout_load_is_the_same_rollout():
    """load_rollout(overlap=True) copies the fields the bonus does not read on a side stream; the learner methods wait for
    them: bonus + GAE + train() give bit-identical losses and weights to the in-order load."""
    import bench
    import ppo_exploration_b200 as ppx
    cfg = dict(bench.CONFIGS["C2"], N=64, T=64, batch=1024)
    host = bench.synth_rollout(cfg, 7, n_envs=64)
    pinned = {k: torch.as_tensor(v).pin_memory() for k, v in host.items() if k in bench.ROLLOUT_FIELDS}
    res = []
    for overlap in (False, True):
        np.random.seed(1); torch.manual_seed(1)
        env = ppx.SyntheticVecEnv(64, cfg["D"], ppx.Box((2,)), seed=0)
        m = ppx.PPO(env=env, nstep=64, batch_size=1024, hidden_size=64, sim_hash=True, hash_bits=64, device="cuda", **cfg["hp"])
        lv = torch.as_tensor(host["last_value"]).cuda()
        dn = torch.as_tensor(host["masks"][-1].copy()).cuda()
        for _ in range(3):                                  # eager, capture, replay
            m.rollout.load_rollout(overlap=overlap, **pinned)
            m.rollout.sim_hash_sharded(m.rollout.observations, m.rollout.rewards)
            m.rollout.compute_returns_and_advantages(lv, dn)
            m.train()
        res.append((m.last_losses.copy(), m.policy.bank.flat.clone(), m.rollout.returns.clone()))
    assert np.array_equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
