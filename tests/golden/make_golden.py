"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) here.

Run in the build container only:   python tests/golden/make_golden.py
The reference has no tests or golden vectors of its own, so these fixtures --
outputs of the reference's live classes on seeded synthetic inputs -- are what pins the oracle
(tests/test_oracle_golden.py) and, through it and directly, the CUDA path (tests/test_gpu_*.py).
Image versions that produced them: numpy 2.3.5, torch 2.11.0 (CPU), scikit-learn 1.9.0.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import algorithms  # noqa: E402
import buffer as refbuf  # noqa: E402
import evolution_strategies as refes  # noqa: E402
import logger as reflogger  # noqa: E402
import models as refmodels  # noqa: E402
import sil_module  # noqa: E402
import util as refutil  # noqa: E402


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, {len(arrs)} arrays")


def sd(module, prefix):
    return {f"{prefix}/{k}": v.detach().numpy().copy() for k, v in module.state_dict().items()}


class Obs:
    def __init__(self, d):
        self.shape = (d,)


class Box:
    def __init__(self, a):
        self.shape = (a,)


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


# ------------------------------------------------------------------ scans
def g_gae():
    out = {}
    for tag, (T, N, gamma, lam, dp) in {"a": (16, 5, 0.99, 0.95, 0.1), "b": (33, 3, 0.999, 0.9, 0.5),
                                        "c": (7, 4, 0.9, 1.0, 0.0), "d": (5, 2, 0.99, 0.95, 1.0)}.items():
        rs = np.random.RandomState(sum(map(ord, tag)))
        buf = refbuf.RolloutStorage(T, N, Obs(3), Box(2), gae_lam=lam, gamma=gamma)
        rew = rs.randn(T, N).astype(np.float32)
        val = rs.randn(T, N).astype(np.float32)
        msk = (rs.rand(T, N) < dp)
        for t in range(T):
            buf.add(rs.randn(N, 3).astype(np.float32), rs.randn(N, 2), rew[t], torch.tensor(val[t]), msk[t],
                    torch.zeros(N, 2))
        last_v = rs.randn(N).astype(np.float32)
        dones = msk[-1] if tag != "b" else (rs.rand(N) < 0.5)
        buf.compute_returns_and_advantages(torch.tensor(last_v), dones)
        out.update({f"{tag}/rewards": rew, f"{tag}/values": val, f"{tag}/masks": msk.astype(np.uint8),
                    f"{tag}/last_value": last_v, f"{tag}/dones": np.asarray(dones).astype(np.uint8),
                    f"{tag}/hp": np.array([gamma, lam]), f"{tag}/adv": buf.advantages, f"{tag}/ret": buf.returns})
    save("gae_single", **out)


def g_gae_dual():
    out = {}
    for tag, (T, N, gamma, ig, lam, dp) in {"a": (16, 5, 0.999, 0.99, 0.95, 0.1),
                                            "b": (40, 3, 0.99, 0.9, 0.8, 0.3)}.items():
        rs = np.random.RandomState(7 + len(out))
        buf = refbuf.IntrinsicStorage(T, N, Obs(3), Box(2), gae_lam=lam, gamma=gamma, int_gamma=ig)
        buf.reset()
        rew, irew = rs.randn(T, N).astype(np.float32), np.abs(rs.randn(T, N)).astype(np.float32)
        val, ival = rs.randn(T, N).astype(np.float32), rs.randn(T, N).astype(np.float32)
        msk = rs.rand(T, N) < dp
        for t in range(T):
            buf.add(rs.randn(N, 3).astype(np.float32), rs.randn(N, 2), rew[t], irew[t], torch.tensor(val[t]),
                    torch.tensor(ival[t]), msk[t], torch.zeros(N, 2))
        lv, liv = rs.randn(N).astype(np.float32), rs.randn(N).astype(np.float32)
        buf.compute_returns_and_advantages(torch.tensor(lv), torch.tensor(liv), msk[-1])
        out.update({f"{tag}/rewards": rew, f"{tag}/int_rewards": irew, f"{tag}/values": val,
                    f"{tag}/int_values": ival, f"{tag}/masks": msk.astype(np.uint8), f"{tag}/last_value": lv,
                    f"{tag}/last_int_value": liv, f"{tag}/dones": msk[-1].astype(np.uint8),
                    f"{tag}/hp": np.array([gamma, ig, lam]), f"{tag}/adv": buf.advantages,
                    f"{tag}/ret": buf.returns, f"{tag}/int_adv": buf.int_advantages,
                    f"{tag}/int_ret": buf.int_returns})
    save("gae_dual", **out)


def g_discount():
    class S:
        gamma = 0.97
    rs = np.random.RandomState(3)
    r = rs.randn(25).tolist()
    d = (rs.rand(25) < 0.2).astype(float).tolist()
    out = sil_module.SilModule.discount_with_dones(S(), r, d)
    save("discount", rewards=np.array(r), dones=np.array(d), gamma=np.array(S.gamma), out=np.array(out))


# ------------------------------------------------------------------ simhash
def g_simhash():
    out = {}
    for tag, (k, D, N, steps, scale) in {"k16": (16, 2, 64, 3, 1.0), "k64": (64, 8, 48, 3, 1.0),
                                         "k8dup": (8, 3, 40, 4, 1.0)}.items():
        np.random.seed(11)
        buf = refbuf.RolloutStorage(steps, N, Obs(D), Box(2), sim_hash=True)
        if k != 16:
            buf.A = np.random.randn(k, D)                       # A is a plain attribute
        rs = np.random.RandomState(5)
        obs_all, rew_in, rew_out = [], [], []
        for t in range(steps):
            obs = rs.randn(N, D).astype(np.float32) * scale
            if t > 0:
                obs[: N // 2] = obs_all[0][: N // 2]            # force cross-step repeats
            obs[1] = obs[0]                                     # and an in-batch duplicate
            r = rs.randn(N).astype(np.float32)
            obs_all.append(obs.copy()); rew_in.append(r.copy())
            rew_out.append(buf.sim_hash(obs, r).copy())
        keys = sorted(buf.count_table.keys())
        bits = np.array([[int(c) for c in key if c in "01"] for key in keys], dtype=np.uint8)
        out.update({f"{tag}/A": buf.A, f"{tag}/obs": np.array(obs_all), f"{tag}/rew_in": np.array(rew_in),
                    f"{tag}/rew_out": np.array(rew_out), f"{tag}/table_bits": bits,
                    f"{tag}/table_counts": np.array([buf.count_table[k_] for k_ in keys], dtype=np.int64)})
    # float64 reward input (env rewards are f64 under VecNormalize): bonus added in f64
    np.random.seed(12)
    buf = refbuf.RolloutStorage(1, 16, Obs(4), Box(2), sim_hash=True)
    rs = np.random.RandomState(6)
    obs = rs.randn(16, 4).astype(np.float32); obs[3] = obs[2]
    r64 = rs.randn(16)
    out.update({"f64/A": buf.A, "f64/obs": obs, "f64/rew_in": r64.copy(), "f64/rew_out": buf.sim_hash(obs, r64).copy()})
    save("simhash", **out)


# ------------------------------------------------------------------ shuffle / gather
def g_get():
    T, N, D, A = 8, 3, 4, 2
    rs = np.random.RandomState(2)
    buf = refbuf.RolloutStorage(T, N, Obs(D), Box(A))
    for t in range(T):
        buf.add(rs.randn(N, D).astype(np.float32), rs.randn(N, A), rs.randn(N).astype(np.float32),
                torch.tensor(rs.randn(N).astype(np.float32)), rs.rand(N) < 0.2,
                torch.tensor(rs.randn(N, A).astype(np.float32)))
    buf.compute_returns_and_advantages(torch.zeros(N), np.zeros(N, bool))
    raw = {k: getattr(buf, k).copy() for k in ("observations", "actions", "values", "action_log_probs",
                                                "advantages", "returns", "rewards", "masks")}
    np.random.seed(123)
    out = {f"raw/{k}": v for k, v in raw.items()}
    for ep in range(2):
        for bi, b in enumerate(buf.get(5)):
            for f in b._fields:
                out[f"ep{ep}/b{bi}/{f}"] = getattr(b, f).numpy()
    np.random.seed(123)
    out["perm0"] = np.random.permutation(T * N)
    out["perm1"] = np.random.permutation(T * N)
    save("get_single", **out)

    buf = refbuf.IntrinsicStorage(T, N, Obs(D), Discrete(3))
    buf.reset()
    for t in range(T):
        buf.add(rs.randn(N, D).astype(np.float32), rs.randint(0, 3, (N, 1)), rs.randn(N).astype(np.float32),
                rs.randn(N).astype(np.float32), torch.tensor(rs.randn(N).astype(np.float32)),
                torch.tensor(rs.randn(N).astype(np.float32)), rs.rand(N) < 0.2,
                torch.tensor(rs.randn(N, 1).astype(np.float32)))
    buf.compute_returns_and_advantages(torch.zeros(N), torch.ones(N), np.zeros(N, bool))
    out = {f"raw/{k}": getattr(buf, k).copy() for k in ("observations", "actions", "values", "int_values",
                                                         "action_log_probs", "advantages", "int_advantages",
                                                         "returns", "int_returns")}
    np.random.seed(77)
    for bi, b in enumerate(buf.get(7)):
        for f in b._fields:
            out[f"b{bi}/{f}"] = getattr(b, f).numpy()
    save("get_dual", **out)


# ------------------------------------------------------------------ learners
class LogCapture:
    def __init__(self):
        self.rec = {}
        self._orig = reflogger.record

    def __enter__(self):
        reflogger.record = lambda k, v, *a, **kw: self.rec.__setitem__(k, v)
        algorithms.logger.record = reflogger.record
        refbuf.logger.record = reflogger.record
        return self

    def __exit__(self, *a):
        reflogger.record = self._orig
        algorithms.logger.record = self._orig
        refbuf.logger.record = self._orig


def rollout_arrays(ro, dual=False):
    keys = ["observations", "actions", "rewards", "values", "masks", "action_log_probs", "advantages", "returns"]
    if dual:
        keys += ["int_rewards", "int_values", "int_advantages", "int_returns"]
    return {f"ro/{k}": getattr(ro, k).copy() for k in keys}


def g_ppo(tag, n_envs, obs_dim, space, seed, **kw):
    ref_shim.set_env_factory(lambda: ref_shim.FakeVecEnv(n_envs, obs_dim, space, seed=seed))
    np.random.seed(seed); torch.manual_seed(seed)
    m = algorithms.PPO(env_id="synthetic", **kw)
    m.collect_samples()
    out = rollout_arrays(m.rollout)
    out.update(sd(m.policy.net, "init"))
    out["A"] = m.rollout.A
    np.random.seed(seed + 100)
    with LogCapture() as lc:
        m.train()
    out.update(sd(m.policy.net, "final"))
    out["log"] = np.array([lc.rec[k] for k in ("train/total_loss", "train/policy_gradient_loss",
                                               "train/value_loss", "train/entropy_loss")])
    out["train_seed"] = np.array(seed + 100)
    save(tag, **out)


def g_rnd(tag, n_envs, obs_dim, space, seed, **kw):
    ref_shim.set_env_factory(lambda: ref_shim.FakeVecEnv(n_envs, obs_dim, space, seed=seed))
    np.random.seed(seed); torch.manual_seed(seed)
    m = algorithms.PPO_RND(env_id="synthetic", rnd_start=kw["nstep"], **kw)
    with LogCapture():
        m.collect_samples()                                      # rollout 1: warm-up (obs_rms only) except last step
        out = {"rms0/obs_mean": np.array(m.obs_rms.mean), "rms0/obs_var": np.array(m.obs_rms.var),
               "rms0/obs_count": np.array(m.obs_rms.count), "rms0/int_mean": np.array(m.int_rew_rms.mean),
               "rms0/int_var": np.array(m.int_rew_rms.var), "rms0/int_count": np.array(m.int_rew_rms.count),
               "first_obs": m.last_obs.copy()}
        m.collect_samples()                                      # rollout 2: bonus on every step
    out.update(rollout_arrays(m.rollout, dual=True))
    out["final_obs"] = m.last_obs.copy()
    out.update({"rms1/int_mean": np.array(m.int_rew_rms.mean), "rms1/int_var": np.array(m.int_rew_rms.var),
                "rms1/int_count": np.array(m.int_rew_rms.count)})
    out.update(sd(m.policy.net, "init")); out.update(sd(m.rnd, "rnd_init"))
    np.random.seed(seed + 100)
    with LogCapture() as lc:
        m.train()
    out.update(sd(m.policy.net, "final")); out.update(sd(m.rnd, "rnd_final"))
    out["log"] = np.array([lc.rec[k] for k in ("train/total_loss", "train/policy_gradient_loss", "train/value_loss",
                                               "train/entropy_loss", "train/intrinsic_loss")])
    out["train_seed"] = np.array(seed + 100)
    save(tag, **out)


def g_icm(tag, n_envs, obs_dim, space, seed, **kw):
    ref_shim.set_env_factory(lambda: ref_shim.FakeVecEnv(n_envs, obs_dim, space, seed=seed))
    np.random.seed(seed); torch.manual_seed(seed)
    m = algorithms.PPO_ICM(env_id="synthetic", **kw)
    # stand-alone bonus fixture (models.py:311-320)
    rs = np.random.RandomState(seed)
    s, ns = rs.randn(n_envs, obs_dim).astype(np.float32), rs.randn(n_envs, obs_dim).astype(np.float32)
    if space.__class__.__name__ == "Discrete":
        a = torch.tensor(rs.randint(0, space.n, n_envs))
    else:
        a = torch.tensor(rs.randn(n_envs, space.shape[0]).astype(np.float32))
    out = {"bonus/s": s, "bonus/ns": ns, "bonus/a": a.numpy(),
           "bonus/r": m.intrinsic_module.int_reward(torch.Tensor(s), torch.Tensor(ns), a).detach().numpy()}
    with LogCapture():
        m.collect_samples()
    out.update(rollout_arrays(m.rollout))
    out.update(sd(m.policy.net, "init")); out.update(sd(m.intrinsic_module, "icm_init"))
    np.random.seed(seed + 100)
    with LogCapture() as lc:
        m.train()
    out.update(sd(m.policy.net, "final")); out.update(sd(m.intrinsic_module, "icm_final"))
    out["log"] = np.array([lc.rec[k] for k in ("train/total_loss", "train/policy_gradient_loss", "train/value_loss",
                                               "train/entropy_loss", "train/icm_loss")])
    out["train_seed"] = np.array(seed + 100)
    save(tag, **out)


def g_rnd_bonus():
    torch.manual_seed(0)
    rs = np.random.RandomState(0)
    net = refmodels.RndNetwork(8, hidden_size=16)
    with torch.no_grad():                                        # non-constant weights, same architecture
        for prm in net.parameters():
            prm.add_(torch.randn_like(prm) * 0.05)
    obs = rs.randn(32, 8)
    out = sd(net, "rnd")
    out["obs"] = obs
    out["r"] = net.int_reward(obs).detach().numpy()
    net0 = refmodels.RndNetwork(8, hidden_size=16)               # reference constant init
    out.update(sd(net0, "rnd0"))
    out["r0"] = net0.int_reward(obs).detach().numpy()
    rms = refutil.RunningMeanStd()
    seq = []
    for i in range(4):
        x = np.abs(rs.randn(32)).astype(np.float32) * (i + 1)
        rms.update(x)
        seq.append(np.concatenate([x.astype(np.float64), [rms.mean, rms.var, rms.count]]))
    out["rms_seq"] = np.array(seq)
    rmsv = refutil.RunningMeanStd()
    xs = rs.randn(3, 16, 8)
    for x in xs:
        rmsv.update(x)
    out["rmsv_in"] = xs; out["rmsv_mean"] = rmsv.mean; out["rmsv_var"] = rmsv.var; out["rmsv_count"] = np.array(rmsv.count)
    save("rnd_bonus", **out)


# ------------------------------------------------------------------ logger CSV schema
LOGGER_DUMPS = [{"time/total timesteps": 2048, "rollout/ep_rew_mean": 1.5},
                {"time/total timesteps": 4096, "train/value_loss": 0.25, "train/entropy_loss": -1.4, "plain": 3},
                {"rollout/ep_rew_mean": 2.5, "train/value_loss": 0.125}]


def g_logger():
    """The reference's CSVOutputFormat (logger.py:13-52) fed a fixed record sequence -> tests/golden/logger_ref.csv"""
    path = os.path.join(HERE, "logger_ref.csv")
    w = reflogger.CSVOutputFormat(path)
    for d in LOGGER_DUMPS:
        w.write(dict(d))
    w.close()


# ------------------------------------------------------------------ ES
def g_es():
    class E:
        pass
    es = E.__new__(E)
    ES = refes.EvolutionStrategy
    shapes = [(8, 16), (16, 16), (16, 2)]
    np.random.seed(4)
    es.weights = [np.random.randn(*s) for s in shapes]
    es.POPULATION_SIZE, es.SIGMA, es.learning_rate, es.decay, es.novelty_param, es.K = 24, 0.1, 0.01, 0.9995, 0.5, 10
    out = {f"w0/{i}": w.copy() for i, w in enumerate(es.weights)}
    pop = ES._get_population(es)
    # keep epsilon f32-representable so the device noise table can hold the identical values
    pop = [[l.astype(np.float32).astype(np.float64) for l in m] for m in pop]
    out["eps"] = np.array([np.concatenate([l.ravel() for l in m]) for m in pop])
    wt = ES._get_weights_try(es, es.weights, pop[3])
    out["try3"] = np.concatenate([l.ravel() for l in wt])
    rewards = np.random.randn(24)
    out["rewards"] = rewards
    ES._update_weights(es, rewards, pop, 0.37)
    for i, w in enumerate(es.weights):
        out[f"w1/{i}"] = w.copy()
    out["lr1"] = np.array(es.learning_rate)
    ES._update_weights(es, rewards * 0 + 2.5, pop, 0.37)           # std == 0 -> no-op
    out["lr2"] = np.array(es.learning_rate)
    es.novelty_param = 0.2
    ES._update_weights(es, rewards[::-1].copy(), pop)              # novelty=None branch
    for i, w in enumerate(es.weights):
        out[f"w3/{i}"] = w.copy()
    out["lr3"] = np.array(es.learning_rate)
    # kNN
    for M in (1, 5, 10, 30, 500):
        arch = [np.random.randn(1, 2) for _ in range(M)]
        q = np.random.randn(1, 2)
        S = int(np.minimum(10, M))
        out[f"knn{M}/archive"] = np.concatenate(arch); out[f"knn{M}/q"] = q
        out[f"knn{M}/sum"] = np.array(ES.get_kNN(es, arch, q, S))
    out["probs_in"] = np.array([0.3123, 1.77])
    out["probs"] = np.array(ES.calc_noveltiy_distribution(es, [0.3123, 1.77]))
    save("es", **out)


def g_es_predict():
    """FeedForwardNetwork.predict (evolution_strategies.py:48-61) of the unmodified reference on a Box and a Discrete env."""
    out = {}
    for tag, space in (("box", ref_shim.Box((2,))), ("disc", ref_shim.Discrete(3))):
        env = ref_shim.FakeVecEnv(1, 8, space, seed=11)
        np.random.seed(21)
        net = refes.FeedForwardNetwork(env, hidden_sizes=[16, 16])
        for i, w in enumerate(net.weights):
            out[f"{tag}/w/{i}"] = w.copy()
        obs = np.random.randn(6, 8)
        out[f"{tag}/obs"] = obs
        np.random.seed(31)
        out[f"{tag}/actions"] = np.array([net.predict(o) for o in obs]).reshape(6, -1)
    save("es_predict", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "es_predict":           # added later: regenerate this fixture alone
        g_es_predict()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "logger":               # added in round 2
        g_logger()
        sys.exit(0)
    g_es_predict()
    g_logger()
    g_gae(); g_gae_dual(); g_discount(); g_simhash(); g_get(); g_rnd_bonus(); g_es()
    # C1: the reference's own CPU-runnable case (8 envs x 128 steps, obs 4, Discrete(2), defaults)
    g_ppo("ppo_c1_discrete", 8, 4, ref_shim.Discrete(2), seed=1, nstep=128)
    g_ppo("ppo_box_small", 4, 8, ref_shim.Box((2,)), seed=2, nstep=32, batch_size=40, n_epochs=3, hidden_size=64,
          lr=3e-4, gamma=0.999, vf_coef=1, max_grad_norm=5, ent_coef=0.0)
    g_ppo("ppo_box_simhash", 4, 8, ref_shim.Box((2,)), seed=3, nstep=16, batch_size=32, n_epochs=2, hidden_size=32,
          sim_hash=True, ent_coef=0.01)
    g_rnd("rnd_box_small", 4, 8, ref_shim.Box((2,)), seed=4, nstep=32, batch_size=32, n_epochs=3, hidden_size=32,
          int_hidden_size=16, gamma=0.999, max_grad_norm=5)
    g_rnd("rnd_discrete_small", 4, 6, ref_shim.Discrete(3), seed=5, nstep=16, batch_size=16, n_epochs=2,
          hidden_size=32, int_hidden_size=16)
    g_icm("icm_box_small", 4, 8, ref_shim.Box((2,)), seed=6, nstep=32, batch_size=32, n_epochs=2, hidden_size=32,
          int_hidden_size=16)
    g_icm("icm_discrete_small", 4, 6, ref_shim.Discrete(3), seed=7, nstep=16, batch_size=16, n_epochs=2,
          hidden_size=32, int_hidden_size=16)
