"""GPU parity: GAE / dual GAE / discounted-return scans vs the reference fixtures and the oracle."""
import numpy as np
import pytest
import torch

from conftest import Golden
from oracle import rollout as OR

pytestmark = pytest.mark.gpu


def _spaces(D=3, A=2):
    import ppo_exploration_b200 as ppx
    return ppx.Box((D,)), ppx.Box((A,))


def _run_single(rew, val, msk, lv, dones, gamma, lam):
    import ppo_exploration_b200 as ppx
    T, N = rew.shape
    o, a = _spaces()
    buf = ppx.RolloutStorage(T, N, o, a, gae_lam=lam, gamma=gamma)
    buf.load_rollout(rewards=rew, values=val, masks=msk)
    buf.compute_returns_and_advantages(torch.tensor(lv), dones)
    torch.cuda.synchronize()
    return buf.advantages.cpu().numpy(), buf.returns.cpu().numpy()


def _close(got, want, tag):
    # fp32 outputs of an f64-carried scan: 1e-5 relative (north_star), with an absolute floor of 1e-6 x scale
    scale = max(1.0, float(np.abs(want).max()))
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6 * scale, err_msg=tag)


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_gae_single_golden(tag):
    g = Golden("gae_single").group(tag)
    adv, ret = _run_single(g["rewards"], g["values"], g["masks"], g["last_value"], g["dones"], g["hp"][0], g["hp"][1])
    _close(adv, g["adv"], "adv"); _close(ret, g["ret"], "ret")
    assert (adv == g["adv"]).mean() > 0.99          # the f64 scan re-associates; after f32 rounding nearly all bits agree


@pytest.mark.parametrize("tag", ["a", "b"])
def test_gae_dual_golden(tag):
    import ppo_exploration_b200 as ppx
    g = Golden("gae_dual").group(tag)
    T, N = g["rewards"].shape
    o, a = _spaces()
    buf = ppx.IntrinsicStorage(T, N, o, a, gae_lam=g["hp"][2], gamma=g["hp"][0], int_gamma=g["hp"][1])
    buf.load_rollout(rewards=g["rewards"], values=g["values"], masks=g["masks"], int_rewards=g["int_rewards"],
                     int_values=g["int_values"])
    mean_int = buf.compute_returns_and_advantages(torch.tensor(g["last_value"]), torch.tensor(g["last_int_value"]), g["dones"])
    for name, key in (("advantages", "adv"), ("returns", "ret"), ("int_advantages", "int_adv"), ("int_returns", "int_ret")):
        _close(getattr(buf, name).cpu().numpy(), g[key], key)
    np.testing.assert_allclose(float(mean_int), g["int_rewards"].mean(), rtol=1e-6)


@pytest.mark.parametrize("T,N,dp", [(1, 1, 0.0), (2, 7, 0.5), (31, 8, 0.1), (32, 33, 0.02), (33, 200, 0.02), (64, 5, 1.0),
                                    (65, 16, 0.0), (300, 40, 0.05), (256, 2048, 0.02), (2048, 4, 0.01), (128, 8, 0.02)])
def test_gae_single_vs_oracle(T, N, dp):
    rs = np.random.RandomState(T * 1000 + N)
    rew, val = rs.randn(T, N).astype(np.float32), rs.randn(T, N).astype(np.float32)
    msk = (rs.rand(T, N) < dp).astype(np.uint8)
    lv = rs.randn(N).astype(np.float32)
    dones = msk[-1].copy()
    want_adv, want_ret = OR.gae(rew, val, msk.astype(np.int64), lv, dones, 0.999, 0.95)
    adv, ret = _run_single(rew, val, msk, lv, dones, 0.999, 0.95)
    _close(adv, want_adv, "adv"); _close(ret, want_ret, "ret")


@pytest.mark.parametrize("T,N", [(128, 128), (37, 9), (512, 3)])
def test_gae_dual_vs_oracle(T, N):
    import ppo_exploration_b200 as ppx
    rs = np.random.RandomState(T + N)
    f = lambda: rs.randn(T, N).astype(np.float32)
    rew, val, irew, ival = f(), f(), np.abs(f()), f()
    msk = (rs.rand(T, N) < 0.02).astype(np.uint8)
    lv, liv = rs.randn(N).astype(np.float32), rs.randn(N).astype(np.float32)
    want = OR.gae_dual(rew, val, msk.astype(np.int64), lv, msk[-1], 0.999, 0.95, irew, ival, liv, 0.99)
    o, a = _spaces()
    buf = ppx.IntrinsicStorage(T, N, o, a, gae_lam=0.95, gamma=0.999, int_gamma=0.99)
    buf.load_rollout(rewards=rew, values=val, masks=msk, int_rewards=irew, int_values=ival)
    buf.compute_returns_and_advantages(lv, liv, msk[-1])
    for name, w in zip(("advantages", "returns", "int_advantages", "int_returns"), want):
        _close(getattr(buf, name).cpu().numpy(), w, name)


def test_gae_linearity_full_size():
    """Size-independent property at C2 size: with no terminals the scan is linear in the rewards."""
    T, N = 256, 2048
    rs = np.random.RandomState(0)
    r1, r2 = rs.randn(T, N).astype(np.float32), rs.randn(T, N).astype(np.float32)
    z = np.zeros((T, N), np.float32)
    m = np.zeros((T, N), np.uint8)
    lv = np.zeros(N, np.float32)
    a1, _ = _run_single(r1, z, m, lv, m[-1], 0.99, 0.95)
    a2, _ = _run_single(r2, z, m, lv, m[-1], 0.99, 0.95)
    a12, _ = _run_single(r1 + r2, z, m, lv, m[-1], 0.99, 0.95)
    np.testing.assert_allclose(a12, a1 + a2, rtol=1e-5, atol=1e-5)


def test_discount_with_dones_bit_exact():
    import ppo_exploration_b200 as ppx
    g = Golden("discount")
    out = ppx.discount_with_dones(g["rewards"], g["dones"], float(g["gamma"])).cpu().numpy()
    assert np.array_equal(out, g["out"])
    rs = np.random.RandomState(1)
    r, d = rs.randn(50, 7), (rs.rand(50, 7) < 0.1)
    out = ppx.discount_with_dones(r, d, 0.99).cpu().numpy()
    assert np.array_equal(out, OR.discount_with_dones(r, d, 0.99))


@pytest.mark.parametrize("T,N,dp", [(16, 65536, 0.05), (37, 70001, 0.02)])
def test_gae_wide_rollout_bit_exact(T, N, dp):
    """N >= 64k columns takes the thread-per-column kernel: the reference's own operation order -> bit-exact."""
    import ppo_exploration_b200 as ppx
    rs = np.random.RandomState(N)
    f = lambda: rs.randn(T, N).astype(np.float32)
    rew, val, irew, ival = f(), f(), np.abs(f()), f()
    msk = (rs.rand(T, N) < dp).astype(np.uint8)
    lv, liv = rs.randn(N).astype(np.float32), rs.randn(N).astype(np.float32)
    want_adv, want_ret = OR.gae(rew, val, msk.astype(np.int64), lv, msk[-1], 0.999, 0.95)
    adv, ret = _run_single(rew, val, msk, lv, msk[-1], 0.999, 0.95)
    assert np.array_equal(adv, want_adv) and np.array_equal(ret, want_ret)
    want = OR.gae_dual(rew, val, msk.astype(np.int64), lv, msk[-1], 0.999, 0.95, irew, ival, liv, 0.99)
    o, a = _spaces()
    buf = ppx.IntrinsicStorage(T, N, o, a, gae_lam=0.95, gamma=0.999, int_gamma=0.99)
    buf.load_rollout(rewards=rew, values=val, masks=msk, int_rewards=irew, int_values=ival)
    buf.compute_returns_and_advantages(lv, liv, msk[-1])
    for name, w in zip(("advantages", "returns", "int_advantages", "int_returns"), want):
        assert np.array_equal(getattr(buf, name).cpu().numpy(), w), name
