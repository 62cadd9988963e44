"""GPU parity: ES-NSRA perturb / update / ranks / novelty kNN vs the reference fixtures and the oracle."""
import numpy as np
import pytest
import torch

from conftest import Golden
from oracle import es as OE

pytestmark = pytest.mark.gpu

SHAPES = [(8, 16), (16, 16), (16, 2)]


def _es(g, **kw):
    import ppo_exploration_b200 as ppx
    np.random.seed(0)
    es = ppx.EvolutionStrategy(hidden_sizes=[16, 16], obs_dim=8, n_actions=2, population_size=24, **kw)
    es.set_weights([g[f"w0/{i}"] for i in range(3)])
    return es


def _flat(g, pfx):
    return np.concatenate([g[f"{pfx}/{i}"].ravel() for i in range(3)])


def test_es_golden_update_sequence():
    g = Golden("es")
    es = _es(g)
    eps = g["eps"].astype(np.float32)
    assert np.array_equal(eps.astype(np.float64), g["eps"])
    # _get_weights_try: bit-exact (separately rounded product and sum, f64)
    assert np.array_equal(es.perturb_all(eps, out_f64=True)[3].cpu().numpy(), g["try3"])
    sizes = np.cumsum([0] + [a * b for a, b in SHAPES])
    member3 = [g["eps"][3][sizes[l]:sizes[l + 1]].reshape(SHAPES[l]) for l in range(3)]
    wt = es._get_weights_try(None, member3)
    assert np.array_equal(np.concatenate([w.ravel() for w in wt]), g["try3"])
    es._update_weights(g["rewards"], eps, 0.37)
    np.testing.assert_allclose(es.theta.cpu().numpy(), _flat(g, "w1"), rtol=1e-12, atol=1e-14)
    assert not es.update_skipped and np.isclose(es.learning_rate, float(g["lr1"]), rtol=1e-15)
    es._update_weights(g["rewards"] * 0 + 2.5, eps, 0.37)                      # std == 0: weights and lr untouched
    assert es.update_skipped and es.learning_rate == float(g["lr2"])
    np.testing.assert_allclose(es.theta.cpu().numpy(), _flat(g, "w1"), rtol=1e-12, atol=1e-14)
    es.novelty_param = 0.2
    es._update_weights(g["rewards"][::-1].copy(), eps)                         # novelty=None branch
    np.testing.assert_allclose(es.theta.cpu().numpy(), _flat(g, "w3"), rtol=1e-12, atol=1e-14)
    assert np.isclose(es.learning_rate, float(g["lr3"]), rtol=1e-15)
    # the reference's nested-list population is accepted too
    es2 = _es(g)
    pop = [[e[sizes[l]:sizes[l + 1]].reshape(SHAPES[l]) for l in range(3)] for e in g["eps"]]
    es2._update_weights(g["rewards"], pop, 0.37)
    np.testing.assert_allclose(es2.theta.cpu().numpy(), _flat(g, "w1"), rtol=1e-12, atol=1e-14)


def test_knn_golden_and_oracle():
    g = Golden("es")
    es = _es(g)
    for M in (1, 5, 10, 30, 500):
        s = es.get_kNN(g[f"knn{M}/archive"], g[f"knn{M}/q"], min(10, M))
        np.testing.assert_allclose(s, float(g[f"knn{M}/sum"]), rtol=1e-13)
        assert s == OE.knn_sum(g[f"knn{M}/archive"], g[f"knn{M}/q"], min(10, M))   # bit-exact vs the direct-distance oracle
    rs = np.random.RandomState(0)
    arch = rs.randn(10000, 2)
    qs = rs.randn(64, 2)
    sums, nov = es.novelty_batch(arch, qs)
    for i in range(0, 64, 7):
        assert sums[i].item() == OE.knn_sum(arch, qs[i], 10)
        assert nov[i].item() == OE.novelty(list(arch[:, None, :]), qs[i])
    # floor: a query sitting on >= K identical archive points has novelty 0 -> 5e-3
    same = np.zeros((20, 2))
    _, nv = es.novelty_batch(same, np.zeros((1, 2)))
    assert nv.item() == 5e-3
    # list-of-[1,2] archives like the reference builds them (evolution_strategies.py:365)
    assert es.get_kNN([a[None] for a in arch[:50]], qs[:1], 10) == OE.knn_sum(arch[:50], qs[0], 10)


def test_centered_ranks_bit_exact():
    g = Golden("es")
    es = _es(g)
    rs = np.random.RandomState(1)
    r = rs.randn(10000)
    r[::5] = r[0]                                                              # ties -> index order
    ranks, cen = es.centered_ranks(r)
    want_r, want_c = OE.centered_ranks(r)
    assert np.array_equal(ranks.cpu().numpy(), want_r)
    assert np.array_equal(cen.cpu().numpy(), want_c)


def test_noise_table_population_matches_dense():
    import ppo_exploration_b200 as ppx
    np.random.seed(0)
    es = ppx.EvolutionStrategy(hidden_sizes=[64, 64], obs_dim=8, n_actions=2, population_size=300, noise_table_size=1 << 22)
    assert es.D == 4736
    table = es.noise_table()
    t = table.cpu().numpy()
    assert abs(t.mean()) < 5e-3 and abs(t.std() - 1) < 5e-3 and np.isfinite(t).all()
    k = ((t - t.mean()) ** 4).mean() / t.var() ** 2
    assert abs(k - 3) < 0.05                                                   # gaussian kurtosis
    es_b = ppx.EvolutionStrategy(hidden_sizes=[64, 64], obs_dim=8, n_actions=2, population_size=300, noise_table_size=1 << 22)
    assert torch.equal(es_b.noise_table(), table)                              # counter-based: reproducible
    off = es._get_population()
    assert off.dtype == torch.int64 and (off % 4 == 0).all() and off.max() + es.D <= table.numel()
    dense = torch.stack([table[o:o + es.D] for o in off.tolist()])
    a = es.perturb_all(off)
    b = es.perturb_all(dense.cpu().numpy())
    assert torch.equal(a, b)
    theta0 = es.theta.clone()
    rewards = np.random.randn(300)
    es._update_weights(rewards, off, 0.5)
    th_a = es.theta.clone()
    es.theta.copy_(theta0); es.learning_rate = 0.01
    es._update_weights(rewards, dense.cpu().numpy(), 0.5)
    assert torch.equal(th_a, es.theta)
    # vs oracle in f64
    w = [theta0.cpu().numpy()[:512].reshape(8, 64), theta0.cpu().numpy()[512:4608].reshape(64, 64), theta0.cpu().numpy()[4608:].reshape(64, 2)]
    d = dense.cpu().numpy().astype(np.float64)
    pop = [[e[:512].reshape(8, 64), e[512:4608].reshape(64, 64), e[4608:].reshape(64, 2)] for e in d]
    w1, lr1 = OE.update_weights(w, rewards, pop, 0.01, 0.1, 0.5, 0.9995, novelty=0.5)
    np.testing.assert_allclose(th_a.cpu().numpy(), np.concatenate([x.ravel() for x in w1]), rtol=1e-12, atol=1e-14)


def test_centered_rank_update_mode():
    import ppo_exploration_b200 as ppx
    np.random.seed(2)
    es = ppx.EvolutionStrategy(hidden_sizes=[16], obs_dim=4, n_actions=2, population_size=64, fitness_shaping="centered_rank")
    eps = np.random.randn(64, es.D).astype(np.float32)
    r = np.random.randn(64)
    th0 = es.theta.cpu().numpy().copy()
    es._update_weights(r, eps)
    _, c = OE.centered_ranks(r)
    want = th0 + 0.01 / (64 * 0.1) * (eps.astype(np.float64).T @ c)
    np.testing.assert_allclose(es.theta.cpu().numpy(), want, rtol=1e-12, atol=1e-14)


def test_predict_matches_reference_fixture():
    """Population forward with sigma-free weights == FeedForwardNetwork.predict of the reference (Box: tanh(logits);
    Discrete: logits on the device, the categorical draw on the host RNG -> the reference's actions under its seed)."""
    import ppo_exploration_b200 as ppx
    g = Golden("es_predict")
    for tag, disc, nact in (("box", False, 2), ("disc", True, 3)):
        np.random.seed(0)
        es = ppx.EvolutionStrategy(hidden_sizes=[16, 16], obs_dim=8, n_actions=nact, population_size=6)
        es.set_weights([g[f"{tag}/w/{i}"] for i in range(3)])
        out = es.predict(g[f"{tag}/obs"], discrete=disc)
        if disc:
            np.random.seed(31)
            assert np.array_equal(ppx.EvolutionStrategy.discrete_action(out).reshape(6, 1), g[f"{tag}/actions"])
        else:
            np.testing.assert_allclose(out.cpu().numpy(), g[f"{tag}/actions"], rtol=1e-5, atol=1e-6)      # fp32 arithmetic on the device


@pytest.mark.parametrize("P,hidden,table", [(1, [16, 16], False), (37, [64, 64], False), (1000, [64, 64], True), (50, [40, 130, 7], True)])
def test_predict_population_matches_oracle(P, hidden, table):
    """Member p = MLP(theta + sigma*eps_p) on obs[p], weights formed on the fly (dense eps rows or noise-table offsets)."""
    import ppo_exploration_b200 as ppx
    np.random.seed(5)
    es = ppx.EvolutionStrategy(hidden_sizes=hidden, obs_dim=8, n_actions=2, population_size=P, sigma=0.1,
                               noise_table_size=1 << 20)
    rs = np.random.RandomState(6)
    obs = rs.randn(P, 8)
    if table:
        pop = es._get_population()
        eps = es.noise_table()[(pop[:, None] + torch.arange(es.D, device=pop.device)[None, :])].cpu().numpy().astype(np.float64)
    else:
        pop = rs.randn(P, es.D).astype(np.float32)
        eps = pop.astype(np.float64)
    got = es.predict_population(pop, obs).cpu().numpy()
    theta = es.theta.cpu().numpy()
    sizes = np.cumsum([0] + es.layer_sizes)
    for p in range(0, P, max(1, P // 40)):
        w = [(theta[sizes[l]:sizes[l + 1]] + 0.1 * eps[p, sizes[l]:sizes[l + 1]]).reshape(es.shapes[l]) for l in range(len(es.shapes))]
        # fp32 arithmetic on the device: 1e-5 of the output scale (tanh outputs are bounded by 1; the logits behind them
        # are O(10) sums of 64 products, so their fp32 rounding already is ~1e-6 absolute)
        np.testing.assert_allclose(got[p], OE.predict(w, obs[p])[0], rtol=1e-5, atol=1e-5)
