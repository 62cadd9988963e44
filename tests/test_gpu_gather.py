"""GPU parity: shuffle indices and minibatch gather are bit-exact vs the reference fixtures."""
import numpy as np
import pytest
import torch

from conftest import Golden
from oracle import rollout as OR

pytestmark = pytest.mark.gpu


def test_get_single_golden():
    import ppo_exploration_b200 as ppx
    g = Golden("get_single")
    raw = g.group("raw")
    buf = ppx.RolloutStorage(8, 3, ppx.Box((4,)), ppx.Box((2,)))
    buf.load_rollout(**{k: raw[k] for k in ("observations", "actions", "values", "action_log_probs", "advantages",
                                           "returns", "rewards", "masks")})
    np.random.seed(123)
    for ep in range(2):
        batches = list(buf.get(5))
        assert len(batches) == 5 and batches[-1].observations.shape[0] == 4          # ragged last minibatch
        for bi, b in enumerate(batches):
            for f in b._fields:
                want = g[f"ep{ep}/b{bi}/{f}"]
                got = getattr(b, f).cpu().numpy()
                assert got.shape == want.shape and got.dtype == want.dtype and np.array_equal(got, want), (ep, bi, f)
    assert buf.generator_ready
    assert np.array_equal(buf.flat("values").cpu().numpy(), OR.swap_and_flatten(raw["values"]))


def test_get_dual_golden():
    import ppo_exploration_b200 as ppx
    g = Golden("get_dual")
    raw = g.group("raw")
    buf = ppx.IntrinsicStorage(8, 3, ppx.Box((4,)), ppx.Discrete(3))
    buf.load_rollout(**raw)
    np.random.seed(77)
    for bi, b in enumerate(buf.get(7)):
        assert b._fields == ('observations', 'actions', 'old_values', 'int_values', 'old_log_probs', 'advantages',
                             'int_advantages', 'returns', 'int_returns')
        for f in b._fields:
            want = g[f"b{bi}/{f}"]
            got = getattr(b, f).cpu().numpy()
            assert got.shape == want.shape and np.array_equal(got, want), (bi, f)


def test_get_requires_full():
    import ppo_exploration_b200 as ppx
    buf = ppx.RolloutStorage(4, 2, ppx.Box((3,)), ppx.Box((1,)))
    with pytest.raises(AssertionError):
        next(buf.get(2))


@pytest.mark.parametrize("T,N,D,A,B", [(256, 2048, 8, 2, 131072), (16, 5, 28224, 1, 17), (128, 8, 4, 1, 128)])
def test_gather_vs_oracle_sizes(T, N, D, A, B):
    import ppo_exploration_b200 as ppx
    rs = np.random.RandomState(T)
    buf = ppx.RolloutStorage(T, N, ppx.Box((D,)), ppx.Box((A,)))
    arrs = dict(observations=rs.randn(T, N, D).astype(np.float32), actions=rs.randn(T, N, A),
                values=rs.randn(T, N).astype(np.float32), action_log_probs=rs.randn(T, N, A).astype(np.float32),
                advantages=rs.randn(T, N).astype(np.float32), returns=rs.randn(T, N).astype(np.float32))
    buf.load_rollout(**arrs)
    np.random.seed(5)
    it = buf.get(B)
    b0 = next(it)
    np.random.seed(5)
    perm = OR.epoch_permutation(T, N)
    want = OR.gather_single(arrs, perm[:B])
    for f in b0._fields:
        assert np.array_equal(getattr(b0, f).cpu().numpy(), want[f]), f
    # a permutation gathers every row exactly once: checksum over the whole epoch
    tot = b0.returns.double().sum().item() + sum(b.returns.double().sum().item() for b in it)
    np.testing.assert_allclose(tot, arrs["returns"].astype(np.float64).sum(), rtol=1e-9)
