"""GPU parity: SimHash codes, sequential-semantics count table and bonus (bit-exact integer work)."""
import numpy as np
import pytest
import torch

from conftest import Golden
from oracle import rollout as OR

pytestmark = pytest.mark.gpu


def _buf(k, D, N, steps=1):
    import ppo_exploration_b200 as ppx
    np.random.seed(0)
    return ppx.RolloutStorage(steps, N, ppx.Box((D,)), ppx.Box((2,)), sim_hash=True, hash_bits=k, table_capacity=1024)


@pytest.mark.parametrize("tag,k", [("k16", 16), ("k64", 64), ("k8dup", 8)])
def test_simhash_golden_per_step(tag, k):
    g = Golden("simhash").group(tag)
    steps, N, D = g["obs"].shape
    buf = _buf(k, D, N)
    buf.A = g["A"]
    for t in range(steps):
        r = g["rew_in"][t].copy()
        out = buf.sim_hash(g["obs"][t], r)
        assert out is r and np.array_equal(r, g["rew_out"][t]), t           # bonuses land on the right envs, bit-exact
        codes = buf.sim_hash_codes(g["obs"][t]).cpu().numpy().view(np.uint64)
        assert np.array_equal(codes, OR.pack_bits(OR.simhash_bits(g["A"], g["obs"][t])))
    want = {int(OR.pack_bits(b[None].astype(int))[0]): int(c) for b, c in zip(g["table_bits"], g["table_counts"])}
    assert buf.count_table.items() == want
    assert len(buf.count_table) == len(want)


@pytest.mark.parametrize("tag,k", [("k16", 16), ("k64", 64)])
def test_simhash_golden_whole_rollout(tag, k):
    """One launch over [T,N,D] == the reference's T sequential calls (order is t-major, env-minor)."""
    g = Golden("simhash").group(tag)
    steps, N, D = g["obs"].shape
    buf = _buf(k, D, N, steps)
    buf.A = g["A"]
    r = torch.tensor(g["rew_in"]).cuda()
    buf.sim_hash(torch.tensor(g["obs"]).cuda(), r)
    assert np.array_equal(r.cpu().numpy(), g["rew_out"])


def test_simhash_f64_rewards():
    g = Golden("simhash").group("f64")
    buf = _buf(16, 4, 16)
    buf.A = g["A"]
    r = g["rew_in"].copy()
    buf.sim_hash(g["obs"], r)
    assert r.dtype == np.float64 and np.array_equal(r, g["rew_out"])


@pytest.mark.parametrize("n,k,D,calls", [(5000, 8, 3, 3), (2048, 4, 2, 2), (2049, 6, 2, 1), (1, 16, 5, 4), (20000, 64, 8, 2),
                                          (4096, 5, 2, 2), (4097, 12, 3, 2), (50000, 3, 2, 2), (1200000, 10, 4, 1)])
def test_count_update_heavy_collisions_vs_oracle(n, k, D, calls):
    """Many duplicates inside a batch, across partition tiles / buckets / launches (n > 2^20) and across calls; skewed
    buckets (8 distinct codes over 50 000 elements) run as sequential batches of one CTA; table grows from 1024 slots."""
    rs = np.random.RandomState(n + k)
    buf = _buf(k, D, n)
    table = {}
    for c in range(calls):
        obs = rs.randn(n, D).astype(np.float32)
        codes = buf.sim_hash_codes(obs)
        want_codes = OR.pack_bits(OR.simhash_bits(buf.A, obs))
        assert np.array_equal(codes.cpu().numpy().view(np.uint64), want_codes)
        counts = buf.count_table.update_codes(codes).cpu().numpy().astype(np.uint32)
        assert np.array_equal(counts, OR.count_update_codes(table, want_codes)), c
    assert buf.count_table.items() == {int(a): b for a, b in table.items()}


def test_all_ones_code_and_growth():
    import ppo_exploration_b200 as ppx
    tab = ppx.CountTable(1024)
    rs = np.random.RandomState(3)
    codes = rs.randint(0, 2 ** 62, size=6000, dtype=np.int64).astype(np.uint64)
    codes[::7] = np.uint64(0xFFFFFFFFFFFFFFFF)                          # the EMPTY sentinel is a valid key
    codes[1::7] = np.uint64(0)
    ref = {}
    want = OR.count_update_codes(ref, codes)
    got = tab.update_codes(torch.as_tensor(codes.view(np.int64)).cuda()).cpu().numpy().astype(np.uint32)
    assert np.array_equal(got, want)
    assert tab.items() == {int(a): b for a, b in ref.items()} and len(tab) == len(ref)
    tab.clear()
    assert len(tab) == 0


def test_full_size_c2_checksum():
    """C2 size (2048 envs x 256 steps, D=8, k=64) in one launch: sum(counts) identity and idempotent table."""
    T, N, D, k = 256, 2048, 8, 64
    rs = np.random.RandomState(0)
    obs = (rs.randn(T * N, D) * 0.5).astype(np.float32)
    obs[::3] = obs[0]                                                   # a third of all visits share one state
    buf = _buf(k, D, N, T)
    codes = buf.sim_hash_codes(obs)
    counts = buf.count_table.update_codes(codes).cpu().numpy().astype(np.int64)
    table = buf.count_table.items()
    # each key contributes 1+2+...+c to sum(counts)
    assert counts.sum() == sum(c * (c + 1) // 2 for c in table.values())
    assert sum(table.values()) == T * N
    ref = {}
    want = OR.count_update_codes(ref, codes.cpu().numpy().view(np.uint64)[:50000])
    assert np.array_equal(counts[:50000].astype(np.uint32), want)


@pytest.mark.parametrize("D,k", [(8, 64), (13, 32), (16, 64)])
def test_codes_exact_when_projections_are_nearly_zero(D, k):
    """The codes kernel evaluates the projections in f32 and falls back to the reference's f64 chain when the f32 value
    cannot decide the sign.  Adversarial inputs: observations (almost) orthogonal to rows of A, i.e. projections of the
    order of the f32 rounding error and below, of both signs, and exact zeros -- every bit must still equal
    (np.dot(A, obs.T).T > 0) of buffer.py:194."""
    rs = np.random.RandomState(D * 100 + k)
    buf = _buf(k, D, 4)
    A = rs.randn(k, D)
    buf.A = A
    n = 20000
    obs = rs.randn(n, D)
    rows = rs.randint(0, k, size=n)
    a = A[rows]
    obs -= (np.sum(obs * a, axis=1) / np.sum(a * a, axis=1))[:, None] * a           # orthogonal to one row of A (in f64)
    eps = np.concatenate([np.zeros(n // 5), 10.0 ** rs.uniform(-12, -5, size=n - n // 5) * rs.choice([-1, 1], size=n - n // 5)])
    obs += eps[:, None] * a
    obs = obs.astype(np.float32)
    obs[:50] = 0.0                                                                 # all projections exactly zero
    want = OR.pack_bits(OR.simhash_bits(A, obs))
    got = buf.sim_hash_codes(obs).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, want)
    proj = np.abs(np.dot(A, obs.astype(np.float64).T))
    assert (proj < 1e-6).sum() > n // 2, "the test must exercise the near-zero region"
