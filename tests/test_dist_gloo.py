"""World-size-2 gloo tests (CPU) of the shard-boundary host logic in ppo-exploration_b200/dist.py:
the row split of a global minibatch over the ranks, the numpy-state digest that guards the shared permutation stream,
env-shard interleaving for the sharded SimHash update, and the gather / reduce wrappers."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib.util
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        spec = importlib.util.spec_from_file_location("ppx_dist", os.path.join(root, "ppo-exploration_b200", "dist.py"))
        D = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(D)
        assert D.world_size() == world and D.rank() == rank
        rs = np.random.RandomState(0)
        T, n_local = 16, 6
        # ---- global_slice: every row of a global minibatch lands on exactly one rank, contiguous, in rank order ----
        for bg in (50, 51, 4, 131072 * world):
            spans = [D.global_slice(bg, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(b for _, b in spans) == bg
            assert all(spans[r][0] + spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            assert max(b for _, b in spans) - min(b for _, b in spans) <= 1
        lo, b = D.global_slice(101, world, rank)
        counts = D.all_gather_cat(torch.tensor([b]))
        assert int(counts.sum()) == 101
        # ---- rng_digest: equal states agree across ranks, a rank that drew on its own is found out ----
        np.random.seed(7)
        same = D.all_gather_cat(torch.tensor([D.rng_digest(np.random.get_state())], dtype=torch.int64))
        assert bool((same == same[0]).all())
        if rank == 1:
            np.random.rand(3)
        diff = D.all_gather_cat(torch.tensor([D.rng_digest(np.random.get_state())], dtype=torch.int64))
        assert not bool((diff == diff[0]).all())
        # ---- interleave_env_shards: gathered per-rank [T, n_local] blocks -> global env order ----
        full = rs.randn(T, n_local * world).astype(np.float32)
        block = torch.tensor(full[:, rank * n_local:(rank + 1) * n_local])
        glob = D.interleave_env_shards(D.all_gather_cat(block))
        assert np.array_equal(glob.numpy(), full)
        # ---- all_reduce_sum_ / all_gather_into ----
        x = torch.full((4,), float(rank + 1), dtype=torch.float64)
        D.all_reduce_sum_(x)
        assert torch.equal(x, torch.full((4,), 3.0, dtype=torch.float64))
        out = torch.zeros(world, 3, dtype=torch.float64)
        D.all_gather_into(out, torch.tensor([rank, rank + 10.0, rank + 20.0], dtype=torch.float64))
        assert out[:, 0].tolist() == [0.0, 1.0] and out[:, 2].tolist() == [20.0, 21.0]
        q.put((rank, "ok"))
    except Exception as e:                      # surface the failure in the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_shard_boundary_helpers_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
    assert res == {0: "ok", 1: "ok"}, res
