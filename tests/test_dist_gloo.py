"""World-size-2 gloo tests (CPU) of the shard-boundary host logic in ppo-exploration_b200/dist.py:
moment merging for the advantage normalisation, owner-computes index slicing of a global permutation, env-shard
interleaving for the replicated SimHash update, and the gather / reduce wrappers."""
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
        # ---- merge_mean_std: ragged local counts (owner-computes minibatches are ragged) ----
        data = rs.randn(101) * 3 + 1
        cut = 37
        mine = data[:cut] if rank == 0 else data[cut:]
        loc = torch.tensor([mine.mean(), mine.std(ddof=1)], dtype=torch.float64)
        got = D.merge_mean_std(loc, len(mine))
        assert abs(float(got[0]) - data.mean()) < 1e-12 and abs(float(got[1]) - data.std(ddof=1)) < 1e-12
        # ---- owned_slice: every global index lands on exactly one rank, as the right local index ----
        perm = np.random.RandomState(1).permutation(T * n_local * world)
        g = perm[:50]
        loc_idx = D.owned_slice(g, T, n_local, rank)
        env, t = g // T, g % T
        want = [(e - rank * n_local) * T + tt for e, tt in zip(env, t) if e // n_local == rank]
        assert list(loc_idx) == want
        counts = D.all_gather_cat(torch.tensor([len(loc_idx)]))
        assert int(counts.sum()) == len(g)
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
