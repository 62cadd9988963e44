"""Shard-boundary collectives for the learner hot path.

One process per GPU; torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) is
plumbing only.  Envs / population members shard across ranks with no data-path collective; the
exchanges are exactly the points where the reference's arithmetic is global:
  * SimHash: all-gather of the packed codes so every rank replays the identical, env-ordered
    count-table update (bit-identical to one GPU);
  * PPO update: all-gather of (n, mean, M2) for the advantage normalisation, one 256-byte all-reduce
    of the loss partial sums (max-of-means value branch), one flat-gradient all-reduce before Adam;
  * ES: all-gather of fitness values, every rank applies the same update from the shared noise table.
"""
import numpy as np
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def all_reduce_sum_(t):
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_gather_into(out, t):
    """out [W, ...] <- t [...] from every rank (static buffers: capturable in a CUDA graph)."""
    if world_size() == 1:
        out[0].copy_(t)
    else:
        dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1))
    return out


def all_gather_cat(t):
    """[n, ...] per rank -> [W, n, ...] (same n on every rank)."""
    W = world_size()
    if W == 1:
        return t.unsqueeze(0)
    t = t.contiguous()
    parts = [torch.empty_like(t) for _ in range(W)]
    dist.all_gather(parts, t)
    return torch.stack(parts)


def merge_mean_std(stats_local, n_local):
    """Global mean / unbiased std from per-rank {mean, std(ddof=1)} of n_local samples each (n may differ).
    stats_local: f64 tensor [2] on the compute device.  Returns f64 [2] (device), merged in rank order."""
    W = world_size()
    if W == 1:
        return stats_local
    n = torch.tensor([float(n_local)], dtype=torch.float64, device=stats_local.device)
    m2 = stats_local[1:2] ** 2 * (n - 1.0)
    rec = torch.cat([n, stats_local[0:1], m2])
    allr = all_gather_cat(rec)                                    # [W, 3]
    ns, means, m2s = allr[:, 0], allr[:, 1], allr[:, 2]
    tot = ns.sum()
    mean = (ns * means).sum() / tot
    M2 = (m2s + ns * (means - mean) ** 2).sum()
    return torch.stack([mean, torch.sqrt(M2 / (tot - 1.0))])


def owned_slice(global_idx, T, n_local, r):
    """Rows of a global minibatch that live on rank r.  global_idx: numpy int64 flat indices over
    [T, W*n_local] in the reference's env-major flatten; returns LOCAL flat indices (numpy int64)."""
    env = global_idx // T
    mine = (env // n_local) == r
    g = global_idx[mine]
    return (g // T - r * n_local) * T + g % T


def interleave_env_shards(x):
    """[W, T, n_local, ...] gathered per-rank blocks -> [T, W*n_local, ...] in global env order."""
    W, T, n = x.shape[:3]
    return x.transpose(0, 1).reshape(T, W * n, *x.shape[3:])
