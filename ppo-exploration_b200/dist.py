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


def global_slice(bg, W, r):
    """Rows [lo, lo + b) of a global minibatch of bg rows that rank r of W processes ("global" shard mode): contiguous,
    in rank order, sizes differing by at most one.  Returns (lo, b)."""
    base, rem = divmod(int(bg), int(W))
    return r * base + min(r, rem), base + (1 if r < rem else 0)


def rng_digest(state):
    """48-bit digest of a numpy legacy RandomState state tuple (MT19937 key + position): the ranks of a "global" sharded
    learner exchange it to check that they are about to draw the same permutation."""
    import zlib
    import numpy as np
    return (zlib.crc32(np.ascontiguousarray(state[1]).tobytes()) << 16) ^ int(state[2])


def interleave_env_shards(x):
    """[W, T, n_local, ...] gathered per-rank blocks -> [T, W*n_local, ...] in global env order."""
    W, T, n = x.shape[:3]
    return x.transpose(0, 1).reshape(T, W * n, *x.shape[3:])


class PeerExchange:
    """Symmetric (peer-mapped) scratch of one sharded learner: the policy gradient vector, the loss partial sums, the
    advantage-moment records and three barrier channels, visible to every rank over NVLink (p2p.cu).  The allocation and
    handle exchange use torch.distributed._symmetric_memory; the kernels are ppx's own."""
    FLAGS, SEQ, STATUS, REC, SUMS, GRAD = 0, 256, 272, 512, 1024, 4096       # byte offsets
    MAXW = 16

    def __init__(self, n_params, device, flag_blocks=0):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        W, r = world_size(), rank()
        assert 2 <= W <= self.MAXW
        n4 = (int(n_params) + 3) // 4 * 4
        # after the gradient vector: the push staging of the fused optimiser tail, 2 parities x W source ranks x n words of
        # {f32 value, u32 sequence} (ppx_fused_adam); flag_blocks > 0 says the tail is available for the policy's shape
        self.XG = self.GRAD + 4 * n4
        self.flag_blocks = int(flag_blocks)
        self.XS = self.XG + 8 * 2 * W * int(n_params)            # the loss head's sums exchange: 2 x W x 64 words (ppx_ppo_cfg)
        self.XM = self.XS + 8 * 2 * W * 64                       # the gather's moments exchange: 2 slots x 2 x W x 8 words (ppx_gather_opts)
        words = self.XM // 4 + 2 * (2 * 2 * W * 8) + 4
        self.buf = symm_mem.empty(words, dtype=torch.float32, device=device)
        self.buf.zero_()
        group = dist.group.WORLD
        self.hdl = symm_mem.rendezvous(self.buf, group.group_name if hasattr(group, "group_name") else group)
        torch.cuda.synchronize(device)
        self.hdl.barrier()                                   # every rank's flags are zero before anyone signals
        base = [int(p) for p in self.hdl.buffer_ptrs]
        arr = lambda off: (C.c_void_p * W)(*[b + off for b in base])
        self.W, self.rank = W, r
        self.peer_grad, self.peer_sums, self.peer_rec = arr(self.GRAD), arr(self.SUMS), arr(self.REC)
        self.peer_flags = [arr(self.FLAGS + 64 * ch) for ch in range(3)]     # 16 uint32 per channel
        w = lambda off, n: self.buf[off // 4: off // 4 + n]
        self.grad = w(self.GRAD, int(n_params))
        self.sums = w(self.SUMS, 64).view(torch.float64)
        self.rec = w(self.REC, 12).view(torch.float64)
        self.seq = [self.buf.data_ptr() + self.SEQ + 4 * ch for ch in range(4)]
        # [1]: the loss head's sums exchange, [3]: the fused optimiser tail; two more words for the gather's moments exchange
        self.seq_moments = self.buf.data_ptr() + self.STATUS + 8
        self.peer_xg, self.peer_xs, self.peer_xm = arr(self.XG), arr(self.XS), arr(self.XM)
        self.status_ptr = self.buf.data_ptr() + self.STATUS
        self.status = w(self.STATUS, 1).view(torch.int32)


def peer_exchange_or_none(n_params, device, max_params, flag_blocks=0):
    """A PeerExchange when every rank can build one (NVLink P2P, symmetric memory, bank small enough), else None --
    decided collectively so all ranks take the same path."""
    import os
    ok, px = 1, None
    if world_size() < 2 or world_size() > PeerExchange.MAXW or os.environ.get("PPX_P2P", "1") == "0" or n_params > max_params \
            or not torch.cuda.is_available():
        ok = 0
    if ok:
        try:
            px = PeerExchange(n_params, device, flag_blocks)
        except Exception:                                     # no P2P / symmetric memory on this box
            ok, px = 0, None
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    if world_size() > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return px if int(flag.item()) == 1 else None
