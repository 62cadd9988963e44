"""PPO / PPO-SimHash / RND / ICM learners with the reference's entry points, on the GPU.

Mirrors algorithms.py of the reference: BaseAlgorithm (:22-118), PPO (:121-308), PPO_RND (:310-543),
PPO_ICM (:546-756).  Constructor keyword names and defaults, `collect_samples()`, `train()` and
`learn()` keep their meaning; hyper-parameter dictionaries from the reference's hyperparameters.py
splat into these constructors unchanged.  What differs:
  * the env layer is out of scope: pass `env=` (anything with num_envs,
    observation_space, action_space, reset(), step()) instead of `env_id` only;
  * train() runs the fused path  shuffle-gather -> advantage stats -> batched MLP forward ->
    fused PPO loss fwd+bwd -> MLP backward -> clip+Adam, all in libppx.so, and reads the loss
    scalars back ONCE per train() (the reference syncs 4-5 .item() per minibatch);
  * losses are returned per minibatch in `self.last_losses` ([steps, 5] numpy) and their means in
    `self.train_stats` under the reference's logger keys; pass `logger=` (a module/object with
    .record) to keep logging through the reference's logger.
The host numpy RNG is consumed in exactly the reference's order: one permutation per epoch
(buffer.py:239) and, for RND, one randn() per minibatch (algorithms.py:468).
"""
import ctypes as C
import time
from collections import deque

import numpy as np
import torch

from . import _lib as L
from . import dist as D
from .buffer import (RolloutStorage, IntrinsicStorage, HostRngStream, DevicePartners, device_shuffle_default,
                     rng_states_equal, _dev)
from .models import Policy, RndNetwork, IntrinsicCuriosityModule, ActionConverter, _Scratch
from .util import RunningMeanStd, normalize_obs


class BaseAlgorithm(object):
    """algorithms.py:22-118."""

    def __init__(self, env_id, lr, nstep, batch_size, n_epochs, gamma, gae_lam, clip_range, ent_coef, vf_coef,
                 max_grad_norm, env=None, device="cuda", logger=None):
        if env is None:
            raise ValueError("ppx accelerates the learner hot path only: construct the vectorised env yourself "
                             "and pass env=... (the reference builds it from env_id, algorithms.py:52)")
        self.env_id = env_id
        self.env = env
        self.device = torch.device(device)
        self.num_envs = env.num_envs
        self.state_dim = env.observation_space.shape[0]
        self.action_converter = ActionConverter(env.action_space)
        self.discrete = self.action_converter.action_type == "Discrete"
        self.lr, self.nstep, self.batch_size, self.n_epochs = lr, nstep, batch_size, n_epochs
        self.gamma, self.gae_lam, self.clip_range = gamma, gae_lam, clip_range
        self.ent_coef, self.vf_coef, self.max_grad_norm = ent_coef, vf_coef, max_grad_norm
        self.ep_info_buffer = deque(maxlen=50)
        self._n_updates = 0
        self.num_timesteps = 0
        self.num_episodes = 0
        self.obs_rms = RunningMeanStd(shape=(self.state_dim,), device=self.device)
        self.logger = logger
        self.train_stats = {}
        self.last_losses = None
        self._scratch = _Scratch(self.device)
        self._stats = torch.zeros(4, dtype=torch.float64, device=self.device)
        self._sums = torch.zeros(32, dtype=torch.float64, device=self.device)
        self._branch = torch.zeros(4, dtype=torch.float64, device=self.device)
        self.scale_batch_with_world = True     # sharded runs: batch_size is per rank (weak scaling)
        # sharded minibatch composition: "global" = every rank draws the SAME permutation over the global
        # [T, W*N] index space and keeps the rows it owns (bit-identical to one GPU holding all envs; host cost
        # grows with W); "local" = every rank shuffles its own rollout with its own numpy stream (the reference's
        # buffer semantics per rank; static shapes -> CUDA graphs, scales)
        self.shard_shuffle = "global"
        self._mrec = None
        self._spec = None                      # (HostRngStream, rng snapshot) pre-drawn for the next train() call
        self.speculative_shuffle = True
        self._px = False                       # PeerExchange (NVLink peer-memory kernels) | None; False = not probed yet
        self.use_cuda_graph = True             # replay the per-minibatch launch sequence as one CUDA graph
        self._graphs = {}
        self._loss_row = torch.zeros(8, dtype=torch.float64, device=self.device)
        self._perm_bufs, self._perm_ready, self._perm_free, self._copy_stream = None, [None, None], [None, None], None
        self._perm_j, self._perm_ws = None, None
        self.device_shuffle = device_shuffle_default()         # swaps of the epoch shuffle applied on the GPU (shuffle_dev.cu)

    def __del__(self):
        sp = getattr(self, "_spec", None)
        if sp is not None:
            sp[0].cancel()

    # ---- shared pieces of the fused update --------------------------------------------------
    def _record(self, key, value):
        self.train_stats[key] = value
        if self.logger is not None:
            self.logger.record(key, value)

    def update_info_buffer(self, infos, dones=None):
        for info in infos:
            ep = info.get('episode')
            if ep is not None:
                self.ep_info_buffer.extend([ep])

    def _loss_workspace(self):
        return self._scratch.get("ppo_loss_ws", L.call("ppx_ppo_loss_workspace", 0, 0) // 8 + 1, torch.float64)

    def _gather_with_stats(self, ro, sl, bufs, dual=False):
        """Minibatch gather; the advantage moments (algorithms.py:219, :431-434) come out of the same launch."""
        stats = [('advantages', self._stats.data_ptr())] + ([('int_advantages', self._stats.data_ptr() + 16)] if dual else [])
        ro.gather_into(sl, bufs, stats=stats if sl.numel() >= 2 else None)
        return sl.numel() >= 2

    def _policy_step(self, bufs, B, losses_row, dual=False, policy_weight=1.0, int_vf_coef=0.0, B_total=0, stats_ready=False):
        """One minibatch: forward, fused loss fwd+bwd, backward.  Gradients land in policy.bank.grad.
        Sharded runs (B_total = rows of the global minibatch) exchange only the advantage moments and the
        32 loss partial sums; the caller all-reduces the flat gradient."""
        pol, sc = self.policy, self._scratch
        A = pol.action_dim
        obs = bufs['observations'][:B]
        outs = pol.forward_raw(obs)
        adv = bufs['advantages'][:B]
        sharded = D.world_size() > 1
        if not stats_ready:
            L.call("ppx_mean_std", adv.data_ptr(), B, self._stats.data_ptr(), L.stream())
            if dual:
                iadv = bufs['int_advantages'][:B]
                L.call("ppx_mean_std", iadv.data_ptr(), B, self._stats.data_ptr() + 16, L.stream())
        if sharded:
            self._merge_stats(B, dual)
        d_actor = sc.get("d_actor", B * A)[:B * A].view(B, A)
        cfg = L.PpoCfg(B, int(B_total), A, int(self.discrete), int(dual), float(self.clip_range), float(self.ent_coef),
                       float(self.vf_coef), float(int_vf_coef), float(policy_weight))
        g = lambda k: bufs[k][:B].data_ptr() if k in bufs else None
        ws = self._loss_workspace().data_ptr()
        head_args = (C.byref(cfg), outs[0].data_ptr(), pol.bank.p("action_log_std"), g('actions'),
                     g('old_log_probs'), adv.data_ptr(), self._stats.data_ptr(), outs[1].data_ptr(), g('old_values'),
                     g('returns'), g('int_advantages'), self._stats.data_ptr() + 16,
                     outs[2].data_ptr() if dual else None, g('int_values'), g('int_returns'), d_actor.data_ptr())
        if pol.mlp.fused():
            # head + partial sums (+ loss scalars / branch when single-GPU) in ONE launch; the value-head gradients are
            # evaluated inside the fused MLP backward from the branch weights.  Sharded: the 32 partial sums are
            # all-reduced (256 bytes) before the finalize kernel so the max-of-means branch is the global one.
            Bt = int(B_total) if B_total else B
            if not sharded:
                L.call("ppx_ppo_loss_head_final", *head_args, pol.bank.g("action_log_std"), losses_row,
                       self._branch.data_ptr(), ws, L.stream())
            else:
                px = self._peer_exchange()
                L.call("ppx_ppo_loss_head", *head_args, self._sums.data_ptr(), ws, L.stream())
                if px is not None:
                    L.call("ppx_p2p_sums_allreduce", px.peer_sums, px.peer_flags[1], px.W, px.rank, px.seq[1], px.status_ptr,
                           self._sums_global.data_ptr(), L.stream())
                    gsums = self._sums_global
                else:
                    D.all_reduce_sum_(self._sums)
                    gsums = self._sums
                L.call("ppx_ppo_loss_finalize", C.byref(cfg), gsums.data_ptr(), pol.bank.p("action_log_std"),
                       pol.bank.g("action_log_std"), losses_row, self._branch.data_ptr(), L.stream())
            vh = {1: (outs[1], bufs['old_values'][:B], bufs['returns'][:B], self._branch.data_ptr(),
                      float(policy_weight) * float(self.vf_coef))}
            if dual:
                vh[2] = (outs[2], bufs['int_values'][:B], bufs['int_returns'][:B], self._branch.data_ptr() + 16,
                         float(int_vf_coef))
            # single GPU with clipping: the reduce kernel also leaves the clip_grad_norm_ partials (no sumsq launch)
            self._pre_sumsq = (not sharded) and self.max_grad_norm > 0
            pol.mlp.backward([d_actor, None] + ([None] if dual else []), value_heads=vh, clip_range=self.clip_range,
                             B_total=Bt, with_sumsq=self._pre_sumsq)
            return
        self._pre_sumsq = False
        d_val = sc.get("d_val", B)[:B].view(B, 1)
        d_ival = sc.get("d_ival", B)[:B].view(B, 1)
        L.call("ppx_ppo_loss_head", *head_args, self._sums.data_ptr(), ws, L.stream())
        if sharded:
            D.all_reduce_sum_(self._sums)
        L.call("ppx_ppo_loss_finish", C.byref(cfg), self._sums.data_ptr(), pol.bank.p("action_log_std"),
               outs[1].data_ptr(), g('old_values'), g('returns'), outs[2].data_ptr() if dual else None,
               g('int_values'), g('int_returns'), pol.bank.g("action_log_std"), d_val.data_ptr(),
               d_ival.data_ptr() if dual else None, losses_row, ws, L.stream())
        pol.mlp.backward([d_actor, d_val] + ([d_ival] if dual else []))

    def _peer_exchange(self):
        """Lazily move the policy bank's gradient vector, the loss partial sums and the moment records into symmetric
        peer memory so the three per-minibatch exchanges run as ppx kernels over NVLink (p2p.cu) instead of NCCL."""
        if self._px is False:
            self._px = None
            if D.world_size() > 1 and self.policy.mlp.fused():
                bank = self.policy.bank
                px = D.peer_exchange_or_none(bank.size, self.device, L.call("ppx_p2p_max_params"))
                if px is not None:
                    px.grad.copy_(bank.grad)
                    bank.grad = px.grad                         # kernels now write the local gradient into peer-visible memory
                    self.policy.mlp._fa = None                  # cached gradient pointers are stale
                    self._sums = px.sums[:32]
                    self._sums_global = torch.zeros(32, dtype=torch.float64, device=self.device)
                    self._px = px
        return self._px

    def _merge_stats(self, B, dual):
        """Sharded minibatch: local {mean, std} -> global, through one exchange of {n, mean, M2} records."""
        W = D.world_size()
        px = self._peer_exchange()
        if px is not None:
            L.call("ppx_moments_pack", self._stats.data_ptr(), B, px.rec.data_ptr(), L.stream())
            if dual:
                L.call("ppx_moments_pack", self._stats.data_ptr() + 16, B, px.rec.data_ptr() + 24, L.stream())
            L.call("ppx_p2p_moments_merge", px.peer_rec, px.peer_flags[0], W, px.rank, px.seq[0], px.status_ptr,
                   2 if dual else 1, self._stats.data_ptr(), L.stream())
            return
        if self._mrec is None or self._mrec_all.shape[0] != W:
            self._mrec = torch.zeros(6, dtype=torch.float64, device=self.device)
            self._mrec_all = torch.zeros(W, 6, dtype=torch.float64, device=self.device)
        L.call("ppx_moments_pack", self._stats.data_ptr(), B, self._mrec.data_ptr(), L.stream())
        if dual:
            L.call("ppx_moments_pack", self._stats.data_ptr() + 16, B, self._mrec.data_ptr() + 24, L.stream())
        D.all_gather_into(self._mrec_all, self._mrec)
        # records of one stream are strided by 6 doubles in the gathered buffer: merge from a compact copy
        a = self._mrec_all[:, 0:3].contiguous()
        L.call("ppx_moments_merge", a.data_ptr(), W, self._stats.data_ptr(), L.stream())
        if dual:
            b = self._mrec_all[:, 3:6].contiguous()
            L.call("ppx_moments_merge", b.data_ptr(), W, self._stats.data_ptr() + 16, L.stream())

    def _graph_call(self, key, fn):
        """Run fn() -- a fixed sequence of libppx launches (and, when sharded with local shuffles, NCCL collectives)
        on static buffers -- through a CUDA graph: eager the first time a key is seen (allocations settle),
        captured the second time, replayed afterwards."""
        if not self.use_cuda_graph or (D.world_size() > 1 and self.shard_shuffle != "local"):
            return fn()
        ent = self._graphs.get(key)
        if ent is None:
            self._graphs[key] = "warm"
            return fn()
        if ent == "warm":
            g = torch.cuda.CUDAGraph()
            n0 = L.launch_count()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):   # the RNG worker thread may pin memory meanwhile
                fn()
            ent = (g, L.launch_count() - n0)
            L.extra_launches -= ent[1]                      # the capture pass launched nothing
            self._graphs[key] = ent
        ent[0].replay()
        L.extra_launches += ent[1]

    def _rng_script(self, ro, randn_per_minibatch=False):
        W = D.world_size() if self.shard_shuffle != "local" else 1
        total = ro.buffer_size * ro.n_envs * W
        Bg = min(self.batch_size * (W if self.scale_batch_with_world else 1), total)
        n_mb = -(-total // Bg)
        script = []
        for _ in range(self.n_epochs):
            script.append(('perm', total))
            script += [('randn',)] * (n_mb if randn_per_minibatch else 0)
        return script

    def _rng_open(self, script):
        """The RNG stream of this train() call.  If the stream started speculatively at the end of the previous call was
        seeded with exactly the state np.random is in now (nobody drew in between) and has the same script, its
        permutations are already waiting; otherwise it is dropped and a fresh stream starts from the current state.
        Either way the draws are those the reference would make from this state."""
        cur = np.random.get_state()
        sp, self._spec = self._spec, None
        if sp is not None:
            stream, snapshot = sp
            if (stream.script == list(script) and rng_states_equal(cur, snapshot) and stream.err is None
                    and stream.device_apply == self._device_apply()):
                return stream
            stream.cancel()
        return HostRngStream(script, state=cur, device_apply=self._device_apply())

    def _device_apply(self):
        """Device-side swaps need the permutation only on the device: single GPU or per-rank ("local") shuffles."""
        return bool(self.device_shuffle) and (D.world_size() == 1 or self.shard_shuffle == "local")

    def _rng_close(self, rng, speculate=True):
        """Commit the consumed draws to the global numpy RNG and pre-draw the next call's stream from there."""
        final = rng.final_state()
        np.random.set_state(final)
        if speculate and self.speculative_shuffle:
            self._spec = (HostRngStream(rng.script, state=final, device_apply=self._device_apply()), final)

    def _perm_prefetch(self, rng, total, slot):
        """Upload the next epoch's permutation on a side stream into one of two static device buffers, so the
        4 MB H2D copy overlaps the previous epoch's kernels instead of sitting in the compute stream."""
        perm = rng.next()                                       # pinned int64 [total], or the partner list (DevicePartners)
        if self._perm_bufs is None or self._perm_bufs[0].numel() != total:
            self._perm_bufs = [torch.empty(total, dtype=torch.int64, device=self.device) for _ in range(2)]
            self._perm_free = [None, None]
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._perm_j = None
        on_device = isinstance(perm, DevicePartners)
        if on_device and self._perm_j is None:
            self._perm_j = [torch.empty(total, dtype=torch.int32, device=self.device) for _ in range(2)]
            self._perm_ws = torch.empty(L.call("ppx_np_shuffle_apply_device_workspace", total), dtype=torch.uint8, device=self.device)
        with torch.cuda.stream(self._copy_stream):
            if self._perm_free[slot] is not None:               # the epoch that last read this buffer must be done
                self._copy_stream.wait_event(self._perm_free[slot])
            if on_device:                                       # 2 MB up instead of 4, swaps resolved in parallel on the copy stream
                self._perm_j[slot].copy_(perm.j, non_blocking=True)
                L.call("ppx_np_shuffle_apply_device", self._perm_j[slot].data_ptr(), total, 1, self._perm_ws.data_ptr(),
                       self._perm_bufs[slot].data_ptr(), L.stream())
            else:
                self._perm_bufs[slot].copy_(perm, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._perm_ready[slot] = (ev, perm)                     # keep the pinned source alive until the copy ran

    def _epoch_minibatches(self, ro, rng, epoch=0, n_epochs=1):
        """Yields (idx_dev, B_local, B_total, key) for one epoch.  Single GPU: slices of the reference's
        permutation (buffer.py:239,251-254), staged in a static device buffer (double-buffered, prefetched on a
        copy stream).  Sharded: every rank draws the SAME permutation over the global [T, W*N] index space and
        keeps the rows of each global minibatch whose env it owns (owner-computes)."""
        W, r = D.world_size(), D.rank()
        T, N = ro.buffer_size, ro.n_envs
        local = W > 1 and self.shard_shuffle == "local"
        if local:
            W_eff, W = W, 1                                     # per-rank shuffle: the single-GPU path + B_total
        total = T * N * W
        Bg = min(self.batch_size * (W if self.scale_batch_with_world else 1), total)
        if W == 1:
            slot = epoch & 1
            if epoch == 0 or self._perm_ready[slot] is None:
                self._perm_prefetch(rng, total, slot)
            ev, _keep = self._perm_ready[slot]
            torch.cuda.current_stream().wait_event(ev)
            buf = self._perm_bufs[slot]
            for s in range(0, total, Bg):
                sl = buf[s:s + Bg]
                yield sl, sl.numel(), (sl.numel() * W_eff if local else 0), (slot, s)
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream())
            self._perm_free[slot] = done
            self._perm_ready[slot] = None
            if epoch + 1 < n_epochs:                            # after this epoch's randn()s were consumed (RND)
                self._perm_prefetch(rng, total, slot ^ 1)
            return
        perm = rng.next().numpy()
        for s in range(0, total, Bg):
            g = perm[s:s + Bg]
            loc = D.owned_slice(g, T, N, r)
            yield torch.as_tensor(loc).to(self.device, non_blocking=True), len(loc), len(g), s

    def _sync_grads(self, bank):
        if D.world_size() > 1:
            D.all_reduce_sum_(bank.grad)

    def _policy_optim_step(self):
        """Gradient exchange + clip + Adam for the policy bank.  Sharded with peer memory: ONE kernel reads every rank's
        gradient over NVLink, sums in rank order, clips and applies Adam (weights stay bit-identical replicas)."""
        bank = self.policy.bank
        px = self._peer_exchange() if D.world_size() > 1 else None
        if px is not None:
            n_clip = bank.size if self.max_grad_norm > 0 else 0
            L.call("ppx_p2p_clip_adam", bank.flat.data_ptr(), px.peer_grad, px.peer_flags[2], px.W, px.rank, px.seq[2],
                   px.status_ptr, bank.exp_avg.data_ptr(), bank.exp_avg_sq.data_ptr(), bank.size, float(self.max_grad_norm),
                   n_clip, float(self.lr), 0.9, 0.999, 1e-8, bank.step_dev.data_ptr(), bank.norm_dev.data_ptr(), None,
                   L.stream())
            bank.refresh_tc()
            return
        if getattr(self, "_pre_sumsq", False):
            ss, n_ss = self.policy.mlp.sumsq
            bank.adam_step_pre(self.lr, self.max_grad_norm, ss, n_ss, extra_name="action_log_std")
            return
        self._sync_grads(bank)
        bank.adam_step(self.lr, self.max_grad_norm)

    def _finish_train(self, losses_dev, keys):
        losses = losses_dev.cpu().numpy()                         # the only D2H sync of train()
        if self._px and int(self._px.status.item()) != 0:
            raise RuntimeError("ppx: a peer-memory barrier timed out (a rank fell out of the sharded update)")
        self.last_losses = losses
        for i, k in enumerate(keys):
            self._record(k, float(np.mean(losses[:, i])))
        self._n_updates += self.n_epochs

    def _learn_loop(self, total_timesteps, log_interval, reward_target):
        """algorithms.py:277-308 / 514-543 / 725-756 (logging lines trimmed to record())."""
        start, iteration = time.time(), 0
        while self.num_timesteps < total_timesteps:
            self.collect_samples()
            iteration += 1
            if log_interval is not None and iteration % log_interval == 0:
                self._record("time/total timesteps", self.num_timesteps)
                if len(self.ep_info_buffer) > 0 and len(self.ep_info_buffer[0]) > 0:
                    self._record("rollout/ep_rew_mean", np.mean([e["r"] for e in self.ep_info_buffer]))
                    self._record("rollout/num_episodes", self.num_episodes)
                self._record("time/total_time", time.time() - start)
                if self.logger is not None and hasattr(self.logger, "dump"):
                    self.logger.dump(step=self.num_timesteps)
            self.train()
            if reward_target is not None and len(self.ep_info_buffer) > 0 and \
                    np.mean([e["r"] for e in self.ep_info_buffer]) > reward_target:
                break
        return self


class PPO(BaseAlgorithm):
    """algorithms.py:121-308."""

    def __init__(self, *, env_id=None, lr=3e-4, nstep=128, batch_size=128, n_epochs=10, gamma=0.99, gae_lam=0.95,
                 clip_range=0.2, ent_coef=.01, vf_coef=1, max_grad_norm=0.2, hidden_size=128, sim_hash=False,
                 sil=False, env=None, device="cuda", logger=None, hash_bits=16):
        super().__init__(env_id, lr, nstep, batch_size, n_epochs, gamma, gae_lam, clip_range, ent_coef, vf_coef,
                         max_grad_norm, env=env, device=device, logger=logger)
        if sil:
            raise NotImplementedError("self-imitation is dead code in the reference (SURVEY §2 rows 16-17)")
        self.policy = Policy(self.env, hidden_size, device=self.device)
        self.rollout = RolloutStorage(nstep, self.num_envs, self.env.observation_space, self.env.action_space,
                                      gae_lam=gae_lam, gamma=gamma, sim_hash=sim_hash, device=self.device,
                                      hash_bits=hash_bits)
        self.last_obs = self.env.reset()
        self.sim_hash = sim_hash
        self.sil = sil

    def collect_samples(self):
        """algorithms.py:166-198."""
        assert self.last_obs is not None
        self.rollout.reset()
        for _ in range(self.nstep):
            actions, values, log_probs = self.policy.act(self.last_obs)
            obs, rewards, dones, infos = self.env.step(actions.cpu().numpy())
            if any(dones):
                self.num_episodes += sum(dones)
            self.num_timesteps += self.num_envs
            self.update_info_buffer(infos)
            A = self.action_converter.action_output
            self.rollout.add(self.last_obs, actions.reshape(self.num_envs, A), rewards, values, dones,
                             log_probs.reshape(self.num_envs, A))
            self.last_obs = obs
        self.rollout.compute_returns_and_advantages(values, dones=dones)
        return True

    def train(self):
        """algorithms.py:200-259."""
        ro = self.rollout
        total = ro.buffer_size * ro.n_envs
        B = min(self.batch_size, total)
        n_mb = -(-total // B)
        losses = torch.zeros(self.n_epochs * n_mb, 8, dtype=torch.float64, device=self.device)
        bufs = ro._minibatch_buffers(2 * B if D.world_size() > 1 else B)
        step = 0
        rng = self._rng_open(self._rng_script(ro))
        self._perm_ready = [None, None]
        for ep in range(self.n_epochs):
            for sl, b, bt, off in self._epoch_minibatches(ro, rng, ep, self.n_epochs):
                def fn(sl=sl, b=b, bt=bt):
                    ok = self._gather_with_stats(ro, sl, bufs)
                    self._policy_step(bufs, b, self._loss_row.data_ptr(), B_total=bt, stats_ready=ok)
                    self._policy_optim_step()
                self._graph_call(("ppo", off, b, bt), fn)
                losses[step].copy_(self._loss_row)
                step += 1
        ro.generator_ready = True
        self._rng_close(rng)
        self._finish_train(losses[:step], ("train/total_loss", "train/policy_gradient_loss", "train/value_loss",
                                           "train/entropy_loss"))

    def learn(self, total_timesteps, log_interval, reward_target=None, log_to_file=False):
        return self._learn_loop(total_timesteps, log_interval, reward_target)


class PPO_RND(BaseAlgorithm):
    """algorithms.py:310-543."""

    def __init__(self, *, env_id=None, lr=3e-4, nstep=128, batch_size=128, n_epochs=10, gamma=0.99, int_gamma=0.99,
                 gae_lam=0.95, clip_range=0.2, ent_coef=.01, vf_coef=0.5, int_vf_coef=0.5, max_grad_norm=0.2,
                 hidden_size=128, int_hidden_size=128, int_lr=3e-4, rnd_start=1e+3, env=None, device="cuda",
                 logger=None):
        super().__init__(env_id, lr, nstep, batch_size, n_epochs, gamma, gae_lam, clip_range, ent_coef, vf_coef,
                         max_grad_norm, env=env, device=device, logger=logger)
        self.policy = Policy(self.env, hidden_size, intrinsic_model=True, device=self.device)
        self.rnd = RndNetwork(self.state_dim, hidden_size=int_hidden_size, device=self.device)
        self.rollout = IntrinsicStorage(nstep, self.num_envs, self.env.observation_space, self.env.action_space,
                                        gae_lam=gae_lam, gamma=gamma, int_gamma=int_gamma, device=self.device)
        self.int_lr = int_lr
        self.rnd_start = rnd_start
        self.int_vf_coef = int_vf_coef
        self.last_obs = self.env.reset()
        self.int_rew_rms = RunningMeanStd(device=self.device)
        self.normalize = True
        self.last_dones = np.array([0 for _ in range(self.num_envs)])
        self._rnd_loss = torch.zeros(1, dtype=torch.float64, device=self.device)

    def rnd_bonus(self, obs):
        """The bonus lines of collect_samples after warm-up (algorithms.py:394-398) for one env step:
        normalise -> (pred-target)^2 -> int_rew_rms.update -> divide.  obs [N,D]; returns [N] f32 CUDA."""
        o = _dev(obs, torch.float32, self.device)
        r = self.rnd.int_reward(o, rms=self.obs_rms)
        L.call("ppx_rnd_normalize_rollout", r.data_ptr(), 1, r.numel(), self.int_rew_rms.mean_dev.data_ptr(),
               self.int_rew_rms.var_dev.data_ptr(), self.int_rew_rms.count_dev.data_ptr(), L.stream())
        return r

    def rnd_bonus_rollout(self, next_obs):
        """Same arithmetic for a whole rollout at once: next_obs [T,N,D] -> [T,N].  Exact because obs_rms is
        frozen after warm-up and the int-reward moments are merged step by step in t order on device."""
        o = _dev(next_obs, torch.float32, self.device)
        T, N = o.shape[0], o.shape[1]
        r = self.rnd.int_reward(o.reshape(T * N, -1), rms=self.obs_rms)
        L.call("ppx_rnd_normalize_rollout", r.data_ptr(), T, N, self.int_rew_rms.mean_dev.data_ptr(),
               self.int_rew_rms.var_dev.data_ptr(), self.int_rew_rms.count_dev.data_ptr(), L.stream())
        return r.view(T, N)

    def collect_samples(self):
        """algorithms.py:367-407."""
        assert self.last_obs is not None
        self.rollout.reset()
        for _ in range(self.nstep):
            actions, values, int_values, log_probs = self.policy.act(self.last_obs)
            obs, rewards, dones, infos = self.env.step(actions.cpu().numpy())
            if any(dones):
                self.num_episodes += sum(dones)
            self.num_timesteps += self.num_envs
            self.update_info_buffer(infos)
            A = self.action_converter.action_output
            if (self.num_timesteps / self.num_envs) < self.rnd_start:
                int_rewards = torch.zeros(self.num_envs, device=self.device)
                self.obs_rms.update(self.env.unnormalize_obs(self.last_obs))
            else:
                int_rewards = self.rnd_bonus(obs)
            self.rollout.add(self.last_obs, actions.reshape(self.num_envs, A), rewards, int_rewards, values,
                             int_values, dones, log_probs.reshape(self.num_envs, A))
            self.last_obs = obs
            self.last_dones = dones
        mean_int = self.rollout.compute_returns_and_advantages(values, int_values, dones)
        self._mean_int_reward = mean_int                            # logged lazily (rollout/mean_int_reward, buffer.py:335)
        return True

    def train_rnd(self, obs, B_total=0):
        """algorithms.py:487-502 on a gathered minibatch of raw observations [B,D]."""
        x = normalize_obs(obs, self.obs_rms, out=self._scratch.get("rnd_nobs", obs.numel())[:obs.numel()].view(obs.shape))
        self.rnd.train_step(x, self._rnd_loss, B_total)
        self._sync_grads(self.rnd.bank)
        self.rnd.bank.adam_step(self.int_lr, self.max_grad_norm)

    def train(self):
        """algorithms.py:409-485."""
        ro = self.rollout
        total = ro.buffer_size * ro.n_envs
        B = min(self.batch_size, total)
        n_mb = -(-total // B)
        losses = torch.zeros(self.n_epochs * n_mb, 8, dtype=torch.float64, device=self.device)
        bufs = ro._minibatch_buffers(2 * B if D.world_size() > 1 else B)
        step = 0
        self.rnd_trained_steps = 0
        rng = self._rng_open(self._rng_script(ro, randn_per_minibatch=True))
        self._perm_ready = [None, None]
        for ep in range(self.n_epochs):
            for sl, b, bt, off in self._epoch_minibatches(ro, rng, ep, self.n_epochs):
                def fn(sl=sl, b=b, bt=bt):
                    ok = self._gather_with_stats(ro, sl, bufs, dual=True)
                    self._policy_step(bufs, b, self._loss_row.data_ptr(), dual=True, int_vf_coef=self.int_vf_coef,
                                      B_total=bt, stats_ready=ok)
                    self._policy_optim_step()
                self._graph_call(("rnd_policy", off, b, bt), fn)
                losses[step].copy_(self._loss_row)
                if rng.next() < 0.25:                               # algorithms.py:468, same host RNG stream
                    self._graph_call(("rnd_pred", b, bt), lambda b=b, bt=bt: self.train_rnd(bufs['observations'][:b], bt))
                    self.rnd_trained_steps += 1
                step += 1
        ro.generator_ready = True
        self._rng_close(rng)
        self._finish_train(losses[:step], ("train/total_loss", "train/policy_gradient_loss", "train/value_loss",
                                           "train/entropy_loss", "train/intrinsic_loss"))

    def learn(self, total_timesteps, log_interval, reward_target=None, log_to_file=False):
        return self._learn_loop(total_timesteps, log_interval, reward_target)


class PPO_ICM(BaseAlgorithm):
    """algorithms.py:546-756."""

    def __init__(self, *, env_id=None, lr=3e-4, int_lr=3e-4, nstep=128, batch_size=128, n_epochs=10, gamma=0.99,
                 gae_lam=0.95, clip_range=0.2, ent_coef=.01, vf_coef=0.5, max_grad_norm=0.2, hidden_size=128,
                 int_hidden_size=32, int_rew_integration=0.05, beta=0.2, policy_weight=1, env=None, device="cuda",
                 logger=None):
        super().__init__(env_id, lr, nstep, batch_size, n_epochs, gamma, gae_lam, clip_range, ent_coef, vf_coef,
                         max_grad_norm, env=env, device=device, logger=logger)
        self.int_rew_integration = int_rew_integration
        self.policy = Policy(self.env, hidden_size, device=self.device)
        # the reference builds this buffer WITHOUT gamma -> 0.99 always (algorithms.py:591)
        self.rollout = RolloutStorage(nstep, self.num_envs, self.env.observation_space, self.env.action_space,
                                      gae_lam=gae_lam, device=self.device)
        self.intrinsic_module = IntrinsicCuriosityModule(self.state_dim, self.action_converter,
                                                         hidden_size=int_hidden_size, device=self.device)
        self.int_lr = int_lr
        self.last_obs = self.env.reset()
        self.policy_weight = policy_weight
        self.beta = 0.2                                             # hard-wired, algorithms.py:600
        self._icm_loss = torch.zeros(1, dtype=torch.float64, device=self.device)

    def icm_bonus(self, last_obs, obs, actions, rewards):
        """algorithms.py:629-630: r = (1-eta) r + eta * clamp(mean_f((fwd(phi(s),a) - phi(s'))^2), -5, 5).
        Returns (blended rewards [N] f32 CUDA, raw int_rewards [N])."""
        r = _dev(rewards, torch.float32, self.device).clone()
        ri = self.intrinsic_module.int_reward(last_obs, obs, actions, rewards=r, eta=self.int_rew_integration)
        return r, ri

    def collect_samples(self):
        """algorithms.py:603-649."""
        assert self.last_obs is not None
        self.rollout.reset()
        ri_means = []
        for _ in range(self.nstep):
            actions, values, log_probs = self.policy.act(self.last_obs)
            obs, rewards, dones, infos = self.env.step(actions.cpu().numpy())
            if any(dones):
                self.num_episodes += sum(dones)
            self.num_timesteps += self.num_envs
            self.update_info_buffer(infos)
            rewards, ri = self.icm_bonus(self.last_obs, obs, actions, rewards)
            ri_means.append(ri.mean())
            A = self.action_converter.action_output
            self.rollout.add(self.last_obs, actions.reshape(self.num_envs, A), rewards, values, dones,
                             log_probs.reshape(self.num_envs, A))
            self.last_obs = obs
        self._mean_int_reward = torch.stack(ri_means).mean()
        self.rollout.compute_returns_and_advantages(values, dones=dones)
        return True

    def train(self):
        """algorithms.py:651-713.  (Single-GPU only for now: the shuffled-consecutive row pairing of :684 spans
        the whole minibatch, so sharding it needs a halo row -- see DESIGN.md.)"""
        if D.world_size() > 1:
            raise NotImplementedError("PPO_ICM.train is not sharded yet")
        ro = self.rollout
        total = ro.buffer_size * ro.n_envs
        B = min(self.batch_size, total)
        n_mb = -(-total // B)
        losses = torch.zeros(self.n_epochs * n_mb, 8, dtype=torch.float64, device=self.device)
        icm_losses = torch.zeros(self.n_epochs * n_mb, dtype=torch.float64, device=self.device)
        bufs = ro._minibatch_buffers(B)
        step = 0
        rng = self._rng_open(self._rng_script(ro))
        icm_row = torch.zeros(1, dtype=torch.float64, device=self.device) if not hasattr(self, "_icm_row") else self._icm_row
        self._icm_row = icm_row
        self._perm_ready = [None, None]
        for ep in range(self.n_epochs):
            for sl, b, bt, off in self._epoch_minibatches(ro, rng, ep, self.n_epochs):
                def fn(sl=sl, b=b):
                    ok = self._gather_with_stats(ro, sl, bufs)
                    self._policy_step(bufs, b, self._loss_row.data_ptr(), policy_weight=float(self.policy_weight),
                                      stats_ready=ok)
                    icm_row.zero_()
                    self.intrinsic_module.train_step(bufs['observations'][:b], bufs['actions'][:b], self.beta, icm_row)
                    self._policy_optim_step()                                   # only policy grads are clipped (:697)
                    self.intrinsic_module.bank.adam_step(self.int_lr, 0.0)
                self._graph_call(("icm", off, b), fn)
                losses[step].copy_(self._loss_row)
                icm_losses[step:step + 1].copy_(icm_row)
                step += 1
        ro.generator_ready = True
        self._rng_close(rng)
        losses[:, 5] = icm_losses
        losses[:, 0] += icm_losses                                  # total = pw*(...) + icm_loss (:692)
        keys = ("train/total_loss", "train/policy_gradient_loss", "train/value_loss", "train/entropy_loss")
        self._finish_train(losses[:step], keys)
        self._record("train/icm_loss", float(np.mean(self.last_losses[:, 5])))

    def learn(self, total_timesteps, log_interval=5, reward_target=None, log_to_file=False):
        return self._learn_loop(total_timesteps, log_interval, reward_target)
