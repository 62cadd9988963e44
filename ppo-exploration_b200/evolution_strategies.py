"""ES-NSRA step on the GPU with the reference's method names.

Mirrors EvolutionStrategy of evolution_strategies.py (:100-384) for the parts on the learner hot
path: _get_population (:172-182), _get_weights_try (:137-145), _update_weights (:217-239), get_kNN
(:264-281), the novelty lines (:318-325) and calc_noveltiy_distribution (:283-290), plus the policy forward
of episode evaluation (FeedForwardNetwork.predict :48-61) batched over the population
(`predict_population`).  Stepping the environments (evaluate / get_behavior_char / run's env
loop) is the simulator side and stays with the caller: feed the fitness vector back through
`_update_weights`.

Build-side design: the population is a vector of offsets into ONE resident f32 noise
table instead of P fresh f64 randn tensors; theta is one flat f64 device vector over all layers.
Parity mode: pass a dense eps array ([P, D] float32-representable values, or the reference's nested
list-of-layers population) wherever a population is expected.
"""
import numpy as np
import torch

from . import _lib as L
from . import dist as D


class EvolutionStrategy(object):
    def __init__(self, env_id=None, hidden_sizes=(64, 64), nsr_plateu=1.5, nsr_range=(0, 1), nsr_update=0.05,
                 population_size=50, sigma=0.1, learning_rate=0.01, decay=0.9995, novelty_param=0.5, num_threads=1,
                 obs_dim=None, n_actions=None, device="cuda", noise_table_size=1 << 24, noise_seed=0,
                 fitness_shaping="zscore"):
        if obs_dim is None or n_actions is None:
            raise ValueError("pass obs_dim= and n_actions= (the reference reads them from gym.make(env_id), "
                             "evolution_strategies.py:121; the env layer is out of scope here)")
        self.env_id = env_id
        self.device = torch.device(device)
        self.hidden_sizes = list(hidden_sizes)
        sizes = [obs_dim, *hidden_sizes, n_actions]
        self.shapes = [(sizes[i], sizes[i + 1]) for i in range(len(sizes) - 1)]       # :33-35, bias-free
        self.layer_sizes = [a * b for a, b in self.shapes]
        self.D = int(sum(self.layer_sizes))
        # same global-RNG draws as FeedForwardNetwork.__init__ (:34-35)
        w0 = [np.random.randn(*s) for s in self.shapes]
        self.theta = torch.as_tensor(np.concatenate([w.ravel() for w in w0])).to(self.device)
        self.POPULATION_SIZE = population_size
        self.SIGMA = sigma
        self._lr = torch.full((1,), float(learning_rate), dtype=torch.float64, device=self.device)
        self.decay = decay
        self.novelty_param = novelty_param
        self.K = 10
        self.nsr_plateu, self.nsr_range, self.nsr_update = nsr_plateu, list(nsr_range), nsr_update
        self.fitness_shaping = fitness_shaping
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._ws = None
        self._noise = None
        self._noise_size, self._noise_seed = int(noise_table_size), int(noise_seed)
        self._draw = torch.zeros(2, dtype=torch.int64, device=self.device)        # [population draws made so far, ticket word]
        self._px = False                                                          # PeerExchange for the sharded update | None
        self._graphs = {}
        self.update_mode = "single"

    # ---- state views ----
    @property
    def learning_rate(self):
        return float(self._lr.item())

    @learning_rate.setter
    def learning_rate(self, v):
        self._lr.fill_(float(v))

    @property
    def weights(self):
        return self.get_weights()

    def get_weights(self):
        """list of [in,out] f64 numpy arrays, like FeedForwardNetwork.get_weights (:63-64)."""
        flat = self.theta.cpu().numpy()
        out, o = [], 0
        for s, n in zip(self.shapes, self.layer_sizes):
            out.append(flat[o:o + n].reshape(s).copy())
            o += n
        return out

    def set_weights(self, weights):
        self.theta.copy_(torch.as_tensor(np.concatenate([np.asarray(w, np.float64).ravel() for w in weights])))

    # ---- noise table / population ----
    def noise_table(self):
        if self._noise is None:
            self._noise = torch.empty(self._noise_size, dtype=torch.float32, device=self.device)
            L.call("ppx_noise_fill", self._noise.data_ptr(), self._noise_size, self._noise_seed, L.stream())
        return self._noise

    def _get_population(self, out=None):
        """:172-182.  Returns int64 offsets [P] (multiples of 4, so rows are 16-byte aligned) into the table, drawn on the
        device (Philox keyed by the noise seed, counter = (member, draw number)): no host RNG, no H2D copy, and every rank
        of a sharded run draws the identical population."""
        self.noise_table()
        off = torch.empty(self.POPULATION_SIZE, dtype=torch.int64, device=self.device) if out is None else out
        L.call("ppx_es_offsets", self._noise_seed + 1, self._draw.data_ptr(), self.POPULATION_SIZE, self._noise_size, self.D,
               off.data_ptr(), L.stream())
        return off

    def _as_eps(self, population):
        """-> (noise tensor f32, offsets tensor or None)."""
        if isinstance(population, torch.Tensor) and population.dtype == torch.int64:
            return self.noise_table(), population
        if isinstance(population, list):                             # reference nested list [P][L] of arrays
            population = np.array([np.concatenate([np.asarray(l).ravel() for l in m]) for m in population])
        eps = torch.as_tensor(np.ascontiguousarray(population)).to(self.device)
        if eps.dtype != torch.float32:
            e32 = eps.float()
            if not torch.equal(e32.double(), eps.double()):
                raise ValueError("dense eps must be float32-representable (the device noise is f32)")
            eps = e32
        return eps.contiguous(), None

    def perturb_all(self, population, out_f64=False):
        """theta + sigma*eps for every member: [P, D] CUDA (f32 by default, f64 = reference dtype)."""
        noise, off = self._as_eps(population)
        P = off.numel() if off is not None else noise.shape[0]
        out = torch.empty(P, self.D, dtype=torch.float64 if out_f64 else torch.float32, device=self.device)
        L.call("ppx_es_perturb", self.theta.data_ptr(), noise.data_ptr(), off.data_ptr() if off is not None else None,
               float(self.SIGMA), P, self.D, out.data_ptr(), int(out_f64), L.stream())
        return out

    def predict_population(self, population, obs, discrete=False, sigma=None):
        """FeedForwardNetwork.predict (:48-61) for every member at once: member p sees obs[p] and acts with
        theta + sigma*eps_p, formed on the fly from the noise table (the perturbed weights are never materialised).
        obs [P, obs_dim] -> f64 CUDA [P, n_actions]: tanh(logits) for Box (continuous_action :84-89); the raw logits
        with discrete=True (sample them with `discrete_action`, which keeps the reference's host RNG draw)."""
        noise, off = self._as_eps(population)
        P = off.numel() if off is not None else noise.shape[0]
        obs = torch.as_tensor(np.asarray(obs) if not isinstance(obs, torch.Tensor) else obs).to(self.device).double().contiguous()
        sizes = [self.shapes[0][0]] + [sh[1] for sh in self.shapes]
        if tuple(obs.shape) != (P, sizes[0]):
            raise ValueError(f"obs must be [{P}, {sizes[0]}], got {tuple(obs.shape)}")
        out = torch.empty(P, sizes[-1], dtype=torch.float64, device=self.device)
        import ctypes as C
        L.call("ppx_es_forward", self.theta.data_ptr(), noise.data_ptr(), off.data_ptr() if off is not None else None,
               float(self.SIGMA if sigma is None else sigma), P, (C.c_int * len(sizes))(*sizes), len(sizes) - 1,
               obs.data_ptr(), 0 if discrete else 1, out.data_ptr(), L.stream())
        return out

    def predict(self, obs, discrete=False):
        """FeedForwardNetwork.predict (:48-61) of the CURRENT weights for one observation (or a batch [n, obs_dim])."""
        obs = np.asarray(obs, dtype=np.float64)
        obs = obs.reshape(1, -1) if obs.ndim == 1 or obs.size == self.shapes[0][0] else obs
        zero = np.zeros((obs.shape[0], self.D), dtype=np.float32)
        return self.predict_population(zero, obs, discrete=discrete, sigma=0.0)

    @staticmethod
    def discrete_action(logits):
        """discrete_action (:81-82) on the host: softmax then one np.random.choice per member (the reference's RNG)."""
        out = []
        for row in np.asarray(logits.cpu() if isinstance(logits, torch.Tensor) else logits, dtype=np.float64):
            p = np.exp(row) / np.sum(np.exp(row))
            out.append(int(np.random.choice(np.arange(row.size), p=p.squeeze())))
        return np.array(out)

    def _get_weights_try(self, w, p):
        """:137-145 for ONE member given as the reference's list of per-layer eps arrays; returns a list of
        f64 numpy arrays.  (`w` is accepted for signature compatibility; theta on device is what is used.)"""
        flat = self.perturb_all([p], out_f64=True)[0].cpu().numpy()
        out, o = [], 0
        for s, n in zip(self.shapes, self.layer_sizes):
            out.append(flat[o:o + n].reshape(s))
            o += n
        return out

    # ---- population sharding: members split across ranks, update replicated ----
    def shard_population(self, population):
        """The contiguous slice of `population` (offsets [P] or dense eps [P,D]) this rank perturbs and evaluates.
        Every rank must hold the same `population` (same noise seed -> `_get_population` draws identical offsets)."""
        W, r = D.world_size(), D.rank()
        P = population.shape[0]
        if P % W:
            raise ValueError(f"population size {P} must be a multiple of the world size {W}")
        return population[r * (P // W):(r + 1) * (P // W)]

    def gather_fitness(self, local_rewards):
        """all-gather of the per-rank fitness slices -> [P] f64 in member order on every rank; each rank then applies
        the identical `_update_weights` from the replicated noise table (no parameter traffic)."""
        r = torch.as_tensor(np.asarray(local_rewards, dtype=np.float64)).to(self.device) \
            if not isinstance(local_rewards, torch.Tensor) else local_rewards.to(self.device, torch.float64)
        W = D.world_size()
        if W == 1:
            return r
        out = torch.empty(W, r.numel(), dtype=torch.float64, device=self.device)
        D.all_gather_into(out, r)
        return out.reshape(-1)

    # ---- update ----
    def _update_weights(self, rewards, population, novelty=None):
        """:217-239.  Asynchronous; `self.update_skipped` reads back the std==0 flag.  `novelty` may be a python float
        (as in the reference) or a 1-element f64 CUDA tensor (e.g. `novelty_batch(...)[1][m:m+1]`), which keeps the
        k-NN result on the device -- no host round trip in the step."""
        noise, off = self._as_eps(population)
        P = off.numel() if off is not None else noise.shape[0]
        r = torch.as_tensor(np.asarray(rewards, dtype=np.float64)).to(self.device) if not isinstance(rewards, torch.Tensor) \
            else rewards.to(self.device, torch.float64)
        need = L.call("ppx_es_update_workspace", P, self.D)
        if self._ws is None or self._ws.numel() * 8 < need:
            self._ws = torch.zeros(need // 8 + 1, dtype=torch.float64, device=self.device)     # holds self re-arming tickets
        nov_dev = None
        if isinstance(novelty, torch.Tensor):
            nov_dev = novelty.to(self.device, torch.float64).reshape(-1)[:1].contiguous()
        W, rk = D.world_size(), D.rank()
        px = self._peer_exchange() if (W > 1 and self.fitness_shaping != "centered_rank" and P % W == 0) else None
        if px is not None:
            # sharded update: identical z-scores everywhere, GEMV over this rank's P/W members only, one fused
            # barrier + rank-ordered sum + apply over NVLink peer memory (theta / lr stay bit-identical replicas)
            self.update_mode = "sharded: partial GEMV over P/W members + fused peer-memory all-reduce of the update"
            L.call("ppx_es_update_sharded", self.theta.data_ptr(), noise.data_ptr(), off.data_ptr() if off is not None else None,
                   r.data_ptr(), P, rk * (P // W), P // W, self.D, float(self.SIGMA), float(self.novelty_param),
                   float(novelty) if (novelty is not None and nov_dev is None) else 0.0,
                   nov_dev.data_ptr() if nov_dev is not None else None, int(novelty is not None), float(self.decay),
                   self._lr.data_ptr(), self._status.data_ptr(), self._ws.data_ptr(), self._dtheta.data_ptr(), px.peer_grad,
                   px.peer_flags[2], px.W, px.rank, px.seq[2], px.status_ptr, L.stream())
            return
        self.update_mode = "single" if W == 1 else "replicated"
        L.call("ppx_es_update", self.theta.data_ptr(), noise.data_ptr(), off.data_ptr() if off is not None else None,
               r.data_ptr(), P, self.D, float(self.SIGMA), float(self.novelty_param),
               float(novelty) if (novelty is not None and nov_dev is None) else 0.0,
               nov_dev.data_ptr() if nov_dev is not None else None, int(novelty is not None),
               int(self.fitness_shaping == "centered_rank"), float(self.decay), self._lr.data_ptr(),
               self._status.data_ptr(), self._ws.data_ptr(), L.stream())

    def _peer_exchange(self):
        if self._px is False:
            self._px = D.peer_exchange_or_none(2 * self.D, self.device, L.call("ppx_p2p_max_params"))
            if self._px is not None:
                self._dtheta = self._px.grad.view(torch.float64)                  # this rank's partial update, peer-visible
        return self._px

    # ---- ask / tell: the two device-side halves of one ES iteration, each replayed as ONE CUDA graph ----
    def _graph(self, key, fn):
        ent = self._graphs.get(key)
        if ent is None:
            self._graphs[key] = "warm"
            return fn()
        if ent == "warm":
            g = torch.cuda.CUDAGraph()
            n0 = L.launch_count()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                out = fn()
            ent = (g, L.launch_count() - n0, out)
            L.extra_launches -= ent[1]
            self._graphs[key] = ent
        ent[0].replay()
        L.extra_launches += ent[1]
        return ent[2]

    def ask(self, archive=None, queries=None):
        """First half of an iteration: draw the population (:172-182), form theta + sigma*eps for THIS rank's members
        (:137-145) and -- as run() does before it evaluates the population (:318-325) -- the novelty of `queries` against
        the behaviour `archive` (k-NN on a forked stream, under the perturbation).  One CUDA graph.  Returns (offsets [P]
        int64, weights [P/W, D] f32, novelties [Q] f64 | None): static buffers, overwritten by the next ask()."""
        W = D.world_size()
        if not hasattr(self, "_ask_off"):
            self.noise_table()
            self._ask_off = torch.empty(self.POPULATION_SIZE, dtype=torch.int64, device=self.device)
            self._side = torch.cuda.Stream(device=self.device)
        key = ("ask", archive.data_ptr() if archive is not None else 0, queries.data_ptr() if queries is not None else 0,
               archive.shape[0] if archive is not None else 0)

        def fn():
            nov = None
            if archive is not None:
                cur = torch.cuda.current_stream()
                self._side.wait_stream(cur)
                with torch.cuda.stream(self._side):
                    _, nov = self.novelty_batch(archive, queries)
            pop = self._get_population(out=self._ask_off)
            w = self.perturb_all(self.shard_population(pop) if W > 1 else pop)
            if archive is not None:
                torch.cuda.current_stream().wait_stream(self._side)
            return pop, w, nov
        out = self._graph(key, fn)
        self._novelty = out[2]
        return out

    def tell(self, local_rewards, brain=0):
        """Second half (:217-239): exchange the fitness of the rank-local members and update the parameters with the
        novelty of meta-population member `brain` computed by the last ask().  `local_rewards` must be a persistent CUDA
        tensor (its address is captured)."""
        key = ("tell", local_rewards.data_ptr(), int(brain), self._novelty is not None)

        def fn():
            r_all = self.gather_fitness(local_rewards)
            nov = self._novelty[brain:brain + 1] if self._novelty is not None else None
            self._update_weights(r_all, self._ask_off, novelty=nov)
        return self._graph(key, fn)

    @property
    def update_skipped(self):
        return bool(self._status.item())

    def centered_ranks(self, rewards):
        r = torch.as_tensor(np.asarray(rewards, dtype=np.float64)).to(self.device)
        ranks = torch.empty(r.numel(), dtype=torch.int64, device=self.device)
        cen = torch.empty(r.numel(), dtype=torch.float64, device=self.device)
        L.call("ppx_rank_center", r.data_ptr(), r.numel(), ranks.data_ptr(), cen.data_ptr(), L.stream())
        return ranks, cen

    # ---- novelty ----
    def novelty_batch(self, archive, queries, K=None):
        """archive [M,dim], queries [Q,dim] -> (knn distance sums [Q], novelties [Q]) f64 CUDA."""
        a = torch.as_tensor(np.concatenate(archive) if isinstance(archive, list) else np.asarray(archive)) \
            if not isinstance(archive, torch.Tensor) else archive
        a = a.to(self.device, torch.float64).contiguous()
        q = (torch.as_tensor(np.asarray(queries)) if not isinstance(queries, torch.Tensor) else queries)
        q = q.to(self.device, torch.float64).reshape(-1, a.shape[1]).contiguous()
        K = self.K if K is None else K
        s = torch.empty(q.shape[0], dtype=torch.float64, device=self.device)
        nov = torch.empty_like(s)
        L.call("ppx_knn_novelty", a.data_ptr(), a.shape[0], q.data_ptr(), q.shape[0], a.shape[1], int(K), s.data_ptr(),
               nov.data_ptr(), L.stream())
        return s, nov

    def get_kNN(self, archive, bc, n_neighbors):
        """:264-281: summed distance to the n_neighbors nearest archive entries (python float)."""
        s, _ = self.novelty_batch(archive, bc, K=n_neighbors)
        return float(s[0].item())

    def get_novelty_from_bc(self, archive, bc):
        """:318-325 (dup :208-214) given the behaviour characterisation instead of a policy rollout."""
        _, nov = self.novelty_batch(archive, bc)
        return float(nov[0].item())

    def calc_noveltiy_distribution(self, novelties):
        """:283-290 (host; MPS = 2 scalars)."""
        return [round((n / (sum(novelties))), 4) for n in novelties]

    def pick_brain(self, novelties):
        """:327-333: the meta-population member to train next -- probabilities from calc_noveltiy_distribution,
        renormalised (the rounding to 4 digits breaks the unit sum), one np.random.choice draw from the reference's host
        RNG.  Returns (brain_idx, novelty of that member)."""
        probs = np.array(self.calc_noveltiy_distribution(novelties))
        probs /= probs.sum()
        idx = int(np.random.choice(list(range(len(probs))), p=probs))
        return idx, novelties[idx]
