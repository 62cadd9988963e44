"""Oracle for the ES-NSRA step (reference: evolution_strategies.py).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  numpy float64 throughout, as the reference.
"""
import numpy as np


def layer_shapes(obs_dim, hidden_sizes, n_actions):
    """Bias-free MLP weight shapes.  evolution_strategies.py:33-35."""
    sizes = [obs_dim, *hidden_sizes, n_actions]
    return [(sizes[i], sizes[i + 1]) for i in range(len(sizes) - 1)]


def get_population(shapes, P):
    """P members x L layers of fresh randn, member-major / layer-minor RNG order.  :172-182."""
    return [[np.random.randn(*s) for s in shapes] for _ in range(P)]


def predict_logits(weights, obs):
    """Bias-free MLP forward: arctan hidden layers, linear last layer.  evolution_strategies.py:48-59."""
    out = np.expand_dims(np.asarray(obs).flatten(), 0)
    for i in range(len(weights) - 1):
        out = np.dot(out, weights[i])
        out = np.arctan(out)
    return np.dot(out, weights[-1]).astype(float)


def predict(weights, obs, discrete=False):
    """FeedForwardNetwork.predict.  :48-61 with get_action :91-95: tanh(logits) for Box (:84-89), one
    np.random.choice over softmax(logits) for Discrete (:75-82)."""
    logits = predict_logits(weights, obs)
    if discrete:
        p = np.exp(logits) / np.sum(np.exp(logits))
        return np.random.choice(np.arange(logits.shape[1]), p=p.squeeze()).astype(int)
    return np.tanh(logits).astype(np.double)


def weights_try(w, member, sigma):
    """theta_l + sigma * eps_l per layer.  :137-145."""
    return [w[l] + sigma * member[l] for l in range(len(w))]


def update_weights(w, rewards, population, lr, sigma, novelty_param, decay, novelty=None):
    """z-score shaped update with the novelty mix.  :217-239.  Returns (new_w, new_lr).

    std==0 -> untouched weights AND untouched lr (early return, :225-226).  `novelty` is one
    scalar broadcast over the population (:233-234); ddof=0 std (:224)."""
    rewards = np.asarray(rewards, np.float64)
    std = rewards.std()
    if std == 0:
        return w, lr
    r = (rewards - rewards.mean()) / std
    P = len(population)
    f = lr / (P * sigma)
    out = []
    for l, wl in enumerate(w):
        E = np.array([m[l] for m in population])                     # [P, in, out]
        if novelty is not None:
            nov = np.zeros(r.shape)
            nov.fill(novelty)
            score = ((1 - novelty_param) * np.dot(E.T, r).T + novelty_param * np.dot(E.T, nov).T) / 2
            out.append(wl + f * score)
        else:
            out.append(wl + f * np.dot(E.T, r).T)
    return out, lr * decay


def knn_sum(archive, query, k):
    """Sum of the k smallest Euclidean distances archive<->query.  :264-281.
    The reference fits sklearn NearestNeighbors per call; semantics = exact k smallest, summed in
    ascending order with python's sum().  Brute force in f64 here (direct differences)."""
    a = np.concatenate(archive) if isinstance(archive, list) else np.asarray(archive)
    q = np.asarray(query, np.float64).reshape(-1)
    d2 = np.zeros(len(a), np.float64)
    for j in range(a.shape[1]):                                       # sequential over dims, no FMA
        diff = a[:, j] - q[j]
        d2 = d2 + diff * diff
    d = np.sort(np.sqrt(d2))[:k]
    return sum(d)


def novelty(archive, query, K=10):
    """S = min(K, M); nu = knn_sum/S; floor nu<=1e-3 -> 5e-3.  :318-325 (dup :208-214)."""
    M = len(archive)
    S = int(np.minimum(K, M))
    nu = knn_sum(archive, query, S) / S
    if nu <= 1e-3:
        nu = 5e-3
    return nu


def novelty_distribution(novelties):
    """round(nu_m / sum, 4) then renormalise.  :283-290, :329-330."""
    probs = np.array([round(n / sum(novelties), 4) for n in novelties])
    return probs / probs.sum()


def centered_ranks(r):
    """Extra (non-reference) fitness shaping named by BASELINE.json: centred ranks in [-0.5, 0.5],
    ties broken by index (stable).  PARITY UNPINNED by the reference (it has no rank transform);
    this numpy definition is the specification."""
    r = np.asarray(r)
    ranks = np.argsort(np.argsort(r, kind='stable'), kind='stable')
    return ranks, ranks / (len(r) - 1) - 0.5
