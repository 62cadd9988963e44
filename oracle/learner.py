"""Oracle for the learner half of the hot path (reference: algorithms.py, models.py, util.py).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  torch-CPU restatement with explicit parameter
dictionaries (keys = the reference modules' state_dict names, so weights move both ways).  The
reference's own third-party calls (torch.optim.Adam, clip_grad_norm_, torch.distributions math)
are kept as calls into torch: they are the pinned dependency, not reference code.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

from . import rollout as R


# --------------------------------------------------------------------------------------
# parameter construction                              models.py:126-213, 216-249, 270-298
# --------------------------------------------------------------------------------------
def _linear(i, o, gain=math.sqrt(2.0)):
    w = torch.empty(o, i)
    torch.nn.init.orthogonal_(w, gain)                      # models.py:133
    return w, torch.zeros(o)                                # models.py:134


def make_policy_params(obs_dim, act_dim, hidden, intrinsic=False):
    """MlpNetwork / MlpIntrinsic parameters (models.py:137-154, 173-195): separate
    Linear-Tanh-Linear-Tanh-Linear stacks for actor, critic (and int_critic), orthogonal(sqrt 2)
    weights, zero biases, action_log_std = zeros[1,A].  Creation order = reference module order."""
    p = {}
    heads = [('actor', act_dim), ('critic', 1)] + ([('int_critic', 1)] if intrinsic else [])
    for name, out in heads:
        for li, (i, o) in zip((0, 2, 4), ((obs_dim, hidden), (hidden, hidden), (hidden, out))):
            w, b = _linear(i, o)
            p[f'{name}.{li}.weight'], p[f'{name}.{li}.bias'] = w, b
    p['action_log_std'] = torch.zeros(1, act_dim)
    return {k: v.requires_grad_(True) for k, v in p.items()}


def make_rnd_params(obs_dim, hidden):
    """RndNetwork (models.py:220-249): predictor D-h-h-h-1 (LeakyReLU, LeakyReLU, ELU) with
    weights 1 / biases 0.01; frozen target D-h-h-1 (LeakyReLU x2) with weights 0.01 / biases 1."""
    p = {}
    for li, (i, o) in zip((0, 2, 4, 6), ((obs_dim, hidden), (hidden, hidden), (hidden, hidden), (hidden, 1))):
        p[f'predictor.{li}.weight'] = torch.full((o, i), 1.0).requires_grad_(True)
        p[f'predictor.{li}.bias'] = torch.full((o,), 0.01).requires_grad_(True)
    for li, (i, o) in zip((0, 2, 4), ((obs_dim, hidden), (hidden, hidden), (hidden, 1))):
        p[f'target.{li}.weight'] = torch.full((o, i), 0.01)
        p[f'target.{li}.bias'] = torch.full((o,), 1.0)
    return p


def make_icm_params(obs_dim, n_actions, hidden, discrete):
    """IntrinsicCuriosityModule (models.py:281-298); feature_size == hidden_size (models.py:275)."""
    f = hidden
    p = {}
    shapes = [('state_encoder', (obs_dim, hidden), (hidden, f)),
              ('forward_model', (n_actions + f, hidden), (hidden, f)),
              ('inverse_model', (2 * f, hidden), (hidden, n_actions))]
    for name, s0, s2 in shapes:
        for li, (i, o) in ((0, s0), (2, s2)):
            w, b = _linear(i, o)
            p[f'{name}.{li}.weight'], p[f'{name}.{li}.bias'] = w, b
    if discrete:
        p['action_encoder.weight'] = torch.randn(n_actions, n_actions)      # nn.Embedding default
    else:
        w, b = _linear(n_actions, n_actions)
        p['action_encoder.weight'], p['action_encoder.bias'] = w, b
    return {k: v.requires_grad_(True) for k, v in p.items()}


# --------------------------------------------------------------------------------------
# forward passes
# --------------------------------------------------------------------------------------
def _mlp3(p, name, x):
    h = torch.tanh(F.linear(x, p[f'{name}.0.weight'], p[f'{name}.0.bias']))
    h = torch.tanh(F.linear(h, p[f'{name}.2.weight'], p[f'{name}.2.bias']))
    return F.linear(h, p[f'{name}.4.weight'], p[f'{name}.4.bias'])


def evaluate(p, obs, actions, discrete, intrinsic=False):
    """Policy.evaluate / evaluate_intrinsic.  models.py:52-73, 101-124.

    Box:  Normal(tanh(actor(x)), exp(log_std)); per-dimension log_prob and entropy [B,A].  The
          stored actions are f64 (buffer.py:154), so log_prob -- and everything downstream of it
          in the policy loss -- is f64.
    Discrete: Categorical(softmax(logits)); log_prob [B,1], entropy [B].
    Returns (values[B], int_values[B] or None, log_probs, entropy)."""
    obs = obs.float()
    a = _mlp3(p, 'actor', obs)
    v = _mlp3(p, 'critic', obs).squeeze()
    iv = _mlp3(p, 'int_critic', obs).squeeze() if intrinsic else None
    if discrete:
        dist = torch.distributions.Categorical(F.softmax(a, dim=-1))
        lp = dist.log_prob(actions.flatten()).unsqueeze(1)
        ent = dist.entropy()
    else:
        mean = a.tanh()
        dist = torch.distributions.Normal(mean, torch.exp(p['action_log_std'].expand_as(mean)))
        lp = dist.log_prob(actions)
        ent = dist.entropy()
    return v, iv, lp, ent


def rnd_forward(p, x):
    """RndNetwork.forward.  models.py:251-256."""
    x = x.float()
    h = F.leaky_relu(F.linear(x, p['predictor.0.weight'], p['predictor.0.bias']))
    h = F.leaky_relu(F.linear(h, p['predictor.2.weight'], p['predictor.2.bias']))
    h = F.elu(F.linear(h, p['predictor.4.weight'], p['predictor.4.bias']))
    pred = F.linear(h, p['predictor.6.weight'], p['predictor.6.bias'])
    t = F.leaky_relu(F.linear(x, p['target.0.weight'], p['target.0.bias']))
    t = F.leaky_relu(F.linear(t, p['target.2.weight'], p['target.2.bias']))
    tgt = F.linear(t, p['target.4.weight'], p['target.4.bias'])
    return pred, tgt


def rnd_int_reward(p, obs):
    """(pred - target)^2, squeezed.  models.py:261-267."""
    pred, tgt = rnd_forward(p, torch.as_tensor(np.asarray(obs)).float())
    return (pred - tgt).pow(2).squeeze()


def rnd_bonus_step(p, obs_next, obs_mean, obs_var, int_rms):
    """The per-env-step RND lines of collect_samples after warm-up.  algorithms.py:394-398."""
    nobs = R.normalize_obs(obs_next, obs_mean, obs_var)
    r = rnd_int_reward(p, nobs).detach().numpy()
    int_rms.update(r)
    r = r / (np.sqrt(int_rms.var) + 1e-08)
    return r


def _seq2(p, name, x):
    h = F.leaky_relu(F.linear(x, p[f'{name}.0.weight'], p[f'{name}.0.bias']))
    return F.linear(h, p[f'{name}.2.weight'], p[f'{name}.2.bias'])


def _icm_encode_action(p, action, discrete, squeeze):
    if discrete:
        a = action.squeeze().long() if squeeze else action.long()
        return F.embedding(a, p['action_encoder.weight'])
    return F.linear(action.float(), p['action_encoder.weight'], p['action_encoder.bias'])


def icm_forward(p, state, next_state, action, discrete):
    """IntrinsicCuriosityModule.forward -> (action_hat, next_feat_hat, next_feat).  models.py:300-309."""
    a = _icm_encode_action(p, action, discrete, squeeze=True)
    s = _seq2(p, 'state_encoder', state)
    ns = _seq2(p, 'state_encoder', next_state)
    a_hat = _seq2(p, 'inverse_model', torch.cat((s, ns), 1))
    ns_hat = _seq2(p, 'forward_model', torch.cat((s, a), 1))
    return a_hat, ns_hat, ns


def icm_int_reward(p, state, next_state, action, discrete):
    """clamp(mean_f((fwd(s,a) - enc(s'))^2), -5, 5).  models.py:311-320."""
    a = _icm_encode_action(p, action, discrete, squeeze=False)
    s = _seq2(p, 'state_encoder', state)
    ns = _seq2(p, 'state_encoder', next_state)
    ns_hat = _seq2(p, 'forward_model', torch.cat((s, a), 1))
    return torch.clamp((ns_hat - ns).pow(2).mean(dim=-1), -5, 5)


# --------------------------------------------------------------------------------------
# losses                                           algorithms.py:216-238, 428-460, 668-692
# --------------------------------------------------------------------------------------
def _normalise(adv):
    return (adv - adv.mean()) / (adv.std() + 1e-8)           # unbiased std, algorithms.py:219


def _clipped_value_loss(returns, v, old_v, clip):
    """max of the two MEANS (not mean of maxes).  algorithms.py:229-232."""
    v_clip = old_v + (v - old_v).clamp(-clip, clip)
    return torch.max(F.mse_loss(returns, v).mean(), F.mse_loss(returns, v_clip).mean()).mean()


def _surrogate(adv, lp, old_lp, clip):
    ratio = torch.exp(lp - old_lp)
    return -torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - clip, 1 + clip)).mean()


def ppo_losses(p, batch, hp, discrete):
    """PPO.train inner loss.  algorithms.py:216-238.  Returns (total, policy, value, entropy)."""
    v, _, lp, ent = evaluate(p, batch['observations'], batch['actions'], discrete)
    adv = _normalise(batch['advantages'])
    pl = _surrogate(adv, lp, batch['old_log_probs'], hp['clip_range'])
    vl = _clipped_value_loss(batch['returns'], v, batch['old_values'], hp['clip_range'])
    el = -torch.mean(ent)
    return pl + hp['ent_coef'] * el + hp['vf_coef'] * vl, pl, vl, el


def rnd_losses(p, batch, hp, discrete):
    """PPO_RND.train inner loss.  algorithms.py:428-460.  Returns (total, pl, vl, el, int_vl)."""
    v, iv, lp, ent = evaluate(p, batch['observations'], batch['actions'], discrete, intrinsic=True)
    adv = _normalise(batch['advantages']) + _normalise(batch['int_advantages'])   # :431-434
    pl = _surrogate(adv, lp, batch['old_log_probs'], hp['clip_range'])
    vl = _clipped_value_loss(batch['returns'], v, batch['old_values'], hp['clip_range'])
    ivl = _clipped_value_loss(batch['int_returns'], iv, batch['int_values'], hp['clip_range'])
    el = -torch.mean(ent)
    total = pl + hp['ent_coef'] * el + hp['vf_coef'] * vl + hp['int_vf_coef'] * ivl
    return total, pl, vl, el, ivl


def icm_losses(p, icm, batch, hp, discrete):
    """PPO_ICM.train inner loss.  algorithms.py:668-692; beta is hard-wired to 0.2 (:600)."""
    total_ppo, pl, vl, el = None, None, None, None
    v, _, lp, ent = evaluate(p, batch['observations'], batch['actions'], discrete)
    adv = _normalise(batch['advantages'])
    pl = _surrogate(adv, lp, batch['old_log_probs'], hp['clip_range'])
    vl = _clipped_value_loss(batch['returns'], v, batch['old_values'], hp['clip_range'])
    obs, act = batch['observations'], batch['actions']
    a_hat, nf_hat, nf = icm_forward(icm, obs[:-1], obs[1:], act[:-1], discrete)   # shuffled-consecutive rows
    fwd = F.mse_loss(nf, nf_hat)
    if discrete:
        inv = F.cross_entropy(a_hat, act[:-1].squeeze().long())                  # util.py:61-78
    else:
        inv = F.mse_loss(a_hat, act[:-1].float())
    icm_loss = (1 - 0.2) * inv + 0.2 * fwd
    el = -torch.mean(ent)
    total = hp.get('policy_weight', 1) * (pl + hp['vf_coef'] * vl + hp['ent_coef'] * el) + icm_loss
    return total, pl, vl, el, icm_loss


# --------------------------------------------------------------------------------------
# train loops
# --------------------------------------------------------------------------------------
def _to_torch(d):
    return {k: torch.tensor(v) for k, v in d.items()}            # buffer.py:107-108


def _policy_param_list(p):
    # reference order of policy.net.parameters(): own Parameter first, then children
    return [p['action_log_std']] + [v for k, v in p.items() if k != 'action_log_std']


def ppo_train(p, opt, buf, hp, discrete, max_steps=None):
    """PPO.train.  algorithms.py:200-259.  `buf` = dict of [T,N,...] numpy arrays (pre-flatten).
    Draws one np.random.permutation per epoch (buffer.py:239).  Returns per-minibatch losses."""
    T, N = buf['rewards'].shape
    log, steps = [], 0
    plist = _policy_param_list(p)
    for _ in range(hp['n_epochs']):
        perm = R.epoch_permutation(T, N)
        for s, e in R.minibatch_slices(T * N, hp['batch_size']):
            batch = _to_torch(R.gather_single(buf, perm[s:e]))
            total, pl, vl, el = ppo_losses(p, batch, hp, discrete)
            opt.zero_grad()
            total.backward()
            torch.nn.utils.clip_grad_norm_(plist, hp['max_grad_norm'])
            opt.step()
            log.append((total.item(), pl.item(), vl.item(), el.item()))
            steps += 1
            if max_steps is not None and steps >= max_steps:
                return np.array(log)
    return np.array(log)


def rnd_train(p, opt, rnd, rnd_opt, buf, hp, discrete, obs_mean, obs_var):
    """PPO_RND.train + train_rnd.  algorithms.py:409-502.  One np.random.randn() per minibatch
    between the epoch permutations (:468) decides whether the predictor trains on that batch."""
    T, N = buf['rewards'].shape
    log, rnd_log = [], []
    plist = _policy_param_list(p)
    rnd_plist = [v for k, v in rnd.items() if k.startswith('predictor')]
    for _ in range(hp['n_epochs']):
        perm = R.epoch_permutation(T, N)
        for s, e in R.minibatch_slices(T * N, hp['batch_size']):
            batch = _to_torch(R.gather_dual(buf, perm[s:e]))
            total, pl, vl, el, ivl = rnd_losses(p, batch, hp, discrete)
            opt.zero_grad()
            total.backward()
            torch.nn.utils.clip_grad_norm_(plist, hp['max_grad_norm'])
            opt.step()
            if np.random.randn() < 0.25:
                nobs = R.normalize_obs(batch['observations'].numpy(), obs_mean, obs_var)   # :494
                pred, tgt = rnd_forward(rnd, torch.from_numpy(nobs).float())
                loss = F.mse_loss(pred, tgt)
                rnd_opt.zero_grad()
                loss.backward()
                torch.nn.utils.clip_grad_norm_(rnd_plist, hp['max_grad_norm'])
                rnd_opt.step()
                rnd_log.append(loss.item())
            else:
                rnd_log.append(float('nan'))
            log.append((total.item(), pl.item(), vl.item(), el.item(), ivl.item()))
    return np.array(log), np.array(rnd_log)


def icm_train(p, opt, icm, icm_opt, buf, hp, discrete):
    """PPO_ICM.train.  algorithms.py:651-713.  One backward over policy+ICM; only the policy
    gradients are norm-clipped (:697); two Adam steps (:698-699)."""
    T, N = buf['rewards'].shape
    log = []
    plist = _policy_param_list(p)
    for _ in range(hp['n_epochs']):
        perm = R.epoch_permutation(T, N)
        for s, e in R.minibatch_slices(T * N, hp['batch_size']):
            batch = _to_torch(R.gather_single(buf, perm[s:e]))
            total, pl, vl, el, il = icm_losses(p, icm, batch, hp, discrete)
            opt.zero_grad()
            icm_opt.zero_grad()
            total.backward()
            torch.nn.utils.clip_grad_norm_(plist, hp['max_grad_norm'])
            opt.step()
            icm_opt.step()
            log.append((total.item(), pl.item(), vl.item(), el.item(), il.item()))
    return np.array(log)
