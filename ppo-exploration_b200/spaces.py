"""Minimal stand-ins for gym.spaces and a synthetic vectorised env (the env layer is out of scope;
these only carry shapes and produce seeded synthetic transitions for tests and the benchmark)."""
import numpy as np


class Box:
    def __init__(self, shape):
        self.shape = tuple(shape) if not isinstance(shape, int) else (shape,)


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()


class SyntheticVecEnv:
    """obs ~ N(0,1) f32, reward ~ N(0,1) f32, done ~ Bernoulli(done_p) from a private RandomState."""

    def __init__(self, n_envs, obs_dim, action_space, seed=0, done_p=0.02):
        self.num_envs = n_envs
        self.observation_space = Box((obs_dim,))
        self.action_space = action_space
        self.rs = np.random.RandomState(seed)
        self.done_p = done_p

    def reset(self):
        return self.rs.randn(self.num_envs, self.observation_space.shape[0]).astype(np.float32)

    def step(self, actions):
        obs = self.rs.randn(self.num_envs, self.observation_space.shape[0]).astype(np.float32)
        rew = self.rs.randn(self.num_envs).astype(np.float32)
        done = self.rs.rand(self.num_envs) < self.done_p
        return obs, rew, done, [{} for _ in range(self.num_envs)]

    def unnormalize_obs(self, obs):
        return obs
