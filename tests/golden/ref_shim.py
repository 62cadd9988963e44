"""Makes the unmodified reference at /root/reference importable in THIS container.

Only used by tests/golden/make_golden.py (fixture generation).  Nothing that runs on the
GPU box imports this file: /root/reference does not exist there.

The reference imports gym / stable_baselines3 / mujoco_py / pybulletgym at module scope
(algorithms.py:2,10,20; evolution_strategies.py:8-9; env.py:1-4).  None is installed and none is
on the learner hot path, so we register empty stand-ins and a FakeVecEnv.
"""
import sys
import types
import numpy as np

REF = "/root/reference"


class Space:
    def __init__(self, shape, n=None):
        self.shape = tuple(shape)
        if n is not None:
            self.n = n


class Box(Space):
    pass


class Discrete(Space):
    def __init__(self, n):
        super().__init__((), n=n)


class VecEnv:
    pass


class FakeVecEnv(VecEnv):
    """Deterministic synthetic env: obs/reward/done streams drawn from its own RandomState."""

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


def install():
    if "algorithms" in sys.modules and getattr(sys.modules["algorithms"], "_ppx_shimmed", False):
        return
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    spaces = mod("gym.spaces", Box=Box, Discrete=Discrete)
    mod("gym", spaces=spaces, make=lambda *a, **k: None)
    mod("mujoco_py")
    mod("pybulletgym")
    mod("stable_baselines3")
    mod("stable_baselines3.common")
    mod("stable_baselines3.common.vec_env", SubprocVecEnv=object, VecFrameStack=object,
        VecTransposeImage=object, VecNormalize=object)
    mod("stable_baselines3.common.vec_env.base_vec_env", VecEnv=VecEnv)
    mod("stable_baselines3.common.cmd_util", make_atari_env=None, make_vec_env=None)
    import algorithms  # noqa: E402
    algorithms._ppx_shimmed = True


def set_env_factory(factory):
    """algorithms.BaseAlgorithm calls make_env(env_id, n_envs=4) (algorithms.py:52)."""
    import algorithms
    algorithms.make_env = lambda env_id, n_envs=4: factory()
