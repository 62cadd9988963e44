"""The UNMODIFIED reference as a runnable CPU baseline  --  TEST / BENCH INFRASTRUCTURE, NOT PRODUCT.

The reference (BoogaQ/PPO-exploration) is ~3.4 kLoC of pure Python.  `vendor()` -- called by
`__graft_entry__.build()` in the build container, where /root/reference exists -- copies the modules of the
learner hot path byte for byte into `oracle/_ref/` (git-ignored build output, like a compiled .so: it travels
to the GPU box with the gpurun snapshot, it never enters the history).  `install()` makes them importable:
the reference imports gym / stable_baselines3 / mujoco_py / pybulletgym at module scope
(algorithms.py:2,10,20; evolution_strategies.py:8-9; env.py:1-4); none is installed and none is on the
learner hot path, so empty stand-ins and a FakeVecEnv are registered first.

Users: tests/golden/make_golden.py (fixture generation, from /root/reference directly) and bench.py's
`cpu_baseline` / `--impl reference` legs (from oracle/_ref, kind = "reference").
"""
import hashlib
import os
import shutil
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference"
MODULES = ["algorithms.py", "buffer.py", "models.py", "util.py", "logger.py", "env.py", "evolution_strategies.py",
           "sil_module.py", "hyperparameters.py"]


def vendor(src=REF_SRC, dst=REF_DIR):
    """Copy the reference's hot-path modules (unmodified) into oracle/_ref.  Returns {file: sha256}."""
    os.makedirs(dst, exist_ok=True)
    digests = {}
    for name in MODULES:
        a = os.path.join(src, name)
        if not os.path.exists(a):
            continue
        shutil.copyfile(a, os.path.join(dst, name))
        digests[name] = hashlib.sha256(open(a, "rb").read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST"), "w") as f:
        f.write("# byte-identical copies of /root/reference/<file>, made by oracle/ref_runtime.vendor()\n")
        for k, v in sorted(digests.items()):
            f.write(f"{v}  {k}\n")
    return digests


def available(ref_dir=REF_DIR):
    return all(os.path.exists(os.path.join(ref_dir, m)) for m in ("algorithms.py", "buffer.py", "models.py", "util.py"))


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


def install(ref_dir=REF_DIR):
    """Register the stand-in modules and put `ref_dir` first on sys.path; imports `algorithms`."""
    if "algorithms" in sys.modules and getattr(sys.modules["algorithms"], "_ppx_shimmed", False):
        return sys.modules["algorithms"]
    sys.dont_write_bytecode = True
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)

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
    return algorithms


def set_env_factory(factory):
    """algorithms.BaseAlgorithm calls make_env(env_id, n_envs=4) (algorithms.py:52)."""
    import algorithms
    algorithms.make_env = lambda env_id, n_envs=4: factory()


def silence_logger():
    """logger.record/dump inside train() are host I/O, not the path being timed."""
    import algorithms
    import buffer
    import logger
    logger.record = lambda *a, **k: None
    algorithms.logger.record = logger.record
    buffer.logger.record = logger.record
