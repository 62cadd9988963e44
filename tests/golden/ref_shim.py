"""Makes the unmodified reference at /root/reference importable in THIS container.

Only used by tests/golden/make_golden.py (fixture generation).  Nothing that runs on the GPU box imports
this file: /root/reference does not exist there.  The stand-in modules (gym, stable_baselines3, ...) and the
FakeVecEnv live in oracle/ref_runtime.py, which bench.py's CPU arm shares (it imports the vendored copy under
oracle/_ref instead).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_runtime as _rt  # noqa: E402

REF = "/root/reference"
Space, Box, Discrete, VecEnv, FakeVecEnv = _rt.Space, _rt.Box, _rt.Discrete, _rt.VecEnv, _rt.FakeVecEnv
set_env_factory = _rt.set_env_factory


def install():
    return _rt.install(REF)
