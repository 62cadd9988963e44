"""Diagnostic: host-side shuffle pipeline timings on this machine (stage times, first-permutation latency)."""
import numpy as np, time, ctypes as C, threading, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppo_exploration_b200 as ppx
from ppo_exploration_b200 import _lib as L
from ppo_exploration_b200.buffer import HostRngStream
if torch.cuda.is_available():
    torch.zeros(1, device="cuda"); torch.empty(1 << 20, pin_memory=True)
n = 524288
np.random.seed(0); st = np.random.get_state()
key = np.ascontiguousarray(st[1], dtype=np.uint32).copy(); pos = C.c_int(int(st[2]))
j = np.empty(n, np.int32); prog = np.zeros(1, np.int64); out = np.empty(n, np.int64); sc = np.empty(n, np.int32)
for _ in range(3):
    prog[0] = 0; t0 = time.perf_counter()
    L.call("ppx_np_shuffle_draws32_stream", key.ctypes.data, C.byref(pos), n, j.ctypes.data, prog.ctypes.data); t1 = time.perf_counter()
    L.call("ppx_np_shuffle_apply32_stream", j.ctypes.data, n, prog.ctypes.data, sc.ctypes.data, out.ctypes.data); t2 = time.perf_counter()
    print("draws_stream %.2f ms, apply_stream (all available) %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
for trial in range(4):
    t0 = time.perf_counter()
    r = HostRngStream([('perm', n)] * 10)
    r.next(); t1 = time.perf_counter()
    for _ in range(9): r.next()
    t2 = time.perf_counter()
    print("HostRngStream: first perm %.2f ms, then %.2f ms each" % ((t1 - t0) * 1e3, (t2 - t1) / 9 * 1e3))
