import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
import ppo_exploration_b200 as ppx
from ppo_exploration_b200 import _lib as L
n = 524288
np.random.seed(0)
st = np.random.get_state()
key = np.ascontiguousarray(st[1], dtype=np.uint32).copy(); pos = C.c_int(int(st[2]))
j = np.zeros(n, np.int32)
L.call("ppx_np_shuffle_draws32", key.ctypes.data, C.byref(pos), n, j.ctypes.data)
jd = torch.as_tensor(j).cuda()
ws = torch.empty(L.call("ppx_np_shuffle_apply_device_workspace", n), dtype=torch.uint8, device="cuda")
out = torch.empty(n, dtype=torch.int64, device="cuda")
f = lambda: L.call("ppx_np_shuffle_apply_device", jd.data_ptr(), n, 0, ws.data_ptr(), out.data_ptr(), L.stream())
for _ in range(3): f()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20): f()
e.record(); torch.cuda.synchronize()
print("device apply n=524288: %.1f us" % (s.elapsed_time(e) / 20 * 1e3))
