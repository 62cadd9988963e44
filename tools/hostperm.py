import numpy as np, time, ctypes as C, threading, os, sys
sys.path.insert(0, os.getcwd())
import ppo_exploration_b200 as ppx
from ppo_exploration_b200 import _lib as L
from ppo_exploration_b200.buffer import HostRngStream
n=524288
def draws(reps, res):
    np.random.seed(0)
    st=np.random.get_state()
    key=np.ascontiguousarray(st[1],dtype=np.uint32).copy(); pos=C.c_int(int(st[2]))
    j=np.empty(n,np.int32)
    t0=time.perf_counter()
    for _ in range(reps): L.call("ppx_np_shuffle_draws32", key.ctypes.data, C.byref(pos), n, j.ctypes.data)
    res['d']=(time.perf_counter()-t0)/reps*1e3
def apply(reps,res):
    j=np.minimum(np.random.randint(0,n,n),np.arange(n)).astype(np.int32)
    out=np.empty(n,np.int64); sc=np.empty(n,np.int32)
    t0=time.perf_counter()
    for _ in range(reps): L.call("ppx_np_shuffle_apply32", j.ctypes.data, n, sc.ctypes.data, out.ctypes.data)
    res['a']=(time.perf_counter()-t0)/reps*1e3
res={}
draws(20,res); apply(20,res); print("alone",res, "cpus", os.cpu_count())
t1=threading.Thread(target=draws,args=(30,res)); t2=threading.Thread(target=apply,args=(30,res))
t1.start(); t2.start(); t1.join(); t2.join(); print("concurrent",res)
r=HostRngStream([('perm',n)]*30)
t0=time.perf_counter()
for _ in range(30): r.next()
print("HostRngStream per perm %.2f ms"%((time.perf_counter()-t0)/30*1e3))
os.system("lscpu | head -20")
