import os, sys, time
sys.path.insert(0, "/root/repo")
os.chdir("/root/repo")
import numpy as np
from libmultiviewnative_b200 import load
from libmultiviewnative_b200.synthetic import make_views_fast
dims=(256,256,256)
lib=load()
d=make_views_fast(dims,6,41,1)
psi=d["psi0"].copy()
for i in range(3):
    np.copyto(psi,d["psi0"])
    t0=time.perf_counter()
    lib.inplace_gpu_deconvolve(psi,d["views"],d["kernels1"],d["kernels2"],d["weights"],50,0.006,1e-4,0)
    print("call %d: %.1f ms"%(i,(time.perf_counter()-t0)*1e3),flush=True)
os.environ["LMVN_TRACE"]="1"
np.copyto(psi,d["psi0"])
t0=time.perf_counter()
lib.inplace_gpu_deconvolve(psi,d["views"],d["kernels1"],d["kernels2"],d["weights"],50,0.006,1e-4,0)
print("traced call: %.1f ms"%((time.perf_counter()-t0)*1e3),flush=True)
t0=time.perf_counter(); q=np.array(d["psi0"],copy=True); print("np.array copy 64MB: %.1f ms"%((time.perf_counter()-t0)*1e3))
