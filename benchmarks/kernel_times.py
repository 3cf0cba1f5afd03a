import sys, os, time, numpy as np
sys.path.insert(0, "tests")
from _pkg import load_pkg; load_pkg()
from b200ls import capi, cases
N = int(sys.argv[1])
capi.init(0)
s = cases.cavity_laplacian(N,N,N)
t=time.time(); mesh, mat = capi.from_system(s); print("mesh+set s", time.time()-t, flush=True)
print("B/SM", os.environ.get("B200LS_SWEEP_BLOCKS_PER_SM"), "N", N,
      "amul ms %.4f" % mat.time_kernel(0, 20), "precond ms %.4f" % mat.time_kernel(1, 10), "gs ms %.4f" % mat.time_kernel(2, 10), flush=True)
