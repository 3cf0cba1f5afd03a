import sys, os, time, numpy as np
sys.path.insert(0, "tests")
from _pkg import load_pkg; load_pkg()
from b200ls import capi, cases
N = int(sys.argv[1]); iters = int(sys.argv[2]) if len(sys.argv)>2 else 10
capi.init(0)
s = cases.cavity_laplacian(N,N,N)
mesh, mat = capi.from_system(s)
ctl = capi.controls("PCG","DIC", tolerance=0.0, relTol=0.0, maxIter=iters)
for rep in range(2):
    psi, perf = mat.solve(ctl, s.source)
    print(N, "PCG iters", perf.nIterations, "solveMs", perf.solveMs, "ms/iter", perf.solveMs/perf.nIterations, "launches", perf.kernelLaunches)
