import sys, time, numpy as np
sys.path.insert(0, "tests")
from _pkg import load_pkg; load_pkg()
from b200ls import capi, cases
capi.init(0)
for N in (64, 128):
    t=time.time(); s = cases.cavity_laplacian(N,N,N); print("gen", time.time()-t)
    t=time.time(); mesh, mat = capi.from_system(s); print("mesh+set", time.time()-t)
    for rep in range(2):
        ctl = capi.controls("PCG","DIC", tolerance=0.0, relTol=0.0, maxIter=50)
        psi, perf = mat.solve(ctl, s.source)
        print(N, "PCG iters", perf.nIterations, "res", perf.finalResidual, "solveMs", perf.solveMs, "ms/iter", perf.solveMs/perf.nIterations, "launches", perf.kernelLaunches)
    print("amul ms", mat.time_kernel(0, 20), "GB/s", 72*s.n_cells/mat.time_kernel(0,20)/1e6)
    print("precond ms", mat.time_kernel(1, 20))
    print("gs ms", mat.time_kernel(2, 20))
