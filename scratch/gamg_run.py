import sys, os, time, numpy as np
sys.path.insert(0, "tests")
from _pkg import load_pkg; load_pkg()
from b200ls import capi, cases
N = int(sys.argv[1])
capi.init(0)
t=time.time(); s = cases.cavity_laplacian(N,N,N); print("gen %.1fs"%(time.time()-t), flush=True)
t=time.time(); mesh = capi.Mesh(s.n_cells, s.lower, s.upper); print("mesh %.1fs"%(time.time()-t), flush=True)
t=time.time(); nc = mesh.agglomerate(s.face_weights); print("agglomerate %.1fs levels %d"%(time.time()-t, nc), flush=True)
mat = capi.Matrix(mesh); t=time.time(); mat.set(s.diag, s.upper_coeffs); print("set %.2fs"%(time.time()-t), flush=True)
for smoother in ("GaussSeidel", "DIC"):
  for rep in range(2):
    ctl = capi.controls("GAMG", smoother=smoother, tolerance=1e-6, relTol=0.01)
    psi, perf = mat.solve(ctl, s.source)
    print(N, "GAMG", smoother, "iters", perf.nIterations, "res %.3e"%perf.finalResidual, "setupMs %.2f solveMs %.2f ms/cycle %.3f launches %d" % (perf.setupMs, perf.solveMs, perf.solveMs/max(perf.nIterations,1), perf.kernelLaunches), flush=True)
ctl = capi.controls("PCG", "DIC", tolerance=1e-6, relTol=0.01)
psi, perf = mat.solve(ctl, s.source)
print(N, "PCG DIC iters", perf.nIterations, "solveMs %.2f"%perf.solveMs)
