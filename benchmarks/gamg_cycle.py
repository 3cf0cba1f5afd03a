import sys, os, time, numpy as np
sys.path.insert(0, "tests")
from _pkg import load_pkg; load_pkg()
from b200ls import capi, cases
N = int(sys.argv[1])
capi.init(0)
s = cases.cavity_laplacian(N,N,N, coeffs="random", rhs_kind="uniform")
mesh = capi.Mesh(s.n_cells, s.lower, s.upper); nc = mesh.agglomerate(s.face_weights)
mat = capi.Matrix(mesh); mat.set(s.diag, s.upper_coeffs)
for smoother in ("GaussSeidel",):
  for rep in range(2):
    ctl = capi.controls("GAMG", smoother=smoother, tolerance=1e-6, relTol=0.0)
    psi, perf = mat.solve(ctl, s.source)
    print(os.environ.get("B200LS_SMALL_LEVEL_CELLS"), N, "GAMG", smoother, "iters", perf.nIterations, "res %.3e"%perf.finalResidual, "setupMs %.2f solveMs %.2f ms/cycle %.3f launches %d" % (perf.setupMs, perf.solveMs, perf.solveMs/max(perf.nIterations,1), perf.kernelLaunches), flush=True)
