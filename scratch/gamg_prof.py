import sys, os, time, numpy as np
sys.path.insert(0, "tests")
from _pkg import load_pkg; load_pkg()
from b200ls import capi, cases
N = int(sys.argv[1])
capi.init(0)
s = cases.cavity_laplacian(N,N,N)
mesh = capi.Mesh(s.n_cells, s.lower, s.upper); nc = mesh.agglomerate(s.face_weights)
print("levels", [int(mesh.get_i32(capi.LEVEL_SIZES, l)[0]) for l in range(nc+1)])
print("fwd wavefronts per level", [int(mesh.get_i32(capi.FWD_LEVEL_OFFSETS, l).size-1) for l in range(nc+1)])
mat = capi.Matrix(mesh); mat.set(s.diag, s.upper_coeffs)
ctl = capi.controls("GAMG", smoother="GaussSeidel", tolerance=0.0, relTol=0.0, maxIter=3)
for rep in range(2):
    psi, perf = mat.solve(ctl, s.source)
    print(N, "GAMG iters", perf.nIterations, "setupMs %.2f solveMs %.2f ms/cycle %.3f launches %d" % (perf.setupMs, perf.solveMs, perf.solveMs/max(perf.nIterations,1), perf.kernelLaunches), flush=True)
