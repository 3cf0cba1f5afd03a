"""Debugging aid: many repetitions of the Gauss-Seidel pencil sweeps on one shape, mismatches vs the C oracle counted per
smoother and described (first bad cell, NaNs) -- to tell a race in the forward half from one in the reverse half."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "oracle"))
os.environ.setdefault("B200LS_PENCIL_MIN_CELLS", "0")
from _pkg import load_pkg
load_pkg()
import ldu_oracle as orc
from b200ls import capi, cases
shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "150,9,5").split(","))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
capi.init(0)
nx, ny, nz = shape
s = cases.cavity_laplacian(nx, ny, nz, coeffs="random")
mesh, mat = capi.from_system(s)
S = orc.System(s)
x = np.cos(0.7 * np.arange(s.n_cells)) + 0.3
seq = [("GaussSeidel", 1), ("GaussSeidel", 3), ("symGaussSeidel", 1), ("symGaussSeidel", 3)]
want = {k: orc.smooth(S, k[0], x, s.source, k[1]) for k in seq}
bad = {k: 0 for k in seq}
desc = {}
for r in range(reps):
    if r % 7 == 0:   # a fresh matrix object now and then, like consecutive tests
        mat.close(); mesh.close()
        mesh, mat = capi.from_system(s)
    for k in seq:
        got = mat.smooth(k[0], x, s.source, k[1])
        m = np.flatnonzero(~(got == want[k]))
        if m.size:
            bad[k] += 1
            if k not in desc:
                c = int(m[0])
                desc[k] = (f" first failure rep {r}: {m.size} cells, first (i {c % nx}, j {(c // nx) % ny}, k {c // (nx * ny)}) got {got[c]!r} want {want[k][c]!r}, "
                           f"nan {int(np.isnan(got).sum())}, bad j {sorted(set(((m // nx) % ny).tolist()))} bad k {sorted(set((m // (nx * ny)).tolist()))} "
                           f"i-range {int((m % nx).min())}..{int((m % nx).max())}")
for k in seq:
    print(f"{shape} {k[0]} x{k[1]}: {bad[k]}/{reps} mismatches{desc.get(k, '')}", flush=True)
