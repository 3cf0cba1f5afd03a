"""Debugging aid for the two-tiles-per-CTA experiment: the FORWARD substitution alone (B200LS_DEBUG_FWD_ONLY) against a
numpy face loop, mismatches listed per tile (k range) and i range."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "oracle"))
os.environ.setdefault("B200LS_PENCIL_MIN_CELLS", "0")
os.environ["B200LS_DEBUG_FWD_ONLY"] = "1"
from _pkg import load_pkg
load_pkg()
import ldu_oracle as orc
from b200ls import capi, cases
shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "20,8,16").split(","))
nx, ny, nz = shape
capi.init(0)
s = cases.cavity_laplacian(*shape, coeffs="random")
mesh, mat = capi.from_system(s)
S = orc.System(s)
rD = orc.reciprocal_d(S)
rA = np.cos(0.37 * np.arange(s.n_cells)) + 0.1
want = rD * rA
for f in range(s.n_faces):
    u, l = s.upper[f], s.lower[f]
    want[u] -= rD[u] * s.upper_coeffs[f] * want[l]
got = mat.precondition("DIC", rA)
bad = np.flatnonzero(~(got == want))
print("forward sweep:", bad.size, "bad of", got.size, "nan", int(np.isnan(got).sum()), "sentinel-like", int((np.abs(got) > 1e300).sum()))
for k0 in range(0, nz, 4):
    m = bad[(bad // (nx * ny)) // 4 == k0 // 4]
    if m.size:
        ii = m % nx
        jj = (m // nx) % ny
        print(f"  tile k {k0}..{k0+3}: {m.size} bad, i {ii.min()}..{ii.max()}, j {sorted(set(jj.tolist()))}, first cell (i {int(ii[0])}, j {int(jj[0])}, k {int(m[0] // (nx*ny))}) got {got[m[0]]!r} want {want[m[0]]!r}")
