"""Debugging aid: one DIC precondition on a small block with a capped number of CTAs (multi-round), checked against the
oracle; B200LS_PENCIL_GROUP_MODES selects which sweeps run with two tiles per CTA."""
import os, sys, signal
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "oracle"))
os.environ.setdefault("B200LS_PENCIL_MIN_CELLS", "0")
from _pkg import load_pkg
load_pkg()
import ldu_oracle as orc
from b200ls import capi, cases
shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "40,33,17").split(","))
capi.init(0)
s = cases.cavity_laplacian(*shape, coeffs="random")
mesh, mat = capi.from_system(s)
S = orc.System(s)
print("rD", np.array_equal(mat.reciprocal_d("DIC"), orc.reciprocal_d(S)), flush=True)
rA = np.cos(0.37 * np.arange(s.n_cells)) + 0.1
for rep in range(2):
    got = mat.precondition("DIC", rA)
    want = orc.precondition(S, "DIC", rA)
    bad = np.flatnonzero(~(got == want))
    nx, ny, nz = shape
    msg = ""
    if bad.size:
        cs = bad[:6]
        msg = " first bad cells (i,j,k): " + str([(int(c % nx), int((c // nx) % ny), int(c // (nx * ny))) for c in cs]) + \
            f" nan {int(np.isnan(got).sum())}; bad k-planes {sorted(set((bad // (nx * ny)).tolist()))[:20]} bad j {sorted(set(((bad // nx) % ny).tolist()))[:40]}"
    print("precondition", rep, bad.size, "bad of", got.size, msg, flush=True)
