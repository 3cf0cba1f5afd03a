#!/usr/bin/env python3
"""Debugging aid: repeat every pencil operator on a few block shapes and count the runs whose result is not
bit-identical to the C oracle (a race shows up as an occasional mismatch).

    python benchmarks/pencil_stress.py [reps]
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "oracle"))
os.environ.setdefault("B200LS_PENCIL_MIN_CELLS", "0")
from _pkg import load_pkg  # noqa: E402

load_pkg()
import ldu_oracle as orc  # noqa: E402
from b200ls import capi, cases  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
capi.init(0)
for shape in [(40, 33, 17), (150, 9, 5), (70, 70, 1), (64, 64, 64)]:
    for sym in (True, False):
        nx, ny, nz = shape
        s = cases.cavity_laplacian(nx, ny, nz, coeffs="random") if sym else \
            cases.convection_diffusion(nx, ny, nz, dt_coeff=50.0)
        mesh, mat = capi.from_system(s)
        S = orc.System(s)
        kind = "DIC" if sym else "DILU"
        x = np.cos(0.7 * np.arange(s.n_cells)) + 0.3
        ops = [("rD", lambda: mat.reciprocal_d(kind), orc.reciprocal_d(S)),
               ("precondition", lambda: mat.precondition(kind, x), orc.precondition(S, kind, x)),
               ("GaussSeidel x1", lambda: mat.smooth("GaussSeidel", x, s.source, 1), orc.smooth(S, "GaussSeidel", x, s.source, 1)),
               ("GaussSeidel x3", lambda: mat.smooth("GaussSeidel", x, s.source, 3), orc.smooth(S, "GaussSeidel", x, s.source, 3)),
               ("symGaussSeidel x1", lambda: mat.smooth("symGaussSeidel", x, s.source, 1), orc.smooth(S, "symGaussSeidel", x, s.source, 1)),
               ("symGaussSeidel x3", lambda: mat.smooth("symGaussSeidel", x, s.source, 3), orc.smooth(S, "symGaussSeidel", x, s.source, 3))]
        line = []
        for name, fn, want in ops:
            bad = 0
            for _ in range(reps):
                if name == "rD":
                    mat.set(s.diag, s.upper_coeffs, s.lower_coeffs)   # forces a new factorisation
                got = fn()
                bad += 0 if np.array_equal(got, want) else 1
            line.append(f"{name} {bad}/{reps}")
        print(shape, "sym" if sym else "asym", "|", ", ".join(line), flush=True)
