#!/usr/bin/env python3
"""Debugging aid: every pencil operator on a list of block shapes against the C oracle, one line per check with the
first mismatching cell (i, j, k) -- quicker to read than a pytest log when a kernel is being changed.

    python benchmarks/pencil_check.py [nx,ny,nz ...]
"""
import os
import sys
import time
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

SHAPES = [(12, 10, 9), (5, 40, 3), (7, 6, 1), (33, 1, 1), (9, 8, 2), (4, 37, 5), (40, 33, 17), (150, 9, 5), (3, 3, 3),
          (70, 70, 1), (17, 16, 12)]


def report(name, got, want, shape):
    nx, ny, nz = shape
    bad = np.flatnonzero(~((got == want) | (np.isnan(got) & np.isnan(want))))
    if bad.size == 0:
        print(f"    ok   {name}")
        return True
    c = int(bad[0])
    print(f"    FAIL {name}: {bad.size}/{got.size} cells differ, first cell {c} = (i {c % nx}, j {(c // nx) % ny}, "
          f"k {c // (nx * ny)}): got {got[c]!r} want {want[c]!r}; nan in result: {int(np.isnan(got).sum())}")
    return False


def main():
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or SHAPES
    capi.init(0)
    n_fail = 0
    for shape in shapes:
        for sym in (True, False):
            nx, ny, nz = shape
            s = cases.cavity_laplacian(nx, ny, nz, coeffs="random") if sym else \
                cases.convection_diffusion(nx, ny, nz, dt_coeff=50.0)
            print(f"{shape} {'sym' if sym else 'asym'}", flush=True)
            t0 = time.time()
            try:
                mesh, mat = capi.from_system(s)
                S = orc.System(s)
                kind = "DIC" if sym else "DILU"
                x = np.cos(0.7 * np.arange(s.n_cells)) + 0.3
                ok = report("rD", mat.reciprocal_d(kind), orc.reciprocal_d(S), shape)
                for seed in (0.37, 0.11):
                    rA = np.cos(seed * np.arange(s.n_cells)) + 0.1
                    ok &= report(f"precondition {seed}", mat.precondition(kind, rA), orc.precondition(S, kind, rA), shape)
                for sm in ("GaussSeidel", "symGaussSeidel"):
                    for n_sweeps in (1, 3):
                        ok &= report(f"{sm} x{n_sweeps}", mat.smooth(sm, x, s.source, n_sweeps),
                                     orc.smooth(S, sm, x, s.source, n_sweeps), shape)
                n_fail += 0 if ok else 1
            except Exception as e:  # noqa: BLE001
                print(f"    ERROR {type(e).__name__}: {e}")
                n_fail += 1
            print(f"    {time.time() - t0:.2f} s", flush=True)
    print("failures:", n_fail)


if __name__ == "__main__":
    main()
