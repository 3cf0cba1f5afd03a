#!/usr/bin/env python3
"""Application-level drop-in measurement: the reference's own icoFoam (oracle/_app/icoFoam, unmodified) on the
lid-driven cavity, once with the reference's solvers and once with `libs ("libB200LinearSolvers.so")` + the B200
solver names.  Prints the two solver logs side by side (summary) and the wall-clock time of each run.

    python benchmarks/icofoam_dropin.py [N=64] [steps=3] [p: pcg|gamg] [U: symgs|bicg]
"""
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
from _pkg import load_pkg  # noqa: E402

load_pkg()
from b200ls import foam_case  # noqa: E402
import _icofoam as ico  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    psel = sys.argv[3] if len(sys.argv) > 3 else "pcg"
    p_ref = {"pcg": "solver PCG; preconditioner DIC; tolerance 1e-06; relTol 0.05;",
             "gamg": "solver GAMG; smoother GaussSeidel; tolerance 1e-06; relTol 0.05;"}[psel]
    usel = sys.argv[4] if len(sys.argv) > 4 else "symgs"
    u_ref = {"symgs": "solver smoothSolver; smoother symGaussSeidel; tolerance 1e-05; relTol 0;",
             "bicg": "solver PBiCGStab; preconditioner DILU; tolerance 1e-05; relTol 0;"}[usel]
    dt = 0.005 * 20 / n          # keep the Courant number of the 20x20 tutorial
    kw = dict(nx=n, ny=n, nz=n, end_time=steps * dt, delta_t=dt)
    out = {}
    for tag, ps, us, libs in (("reference", p_ref, u_ref, None),
                              ("B200", p_ref.replace("solver ", "solver B200"), u_ref.replace("solver ", "solver B200"),
                               f'"{ico.PLUGIN}"')):
        with tempfile.TemporaryDirectory() as td:
            case = foam_case.write_cavity_case(Path(td) / "case", p_solver=ps, u_solver=us, libs=libs, **kw)
            t0 = time.time()
            log = ico.run_icofoam(case, timeout=3000)
            out[tag] = (time.time() - t0, ico.parse_log(log), log)
    tr, sr, _ = out["reference"]
    tb, sb, _ = out["B200"]
    print(f"icoFoam cavity {n}^3 = {n**3} cells, {steps} time steps, p: {psel.upper()}, U: {usel}")
    print(f"wall clock: reference solvers (1 host core) {tr:.2f} s | B200 plugin {tb:.2f} s (includes CUDA/plugin start-up, "
          f"mesh analysis, per-solve H2D/D2H) | ratio {tr / tb:.1f}x")
    print(f"{'#':>3} {'field':5} {'reference':>14} {'iters':>5} | {'B200':>14} {'iters':>5} | initial residual rel diff")
    worst = 0.0
    for k, (a, b) in enumerate(zip(sr, sb)):
        d = abs(a[2] - b[2]) / max(a[2], 1e-300)
        worst = max(worst, d)
        print(f"{k:3d} {a[1]:5} {a[0]:>14} {a[4]:5d} | {b[0]:>14} {b[4]:5d} | {d:.1e}")
    its_r, its_b = np.array([s[4] for s in sr]), np.array([s[4] for s in sb])
    print(f"{len(sr)} solves; iteration counts identical in {int(np.sum(its_r == its_b))}, max |diff| {int(np.max(np.abs(its_r - its_b)))}; "
          f"max initial-residual rel diff {worst:.1e}")


if __name__ == "__main__":
    main()
