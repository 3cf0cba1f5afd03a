#!/usr/bin/env python3
"""Golden log + final fields of the reference's icoFoam on the generated cavity case (build container only):

    python oracle/build_app.py && python tests/golden/make_icofoam_golden.py

Writes tests/golden/icofoam_<name>.b2ls: per linear solve (in log order) solver name, field, initial/final residual,
iteration count, and the p and U fields at the end time.  tests/test_gpu_icofoam.py re-runs the same cases with
`libs ("libB200LinearSolvers.so")` and the B200 solver names and compares."""
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT / "tests"))
from _pkg import load_pkg  # noqa: E402

load_pkg()
from b200ls import foam_case, ldu_io  # noqa: E402
import _icofoam as ico  # noqa: E402

CASES = {
    # the shipped tutorial settings (tutorials/legacy/incompressible/icoFoam/cavity/cavity/system/fvSolution)
    "cavity20_pcg": dict(nx=20, ny=20, nz=1, end_time=0.1),
    # the p solver of the incompressibleFluid/cavity tutorial: GAMG + GaussSeidel
    "cavity20_gamg": dict(nx=20, ny=20, nz=1, end_time=0.1,
                          p_solver="solver GAMG; smoother GaussSeidel; tolerance 1e-06; relTol 0.1;"),
    # 3-D, asymmetric U solve with PBiCGStab+DILU
    "cavity12x12x6_bicg": dict(nx=12, ny=12, nz=6, end_time=0.05,
                               u_solver="solver PBiCGStab; preconditioner DILU; tolerance 1e-05; relTol 0;"),
    # periodic in z: real cyclicFvPatch / cyclicFvPatchField on the finest level, cyclicGAMGInterface on the coarse ones
    "cavity12x12x6_cyclic_gamg": dict(nx=12, ny=12, nz=6, lz=0.05, end_time=0.05, cyclic_z=True,
                                      p_solver="solver GAMG; smoother GaussSeidel; tolerance 1e-06; relTol 0.05;"),
    "cavity12x12x6_cyclic_pcg": dict(nx=12, ny=12, nz=6, lz=0.05, end_time=0.05, cyclic_z=True,
                                     u_solver="solver PBiCGStab; preconditioner DILU; tolerance 1e-05; relTol 0;"),
}


def main():
    if not ico.ICOFOAM.exists():
        sys.exit("oracle/_app/icoFoam missing: run python oracle/build_app.py")
    only = sys.argv[1:]
    for name, kw in CASES.items():
        if only and not any(o in name for o in only):
            continue
        with tempfile.TemporaryDirectory() as td:
            case = foam_case.write_cavity_case(Path(td) / "case", write=True, **kw)
            log = ico.run_icofoam(case)
            solves = ico.parse_log(log)
            t_dir = max((p for p in case.iterdir() if p.name.replace(".", "").isdigit() and p.name != "0"),
                        key=lambda p: float(p.name))
            e = {
                "case": repr(kw),
                "solverNames": "\n".join(s[0] for s in solves),
                "fields": "\n".join(s[1] for s in solves),
                "initialResidual": np.array([s[2] for s in solves]),
                "finalResidual": np.array([s[3] for s in solves]),
                "nIterations": np.array([s[4] for s in solves], dtype=np.int32),
                "p": ico.read_internal_field(t_dir / "p"),
                "U": ico.read_internal_field(t_dir / "U").ravel(),
            }
            ldu_io.write(str(HERE / f"icofoam_{name}.b2ls"), e)
            print(f"{name}: {len(solves)} solves, iterations {sorted(set(s[4] for s in solves))}, end time {t_dir.name}")


if __name__ == "__main__":
    main()
