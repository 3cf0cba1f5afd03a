"""Full-application drop-in test (SURVEY.md 8(c) "full-application oracle"): the reference's own icoFoam
(applications/legacy/incompressible/icoFoam, built unmodified by oracle/build_app.py) runs the lid-driven cavity with
`libs ("libB200LinearSolvers.so")` in controlDict and the B200 solver names in fvSolution -- every p and U solve of
every PISO corrector goes through fvMatrix::solveSegregated -> lduMatrix::solver::New -> the plugin -> the GPU.  The
solver log (name, field, residuals, iteration count of each of the ~100 solves) and the final fields are compared with
the golden run of the same binary with the reference's own solvers (tests/golden/icofoam_*.b2ls).
Skipped where oracle/_app was not built (it needs /root/reference at build time)."""
import ast

import numpy as np
import pytest

import _icofoam as ico
from _util import GOLDEN, ldu_io, load_pkg

load_pkg()
from b200ls import foam_case  # noqa: E402

pytestmark = pytest.mark.gpu

NAMES = sorted(p.stem[len("icofoam_"):] for p in GOLDEN.glob("icofoam_*.b2ls"))


def _b200(d):
    for a, b in (("solver PCG", "solver B200PCG"), ("solver PBiCGStab", "solver B200PBiCGStab"),
                 ("solver GAMG", "solver B200GAMG"), ("solver smoothSolver", "solver B200smoothSolver")):
        d = d.replace(a, b)
    return d


@pytest.mark.parametrize("name", NAMES)
def test_icofoam_with_b200_solvers_reproduces_the_reference_run(name, tmp_path):
    if not ico.ICOFOAM.exists() or not ico.PLUGIN.exists():
        pytest.skip("oracle/_app/icoFoam or the plugin was not built")
    gold = ldu_io.read(str(GOLDEN / f"icofoam_{name}.b2ls"))
    kw = ast.literal_eval(ldu_io.as_str(gold["case"]))
    import inspect

    defaults = {k: v.default for k, v in inspect.signature(foam_case.write_cavity_case).parameters.items()}
    for key in ("p_solver", "u_solver"):
        kw[key] = _b200(kw.get(key, defaults[key]))
    case = foam_case.write_cavity_case(tmp_path / "case", write=True, libs=f'"{ico.PLUGIN}"', **kw)
    solves = ico.parse_log(ico.run_icofoam(case))
    g_names = ldu_io.as_str(gold["solverNames"]).split("\n")
    g_fields = ldu_io.as_str(gold["fields"]).split("\n")
    assert len(solves) == len(g_names) > 20
    assert [s[0] for s in solves] == g_names            # "DICPCG", "smoothSolver", "GAMG", "DILUPBiCGStab"
    assert [s[1] for s in solves] == g_fields           # Ux, Uy, (Uz,) p, p, ...
    its = np.array([s[4] for s in solves])
    assert np.max(np.abs(its - gold["nIterations"])) <= 1, (its, gold["nIterations"])
    assert np.mean(its == gold["nIterations"]) >= 0.95
    ini = np.array([s[2] for s in solves])
    fin = np.array([s[3] for s in solves])
    # every solve starts from the fields the previous solves produced: agreement of the initial residuals to 1e-9
    # pins the whole time-stepping history, not only the individual solves
    assert np.max(np.abs(ini - gold["initialResidual"]) / np.maximum(gold["initialResidual"], 1e-300)) <= 1e-9
    same = its == gold["nIterations"]
    assert np.max(np.abs(fin - gold["finalResidual"])[same] / np.maximum(gold["initialResidual"][same], 1e-300)) <= 1e-9
    t_dir = max((p for p in case.iterdir() if p.name.replace(".", "").isdigit() and p.name != "0"),
                key=lambda p: float(p.name))
    p = ico.read_internal_field(t_dir / "p")
    U = ico.read_internal_field(t_dir / "U").ravel()
    assert np.max(np.abs(p - gold["p"])) <= 1e-9 * np.max(np.abs(gold["p"]))
    assert np.max(np.abs(U - gold["U"])) <= 1e-9 * np.max(np.abs(gold["U"]))
