"""CPU-only check of the drop-in boundary: the unmodified reference loads libB200LinearSolvers.so through `libs`,
selects `B200dump` from its solver table, and B200dump (a) serialises exactly the system the reference handed over and
(b) delegates to the reference's own solver, so the result equals the golden.  The dump replays through the C oracle.
Skipped where oracle/_ref was not built (it needs /root/reference at build time)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from _util import ldu_io, load_fixture, solve_keys, system_from_entries

ROOT = Path(__file__).resolve().parent.parent
HARNESS = ROOT / "oracle/_ref/ref_harness"
PLUGIN = ROOT / "openfoam-dev_b200/libB200LinearSolvers.so"

sys.path.insert(0, str(ROOT / "oracle"))
import ldu_oracle as orc  # noqa: E402


@pytest.mark.parametrize("name,pick", [("block_7x5x3_rand", "solver PCG; preconditioner DIC; tolerance 1e-10"),
                                        ("cyclic_y_convdiff_10x9x6", "solver PBiCGStab; preconditioner DILU; tolerance 1e-10")])
def test_dump_solver_captures_the_system_and_delegates(name, pick, tmp_path):
    if not HARNESS.exists() or not PLUGIN.exists():
        pytest.skip("oracle/_ref or the plugin was not built")
    inp, ref = load_fixture(name)
    i, text = [(i, t) for i, t in solve_keys(inp) if t.startswith(pick)][0]
    e = {k: v for k, v in inp.items() if not k.startswith(("solve.", "smooth.", "agglomerate", "x"))}
    e["libs"] = f'"{PLUGIN}"'
    solver = text.split(";")[0].split()[1]
    e["solve.0.dict"] = text.replace(f"solver {solver};", f'solver B200dump; delegate {solver}; dumpFile "{tmp_path}/sys";')
    ldu_io.write(str(tmp_path / "in.b2ls"), e)
    env = dict(os.environ, WM_PROJECT_DIR=str(ROOT / "oracle/foam_env"), WM_PROJECT="OpenFOAM",
               WM_PROJECT_VERSION="dev")
    r = subprocess.run([str(HARNESS), str(tmp_path / "in.b2ls"), str(tmp_path / "out.b2ls"), str(tmp_path / "case")],
                       env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = ldu_io.read(str(tmp_path / "out.b2ls"))
    # (b) delegated result == the reference's own
    assert np.array_equal(out["solve.0.perf"][:5], ref[f"solve.{i}.perf"][:5])
    assert np.array_equal(out["solve.0.psi"], ref[f"solve.{i}.psi"])
    # (a) the dump holds the system bit for bit
    dump = ldu_io.read(str(tmp_path / "sys.0.b2ls"))
    for k in ("lower", "upper", "diag", "upperCoeffs", "source"):
        assert np.array_equal(dump[k], inp[k]), k
    if "lowerCoeffs" in inp:
        assert np.array_equal(dump["lowerCoeffs"], inp["lowerCoeffs"])
    n_if = int(inp["nIfaces"][0]) if "nIfaces" in inp else 0
    assert int(dump["nIfaces"][0]) == n_if
    for k in range(n_if):
        for f in ("faceCells", "bouCoeffs", "intCoeffs", "nbrPatch"):
            assert np.array_equal(dump[f"iface.{k}.{f}"], inp[f"iface.{k}.{f}"]), (k, f)
    assert solver in ldu_io.as_str(dump["solve.0.dict"]) and "B200dump" not in ldu_io.as_str(dump["solve.0.dict"])
    # ... and replays through the oracle to the same answer
    s = system_from_entries(dump)
    psi, perf = orc.solve(orc.System(s), solver, orc.controls("DIC" if solver == "PCG" else "DILU", tolerance=1e-10),
                          s.source, psi0=dump["psi0"])
    assert np.array_equal(psi, ref[f"solve.{i}.psi"]) and perf["nIterations"] == int(ref[f"solve.{i}.perf"][2])


def test_generated_cavity_mesh_is_what_the_reference_reads(tmp_path):
    """openfoam-dev_b200/foam_case.py writes the polyMesh of the application tests; the reference's own polyMesh must
    read it as a valid mesh with exactly the LDU addressing of cases.block_addressing (upper-triangular order)."""
    if not HARNESS.exists():
        pytest.skip("oracle/_ref was not built")
    from _util import cases, load_pkg

    load_pkg()
    from b200ls import foam_case

    nx, ny, nz = 7, 5, 3
    case = foam_case.write_cavity_case(tmp_path / "case", nx=nx, ny=ny, nz=nz, lz=0.03, cyclic_z=True)
    env = dict(os.environ, WM_PROJECT_DIR=str(ROOT / "oracle/foam_env"), WM_PROJECT="OpenFOAM",
               WM_PROJECT_VERSION="dev")
    r = subprocess.run([str(HARNESS), "--polymesh", str(case), str(tmp_path / "mesh.b2ls")], env=env,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    d = ldu_io.read(str(tmp_path / "mesh.b2ls"))
    lower, upper, _ = cases.block_addressing(nx, ny, nz)
    assert np.array_equal(d["lower"], lower) and np.array_equal(d["upper"], upper)
    assert np.allclose(d["cellVolumes"], (0.1 / nx) * (0.1 / ny) * (0.03 / nz), rtol=1e-12)
    Sf, Cf, C = d["faceAreas"].reshape(-1, 3), d["faceCentres"].reshape(-1, 3), d["cellCentres"].reshape(-1, 3)
    assert np.all(((Cf - C[d["faceOwner"]]) * Sf).sum(1) > 0)          # every face normal points out of its owner
    types = [ldu_io.as_str(d[f"patch.{i}.nameType"]) for i in range(int(d["nPatches"][0]))]
    assert types == ["movingWall wall", "fixedWalls wall", "front cyclic", "back cyclic"]


def test_icofoam_runs_unchanged_with_the_plugin_loaded_on_cpu(tmp_path):
    """The reference's icoFoam (oracle/_app) with libs ("libB200LinearSolvers.so") and B200dump pass-through solvers for
    p and U: the plugin loads inside the real application without a GPU, every solve is captured, and the solver log is
    identical to the run without the plugin."""
    import _icofoam as ico
    from _util import load_pkg

    load_pkg()
    from b200ls import foam_case

    if not ico.ICOFOAM.exists() or not PLUGIN.exists():
        pytest.skip("oracle/_app/icoFoam or the plugin was not built")
    kw = dict(nx=8, ny=8, nz=1, end_time=0.015)
    ref_log = ico.parse_log(ico.run_icofoam(foam_case.write_cavity_case(tmp_path / "ref", **kw)))
    dump = f'solver B200dump; dumpFile "{tmp_path}/'
    case = foam_case.write_cavity_case(
        tmp_path / "dump", libs=f'"{PLUGIN}"',
        p_solver=dump + 'p"; delegate PCG; preconditioner DIC; tolerance 1e-06; relTol 0.05;',
        u_solver=dump + 'U"; delegate smoothSolver; smoother symGaussSeidel; tolerance 1e-05; relTol 0;', **kw)
    log = ico.parse_log(ico.run_icofoam(case))
    assert log == ref_log and len(log) == 12
    dumps = sorted(tmp_path.glob("*.b2ls"))
    assert len(dumps) == 12                                  # 3 steps x (Ux, Uy, p, p)
    d = ldu_io.read(str(tmp_path / "p.2.b2ls"))              # the first p solve
    assert int(d["nCells"][0]) == 64 and "lowerCoeffs" not in d and d["lower"].size == 2 * 8 * 7
