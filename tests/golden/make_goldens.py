#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference solver stack
(oracle/_ref/ref_harness, built from /root/reference by oracle/build_ref.py) on the synthetic systems of
openfoam-dev_b200/cases.py.  Run in the build container only (needs /root/reference to build oracle/_ref):

    python tests/golden/make_goldens.py

Each fixture <name>.b2ls holds the harness INPUT entries (prefix "in.") and its OUTPUT entries (prefix "ref.").
One harness process per fixture (pairGAMGAgglomeration::forward_ is a process-global static).
"""
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT / "tests"))
from _pkg import load_pkg  # noqa: E402

b200ls = load_pkg()
from b200ls import cases, decompose, ldu_io  # noqa: E402


def run_ref(entries):
    env = dict(os.environ, WM_PROJECT_DIR=str(ROOT / "oracle/foam_env"), WM_PROJECT="OpenFOAM",
               WM_PROJECT_VERSION="dev")
    with tempfile.TemporaryDirectory() as td:
        ldu_io.write(f"{td}/in.b2ls", entries)
        r = subprocess.run([str(ROOT / "oracle/_ref/ref_harness"), f"{td}/in.b2ls", f"{td}/out.b2ls", f"{td}/case"],
                           env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stdout[-2000:] + r.stderr[-2000:])
        return ldu_io.read(f"{td}/out.b2ls")


SYM_SOLVES = [
    ("solver PCG; preconditioner DIC; tolerance 1e-6; relTol 0.05;", 30),
    ("solver PCG; preconditioner DIC; tolerance 1e-10; relTol 0;", 80),
    ("solver PCG; preconditioner diagonal; tolerance 1e-8; relTol 0; maxIter 40;", 40),
    ("solver PCG; preconditioner none; tolerance 1e-6; relTol 0; maxIter 25;", 0),
    ("solver PBiCGStab; preconditioner DIC; tolerance 1e-8; relTol 0;", 45),
    ("solver GAMG; smoother GaussSeidel; tolerance 1e-8; relTol 0;", 25),
    ("solver GAMG; smoother DIC; tolerance 1e-8; relTol 0;", 25),
    ("solver GAMG; smoother GaussSeidel; tolerance 1e-6; relTol 0.01; nPreSweeps 1; nFinestSweeps 3;", 0),
    ("solver smoothSolver; smoother GaussSeidel; nSweeps 2; tolerance 1e-3; relTol 0; maxIter 40;", 0),
    ("solver smoothSolver; smoother DIC; nSweeps 1; tolerance 1e-3; relTol 0; maxIter 40;", 0),
    ("solver smoothSolver; smoother symGaussSeidel; nSweeps 1; tolerance 1e-3; relTol 0; maxIter 30;", 0),
    ("solver GAMG; smoother DICGaussSeidel; tolerance 1e-8; relTol 0;", 0),
    ("solver GAMG; smoother symGaussSeidel; tolerance 1e-8; relTol 0;", 0),
    ("solver PCG; preconditioner { preconditioner GAMG; smoother GaussSeidel; nVcycles 2; tolerance 1e-5; "
     "relTol 0; } tolerance 1e-9; relTol 0;", 12),
]

ASYM_SOLVES = [
    ("solver PBiCGStab; preconditioner DILU; tolerance 1e-10; relTol 0;", 30),
    ("solver PBiCGStab; preconditioner DILU; tolerance 1e-6; relTol 0.1;", 0),
    ("solver PBiCGStab; preconditioner diagonal; tolerance 1e-8; relTol 0; maxIter 15;", 15),
    ("solver GAMG; smoother GaussSeidel; tolerance 1e-8; relTol 0;", 20),
    ("solver GAMG; smoother DILU; tolerance 1e-8; relTol 0;", 20),
    ("solver smoothSolver; smoother GaussSeidel; nSweeps 1; tolerance 1e-6; relTol 0; maxIter 60;", 0),
    ("solver smoothSolver; smoother DILU; nSweeps 2; tolerance 1e-6; relTol 0; maxIter 60;", 0),
    ("solver smoothSolver; smoother symGaussSeidel; nSweeps 2; tolerance 1e-6; relTol 0; maxIter 60;", 0),
    ("solver GAMG; smoother DILUGaussSeidel; tolerance 1e-8; relTol 0;", 0),
    ("solver PBiCGStab; preconditioner { preconditioner GAMG; smoother GaussSeidel; nVcycles 1; tolerance 1e-5; "
     "relTol 0; } tolerance 1e-9; relTol 0;", 8),
]


def decomposed_case(kind, n_ranks):
    """The decomposed systems of tests/_mgpu_worker.py folded into one block-diagonal system whose processor patches
    are cyclic pairs (decompose.as_cyclic_blocks): the serial reference then runs its decomposed algorithm."""
    split = decompose.simple_split(n_ranks)
    nx, ny, nz = 10 * split[0], 8 * split[1], 6 * split[2]
    glob = cases.cavity_laplacian(nx, ny, nz, coeffs="random") if kind == "sym" else \
        cases.convection_diffusion(nx, ny, nz, dt_coeff=50.0)
    parts, _ = decompose.decompose_system(glob, decompose.box_cell_ranks(nx, ny, nz, split), n_ranks)
    return decompose.as_cyclic_blocks(parts)


def polymesh_dump(tutorial_mesh):
    """`ref_harness --polymesh` on a mesh shipped with the reference (tutorials/.../constant/polyMesh)."""
    env = dict(os.environ, WM_PROJECT_DIR=str(ROOT / "oracle/foam_env"), WM_PROJECT="OpenFOAM",
               WM_PROJECT_VERSION="dev")
    with tempfile.TemporaryDirectory() as td:
        case = Path(td) / "case"
        (case / "system").mkdir(parents=True)
        (case / "constant").mkdir()
        (case / "constant" / "polyMesh").symlink_to(tutorial_mesh)
        (case / "system" / "controlDict").write_text(
            "FoamFile{format ascii; class dictionary; object controlDict;}\napplication harness;\n"
            "startFrom startTime;\nstartTime 0;\nstopAt endTime;\nendTime 1;\ndeltaT 1;\n"
            "writeControl timeStep;\nwriteInterval 1000000;\n")
        r = subprocess.run([str(ROOT / "oracle/_ref/ref_harness"), "--polymesh", str(case), f"{td}/mesh.b2ls"],
                           env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stdout[-2000:] + r.stderr[-2000:])
        return ldu_io.read(f"{td}/mesh.b2ls")


def icofoam_dumps(nx, ny, nz, steps):
    """Real p and Ux systems of the reference's icoFoam (oracle/_app, oracle/build_app.py) captured at the drop-in
    boundary by the plugin's pass-through solver B200dump during the last time step; returns {field: LduSystem}."""
    sys.path.insert(0, str(ROOT / "tests"))
    import _icofoam as ico
    from _util import system_from_entries
    from b200ls import foam_case

    if not ico.ICOFOAM.exists():
        return None
    with tempfile.TemporaryDirectory() as td:
        dump = f'solver B200dump; dumpFile "{td}/'
        case = foam_case.write_cavity_case(
            Path(td) / "case", nx=nx, ny=ny, nz=nz, end_time=steps * 0.005, libs=f'"{ico.PLUGIN}"',
            p_solver=dump + 'p"; delegate PCG; preconditioner DIC; tolerance 1e-06; relTol 0.05;',
            u_solver=dump + 'U"; delegate smoothSolver; smoother symGaussSeidel; tolerance 1e-05; relTol 0;')
        ico.run_icofoam(case)
        out = {}
        lower, upper, fdir = cases.block_addressing(nx, ny, nz)
        h = np.array([0.1 / nx, 0.1 / ny, 0.01 / nz])
        area = np.array([h[1] * h[2], h[0] * h[2], h[0] * h[1]])
        w = (area / np.sqrt(area) * np.array([1.0, 1.01, 1.02]))[fdir]       # faceAreaPair weights of this mesh
        for fld in ("p", "U"):
            files = sorted(Path(td).glob(f"{fld}.*.b2ls"), key=lambda q: int(q.name.split(".")[1]))
            # the counter is shared by all B200dump solves of the run: take the last p solve / the last Ux solve
            pick = files[-1] if fld == "p" else files[-(3 if nz > 1 else 2)]
            d = ldu_io.read(str(pick))
            s_ = system_from_entries(d)
            assert np.array_equal(s_.lower, lower) and np.array_equal(s_.upper, upper)
            s_.face_weights = w
            out[fld] = (s_, d["psi0"])
        return out


def with_coarsest(solves, n):
    return [(d.replace("solver GAMG;", f"solver GAMG; nCellsInCoarsestLevel {n};")
              .replace("preconditioner GAMG;", f"preconditioner GAMG; nCellsInCoarsestLevel {n};"), h)
            for d, h in solves]


def fixture(name, sys_, solves, smoothers, agglom=True, agglom_dict="solver GAMG;", extra=None):
    e = cases.to_entries(sys_)
    e.update(extra or {})
    n = sys_.n_cells
    e["x"] = np.cos(0.11 * np.arange(n)) + 0.25
    for i, (d, hist) in enumerate(solves):
        e[f"solve.{i}.dict"] = d
        if hist:
            e[f"solve.{i}.history"] = hist
    for i, (d, ns) in enumerate(smoothers):
        e[f"smooth.{i}.dict"] = d
        e[f"smooth.{i}.nSweeps"] = ns
    if agglom:
        e["agglomerate.dict"] = agglom_dict
    if (HERE / f"{name}.b2ls").exists() and "--all" not in sys.argv:
        print(f"{name}: kept (pass --all to regenerate)")
        return
    out = run_ref(e)
    merged = {f"in.{k}": v for k, v in e.items()}
    merged.update({f"ref.{k}": v for k, v in out.items()})
    ldu_io.write(str(HERE / f"{name}.b2ls"), merged)
    its = [int(out[f"solve.{i}.perf"][2]) for i in range(len(solves))]
    print(f"{name}: nCells={n} nFaces={sys_.n_faces} levels={int(out['agg.nLevels'][0]) if agglom else -1} iterations={its}")


def main():
    if not (ROOT / "oracle/_ref/ref_harness").exists():
        sys.exit("oracle/_ref/ref_harness missing: run python oracle/build_ref.py first")
    sym_sm = [("smoother GaussSeidel;", 1), ("smoother GaussSeidel;", 3), ("smoother DIC;", 2),
              ("smoother symGaussSeidel;", 2), ("smoother DICGaussSeidel;", 1)]
    asym_sm = [("smoother GaussSeidel;", 2), ("smoother DILU;", 2), ("smoother symGaussSeidel;", 1),
               ("smoother DILUGaussSeidel;", 2)]
    # config 1: the cavity tutorial mesh, 20x20x1, uniform Laplacian (PCG+DIC as icoFoam/cavity ships it)
    fixture("cavity_20x20x1", cases.cavity_laplacian(20, 20, 1), SYM_SOLVES, sym_sm)
    fixture("block_7x5x3_rand", cases.cavity_laplacian(7, 5, 3, coeffs="random", rhs_kind="uniform"),
            SYM_SOLVES, sym_sm)
    fixture("block_16x16x16_rand", cases.cavity_laplacian(16, 16, 16, coeffs="random"), SYM_SOLVES, sym_sm)
    fixture("block_24x24x24", cases.cavity_laplacian(24, 24, 24, rhs_kind="uniform"), SYM_SOLVES[1:2] + SYM_SOLVES[5:6],
            sym_sm[:1], agglom=False)
    # irregular connectivity: rows with many neighbours, ragged wavefronts, irregular agglomeration
    fixture("random_graph_sym_600", cases.random_graph(600, symmetric=True), SYM_SOLVES, sym_sm)
    fixture("random_graph_asym_500", cases.random_graph(500, symmetric=False, seed=7), ASYM_SOLVES, asym_sm)
    fixture("convdiff_24x18x1", cases.convection_diffusion(24, 18, 1, dt_coeff=50.0), ASYM_SOLVES, asym_sm)
    fixture("convdiff_9x8x7", cases.convection_diffusion(9, 8, 7, dt_coeff=50.0, rhs_kind="uniform"), ASYM_SOLVES, asym_sm)
    # cyclic (periodic) patch pairs: the serial pin of the coupled-interface code that processor patches share
    # (interface terms of Amul/residual/sumA/smoothers, cyclicGAMGInterface agglomeration, coarse interface
    # coefficients, coarsest-level solve with interfaces)
    fixture("cyclic_x_12x10x8_rand", cases.add_cyclic(cases.cavity_laplacian(12, 10, 8, coeffs="random"), 0),
            SYM_SOLVES, sym_sm)
    fixture("cyclic_xz_9x8x7_rand",
            cases.add_cyclic(cases.add_cyclic(cases.cavity_laplacian(9, 8, 7, coeffs="random", rhs_kind="uniform"), 0), 2),
            SYM_SOLVES, sym_sm)
    fixture("cyclic_y_convdiff_10x9x6", cases.add_cyclic(cases.convection_diffusion(10, 9, 6, dt_coeff=50.0), 1),
            ASYM_SOLVES, asym_sm)
    # decomposed runs, executed by the SERIAL reference as block-diagonal systems with cyclic pairs in place of the
    # processor patches; nCellsInCoarsestLevel = 10*nRanks reproduces the decomposed stop criterion
    for kind, n_ranks, solves, sm in (("sym", 2, SYM_SOLVES, sym_sm), ("asym", 2, ASYM_SOLVES, asym_sm),
                                      ("sym", 4, SYM_SOLVES, sym_sm), ("asym", 4, ASYM_SOLVES, asym_sm),
                                      # eight ranks (2 2 2): the Krylov / GAMG runs bench.py --gpus 8 checks before timing
                                      ("sym", 8, SYM_SOLVES[:2] + SYM_SOLVES[5:6], sym_sm[:1])):
        blk, offs = decomposed_case(kind, n_ranks)
        fixture(f"decomp{n_ranks}_{kind}", blk, with_coarsest(solves, 10 * n_ranks), sm,
                agglom_dict=f"solver GAMG; nCellsInCoarsestLevel {10 * n_ranks};",
                extra={"rankOffsets": offs.astype(np.int32)})
    # real meshes shipped with the reference, read by the reference's own polyMesh (addressing + geometry):
    # airFoil2D (10,720 cells, 2-D unstructured C-mesh) and tank3D (3-D)
    tut = Path("/root/reference/tutorials")
    af = polymesh_dump(tut / "incompressibleFluid/airFoil2D/constant/polyMesh")
    fixture("airfoil2d_p", cases.polymesh_laplacian(af), SYM_SOLVES[:2] + SYM_SOLVES[4:7] + SYM_SOLVES[10:], sym_sm)
    fixture("airfoil2d_U", cases.polymesh_convection_diffusion(af), ASYM_SOLVES[:1] + ASYM_SOLVES[3:], asym_sm)
    tk = polymesh_dump(tut / "incompressibleDriftFlux/tank3D/constant/polyMesh")
    fixture("tank3d_p", cases.polymesh_laplacian(tk, rhs_kind="uniform"), SYM_SOLVES[1:2] + SYM_SOLVES[5:6] + SYM_SOLVES[11:12],
            sym_sm[:2], agglom=True)
    # REAL matrices: the p and Ux equations the reference's icoFoam assembles on the cavity (time step 10), captured by
    # B200dump inside the running application; psi0 = the field the application started that solve from
    for nm, dims in (("icofoamsys_20x20x1", (20, 20, 1)), ("icofoamsys_12x12x6", (12, 12, 6))):
        dumps = icofoam_dumps(*dims, steps=10)
        if dumps is None:
            print(f"{nm}: skipped (oracle/_app/icoFoam not built)")
            continue
        fixture(nm + "_p", dumps["p"][0], SYM_SOLVES, sym_sm, extra={"psi0": dumps["p"][1]})
        fixture(nm + "_U", dumps["U"][0], ASYM_SOLVES, asym_sm, extra={"psi0": dumps["U"][1]})


if __name__ == "__main__":
    main()
