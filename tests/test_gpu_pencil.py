"""Pencil sweeps of structured blocks (csrc/pencil.cuh: tile-major layout, warp-owned pencil tiles, bulk-copy operand
ring, neighbour values handed over through L2) against the C oracle (oracle/ldu_oracle.c, itself pinned to the
reference bit for bit): factorisation, DIC/DILU precondition, Gauss-Seidel / symGaussSeidel / DIC / DICGaussSeidel
smoothers bit-exact, solver histories to the usual bars.  Shapes cover partial tiles, odd sizes, 2-D and 1-D blocks,
blocks thinner than a tile, and several ring laps along i."""
import sys
from pathlib import Path

import numpy as np
import pytest

from _util import capi, cases, max_rel_diff

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle"))
import ldu_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu

SHAPES = [(12, 10, 9), (5, 40, 3), (7, 6, 1), (33, 1, 1), (9, 8, 2), (4, 37, 5), (40, 33, 17), (150, 9, 5),
          (3, 3, 3), (70, 70, 1), (17, 16, 12)]


@pytest.fixture(scope="module", autouse=True)
def _init():
    capi.init(0)


def _system(shape, sym):
    nx, ny, nz = shape
    return cases.cavity_laplacian(nx, ny, nz, coeffs="random") if sym else \
        cases.convection_diffusion(nx, ny, nz, dt_coeff=50.0)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("sym", [True, False])
def test_pencil_operators_bit_exact(shape, sym, monkeypatch):
    monkeypatch.setenv("B200LS_PENCIL_MIN_CELLS", "0")
    s = _system(shape, sym)
    mesh, mat = capi.from_system(s)
    assert mesh.get_i32(21, 0).size == 7          # the level has a pencil plan
    S = orc.System(s)
    kind = "DIC" if sym else "DILU"
    x = np.cos(0.7 * np.arange(s.n_cells)) + 0.3
    assert np.array_equal(mat.amul(x), orc.amul(S, x))
    assert np.array_equal(mat.residual(x, s.source), orc.residual(S, x, s.source))
    assert np.array_equal(mat.reciprocal_d(kind), orc.reciprocal_d(S))
    # twice: the sentinel re-arming between calls is exercised
    for seed in (0.37, 0.11):
        rA = np.cos(seed * np.arange(s.n_cells)) + 0.1
        assert np.array_equal(mat.precondition(kind, rA), orc.precondition(S, kind, rA))
    for sm in ("GaussSeidel", "symGaussSeidel", kind, kind + "GaussSeidel"):
        for n_sweeps in (1, 3):
            assert np.array_equal(mat.smooth(sm, x, s.source, n_sweeps), orc.smooth(S, sm, x, s.source, n_sweeps)), \
                (sm, n_sweeps)


@pytest.mark.parametrize("cfg", ["14", "18", "28"])
def test_pencil_ring_configurations_bit_exact(cfg, monkeypatch):
    """Every (skew, ring depth) instantiation of the substitution sweeps gives the same bits."""
    monkeypatch.setenv("B200LS_PENCIL_MIN_CELLS", "0")
    monkeypatch.setenv("B200LS_PENCIL_CFG", cfg)
    s = _system((45, 19, 11), True)
    mesh, mat = capi.from_system(s)
    S = orc.System(s)
    rA = np.cos(0.37 * np.arange(s.n_cells)) + 0.1
    assert np.array_equal(mat.precondition("DIC", rA), orc.precondition(S, "DIC", rA))


@pytest.mark.parametrize("shape", [(12, 10, 9), (40, 33, 17), (70, 70, 1)])
@pytest.mark.parametrize("sym", [True, False])
def test_wavefront_kernels_on_the_tile_major_layout(shape, sym, monkeypatch):
    """B200LS_PENCIL_SWEEPS=0 keeps the tile-major layout but runs the general wavefront kernels through the
    processing-order map: same bits."""
    monkeypatch.setenv("B200LS_PENCIL_MIN_CELLS", "0")
    monkeypatch.setenv("B200LS_PENCIL_SWEEPS", "0")
    s = _system(shape, sym)
    mesh, mat = capi.from_system(s)
    S = orc.System(s)
    kind = "DIC" if sym else "DILU"
    x = np.cos(0.7 * np.arange(s.n_cells)) + 0.3
    assert np.array_equal(mat.reciprocal_d(kind), orc.reciprocal_d(S))
    assert np.array_equal(mat.precondition(kind, x), orc.precondition(S, kind, x))
    for sm in ("GaussSeidel", "symGaussSeidel"):
        assert np.array_equal(mat.smooth(sm, x, s.source, 3), orc.smooth(S, sm, x, s.source, 3)), sm


@pytest.mark.parametrize("shape", [(12, 10, 9), (40, 33, 17), (70, 70, 1)])
@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("gamg_keeps_pencil", [False, True])
def test_pencil_solvers_match_oracle(shape, sym, gamg_keeps_pencil, monkeypatch):
    """Krylov solve on the pencil layout; then the mesh is agglomerated -- which puts the finest level back into the
    wavefront-major layout unless B200LS_PENCIL_GAMG=1 -- and GAMG / smoothSolver run on it."""
    monkeypatch.setenv("B200LS_PENCIL_MIN_CELLS", "0")
    if gamg_keeps_pencil:
        monkeypatch.setenv("B200LS_PENCIL_GAMG", "1")
    s = _system(shape, sym)
    mesh, mat = capi.from_system(s)
    S = orc.System(s)
    kind = "DIC" if sym else "DILU"
    runs = [("PCG" if sym else "PBiCGStab", dict(preconditioner=kind), kind),
            ("GAMG", dict(smoother="GaussSeidel"), "GaussSeidel"),
            ("smoothSolver", dict(smoother="symGaussSeidel", nSweeps=2), "symGaussSeidel")]
    for solver, kw, okind in runs:
        if solver == "GAMG":
            assert mesh.get_i32(21, 0).size == 7
            mesh.agglomerate(s.face_weights)
            assert (mesh.get_i32(21, 0).size == 7) == gamg_keeps_pencil
            mat.set(s.diag, s.upper_coeffs, s.lower_coeffs)
        ctl = capi.controls(solver, tolerance=1e-9, relTol=0.0, maxIter=60, recordHistory=1, **kw)
        psi, perf = mat.solve(ctl, s.source)
        okw = dict(tolerance=1e-9, maxIter=60)
        if "nSweeps" in kw:
            okw["nSweeps"] = kw["nSweeps"]
        xo, po = orc.solve(S, solver, orc.controls(okind, **okw), s.source)
        assert abs(perf.nIterations - po["nIterations"]) <= 1, (solver, perf.nIterations, po["nIterations"])
        assert abs(perf.initialResidual - po["initialResidual"]) <= 1e-9 * po["initialResidual"]
        if perf.nIterations == po["nIterations"]:
            tol = 1e-9 if solver != "PBiCGStab" or perf.nIterations <= 12 else 1e-8
            assert max_rel_diff(psi, xo) <= tol, (solver, max_rel_diff(psi, xo))



@pytest.mark.parametrize("shape", [(20, 8, 16), (40, 33, 17), (9, 9, 9), (24, 8, 16)])
@pytest.mark.parametrize("max_ctas", ["1", "3"])
def test_persistent_ctas_over_many_tiles(shape, max_ctas, monkeypatch):
    """Capping the grid (B200LS_PENCIL_MAX_CTAS) makes every CTA process many tiles in a row -- what happens at 128^3
    and beyond, here on small blocks with even and odd numbers of record groups per tile, partial tiles and partial
    operand chunks: ring stages, barrier phases and the per-tile re-initialisation carry over correctly."""
    monkeypatch.setenv("B200LS_PENCIL_MIN_CELLS", "0")
    monkeypatch.setenv("B200LS_PENCIL_MAX_CTAS", max_ctas)
    for sym in (True, False):
        s = _system(shape, sym)
        mesh, mat = capi.from_system(s)
        S = orc.System(s)
        kind = "DIC" if sym else "DILU"
        assert np.array_equal(mat.reciprocal_d(kind), orc.reciprocal_d(S))
        for seed in (0.37, 0.11):
            rA = np.cos(seed * np.arange(s.n_cells)) + 0.1
            assert np.array_equal(mat.precondition(kind, rA), orc.precondition(S, kind, rA))
        mat.close()
        mesh.close()
