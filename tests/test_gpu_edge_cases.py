"""Edge cases of the CUDA path, checked against the C oracle (oracle/ldu_oracle.c, itself pinned to the reference):
empty and face-less (diagonal) systems, a single cell, disconnected cells, chains (one row per wavefront), rows with
many neighbours, early exit at the initial residual, maxIter/minIter handling, singular systems."""
import sys
from pathlib import Path

import numpy as np
import pytest

from _util import capi, cases, max_rel_diff

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle"))
import ldu_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _init():
    capi.init(0)


def _sys(n, lower, upper, diag, up, lo=None, src=None):
    return cases.LduSystem(n_cells=n, lower=np.array(lower, dtype=np.int32), upper=np.array(upper, dtype=np.int32),
                           diag=np.array(diag, dtype=float), upper_coeffs=np.array(up, dtype=float),
                           lower_coeffs=None if lo is None else np.array(lo, dtype=float),
                           source=np.array(src if src is not None else np.ones(n), dtype=float))


def _compare_ops(s):
    mesh, mat = capi.from_system(s)
    S = orc.System(s)
    x = np.cos(0.7 * np.arange(s.n_cells)) + 0.3
    assert np.array_equal(mat.amul(x), orc.amul(S, x))
    assert np.array_equal(mat.residual(x, s.source), orc.residual(S, x, s.source))
    assert np.array_equal(mat.sum_a(), orc.sum_a(S))
    kind = "DIC" if s.symmetric else "DILU"
    assert np.array_equal(mat.reciprocal_d(kind), orc.reciprocal_d(S))
    assert np.array_equal(mat.precondition(kind, x), orc.precondition(S, kind, x))
    for sm in ("GaussSeidel", "symGaussSeidel", kind, kind + "GaussSeidel"):
        assert np.array_equal(mat.smooth(sm, x, s.source, 2), orc.smooth(S, sm, x, s.source, 2)), sm
    return mesh, mat, S


def test_single_cell_and_diagonal_systems():
    for n in (1, 5):
        s = _sys(n, [], [], 2.0 + np.arange(n), [])
        mesh, mat, S = _compare_ops(s)
        for solver, pre in (("PCG", "DIC"), ("PCG", "diagonal"), ("PCG", "none"), ("PBiCGStab", "DILU")):
            psi, perf = mat.solve(capi.controls(solver, pre, tolerance=1e-12, relTol=0.0), s.source)
            assert np.allclose(psi, s.source / s.diag, rtol=1e-14, atol=0)
            assert perf.converged


def test_empty_system():
    s = _sys(0, [], [], [], [])
    mesh, mat = capi.from_system(s)
    assert mat.amul(np.zeros(0)).size == 0


def test_chain_one_row_per_wavefront():
    n = 300
    s = _sys(n, np.arange(n - 1), np.arange(1, n), np.full(n, 2.5), np.full(n - 1, -1.0),
             src=np.sin(np.arange(n)))
    mesh, mat, S = _compare_ops(s)
    assert mesh.get_i32(capi.FWD_LEVEL_OFFSETS).size - 1 == n
    psi, perf = mat.solve(capi.controls("PCG", "DIC", tolerance=1e-12, relTol=0.0), s.source)
    o_psi, o_perf = orc.solve(S, "PCG", orc.controls("DIC", tolerance=1e-12), s.source)
    assert perf.nIterations == o_perf["nIterations"] and max_rel_diff(psi, o_psi) < 1e-12


def test_disconnected_and_high_degree_rows():
    # a star (cell 0 touches everybody), plus two isolated cells at the end
    n = 70
    lower = [0] * (n - 3)
    upper = list(range(1, n - 2))
    diag = np.full(n, 80.0)
    s = _sys(n, lower, upper, diag, -np.linspace(0.5, 1.5, n - 3), lo=-np.linspace(1.5, 0.5, n - 3),
             src=np.cos(np.arange(n)))
    _compare_ops(s)
    s2 = cases.random_graph(400, avg_degree=12, symmetric=True, seed=3, max_span=40)
    mesh, mat, S = _compare_ops(s2)
    mesh.agglomerate(s2.face_weights)
    mat.set(s2.diag, s2.upper_coeffs)
    psi, perf = mat.solve(capi.controls("GAMG", smoother="GaussSeidel", tolerance=1e-10, relTol=0.0), s2.source)
    o_psi, o_perf = orc.solve(S, "GAMG", orc.controls("GaussSeidel", tolerance=1e-10), s2.source)
    assert abs(perf.nIterations - o_perf["nIterations"]) <= 1
    if perf.nIterations == o_perf["nIterations"]:
        assert max_rel_diff(psi, o_psi) < 1e-9


def test_control_semantics_match_oracle():
    s = cases.cavity_laplacian(9, 7, 5, coeffs="random", rhs_kind="uniform")
    mesh, mat = capi.from_system(s)
    S = orc.System(s)
    exact_ctl = capi.controls("PCG", "DIC", tolerance=1e-13, relTol=0.0)
    psi_exact, _ = mat.solve(exact_ctl, s.source)
    # already converged at the initial residual: no iterations, psi untouched
    psi, perf = mat.solve(capi.controls("PCG", "DIC", tolerance=1e-6, relTol=0.0), s.source, psi0=psi_exact)
    assert perf.nIterations == 0 and perf.converged and np.array_equal(psi, psi_exact)
    # ... unless minIter forces work
    psi, perf = mat.solve(capi.controls("PCG", "DIC", tolerance=1e-6, relTol=0.0, minIter=3), s.source, psi0=psi_exact)
    o_psi, o_perf = orc.solve(S, "PCG", orc.controls("DIC", tolerance=1e-6, minIter=3), s.source, psi0=psi_exact)
    assert perf.nIterations == o_perf["nIterations"] == 3
    # maxIter caps the loop and is not an error
    for solver, pre in (("PCG", "DIC"), ("PBiCGStab", "DIC")):
        psi, perf = mat.solve(capi.controls(solver, pre, tolerance=1e-30, relTol=0.0, maxIter=4), s.source)
        o_psi, o_perf = orc.solve(S, solver, orc.controls(pre, tolerance=1e-30, maxIter=4), s.source)
        assert perf.nIterations == o_perf["nIterations"] == 4 and not perf.converged
        assert max_rel_diff(psi, o_psi) < 1e-10
    # relTol
    psi, perf = mat.solve(capi.controls("PCG", "DIC", tolerance=0.0, relTol=0.1), s.source)
    o_psi, o_perf = orc.solve(S, "PCG", orc.controls("DIC", tolerance=0.0, relTol=0.1), s.source)
    assert perf.nIterations == o_perf["nIterations"] and perf.converged


def test_zero_rhs_is_singular_like_the_reference():
    # source = 0, psi = 0: normFactor = small, wApA = 0 -> the reference breaks out with singular=true after the
    # first Amul (PCG.C:165); the residual stays at 0/1e-20
    s = cases.cavity_laplacian(6, 5, 4)
    s.source = np.zeros(s.n_cells)
    mesh, mat = capi.from_system(s)
    S = orc.System(s)
    psi, perf = mat.solve(capi.controls("PCG", "DIC", tolerance=1e-6, relTol=0.0, minIter=1), s.source)
    o_psi, o_perf = orc.solve(S, "PCG", orc.controls("DIC", tolerance=1e-6, minIter=1), s.source)
    assert bool(perf.singular) == o_perf["singular"] and perf.nIterations == o_perf["nIterations"]
    assert np.array_equal(psi, o_psi)


def test_pcg_rejects_asymmetric_matrix():
    s = cases.convection_diffusion(6, 5, 1)
    mesh, mat = capi.from_system(s)
    with pytest.raises(capi.B200Error, match="symmetric"):
        mat.solve(capi.controls("PCG", "DIC"), s.source)


def test_solve_dev_matches_host_pointer_solve():
    """b200ls_solve_dev (psi/source already resident in HBM, cell order) == b200ls_solve (host pointers)."""
    torch = pytest.importorskip("torch")
    s = cases.cavity_laplacian(12, 9, 7, coeffs="random", rhs_kind="uniform")
    mesh, mat = capi.from_system(s)
    ctl = capi.controls("PCG", "DIC", tolerance=1e-10, relTol=0.0)
    psi_h, perf_h = mat.solve(ctl, s.source)
    d_src = torch.from_numpy(s.source).cuda()
    d_psi = torch.zeros(s.n_cells, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    perf_d = mat.solve_dev(ctl, d_psi.data_ptr(), d_src.data_ptr())
    assert perf_d.nIterations == perf_h.nIterations
    assert np.array_equal(d_psi.cpu().numpy(), psi_h)
    assert perf_d.kernelLaunches > 0 and perf_d.solveMs > 0


def test_diagonal_solver():
    """diagonalSolver.C:62-79: psi = source/diag, zero residuals, zero iterations, converged."""
    s = cases.convection_diffusion(9, 7, 5, dt_coeff=50.0)
    mesh, mat = capi.from_system(s)
    psi, perf = mat.solve(capi.controls("diagonal"), s.source, psi0=np.ones(s.n_cells))
    assert np.array_equal(psi, s.source / s.diag)
    assert (perf.nIterations, perf.initialResidual, perf.finalResidual, perf.converged) == (0, 0.0, 0.0, 1)
