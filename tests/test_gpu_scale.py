"""Parity at BASELINE sizes ("no -m gpu parity test at any BASELINE size"): the 128^3 cavity p-equation
of configs[1] and an unstructured system of 343,000 rows, against the C oracle (oracle/ldu_oracle.c -- pinned bit for
bit to the unmodified reference by tests/test_oracle.py).  At these sizes every warp of the persistent wavefront grids
takes several tasks, the grid-stride reductions wrap, and the pencil kernels run several tiles per CTA.

Bars: Amul / residual / reciprocalD / precondition / Gauss-Seidel sweeps bit-exact (np.array_equal); PCG and GAMG
iteration counts equal, every residual of the history within 1e-9 of the reference on the scale of the initial
residual, solution within 1e-9 max relative difference."""
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


def _unstructured():
    # a 70^3 block with randomly weighted faces, renumbered at random inside windows of 4,096 labels: irregular rows,
    # ragged wavefronts, no structure for the pencil path to find
    return cases.renumbered(cases.cavity_laplacian(70, 70, 70, coeffs="random"), window=4096)


CASES = {
    "cavity128": lambda: cases.cavity_laplacian(128, 128, 128),
    "cavity128_wavefront_layout": lambda: cases.cavity_laplacian(128, 128, 128),
    "unstructured343k": _unstructured,
}


@pytest.fixture(scope="module", params=list(CASES))
def big(request):
    mp = pytest.MonkeyPatch()
    if request.param.endswith("wavefront_layout"):
        mp.setenv("B200LS_PENCIL", "0")         # the general wavefront kernels at the same size
    s = CASES[request.param]()
    mesh, mat = capi.from_system(s)
    assert (mesh.get_i32(21, 0).size == 7) == (request.param == "cavity128")
    yield request.param, s, orc.System(s), mesh, mat
    mat.close()
    mesh.close()
    mp.undo()


def test_operators_bit_exact_at_scale(big):
    name, s, S, mesh, mat = big
    x = np.cos(0.7 * np.arange(s.n_cells)) + 0.3
    assert np.array_equal(mat.amul(x), orc.amul(S, x))
    assert np.array_equal(mat.residual(x, s.source), orc.residual(S, x, s.source))
    assert np.array_equal(mat.sum_a(), orc.sum_a(S))
    assert np.array_equal(mat.reciprocal_d("DIC"), orc.reciprocal_d(S))
    for seed in (0.37, 0.11):      # twice: the sentinel re-arming between calls
        rA = np.cos(seed * np.arange(s.n_cells)) + 0.1
        assert np.array_equal(mat.precondition("DIC", rA), orc.precondition(S, "DIC", rA))
    assert np.array_equal(mat.smooth("GaussSeidel", x, s.source, 2), orc.smooth(S, "GaussSeidel", x, s.source, 2))
    assert np.array_equal(mat.smooth("symGaussSeidel", x, s.source, 1), orc.smooth(S, "symGaussSeidel", x, s.source, 1))


def test_pcg_history_at_scale(big):
    """configs[1] as benchmarked: PCG + DIC, 50 iterations."""
    name, s, S, mesh, mat = big
    ctl = capi.controls("PCG", "DIC", tolerance=0.0, relTol=0.0, maxIter=50, recordHistory=1)
    psi, perf = mat.solve(ctl, s.source)
    xo, po = orc.solve(S, "PCG", orc.controls("DIC", tolerance=0.0, relTol=0.0, maxIter=50), s.source)
    assert perf.nIterations == po["nIterations"] == 50
    assert abs(perf.initialResidual - po["initialResidual"]) <= 1e-9 * po["initialResidual"]
    h = capi.history(perf)
    assert len(h) == len(po["history"]) == 50
    assert np.all(np.abs(h - po["history"]) <= 1e-9 * po["initialResidual"])
    assert max_rel_diff(psi, xo) <= 1e-9


def test_gamg_history_at_scale(big):
    """GAMG + GaussSeidel, three V-cycles (agglomeration included: restrictAddressing of every level bit-exact)."""
    name, s, S, mesh, mat = big
    if name == "cavity128_wavefront_layout":
        pytest.skip("same kernels as cavity128 once the mesh is agglomerated")
    mesh.agglomerate(s.face_weights)
    for k, lev in enumerate(orc.agglomeration(S)):
        assert np.array_equal(mesh.get_i32(7, k), lev[0]), f"restrictAddressing of level {k}"
    mat.set(s.diag, s.upper_coeffs)
    ctl = capi.controls("GAMG", smoother="GaussSeidel", tolerance=0.0, relTol=0.0, maxIter=3, recordHistory=1)
    psi, perf = mat.solve(ctl, s.source)
    xo, po = orc.solve(S, "GAMG", orc.controls("GaussSeidel", tolerance=0.0, relTol=0.0, maxIter=3), s.source)
    assert perf.nIterations == po["nIterations"] == 3
    h = capi.history(perf)
    assert np.all(np.abs(h - po["history"]) <= 1e-9 * po["initialResidual"])
    assert max_rel_diff(psi, xo) <= 1e-9
