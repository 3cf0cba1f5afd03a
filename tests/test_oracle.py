"""Pins the C oracle (oracle/ldu_oracle.c) against outputs of the unmodified reference (tests/golden/*.b2ls).
The oracle restates the reference's sequential face loops and summation order, so everything -- including Krylov
histories and iteration counts -- must agree bit for bit (both are compiled without FMA contraction)."""
import sys
from pathlib import Path

import numpy as np
import pytest

from _util import FIXTURES, ldu_io, load_fixture, min_cells_of, parse_dict, smooth_keys, solve_keys, system_from_entries

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle"))
import ldu_oracle as orc  # noqa: E402


@pytest.fixture(scope="module", params=FIXTURES)
def fx(request):
    inp, ref = load_fixture(request.param)
    s = system_from_entries(inp)
    return request.param, inp, ref, s, orc.System(s, min_cells=min_cells_of(inp))


def test_operators(fx):
    name, inp, ref, s, S = fx
    assert np.array_equal(orc.amul(S, inp["x"]), ref["Amul"])
    assert np.array_equal(orc.residual(S, inp["x"], s.source), ref["residual"])
    assert np.array_equal(orc.sum_a(S), ref["sumA"])
    assert np.array_equal(orc.losort(S), ref["losort"])


def test_preconditioners(fx):
    name, inp, ref, s, S = fx
    kind = "DIC" if s.symmetric else "DILU"
    assert np.array_equal(orc.reciprocal_d(S), ref[f"{kind}_rD"])
    assert np.array_equal(orc.precondition(S, kind, inp["x"]), ref["precondition"])


def test_smoothers(fx):
    name, inp, ref, s, S = fx
    for i, kind, n_sweeps in smooth_keys(inp):
        assert np.array_equal(orc.smooth(S, kind, inp["x"], s.source, n_sweeps), ref[f"smooth.{i}.psi"]), (name, kind)


def test_agglomeration(fx):
    name, inp, ref, s, S = fx
    if "agg.nLevels" not in ref:
        pytest.skip("fixture without agglomeration dump")
    levels = orc.agglomeration(S)
    assert len(levels) == int(ref["agg.nLevels"][0])
    for k, (ra, fra, ff, cl, cu) in enumerate(levels):
        p = f"agg.{k}."
        assert np.array_equal(ra, ref[p + "restrictAddressing"])
        assert np.array_equal(fra, ref[p + "faceRestrictAddressing"])
        assert np.array_equal(ff, ref[p + "faceFlipMap"].astype(np.int32))
        assert np.array_equal(cl, ref[p + "coarseLower"]) and np.array_equal(cu, ref[p + "coarseUpper"])
    # coarse cyclic patches (cyclicGAMGInterface): faceCells + faceRestrictAddressing of every level
    for k, lev in enumerate(orc.interface_agglomeration(S)):
        for i, (fc, fra) in enumerate(lev):
            assert np.array_equal(fc, ref[f"agg.{k}.iface.{i}.faceCells"]), (name, k, i)
            assert np.array_equal(fra, ref[f"agg.{k}.iface.{i}.faceRestrictAddressing"]), (name, k, i)


def test_solvers_bit_exact(fx):
    name, inp, ref, s, S = fx
    for i, text in solve_keys(inp):
        d = parse_dict(text)
        kw = {k: (float(v) if k in ("tolerance", "relTol") else int(v)) for k, v in d.items()
              if k not in ("solver", "preconditioner", "smoother", "nCellsInCoarsestLevel")}
        pre = d.get("preconditioner") or d.get("smoother")
        if isinstance(pre, dict):
            sub, pre = pre, pre["preconditioner"]
            kw.update(precSmoother=sub.get("smoother", "GaussSeidel"), nVcycles=int(sub.get("nVcycles", 2)),
                      precTolerance=float(sub.get("tolerance", 1e-6)), precRelTol=float(sub.get("relTol", 0)))
        ctl = orc.controls(precond=pre, **kw)
        psi, perf = orc.solve(S, d["solver"], ctl, s.source, psi0=inp.get("psi0"))
        rperf = ref[f"solve.{i}.perf"]
        ctx = (name, text, perf["nIterations"], rperf[2], perf["finalResidual"], rperf[1])
        assert perf["nIterations"] == int(rperf[2]), ctx
        assert perf["initialResidual"] == rperf[0], ctx
        assert perf["finalResidual"] == rperf[1], ctx
        assert perf["converged"] == bool(rperf[3]), ctx
        assert np.array_equal(psi, ref[f"solve.{i}.psi"]), ctx
        hkey = f"solve.{i}.historyResiduals"
        if hkey in ref:
            n = min(len(perf["history"]), len(ref[hkey]), int(rperf[2]))
            assert np.array_equal(perf["history"][:n], ref[hkey][:n]), ctx


def test_emulated_ranks_reduce_to_the_reference_at_one_rank():
    """tests/_emulated_ranks.py (decomposed PCG built from the C oracle) equals the reference golden when there is a
    single rank, and converges to the global solution when the system is cut in two."""
    import _emulated_ranks as em
    from _util import cases, max_rel_diff
    from _pkg import load_pkg

    load_pkg()
    from b200ls import decompose

    inp, ref = load_fixture("block_16x16x16_rand")
    s = system_from_entries(inp)
    psi, perf = em.pcg([s], "DIC", tolerance=1e-10)
    r = ref["solve.1.perf"]
    assert perf["nIterations"] == int(r[2])
    assert abs(perf["finalResidual"] - r[1]) <= 1e-12 * r[0]
    assert max_rel_diff(psi[0], ref["solve.1.psi"]) <= 1e-12
    g = cases.cavity_laplacian(12, 10, 8, coeffs="random")
    parts, maps = decompose.decompose_system(g, decompose.box_cell_ranks(12, 10, 8, (2, 1, 1)), 2)
    psi2, perf2 = em.pcg(parts, "DIC", tolerance=1e-12)
    full = np.zeros(g.n_cells)
    for m, x in zip(maps, psi2):
        full[m] = x
    xg, pg = orc.solve(orc.System(g), "PCG", orc.controls("DIC", tolerance=1e-12), g.source)
    assert max_rel_diff(full, xg) <= 1e-10


@pytest.mark.parametrize("name", ["decomp2_sym", "decomp4_sym"])
def test_emulated_ranks_match_the_reference_run_as_cyclic_blocks(name):
    """The decomposed fixtures are the REFERENCE running its decomposed algorithm in serial (processor patches as
    cyclic pairs between diagonal blocks).  The rank-by-rank emulation used by the multi-GPU test must agree with it:
    same iteration count, residual history to ~1e-15 (only the grouping of the global sums differs)."""
    import _emulated_ranks as em
    from _pkg import load_pkg

    load_pkg()
    from b200ls import cases, decompose

    inp, ref = load_fixture(name)
    n_ranks = len(inp["rankOffsets"]) - 1
    split = decompose.simple_split(n_ranks)
    nx, ny, nz = 10 * split[0], 8 * split[1], 6 * split[2]
    glob = cases.cavity_laplacian(nx, ny, nz, coeffs="random")
    parts, _ = decompose.decompose_system(glob, decompose.box_cell_ranks(nx, ny, nz, split), n_ranks)
    blk, offs = decompose.as_cyclic_blocks(parts)
    assert np.array_equal(blk.lower, inp["lower"]) and np.array_equal(blk.diag, inp["diag"])   # same case as the fixture
    i = [k for k, t in solve_keys(inp) if "PCG" in t and "DIC" in t and "1e-10" in t][0]
    e_psi, e_perf = em.pcg(parts, "DIC", tolerance=1e-10)
    rperf = ref[f"solve.{i}.perf"]
    assert e_perf["nIterations"] == int(rperf[2])
    rh = ref[f"solve.{i}.historyResiduals"]
    n = min(len(rh), len(e_perf["history"]), int(rperf[2]))
    assert np.max(np.abs(e_perf["history"][:n] - rh[:n])) <= 1e-12 * rperf[0]
    assert np.max(np.abs(np.concatenate(e_psi) - ref[f"solve.{i}.psi"])) <= 1e-11 * np.max(np.abs(ref[f"solve.{i}.psi"]))
