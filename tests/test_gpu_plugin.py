"""Drop-in test: the UNMODIFIED reference (oracle/_ref/ref_harness -> lduMatrix::solver::New) loads
libB200LinearSolvers.so through its own `libs` mechanism (dlLibraryTable) and selects B200PCG / B200PBiCGStab /
B200GAMG / B200smoothSolver by name from its run-time selection tables -- exactly what a case does with
`libs ("libB200LinearSolvers.so");` in controlDict and `solver B200PCG;` in fvSolution.  Results are compared with
the reference's own solvers on the same matrix (the golden fixtures)."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from _util import FIXTURES, ldu_io, load_fixture, max_rel_diff, solve_keys

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
HARNESS = ROOT / "oracle/_ref/ref_harness"
PLUGIN = ROOT / "openfoam-dev_b200/libB200LinearSolvers.so"


@pytest.mark.parametrize("name", ["cavity_20x20x1", "block_16x16x16_rand", "convdiff_9x8x7",
                                  "cyclic_xz_9x8x7_rand", "cyclic_y_convdiff_10x9x6"])
def test_reference_selects_b200_solvers_by_name(name, tmp_path):
    if not HARNESS.exists() or not PLUGIN.exists():
        pytest.skip("oracle/_ref or the plugin was not built (needs /root/reference at build time)")
    inp, ref = load_fixture(name)
    e = {k: v for k, v in inp.items() if not k.startswith(("solve.", "smooth.", "agglomerate", "x"))}
    e["libs"] = f'"{PLUGIN}"'
    texts = []
    for i, text in solve_keys(inp):
        t = text.replace("solver PCG", "solver B200PCG").replace("solver PBiCGStab", "solver B200PBiCGStab")
        t = t.replace("solver GAMG", "solver B200GAMG").replace("solver smoothSolver", "solver B200smoothSolver")
        e[f"solve.{i}.dict"] = t
        texts.append(t)
    ldu_io.write(str(tmp_path / "in.b2ls"), e)
    env = dict(os.environ, WM_PROJECT_DIR=str(ROOT / "oracle/foam_env"), WM_PROJECT="OpenFOAM",
               WM_PROJECT_VERSION="dev")
    r = subprocess.run([str(HARNESS), str(tmp_path / "in.b2ls"), str(tmp_path / "out.b2ls"), str(tmp_path / "case")],
                       env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = ldu_io.read(str(tmp_path / "out.b2ls"))
    for i, t in enumerate(texts):
        got, want = out[f"solve.{i}.perf"], ref[f"solve.{i}.perf"]
        ctx = (name, t, got[:3], want[:3])
        assert ldu_io.as_str(out[f"solve.{i}.solverName"]) == ldu_io.as_str(ref[f"solve.{i}.solverName"]), ctx
        assert abs(got[2] - want[2]) <= 1, ctx
        assert abs(got[0] - want[0]) <= 1e-9 * abs(want[0]), ctx
        if got[2] == want[2]:
            assert abs(got[1] - want[1]) <= 1e-9 * abs(want[0]), ctx
            assert max_rel_diff(out[f"solve.{i}.psi"], ref[f"solve.{i}.psi"]) <= 1e-9, ctx
            assert got[3] == want[3], ctx


@pytest.mark.parametrize("name", ["block_16x16x16_rand", "convdiff_9x8x7", "cyclic_xz_9x8x7_rand"])
def test_reference_loops_with_b200_preconditioners_and_smoothers(name, tmp_path):
    """Operator-level drop-ins: the reference keeps ITS solver loop (PCG / PBiCGStab / GAMG /
    smoothSolver, on the CPU) and selects `preconditioner B200DIC|B200DILU` or `smoother B200<name>` from its
    preconditioner / smoother tables.  The GPU operators are bit-exact, so every result must EQUAL the all-reference
    run bit for bit."""
    if not HARNESS.exists() or not PLUGIN.exists():
        pytest.skip("oracle/_ref or the plugin was not built (needs /root/reference at build time)")
    inp, ref = load_fixture(name)
    e = {k: v for k, v in inp.items() if not k.startswith(("solve.", "agglomerate"))}
    e["libs"] = f'"{PLUGIN}"'
    i = 0
    while f"smooth.{i}.dict" in inp:
        e[f"smooth.{i}.dict"] = ldu_io.as_str(inp[f"smooth.{i}.dict"]).replace("smoother ", "smoother B200")
        i += 1
    n_smooth = i
    assert n_smooth > 0
    picked = []
    for i, text in solve_keys(inp):
        if "preconditioner {" in text or "nPreSweeps" in text:
            continue
        t = text.replace("preconditioner DIC", "preconditioner B200DIC").replace("preconditioner DILU",
                                                                                 "preconditioner B200DILU")
        t = t.replace("smoother ", "smoother B200")
        if t != text:
            e[f"solve.{len(picked)}.dict"] = t
            picked.append((i, t))
    assert len(picked) >= 5
    ldu_io.write(str(tmp_path / "in.b2ls"), e)
    env = dict(os.environ, WM_PROJECT_DIR=str(ROOT / "oracle/foam_env"), WM_PROJECT="OpenFOAM",
               WM_PROJECT_VERSION="dev")
    r = subprocess.run([str(HARNESS), str(tmp_path / "in.b2ls"), str(tmp_path / "out.b2ls"), str(tmp_path / "case")],
                       env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = ldu_io.read(str(tmp_path / "out.b2ls"))
    for i in range(n_smooth):
        assert np.array_equal(out[f"smooth.{i}.psi"], ref[f"smooth.{i}.psi"]), (name, e[f"smooth.{i}.dict"])
    for k, (i, t) in enumerate(picked):
        got, want = out[f"solve.{k}.perf"], ref[f"solve.{i}.perf"]
        assert np.array_equal(got[:5], want[:5]), (name, t, got[:5], want[:5])
        assert np.array_equal(out[f"solve.{k}.psi"], ref[f"solve.{i}.psi"]), (name, t)

