"""Shared helpers of the parity tests."""
import re
from pathlib import Path

import numpy as np

from _pkg import load_pkg

load_pkg()
from b200ls import capi, cases, ldu_io  # noqa: E402

GOLDEN = Path(__file__).resolve().parent / "golden"
# icofoam_*.b2ls are application logs (tests/test_gpu_icofoam.py), not linear-system fixtures
FIXTURES = sorted(p.stem for p in GOLDEN.glob("*.b2ls") if not p.stem.startswith("icofoam_"))


def load_fixture(name):
    raw = ldu_io.read(str(GOLDEN / f"{name}.b2ls"))
    inp = {k[3:]: v for k, v in raw.items() if k.startswith("in.")}
    ref = {k[4:]: v for k, v in raw.items() if k.startswith("ref.")}
    return inp, ref


def system_from_entries(inp):
    return cases.LduSystem(
        n_cells=int(inp["nCells"][0]), lower=inp["lower"], upper=inp["upper"], diag=inp["diag"],
        upper_coeffs=inp["upperCoeffs"], lower_coeffs=inp.get("lowerCoeffs"), source=inp.get("source"),
        face_weights=inp.get("faceWeights"),
        interfaces=[
            cases.Interface(neighb_rank=-1, face_cells=inp[f"iface.{k}.faceCells"],
                            bou_coeffs=inp[f"iface.{k}.bouCoeffs"], int_coeffs=inp[f"iface.{k}.intCoeffs"],
                            nbr_patch=int(inp[f"iface.{k}.nbrPatch"][0]))
            for k in range(int(inp["nIfaces"][0]) if "nIfaces" in inp else 0)
        ],
    )


def min_cells_of(inp):
    """nCellsInCoarsestLevel of a fixture's GAMG dictionaries (10 = the reference default, GAMGAgglomeration.C:254)."""
    for k, v in inp.items():
        if k.endswith(".dict"):
            m = re.search(r"nCellsInCoarsestLevel\s+(\d+)", v if isinstance(v, str) else ldu_io.as_str(v))
            if m:
                return int(m.group(1))
    return 10


def parse_dict(text):
    """'solver PCG; preconditioner DIC; tolerance 1e-6;' -> {'solver': 'PCG', ...}.  A sub-dictionary
    'preconditioner { preconditioner GAMG; smoother GaussSeidel; ... }' becomes out['preconditioner'] = {...}."""
    out = {}
    sub = None
    if "{" in text:
        head, rest = text.split("{", 1)
        body, tail = rest.split("}", 1)
        key = head.split(";")[-1].split()[0]
        sub = (key, parse_dict(body))
        text = ";".join(head.split(";")[:-1]) + ";" + tail
    for item in text.split(";"):
        parts = item.split()
        if len(parts) >= 2:
            out[parts[0]] = parts[1]
    if sub:
        out[sub[0]] = sub[1]
    return out


_INT_KEYS = {"maxIter", "minIter", "nPreSweeps", "preSweepsLevelMultiplier", "maxPreSweeps", "nPostSweeps",
             "postSweepsLevelMultiplier", "maxPostSweeps", "nFinestSweeps", "nSweeps"}
_FLT_KEYS = {"tolerance", "relTol"}


def controls_from_dict(text, **extra):
    d = parse_dict(text)
    kw = {}
    for k, v in d.items():
        if k in _INT_KEYS:
            kw[k] = int(v)
        elif k in _FLT_KEYS:
            kw[k] = float(v)
    pre = d.get("preconditioner")
    if isinstance(pre, dict):
        # preconditioner GAMG: V-cycle controls come from the sub-dictionary (GAMGPreconditioner.C:47-62)
        sub = pre
        pre = sub["preconditioner"]
        kw["precSmoother"] = sub.get("smoother", "GaussSeidel")
        kw["nVcycles"] = int(sub.get("nVcycles", 2))
        kw["precTolerance"] = float(sub.get("tolerance", 1e-6))
        kw["precRelTol"] = float(sub.get("relTol", 0))
        for k, v in sub.items():
            if k in _INT_KEYS and k not in ("maxIter", "minIter", "nSweeps"):
                kw[k] = int(v)
    kw.update(extra)
    return capi.controls(solver=d["solver"], preconditioner=pre, smoother=d.get("smoother"), **kw)


def solve_keys(inp):
    i = 0
    while f"solve.{i}.dict" in inp:
        yield i, ldu_io.as_str(inp[f"solve.{i}.dict"])
        i += 1


def smooth_keys(inp):
    i = 0
    while f"smooth.{i}.dict" in inp:
        yield i, parse_dict(ldu_io.as_str(inp[f"smooth.{i}.dict"]))["smoother"], int(inp[f"smooth.{i}.nSweeps"][0])
        i += 1


def max_rel_diff(a, b):
    """max |a-b| / max|b| : the 'max relative difference' of north_star, scaled by the field magnitude."""
    scale = np.max(np.abs(b))
    if scale == 0:
        return float(np.max(np.abs(a - b)))
    return float(np.max(np.abs(a - b)) / scale)
