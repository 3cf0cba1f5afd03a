"""In-process emulation of the reference's DECOMPOSED PCG (PCG.C:65-193 on every rank, processor-patch updates as in
processorFvPatchScalarField.C:36-152, reductions as in FieldReductionFunctions.C:190-292 + reduce(sumOp)), built from
the single-rank C oracle (oracle/ldu_oracle.c) for all rank-local arithmetic: the preconditioner is rank-local, the
interface contributions are applied after the face loop in patch order, and rank sums are added in rank order (the
grouping MPI's reduce would use).  It is itself pinned against the reference running the same decomposed algorithm in
serial as cyclic blocks (tests/golden/decomp*_sym.b2ls, tests/test_oracle.py)."""
import sys
from pathlib import Path

import dataclasses

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle"))
import ldu_oracle as orc  # noqa: E402


def _nbr_values(parts, fields, r, itf):
    """patchInternalField of the neighbour, in this patch's face order (identical on both sides)."""
    other = parts[itf.neighb_rank]
    back = [i for i in other.interfaces if i.neighb_rank == r][0]
    return fields[itf.neighb_rank][back.face_cells]


def amul(parts, S, fields):
    out = []
    for r, p in enumerate(parts):
        y = orc.amul(S[r], fields[r])
        for itf in p.interfaces:
            # result[faceCells[i]] -= coeffs[i]*recv[i], sequential (np.subtract.at keeps order)
            np.subtract.at(y, itf.face_cells, itf.bou_coeffs * _nbr_values(parts, fields, r, itf))
        out.append(y)
    return out


def pcg(parts, kind="DIC", tolerance=1e-6, rel_tol=0.0, max_iter=1000):
    # rank-local operators only: the processor-patch terms are added by _amul below
    S = [orc.System(dataclasses.replace(p, interfaces=[])) for p in parts]
    n_tot = sum(p.n_cells for p in parts)
    psi = [np.zeros(p.n_cells) for p in parts]
    src = [p.source for p in parts]
    wA = amul(parts, S, psi)
    rA = [s - w for s, w in zip(src, wA)]
    # normFactor
    sumA = []
    for r, p in enumerate(parts):
        sa = orc.sum_a(S[r])
        for itf in p.interfaces:
            np.subtract.at(sa, itf.face_cells, itf.bou_coeffs)
        sumA.append(sa)
    xbar = sum(float(np.sum(x)) for x in psi) / n_tot
    nf = sum(float(np.sum(np.abs(w - xbar * sa) + np.abs(s - xbar * sa))) for w, sa, s in zip(wA, sumA, src)) + 1e-20
    res0 = sum(float(np.sum(np.abs(r))) for r in rA) / nf
    res, hist, it = res0, [], 0
    pA = [np.zeros_like(x) for x in psi]
    wArA = 1e20

    def conv(v):
        return v < tolerance or (rel_tol > 1e-20 and v < rel_tol * res0)

    if not conv(res):
        while True:
            wArA_old = wArA
            wA = [orc.precondition(S[r], kind, rA[r]) for r in range(len(parts))]
            wArA = sum(float(np.dot(w, r)) for w, r in zip(wA, rA))
            if it == 0:
                pA = [w.copy() for w in wA]
            else:
                beta = wArA / wArA_old
                pA = [w + beta * p for w, p in zip(wA, pA)]
            wA = amul(parts, S, pA)
            wApA = sum(float(np.dot(w, p)) for w, p in zip(wA, pA))
            alpha = wArA / wApA
            psi = [x + alpha * p for x, p in zip(psi, pA)]
            rA = [r - alpha * w for r, w in zip(rA, wA)]
            res = sum(float(np.sum(np.abs(r))) for r in rA) / nf
            hist.append(res)
            it += 1
            if not (it < max_iter and not conv(res)):
                break
    return psi, {"initialResidual": res0, "finalResidual": res, "nIterations": it, "history": np.array(hist)}
