"""ctypes wrapper of oracle/libldu_oracle.so (built from oracle/ldu_oracle.c).  TEST INFRASTRUCTURE ONLY: imported
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg -- never by the product package."""
import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libldu_oracle.so"
PRECONDS = {"none": 0, "diagonal": 1, "DIC": 2, "DILU": 3, "GaussSeidel": 4, "symGaussSeidel": 5,
            "DICGaussSeidel": 6, "DILUGaussSeidel": 7, "GAMG": 8}


class Iface(C.Structure):
    _fields_ = [("n", C.c_int32), ("faceCells", C.c_void_p), ("nbrCells", C.c_void_p), ("bou", C.c_void_p),
                ("inn", C.c_void_p)]


class Ldu(C.Structure):
    _fields_ = [("nCells", C.c_int32), ("nFaces", C.c_int32), ("l", C.c_void_p), ("u", C.c_void_p),
                ("diag", C.c_void_p), ("upper", C.c_void_p), ("lower", C.c_void_p), ("nIfaces", C.c_int32),
                ("ifaces", C.c_void_p)]


class Perf(C.Structure):
    _fields_ = [("initialResidual", C.c_double), ("finalResidual", C.c_double), ("normFactor", C.c_double),
                ("nIterations", C.c_int32), ("converged", C.c_int32), ("singular", C.c_int32),
                ("nHistory", C.c_int32), ("history", C.c_double * 4096)]


class Ctl(C.Structure):
    _fields_ = [("tolerance", C.c_double), ("relTol", C.c_double), ("maxIter", C.c_int32), ("minIter", C.c_int32),
                ("precond", C.c_int32), ("nSweeps", C.c_int32), ("nPreSweeps", C.c_int32),
                ("preSweepsLevelMultiplier", C.c_int32), ("maxPreSweeps", C.c_int32), ("nPostSweeps", C.c_int32),
                ("postSweepsLevelMultiplier", C.c_int32), ("maxPostSweeps", C.c_int32), ("nFinestSweeps", C.c_int32),
                ("scaleCorrection", C.c_int32), ("precSmoother", C.c_int32), ("nVcycles", C.c_int32),
                ("precTolerance", C.c_double), ("precRelTol", C.c_double), ("hierarchy", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} not built (python -c 'import __graft_entry__ as g; g.build()')")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.oracle_norm_factor.restype = C.c_double
        _lib.oracle_gamg_build.restype = C.c_void_p
        _lib.oracle_gamg_build.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _lib.oracle_gamg_build_coupled.restype = C.c_void_p
        _lib.oracle_gamg_build_coupled.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                   C.c_int, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_gamg_set_interface_coeffs.argtypes = [C.c_void_p] * 3
        _lib.oracle_gamg_iface_size.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _lib.oracle_gamg_iface_arrays.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib.oracle_gamg_n_levels.argtypes = [C.c_void_p]
        _lib.oracle_gamg_level_sizes.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib.oracle_gamg_level_arrays.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        _lib.oracle_gamg_solve.argtypes = [C.c_void_p] + [C.c_void_p] * 3 + [C.POINTER(Ctl), C.c_void_p, C.c_void_p,
                                                                             C.POINTER(Perf)]
        _lib.oracle_gamg_free.argtypes = [C.c_void_p]
        _lib.oracle_gamg_set_matrix.argtypes = [C.c_void_p] * 4
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class System:
    """Keeps the numpy arrays alive next to the C view."""

    def __init__(self, s, min_cells=10):
        self.min_cells = min_cells       # nCellsInCoarsestLevel of the GAMG hierarchy built by solve()/agglomeration()
        self.l = np.ascontiguousarray(s.lower, dtype=np.int32)
        self.u = np.ascontiguousarray(s.upper, dtype=np.int32)
        self.diag = np.ascontiguousarray(s.diag, dtype=np.float64)
        self.upper = np.ascontiguousarray(s.upper_coeffs, dtype=np.float64)
        self.lower = self.upper if s.lower_coeffs is None else np.ascontiguousarray(s.lower_coeffs, dtype=np.float64)
        self.symmetric = s.lower_coeffs is None
        self.n = int(s.n_cells)
        # cyclic (same-process) coupled patches; processor patches are emulated by tests/_emulated_ranks.py
        ifs = list(getattr(s, "interfaces", []) or [])
        if any(i.nbr_patch < 0 for i in ifs):
            raise ValueError("the C oracle only takes cyclic interfaces")
        self.n_ifaces = len(ifs)
        self.nbr_patch = np.array([i.nbr_patch for i in ifs], dtype=np.int32)
        self.if_sizes = np.array([i.face_cells.size for i in ifs], dtype=np.int32)
        self.if_cells = [np.ascontiguousarray(i.face_cells, dtype=np.int32) for i in ifs]
        self.if_bou = [np.ascontiguousarray(i.bou_coeffs, dtype=np.float64) for i in ifs]
        self.if_int = [np.ascontiguousarray(i.int_coeffs, dtype=np.float64) for i in ifs]
        self.if_views = (Iface * max(1, len(ifs)))()
        for k in range(len(ifs)):
            self.if_views[k] = Iface(self.if_cells[k].size, _p(self.if_cells[k]), _p(self.if_cells[ifs[k].nbr_patch]),
                                     _p(self.if_bou[k]), _p(self.if_int[k]))
        self.c = Ldu(self.n, self.l.size, _p(self.l), _p(self.u), _p(self.diag), _p(self.upper), _p(self.lower),
                     len(ifs), C.cast(self.if_views, C.c_void_p))
        self.face_weights = s.face_weights

    def _ptr_array(self, arrays):
        return (C.c_void_p * max(1, len(arrays)))(*[a.ctypes.data for a in arrays])

    def gamg_build(self, min_cells=None, forward_start=1):
        """hierarchy_t* with the cyclic patches agglomerated and the finest-level interface coefficients set."""
        min_cells = self.min_cells if min_cells is None else min_cells
        H = lib().oracle_gamg_build_coupled(self.n, self.l.size, _p(self.l), _p(self.u),
                                            _p(np.ascontiguousarray(self.face_weights)), min_cells, forward_start,
                                            self.n_ifaces, _p(self.if_sizes), self._ptr_array(self.if_cells),
                                            _p(self.nbr_patch))
        lib().oracle_gamg_set_interface_coeffs(H, self._ptr_array(self.if_bou), self._ptr_array(self.if_int))
        return H


def controls(precond="DIC", tolerance=1e-6, relTol=0.0, maxIter=1000, minIter=0, nSweeps=1, nPreSweeps=0,
             preSweepsLevelMultiplier=1, maxPreSweeps=4, nPostSweeps=2, postSweepsLevelMultiplier=1, maxPostSweeps=4,
             nFinestSweeps=2, scaleCorrection=-1, precSmoother="GaussSeidel", nVcycles=2, precTolerance=1e-6,
             precRelTol=0.0):
    return Ctl(tolerance, relTol, maxIter, minIter, PRECONDS[precond], nSweeps, nPreSweeps, preSweepsLevelMultiplier,
               maxPreSweeps, nPostSweeps, postSweepsLevelMultiplier, maxPostSweeps, nFinestSweeps, scaleCorrection,
               PRECONDS[precSmoother], nVcycles, precTolerance, precRelTol, None)


def amul(S, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty(S.n)
    lib().oracle_amul(C.byref(S.c), _p(x), _p(y))
    return y


def residual(S, x, b):
    x, b = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    y = np.empty(S.n)
    lib().oracle_residual(C.byref(S.c), _p(x), _p(b), _p(y))
    return y


def sum_a(S):
    y = np.empty(S.n)
    lib().oracle_sum_a(C.byref(S.c), _p(y))
    return y


def losort(S):
    out = np.empty(S.l.size, dtype=np.int32)
    lib().oracle_losort(C.byref(S.c), _p(out))
    return out


def reciprocal_d(S):
    y = np.empty(S.n)
    lib().oracle_calc_reciprocal_d(C.byref(S.c), _p(y))
    return y


def precondition(S, kind, rA):
    rA = np.ascontiguousarray(rA, dtype=np.float64)
    rD = reciprocal_d(S) if kind in ("DIC", "DILU") else 1.0 / S.diag
    w = np.empty(S.n)
    lib().oracle_precondition(C.byref(S.c), PRECONDS[kind], _p(rD), _p(rA), _p(w))
    return w


def smooth(S, kind, psi, source, n_sweeps):
    x = np.ascontiguousarray(psi, dtype=np.float64).copy()
    b = np.ascontiguousarray(source, dtype=np.float64)
    lib().oracle_smooth(C.byref(S.c), PRECONDS[kind], _p(x), _p(b), n_sweeps)
    return x


def _perf(p):
    return {"initialResidual": p.initialResidual, "finalResidual": p.finalResidual, "nIterations": p.nIterations,
            "converged": bool(p.converged), "singular": bool(p.singular), "normFactor": p.normFactor,
            "history": np.array(p.history[: p.nHistory])}


def solve(S, solver, ctl, source, psi0=None):
    b = np.ascontiguousarray(source, dtype=np.float64)
    x = np.zeros(S.n) if psi0 is None else np.ascontiguousarray(psi0, dtype=np.float64).copy()
    p = Perf()
    if solver == "GAMG":
        H = S.gamg_build()
        try:
            lib().oracle_gamg_solve(H, _p(S.diag), _p(S.upper), None if S.symmetric else _p(S.lower), C.byref(ctl),
                                    _p(x), _p(b), C.byref(p))
        finally:
            lib().oracle_gamg_free(H)
    else:
        fn = {"PCG": lib().oracle_pcg, "PBiCGStab": lib().oracle_pbicgstab,
              "smoothSolver": lib().oracle_smooth_solver}[solver]
        H = None
        if ctl.precond == PRECONDS["GAMG"]:     # preconditioner GAMG: build the hierarchy + coarse matrices first
            H = S.gamg_build()
            lib().oracle_gamg_set_matrix(H, _p(S.diag), _p(S.upper), None if S.symmetric else _p(S.lower))
            ctl.hierarchy = H
        try:
            fn(C.byref(S.c), C.byref(ctl), _p(x), _p(b), C.byref(p))
        finally:
            if H:
                lib().oracle_gamg_free(H)
                ctl.hierarchy = None
    return x, _perf(p)


def agglomeration(S, min_cells=None, forward_start=1):
    """Per level: (restrictAddressing, faceRestrictAddressing, faceFlipMap, coarseLower, coarseUpper)."""
    H = S.gamg_build(min_cells, forward_start)
    out = []
    try:
        n_levels = lib().oracle_gamg_n_levels(H)
        for k in range(n_levels - 1):
            nc, nf, ncc, nfc = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
            lib().oracle_gamg_level_sizes(H, k, C.byref(nc), C.byref(nf))
            lib().oracle_gamg_level_sizes(H, k + 1, C.byref(ncc), C.byref(nfc))
            ra = np.empty(nc.value, dtype=np.int32)
            fra = np.empty(nf.value, dtype=np.int32)
            ff = np.empty(nf.value, dtype=np.int32)
            cl = np.empty(nfc.value, dtype=np.int32)
            cu = np.empty(nfc.value, dtype=np.int32)
            lib().oracle_gamg_level_arrays(H, k, _p(ra), _p(fra), _p(ff), _p(cl), _p(cu))
            out.append((ra, fra, ff, cl, cu))
    finally:
        lib().oracle_gamg_free(H)
    return out


def interface_agglomeration(S, min_cells=None, forward_start=1):
    """Per level, per cyclic patch: (coarse faceCells, faceRestrictAddressing of the fine patch)."""
    H = S.gamg_build(min_cells, forward_start)
    out = []
    try:
        for k in range(lib().oracle_gamg_n_levels(H) - 1):
            lev = []
            for i in range(S.n_ifaces):
                fc = np.empty(lib().oracle_gamg_iface_size(H, k + 1, i), dtype=np.int32)
                fra = np.empty(lib().oracle_gamg_iface_size(H, k, i), dtype=np.int32)
                lib().oracle_gamg_iface_arrays(H, k, i, _p(fc), _p(fra))
                lev.append((fc, fra))
            out.append(lev)
    finally:
        lib().oracle_gamg_free(H)
    return out
