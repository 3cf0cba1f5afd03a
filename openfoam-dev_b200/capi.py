"""ctypes binding of include/b200ls.h (test / bench glue; the product is the C-ABI library itself)."""
import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libb200ls.so"

MAX_HISTORY = 4096

# enum b200ls_i32_which
LOSORT, OWNER_START, LOSORT_START = 0, 1, 2
FWD_LEVEL_OFFSETS, FWD_LEVEL_ROWS, BWD_LEVEL_OFFSETS, BWD_LEVEL_ROWS = 3, 4, 5, 6
RESTRICT_ADDRESSING, FACE_RESTRICT_ADDRESSING, FACE_FLIP_MAP = 7, 8, 9
LOWER_ADDR, UPPER_ADDR, LEVEL_SIZES = 10, 11, 12
# enum b200ls_solver / b200ls_precond
PCG, PBICGSTAB, GAMG, SMOOTH_SOLVER = 0, 1, 2, 3
NONE, DIAGONAL, DIC, DILU, GAUSS_SEIDEL, SYM_GAUSS_SEIDEL, DIC_GAUSS_SEIDEL, DILU_GAUSS_SEIDEL, GAMG_PRECOND = range(9)

DIAGONAL_SOLVER = 4
SOLVERS = {"PCG": PCG, "PBiCGStab": PBICGSTAB, "GAMG": GAMG, "smoothSolver": SMOOTH_SOLVER,
           "diagonal": DIAGONAL_SOLVER}
PRECONDS = {"none": NONE, "diagonal": DIAGONAL, "DIC": DIC, "DILU": DILU, "GaussSeidel": GAUSS_SEIDEL,
            "symGaussSeidel": SYM_GAUSS_SEIDEL, "DICGaussSeidel": DIC_GAUSS_SEIDEL,
            "DILUGaussSeidel": DILU_GAUSS_SEIDEL, "GAMG": GAMG_PRECOND}

EXPORTS = [
    "b200ls_init", "b200ls_set_host_comm", "b200ls_nccl_unique_id", "b200ls_finalize", "b200ls_last_error", "b200ls_device_available", "b200ls_device_count",
    "b200ls_mesh_create", "b200ls_mesh_free", "b200ls_mesh_get_i32", "b200ls_mesh_get_iface_i32", "b200ls_mesh_n_levels",
    "b200ls_agglomerate", "b200ls_agglomerate_from_maps", "b200ls_matrix_create", "b200ls_matrix_free", "b200ls_matrix_set",
    "b200ls_matrix_set_dev", "b200ls_matrix_set_if_changed",
    "b200ls_amul", "b200ls_residual", "b200ls_sum_a", "b200ls_precondition", "b200ls_reciprocal_d",
    "b200ls_smooth", "b200ls_controls_default", "b200ls_solve", "b200ls_solve_dev", "b200ls_time_kernel",
]


class Controls(C.Structure):
    _fields_ = [
        ("solver", C.c_int32), ("precond", C.c_int32),
        ("tolerance", C.c_double), ("relTol", C.c_double),
        ("maxIter", C.c_int32), ("minIter", C.c_int32),
        ("nPreSweeps", C.c_int32), ("preSweepsLevelMultiplier", C.c_int32), ("maxPreSweeps", C.c_int32),
        ("nPostSweeps", C.c_int32), ("postSweepsLevelMultiplier", C.c_int32), ("maxPostSweeps", C.c_int32),
        ("nFinestSweeps", C.c_int32), ("scaleCorrection", C.c_int32), ("nSweeps", C.c_int32),
        ("recordHistory", C.c_int32), ("precSmoother", C.c_int32), ("nVcycles", C.c_int32),
        ("precTolerance", C.c_double), ("precRelTol", C.c_double),
    ]


class Perf(C.Structure):
    _fields_ = [
        ("initialResidual", C.c_double), ("finalResidual", C.c_double),
        ("nIterations", C.c_int32), ("converged", C.c_int32), ("singular", C.c_int32), ("nHistory", C.c_int32),
        ("normFactor", C.c_double), ("solveMs", C.c_double), ("setupMs", C.c_double), ("h2dMs", C.c_double),
        ("kernelLaunches", C.c_int64),
        ("history", C.c_double * MAX_HISTORY),
    ]


class B200Error(RuntimeError):
    pass


_lib = None


def lib():
    """Load libb200ls.so.  There is no CPU fallback: a missing library is a hard error."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise B200Error(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(str(LIB_PATH))
    p, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.b200ls_last_error.restype = C.c_char_p
    L.b200ls_init.argtypes = [C.c_int, p, C.c_int, C.c_int]
    L.b200ls_nccl_unique_id.argtypes = [p]
    L.b200ls_mesh_create.restype = p
    L.b200ls_mesh_create.argtypes = [i32, i32, p, p, i32, p, p, p]
    L.b200ls_mesh_free.argtypes = [p]
    L.b200ls_mesh_get_i32.argtypes = [p, C.c_int, C.c_int, C.POINTER(p), C.POINTER(i64)]
    L.b200ls_mesh_n_levels.argtypes = [p]
    L.b200ls_mesh_get_iface_i32.argtypes = [p, C.c_int, C.c_int, C.c_int, C.POINTER(p), C.POINTER(i64)]
    L.b200ls_agglomerate.argtypes = [p, p, i32, i32, i32]
    L.b200ls_agglomerate_from_maps.argtypes = [p, i32, p, p]
    L.b200ls_matrix_create.restype = p
    L.b200ls_matrix_create.argtypes = [p]
    L.b200ls_matrix_free.argtypes = [p]
    L.b200ls_matrix_set.argtypes = [p, p, p, p, p, p]
    L.b200ls_matrix_set_dev.argtypes = [p, p, p, p, p, p]
    L.b200ls_matrix_set_if_changed.argtypes = [p, p, p, p, p, p, p]
    L.b200ls_amul.argtypes = [p, p, p]
    L.b200ls_residual.argtypes = [p, p, p, p]
    L.b200ls_sum_a.argtypes = [p, p]
    L.b200ls_precondition.argtypes = [p, C.c_int, p, p]
    L.b200ls_reciprocal_d.argtypes = [p, C.c_int, p]
    L.b200ls_smooth.argtypes = [p, C.c_int, p, p, i32]
    L.b200ls_controls_default.argtypes = [C.POINTER(Controls)]
    L.b200ls_solve.argtypes = [p, C.POINTER(Controls), p, p, C.POINTER(Perf)]
    L.b200ls_solve_dev.argtypes = [p, C.POINTER(Controls), p, p, C.POINTER(Perf)]
    L.b200ls_time_kernel.argtypes = [p, C.c_int, C.c_int, C.POINTER(dbl)]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise B200Error(lib().b200ls_last_error().decode())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


EXCHANGE_FN = C.CFUNCTYPE(None, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                          C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.POINTER(C.c_int32)))
SUM_FN = C.CFUNCTYPE(C.c_int64, C.c_int64)
_host_comm_refs = []


def set_host_comm(rank, n_ranks, exchange, total):
    """exchange(nbr: list[int], send: list[np.ndarray]) -> list[np.ndarray]; total(v: int) -> int (global sum)."""

    def _ex(n, nbr, sizes, send, recv):
        nb = [nbr[i] for i in range(n)]
        snd = [np.ctypeslib.as_array(send[i], shape=(sizes[i],)).copy() if sizes[i] else np.zeros(0, np.int32)
               for i in range(n)]
        got = exchange(nb, snd)
        for i in range(n):
            if sizes[i]:
                np.ctypeslib.as_array(recv[i], shape=(sizes[i],))[:] = got[i]

    ex_c, sum_c = EXCHANGE_FN(_ex), SUM_FN(lambda v: int(total(int(v))))
    _host_comm_refs[:] = [ex_c, sum_c]
    lib().b200ls_set_host_comm(rank, n_ranks, ex_c, sum_c)


def device_available():
    return bool(lib().b200ls_device_available())


def init(device=0, unique_id=None, rank=0, n_ranks=1):
    buf = None
    if unique_id is not None:
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
    _check(lib().b200ls_init(device, C.cast(buf, C.c_void_p) if buf is not None else None, rank, n_ranks))


def nccl_unique_id():
    buf = (C.c_char * 128)()
    _check(lib().b200ls_nccl_unique_id(C.cast(buf, C.c_void_p)))
    return bytes(buf)


def controls(solver="PCG", preconditioner=None, smoother=None, **kw):
    c = Controls()
    lib().b200ls_controls_default(C.byref(c))
    c.solver = SOLVERS[solver]
    which = smoother if smoother is not None else preconditioner
    if which is None:
        which = "GaussSeidel" if solver in ("GAMG", "smoothSolver") else "DIC"
    c.precond = PRECONDS[which]
    for k, v in kw.items():
        if not hasattr(c, k):
            raise KeyError(k)
        if k == "precSmoother" and isinstance(v, str):
            v = PRECONDS[v]
        setattr(c, k, v)
    return c


class Mesh:
    """lduAddressing (+ coupled interfaces) analysed once on the host; device upload is lazy."""

    def __init__(self, n_cells, lower, upper, interfaces=()):
        self.lower = _i32(lower)
        self.upper = _i32(upper)
        self.n_cells = int(n_cells)
        self.interfaces = list(interfaces)
        n_if = len(self.interfaces)
        sizes = _i32([len(i.face_cells) for i in self.interfaces])
        # cyclic halves are passed as B200LS_CYCLIC(partner) = -(1 + partner)
        nbr = _i32([-(1 + i.nbr_patch) if getattr(i, "nbr_patch", -1) >= 0 else i.neighb_rank
                    for i in self.interfaces])
        self._fc = [_i32(i.face_cells) for i in self.interfaces]
        fc_ptrs = (C.c_void_p * max(n_if, 1))(*[a.ctypes.data for a in self._fc])
        self.h = lib().b200ls_mesh_create(self.n_cells, self.lower.size, _ptr(self.lower), _ptr(self.upper),
                                          n_if, _ptr(sizes), C.cast(fc_ptrs, C.c_void_p), _ptr(nbr))
        if not self.h:
            raise B200Error(lib().b200ls_last_error().decode())

    def close(self):
        if self.h:
            lib().b200ls_mesh_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_i32(self, which, level=0):
        data = C.c_void_p()
        n = C.c_int64()
        _check(lib().b200ls_mesh_get_i32(self.h, which, level, C.byref(data), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=np.int32)
        arr = np.ctypeslib.as_array(C.cast(data, C.POINTER(C.c_int32)), shape=(n.value,))
        return arr.copy()

    def get_iface_i32(self, which, level, iface):
        data = C.c_void_p()
        n = C.c_int64()
        _check(lib().b200ls_mesh_get_iface_i32(self.h, which, level, iface, C.byref(data), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=np.int32)
        return np.ctypeslib.as_array(C.cast(data, C.POINTER(C.c_int32)), shape=(n.value,)).copy()

    @property
    def n_levels(self):
        return lib().b200ls_mesh_n_levels(self.h)

    def agglomerate(self, face_weights, min_cells_per_processor=10, merge_levels=1, forward_start=1):
        w = _f64(face_weights)
        n = lib().b200ls_agglomerate(self.h, _ptr(w), min_cells_per_processor, merge_levels, forward_start)
        if n < 0:
            raise B200Error(lib().b200ls_last_error().decode())
        return n


    def agglomerate_from_maps(self, restrict_maps):
        maps = [_i32(m) for m in restrict_maps]
        n_coarse = _i32([int(m.max()) + 1 for m in maps])
        ptrs = (C.c_void_p * max(len(maps), 1))(*[m.ctypes.data for m in maps])
        n = lib().b200ls_agglomerate_from_maps(self.h, len(maps), C.cast(ptrs, C.c_void_p), _ptr(n_coarse))
        if n < 0:
            raise B200Error(lib().b200ls_last_error().decode())
        return n


class Matrix:
    """lduMatrix coefficients resident in HBM."""

    def __init__(self, mesh: Mesh):
        self.mesh = mesh
        self.h = lib().b200ls_matrix_create(mesh.h)
        if not self.h:
            raise B200Error(lib().b200ls_last_error().decode())

    def close(self):
        if self.h:
            lib().b200ls_matrix_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, diag, upper, lower=None, bou_coeffs=(), int_coeffs=()):
        d, u = _f64(diag), _f64(upper)
        lo = _f64(lower) if lower is not None else None
        bou = [_f64(b) for b in bou_coeffs]
        inn = [_f64(b) for b in int_coeffs]
        n_if = len(bou)
        bp = (C.c_void_p * max(n_if, 1))(*[a.ctypes.data for a in bou])
        ip = (C.c_void_p * max(n_if, 1))(*[a.ctypes.data for a in inn])
        _check(lib().b200ls_matrix_set(self.h, _ptr(d), _ptr(u), _ptr(lo), C.cast(bp, C.c_void_p),
                                       C.cast(ip, C.c_void_p)))

    def set_if_changed(self, diag, upper, lower=None, bou_coeffs=(), int_coeffs=()):
        """b200ls_matrix_set_if_changed: returns True when the coefficients were uploaded, False when the matrix
        already held exactly these."""
        d, u = _f64(diag), _f64(upper)
        lo = _f64(lower) if lower is not None else None
        bou = [_f64(b) for b in bou_coeffs]
        inn = [_f64(b) for b in int_coeffs]
        n_if = len(bou)
        bp = (C.c_void_p * max(n_if, 1))(*[a.ctypes.data for a in bou])
        ip = (C.c_void_p * max(n_if, 1))(*[a.ctypes.data for a in inn])
        changed = C.c_int32(-1)
        _check(lib().b200ls_matrix_set_if_changed(self.h, _ptr(d), _ptr(u), _ptr(lo), C.cast(bp, C.c_void_p),
                                                  C.cast(ip, C.c_void_p), C.byref(changed)))
        return bool(changed.value)

    def set_dev(self, diag_ptr, upper_ptr, lower_ptr=None, bou_ptrs=(), int_ptrs=()):
        """b200ls_matrix_set_dev: raw device addresses (e.g. torch tensors' data_ptr()) in reference order."""
        n_if = len(bou_ptrs)
        bp = (C.c_void_p * max(n_if, 1))(*bou_ptrs)
        ip = (C.c_void_p * max(n_if, 1))(*int_ptrs)
        _check(lib().b200ls_matrix_set_dev(self.h, C.c_void_p(diag_ptr), C.c_void_p(upper_ptr),
                                           C.c_void_p(lower_ptr) if lower_ptr else None,
                                           C.cast(bp, C.c_void_p), C.cast(ip, C.c_void_p)))

    def amul(self, psi):
        x = _f64(psi)
        out = np.empty_like(x)
        _check(lib().b200ls_amul(self.h, _ptr(x), _ptr(out)))
        return out

    def residual(self, psi, source):
        x, b = _f64(psi), _f64(source)
        out = np.empty_like(x)
        _check(lib().b200ls_residual(self.h, _ptr(x), _ptr(b), _ptr(out)))
        return out

    def sum_a(self):
        out = np.empty(self.mesh.n_cells)
        _check(lib().b200ls_sum_a(self.h, _ptr(out)))
        return out

    def precondition(self, kind, rA):
        x = _f64(rA)
        out = np.empty_like(x)
        _check(lib().b200ls_precondition(self.h, PRECONDS[kind], _ptr(x), _ptr(out)))
        return out

    def reciprocal_d(self, kind):
        out = np.empty(self.mesh.n_cells)
        _check(lib().b200ls_reciprocal_d(self.h, PRECONDS[kind], _ptr(out)))
        return out

    def smooth(self, kind, psi, source, n_sweeps):
        x = _f64(psi).copy()
        b = _f64(source)
        _check(lib().b200ls_smooth(self.h, PRECONDS[kind], _ptr(x), _ptr(b), n_sweeps))
        return x

    def solve(self, ctl: Controls, source, psi0=None):
        b = _f64(source)
        x = np.zeros_like(b) if psi0 is None else _f64(psi0).copy()
        perf = Perf()
        _check(lib().b200ls_solve(self.h, C.byref(ctl), _ptr(x), _ptr(b), C.byref(perf)))
        return x, perf

    def solve_dev(self, ctl: Controls, psi_dev_ptr, source_dev_ptr):
        perf = Perf()
        _check(lib().b200ls_solve_dev(self.h, C.byref(ctl), C.c_void_p(psi_dev_ptr), C.c_void_p(source_dev_ptr),
                                      C.byref(perf)))
        return perf

    def time_kernel(self, which, reps=20):
        ms = C.c_double()
        _check(lib().b200ls_time_kernel(self.h, which, reps, C.byref(ms)))
        return ms.value


def from_system(sys_):
    """Mesh + Matrix from a cases.LduSystem (matrix coefficients uploaded)."""
    mesh = Mesh(sys_.n_cells, sys_.lower, sys_.upper, sys_.interfaces)
    mat = Matrix(mesh)
    mat.set(sys_.diag, sys_.upper_coeffs, sys_.lower_coeffs,
            [i.bou_coeffs for i in sys_.interfaces], [i.int_coeffs for i in sys_.interfaces])
    return mesh, mat


def history(perf: Perf):
    return np.array(perf.history[: perf.nHistory])
