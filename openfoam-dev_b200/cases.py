"""Synthetic LDU systems for the configurations of BASELINE.json.

Meshes are blockMesh-equivalent single hex blocks: cell c = i + nx*(j + ny*k) (i fastest);
internal faces in upper-triangular order: for each cell ascending, faces to c+1, c+nx,
c+nx*ny when they exist (the ordering polyMeshFromShapeMesh produces for one block,
reference src/OpenFOAM/meshes/polyMesh/polyMeshFromShapeMesh.C:176-269).

All arrays are numpy: int32 addressing ("label", WM_LABEL_SIZE=32) and float64 coefficients.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


@dataclass
class Interface:
    """One coupled patch: a processor patch of a decomposed system (reference: processorFvPatch) when
    nbr_patch < 0, or one half of a cyclic pair on the same rank (reference: cyclicFvPatch; face i is coupled
    to face i of patch `nbr_patch`) when nbr_patch >= 0."""
    neighb_rank: int
    face_cells: np.ndarray          # int32 [nPatchFaces], local cell of each patch face
    bou_coeffs: np.ndarray          # float64, interfaceBouCoeffs
    int_coeffs: np.ndarray          # float64, interfaceIntCoeffs
    nbr_patch: int = -1


@dataclass
class LduSystem:
    n_cells: int
    lower: np.ndarray               # int32 [nFaces]  owner ("l") of each face
    upper: np.ndarray               # int32 [nFaces]  neighbour ("u") of each face
    diag: np.ndarray                # float64 [nCells]
    upper_coeffs: np.ndarray        # float64 [nFaces]
    lower_coeffs: Optional[np.ndarray] = None   # None => symmetric
    source: Optional[np.ndarray] = None
    face_weights: Optional[np.ndarray] = None   # faceAreaPair agglomeration weights
    face_dir: Optional[np.ndarray] = None       # 0/1/2 for block meshes
    interfaces: List[Interface] = field(default_factory=list)
    shape: tuple = ()

    @property
    def n_faces(self):
        return int(self.lower.size)

    @property
    def symmetric(self):
        return self.lower_coeffs is None


def block_addressing(nx, ny, nz):
    """lower/upper addressing + face direction for an nx*ny*nz block."""
    n = nx * ny * nz
    c = np.arange(n, dtype=np.int64)
    i = c % nx
    j = (c // nx) % ny
    k = c // (nx * ny)
    cand_u = np.stack([c + 1, c + nx, c + nx * ny], axis=1)
    ok = np.stack([i < nx - 1, j < ny - 1, k < nz - 1], axis=1)
    own = np.repeat(c[:, None], 3, axis=1)
    dirs = np.broadcast_to(np.arange(3, dtype=np.int8), (n, 3))
    lower = own[ok].astype(np.int32)
    upper = cand_u[ok].astype(np.int32)
    fdir = dirs[ok].astype(np.int8)
    return lower, upper, fdir


def face_area_pair_weights(nx, ny, nz, fdir):
    """mag(cmptMultiply(Sf/sqrt(magSf), (1, 1.01, 1.02))) on a unit-cube block
    (reference faceAreaPairGAMGAgglomeration.C:66-79)."""
    dx, dy, dz = 1.0 / nx, 1.0 / ny, 1.0 / nz
    area = np.array([dy * dz, dx * dz, dx * dy])
    scale = np.array([1.0, 1.01, 1.02])
    a = area / np.sqrt(area)
    w = np.sqrt((a * scale) ** 2)
    return w[fdir]


def negsum_diag(n_cells, lower, upper, upper_coeffs, lower_coeffs):
    """lduMatrix::negSumDiag (reference lduMatrixOperations.C): Diag[l] -= Lower, Diag[u] -= Upper,
    sequential face order (np.subtract.at keeps index order)."""
    diag = np.zeros(n_cells)
    lo = upper_coeffs if lower_coeffs is None else lower_coeffs
    np.subtract.at(diag, lower, lo)
    np.subtract.at(diag, upper, upper_coeffs)
    return diag


def rhs(n, kind="sin", seed=20261017):
    if kind == "sin":
        return np.sin(0.37 * np.arange(n, dtype=np.float64))
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.uniform(-1.0, 1.0, n)


def cavity_laplacian(nx, ny, nz, coeffs="uniform", seed=20261017, rhs_kind="sin"):
    """p-equation stand-in: fvm::laplacian on a uniform block, all-Neumann + setReference(0, 0)
    (reference gaussLaplacianScheme.C:63-64, fvMatrix.C:553-565)."""
    lower, upper, fdir = block_addressing(nx, ny, nz)
    nf = lower.size
    if coeffs == "uniform":
        up = np.ones(nf)
    else:
        rng = np.random.Generator(np.random.PCG64(seed))
        up = 1.0 + 0.5 * rng.random(nf)
    n = nx * ny * nz
    diag = negsum_diag(n, lower, upper, up, None)
    diag[0] += diag[0]
    return LduSystem(
        n_cells=n, lower=lower, upper=upper, diag=diag, upper_coeffs=up,
        source=rhs(n, rhs_kind, seed + 1),
        face_weights=face_area_pair_weights(nx, ny, nz, fdir), face_dir=fdir,
        shape=(nx, ny, nz),
    )


def convection_diffusion(nx, ny, nz=1, nu=0.01, dt_coeff=0.5, seed=20261017, rhs_kind="sin"):
    """Asymmetric U/k/epsilon stand-in: upwind convection + diffusion + ddt on a uniform block
    (reference gaussConvectionScheme.C:140-142: lower = -w*phi, upper = lower + phi, negSumDiag;
    laplacian as above with opposite sign)."""
    lower, upper, fdir = block_addressing(nx, ny, nz)
    n = nx * ny * nz
    h = np.array([1.0 / nx, 1.0 / ny, 1.0 / max(nz, 1)])
    # smooth solenoidal-ish flux: rotation about the block centre + seeded perturbation
    cl = lower.astype(np.int64)
    i = cl % nx
    j = (cl // nx) % ny
    xc = (i + 0.5) / nx
    yc = (j + 0.5) / ny
    rng = np.random.Generator(np.random.PCG64(seed))
    pert = 0.1 * (rng.random(lower.size) - 0.5)
    ux = -(yc - 0.5) + pert
    uy = (xc - 0.5) + pert
    uz = 0.3 * np.sin(6.0 * xc) + pert
    vel = np.stack([ux, uy, uz], axis=1)
    area = np.array([h[1] * h[2], h[0] * h[2], h[0] * h[1]])
    phi = vel[np.arange(lower.size), fdir] * area[fdir]
    diff = nu * area[fdir] / h[fdir]
    lo = -np.maximum(phi, 0.0) - diff
    up = np.minimum(phi, 0.0) - diff
    diag = negsum_diag(n, lower, upper, up, lo)
    vol = h[0] * h[1] * h[2]
    diag = diag + vol / (dt_coeff * h.min())
    return LduSystem(
        n_cells=n, lower=lower, upper=upper, diag=diag, upper_coeffs=up, lower_coeffs=lo,
        source=rhs(n, rhs_kind, seed + 1) * vol,
        face_weights=face_area_pair_weights(nx, ny, max(nz, 1), fdir), face_dir=fdir,
        shape=(nx, ny, nz),
    )


def _axis_planes(shape, axis):
    """Cells of the low and high boundary planes normal to `axis`, both in ascending cell order (so face i of
    one plane faces face i of the other)."""
    nx, ny, nz = shape
    c = np.arange(nx * ny * nz, dtype=np.int64)
    idx = [c % nx, (c // nx) % ny, c // (nx * ny)][axis]
    lo = c[idx == 0].astype(np.int32)
    hi = c[idx == shape[axis] - 1].astype(np.int32)
    return lo, hi


def add_cyclic(sys_, axis, seed=1):
    """Make the block periodic along `axis`: a cyclic patch pair (low plane, high plane) whose faces carry the
    same kind of coefficients as the internal faces normal to `axis` (the wrap-around face is owned by the
    high-plane cell).  Coupled-patch convention of the reference (fvMatrix.C:112-173 addBoundaryDiag,
    lduMatrixATmul.C:82 updateMatrixInterfaces): row c gets  -bouCoeffs * psi[neighbour cell],  the diagonal
    gets internalCoeffs; with negSumDiag semantics internalCoeffs = -(off-diagonal coefficient)."""
    lo_cells, hi_cells = _axis_planes(sys_.shape, axis)
    n = lo_cells.size
    rng = np.random.Generator(np.random.PCG64(seed + 17 * axis))
    base = np.abs(sys_.upper_coeffs[sys_.face_dir == axis]).mean() if np.any(sys_.face_dir == axis) else 1.0
    sign = -1.0 if sys_.upper_coeffs.mean() < 0 else 1.0
    # coefficient multiplying psi[high] in the low rows ("lower" of the wrap face) and psi[low] in the high rows
    c_hi_rows = sign * base * (0.75 + 0.5 * rng.random(n))
    c_lo_rows = c_hi_rows if sys_.symmetric else sign * base * (0.75 + 0.5 * rng.random(n))
    diag = sys_.diag.copy()
    # negSumDiag: Diag[l] -= Lower, Diag[u] -= Upper (l = high-plane owner, u = low-plane neighbour)
    np.subtract.at(diag, hi_cells, c_lo_rows)
    np.subtract.at(diag, lo_cells, c_hi_rows)
    first = len(sys_.interfaces)
    sys_.diag = diag
    sys_.interfaces = list(sys_.interfaces) + [
        Interface(neighb_rank=-1, face_cells=lo_cells, bou_coeffs=-c_lo_rows, int_coeffs=-c_hi_rows,
                  nbr_patch=first + 1),
        Interface(neighb_rank=-1, face_cells=hi_cells, bou_coeffs=-c_hi_rows, int_coeffs=-c_lo_rows,
                  nbr_patch=first),
    ]
    return sys_


def random_graph(n_cells, avg_degree=5, symmetric=True, seed=20261017, max_span=None):
    """Unstructured stand-in: a random upper-triangular-ordered LDU graph (every cell owns a few faces to random
    higher-numbered cells within `max_span`), random coefficients, strictly diagonally dominant.  Exercises rows
    with many neighbours, irregular wavefronts and irregular agglomeration."""
    rng = np.random.Generator(np.random.PCG64(seed))
    span = max_span or max(8, n_cells // 6)
    lower, upper = [], []
    for c in range(n_cells - 1):
        k = min(n_cells - 1 - c, rng.integers(1, 2 * avg_degree // 2 + 2))
        hi = min(n_cells, c + 1 + span)
        nb = np.unique(rng.integers(c + 1, hi, size=k))
        if c + 1 not in nb and rng.random() < 0.7:
            nb = np.unique(np.append(nb, c + 1))      # keep the graph connected-ish
        lower.extend([c] * nb.size)
        upper.extend(nb.tolist())
    lower = np.array(lower, dtype=np.int32)
    upper = np.array(upper, dtype=np.int32)
    nf = lower.size
    up = -(0.5 + rng.random(nf))
    lo = None if symmetric else -(0.5 + rng.random(nf))
    diag = np.zeros(n_cells)
    np.add.at(diag, lower, np.abs(up))
    np.add.at(diag, upper, np.abs(up if lo is None else lo))
    diag = diag * (1.0 + 0.05 * rng.random(n_cells)) + 0.01
    return LduSystem(n_cells=n_cells, lower=lower, upper=upper, diag=diag, upper_coeffs=up, lower_coeffs=lo,
                     source=rng.uniform(-1.0, 1.0, n_cells), face_weights=0.5 + rng.random(nf))


def renumbered(sys_: LduSystem, seed=20261017, window=None):
    """The same matrix under a random renumbering of the cells (within windows of `window` consecutive labels, or
    globally): an unstructured LDU system -- faces re-oriented so that lower < upper and re-sorted into
    upper-triangular order, coefficients following their faces (a flipped face swaps upper and lower coefficient).
    Used for parity tests at sizes no shipped mesh reaches."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = sys_.n_cells
    if window is None:
        perm = rng.permutation(n)
    else:
        perm = np.arange(n)
        for a in range(0, n, window):
            b = min(n, a + window)
            perm[a:b] = a + rng.permutation(b - a)
    new_of = np.empty(n, dtype=np.int64)
    new_of[perm] = np.arange(n)                     # old cell perm[i] gets label i
    lo, up = new_of[sys_.lower], new_of[sys_.upper]
    flip = lo > up
    l2, u2 = np.where(flip, up, lo), np.where(flip, lo, up)
    lower_c = sys_.upper_coeffs if sys_.lower_coeffs is None else sys_.lower_coeffs
    cu = np.where(flip, lower_c, sys_.upper_coeffs)
    cl = np.where(flip, sys_.upper_coeffs, lower_c)
    order = np.lexsort((u2, l2))
    out = LduSystem(
        n_cells=n, lower=l2[order].astype(np.int32), upper=u2[order].astype(np.int32), diag=sys_.diag[perm].copy(),
        upper_coeffs=cu[order].copy(), lower_coeffs=None if sys_.lower_coeffs is None else cl[order].copy(),
        source=None if sys_.source is None else sys_.source[perm].copy(),
        face_weights=None if sys_.face_weights is None else sys_.face_weights[order].copy(),
    )
    return out


def to_entries(sys_: LduSystem):
    """B2LS entries understood by oracle/ref_harness.C and the C oracle."""
    e = {
        "nCells": int(sys_.n_cells),
        "lower": sys_.lower.astype(np.int32),
        "upper": sys_.upper.astype(np.int32),
        "diag": sys_.diag.astype(np.float64),
        "upperCoeffs": sys_.upper_coeffs.astype(np.float64),
    }
    if sys_.lower_coeffs is not None:
        e["lowerCoeffs"] = sys_.lower_coeffs.astype(np.float64)
    if sys_.source is not None:
        e["source"] = sys_.source.astype(np.float64)
    if sys_.face_weights is not None:
        e["faceWeights"] = sys_.face_weights.astype(np.float64)
    if sys_.interfaces:
        if any(i.nbr_patch < 0 for i in sys_.interfaces):
            raise ValueError("only cyclic interfaces can be written to a single-process B2LS case")
        e["nIfaces"] = len(sys_.interfaces)
        for k, i in enumerate(sys_.interfaces):
            e[f"iface.{k}.faceCells"] = i.face_cells.astype(np.int32)
            e[f"iface.{k}.nbrPatch"] = int(i.nbr_patch)
            e[f"iface.{k}.bouCoeffs"] = i.bou_coeffs.astype(np.float64)
            e[f"iface.{k}.intCoeffs"] = i.int_coeffs.astype(np.float64)
    return e


# ---- systems on REAL meshes: geometry and addressing dumped by `ref_harness --polymesh` (the reference's own polyMesh
# reader and geometry engine); coefficients as finiteVolume builds them -------------------------------------------

def _polymesh_geometry(d):
    n_cells, n_int, _ = (int(v) for v in d["sizes"])
    Sf = d["faceAreas"].reshape(-1, 3)
    Cf = d["faceCentres"].reshape(-1, 3)
    C = d["cellCentres"].reshape(-1, 3)
    lower, upper = d["lower"].astype(np.int32), d["upper"].astype(np.int32)
    magSf = np.sqrt((Sf * Sf).sum(1))
    nf = Sf / magSf[:, None]
    # surfaceInterpolation::makeNonOrthDeltaCoeffs (finiteVolume/interpolation/surfaceInterpolation/
    # surfaceInterpolation/surfaceInterpolation.C): 1/max(nf & delta, 0.05*mag(delta))
    delta = C[upper] - C[lower]
    dc = 1.0 / np.maximum((nf[:n_int] * delta).sum(1), 0.05 * np.sqrt((delta * delta).sum(1)))
    # faceAreaPairGAMGAgglomeration.C:66-79
    w = np.sqrt((((Sf[:n_int] / np.sqrt(magSf[:n_int])[:, None]) * np.array([1.0, 1.01, 1.02])) ** 2).sum(1))
    patches = []
    for i in range(int(d["nPatches"][0])):
        name, typ = bytes(d[f"patch.{i}.nameType"].astype(np.uint8)).decode().split()
        start, size = (int(v) for v in d[f"patch.{i}.startSize"])
        patches.append((name, typ, start, size))
    return n_cells, n_int, lower, upper, Sf, Cf, C, magSf, nf, dc, w, patches


def _fixed_value_patches(patches):
    sel = [p for p in patches if p[1] == "patch" and p[0].lower().startswith("outlet")]
    return sel or [p for p in patches if p[1] == "patch"][:1]


def polymesh_laplacian(d, rhs_kind="sin", seed=20261017):
    """p-equation stand-in on a real mesh: fvm::laplacian with gamma = 1 (gaussLaplacianScheme.C:52-81:
    upper = deltaCoeffs*magSf, negSumDiag), zeroGradient everywhere except fixedValue 0 on the outlet patch(es), whose
    internalCoeffs = -magSf*deltaCoeffs go into the diagonal (fvMatrix::addBoundaryDiag, fvMatrix.C:112-131)."""
    n, n_int, lower, upper, Sf, Cf, C, magSf, nf, dc, w, patches = _polymesh_geometry(d)
    up = dc * magSf[:n_int]
    diag = negsum_diag(n, lower, upper, up, None)
    own = d["faceOwner"]
    for name, typ, start, size in _fixed_value_patches(patches):
        f = np.arange(start, start + size)
        db = Cf[f] - C[own[f]]
        dcb = 1.0 / np.maximum((nf[f] * db).sum(1), 0.05 * np.sqrt((db * db).sum(1)))
        np.add.at(diag, own[f], -(magSf[f] * dcb))
    return LduSystem(n_cells=n, lower=lower, upper=upper, diag=diag, upper_coeffs=up, source=rhs(n, rhs_kind, seed + 1),
                     face_weights=w)


def polymesh_convection_diffusion(d, nu=0.01, rhs_kind="sin", seed=20261017):
    """U/k/epsilon stand-in on a real mesh: upwind convection of a smooth velocity field + diffusion + an implicit
    time-derivative term (gaussConvectionScheme.C:140-142, gaussLaplacianScheme.C:52-81, EulerDdtScheme)."""
    n, n_int, lower, upper, Sf, Cf, C, magSf, nf, dc, w, patches = _polymesh_geometry(d)
    x, y = Cf[:n_int, 0], Cf[:n_int, 1]
    span = max(np.ptp(C[:, 0]), np.ptp(C[:, 1]), 1e-30)
    U = np.stack([1.0 + 0.3 * np.sin(3.0 * y / span), 0.2 * np.cos(2.0 * x / span), 0.1 * np.ones(n_int)], axis=1)
    phi = (U * Sf[:n_int]).sum(1)
    diff = nu * span * dc * magSf[:n_int]
    lo = -np.maximum(phi, 0.0) - diff
    up = np.minimum(phi, 0.0) - diff
    vol = d["cellVolumes"]
    diag = negsum_diag(n, lower, upper, up, lo) + vol / (0.05 * span)
    return LduSystem(n_cells=n, lower=lower, upper=upper, diag=diag, upper_coeffs=up, lower_coeffs=lo,
                     source=rhs(n, rhs_kind, seed + 1) * vol, face_weights=w)
