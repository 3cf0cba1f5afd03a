"""Domain decomposition of LDU systems into per-rank subdomains with processor interfaces.

Mirrors what decomposePar `simple` + the processor patches give the solver (reference
src/parallel/decompose/decompositionMethods/simple/simple.C:62-208,
src/parallel/parallel/domainDecomposition/domainDecompositionDecompose.C:403-529):
  * processor id = gx + px*(gy + py*gz); equal boxes
  * local cells and local internal faces keep their relative global order
  * a rank's processor patches are sorted by neighbour rank; faces inside a patch follow the global internal-face
    order, identically on both sides
  * on the owner side the patch coefficient is -upper[f] (row l gets upper[f]*psi[u] through
    result[faceCells] -= bouCoeffs*psi_nbr), on the neighbour side -lower[f]
    (processorFvPatchScalarField.C:133-136, coupledFvPatchField.C:179-205)
  * the matrix handed to the solver already holds the boundary contributions in its diagonal
    (fvMatrix::addBoundaryDiag, fvScalarMatrix.C:155-156), so the local diagonal is the global one.
"""
import numpy as np

from .cases import Interface, LduSystem, block_addressing, face_area_pair_weights, rhs


def simple_split(n_ranks):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n_ranks]


def box_cell_ranks(nx, ny, nz, split):
    """cell -> rank for a `simple` (px py pz) decomposition of an nx*ny*nz block into equal boxes."""
    px, py, pz = split
    c = np.arange(nx * ny * nz, dtype=np.int64)
    i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
    gx, gy, gz = i // (nx // px), j // (ny // py), k // (nz // pz)
    return (gx + px * (gy + py * gz)).astype(np.int32)


def decompose_system(sys_: LduSystem, cell_rank, n_ranks):
    """General decomposition of any LDU system.  Returns (list of per-rank LduSystem, list of local->global maps)."""
    cell_rank = np.asarray(cell_rank)
    lo, up = sys_.lower.astype(np.int64), sys_.upper.astype(np.int64)
    rl, ru = cell_rank[lo], cell_rank[up]
    lower_c = sys_.upper_coeffs if sys_.lower_coeffs is None else sys_.lower_coeffs
    out, maps = [], []
    for r in range(n_ranks):
        cells = np.nonzero(cell_rank == r)[0]
        g2l = -np.ones(sys_.n_cells, dtype=np.int64)
        g2l[cells] = np.arange(cells.size)
        inner = (rl == r) & (ru == r)
        ifaces = []
        nbrs = sorted(set(ru[(rl == r) & (ru != r)].tolist()) | set(rl[(ru == r) & (rl != r)].tolist()))
        for s in nbrs:
            own_side = (rl == r) & (ru == s)       # this rank holds the owner (lower) cell
            nei_side = (ru == r) & (rl == s)       # this rank holds the neighbour (upper) cell
            sel = np.nonzero(own_side | nei_side)[0]            # global face order
            fc = np.where(own_side[sel], g2l[lo[sel]], g2l[up[sel]])
            coeff = np.where(own_side[sel], sys_.upper_coeffs[sel], lower_c[sel])
            ifaces.append(Interface(neighb_rank=int(s), face_cells=fc.astype(np.int32), bou_coeffs=-coeff,
                                    int_coeffs=-coeff))
        # cyclic pairs of the undecomposed system stay cyclic on the rank that holds both cells of a face pair
        # (decomposePar turns pairs that straddle ranks into processorCyclic patches: not generated here)
        n_proc = len(ifaces)
        kept = {}
        for gi, itf in enumerate(sys_.interfaces):
            if itf.nbr_patch < 0:
                raise ValueError("decompose_system: the input system already has processor patches")
            other = sys_.interfaces[itf.nbr_patch]
            here, there = cell_rank[itf.face_cells] == r, cell_rank[other.face_cells] == r
            if np.any(here != there):
                raise ValueError("decompose_system: a cyclic face pair straddles two ranks")
            if np.any(here):
                kept[gi] = len(kept)
        for gi, k in kept.items():
            itf = sys_.interfaces[gi]
            sel = cell_rank[itf.face_cells] == r
            ifaces.append(Interface(neighb_rank=-1, face_cells=g2l[itf.face_cells[sel]].astype(np.int32),
                                    bou_coeffs=itf.bou_coeffs[sel].copy(), int_coeffs=itf.int_coeffs[sel].copy(),
                                    nbr_patch=n_proc + kept[itf.nbr_patch]))
        out.append(LduSystem(
            n_cells=int(cells.size), lower=g2l[lo[inner]].astype(np.int32), upper=g2l[up[inner]].astype(np.int32),
            diag=sys_.diag[cells].copy(), upper_coeffs=sys_.upper_coeffs[inner].copy(),
            lower_coeffs=None if sys_.lower_coeffs is None else sys_.lower_coeffs[inner].copy(),
            source=None if sys_.source is None else sys_.source[cells].copy(),
            face_weights=None if sys_.face_weights is None else sys_.face_weights[inner].copy(),
            interfaces=ifaces))
        maps.append(cells)
    return out, maps


def cavity_subdomain(nx, ny, nz, split, rank):
    """Rank `rank`'s subdomain of the uniform-coefficient cavity Laplacian on an nx*ny*nz block, built directly
    (no global matrix): identical to decompose_system(cavity_laplacian(nx, ny, nz), box_cell_ranks(...))[rank]."""
    px, py, pz = split
    lx, ly, lz = nx // px, ny // py, nz // pz
    gx, gy, gz = rank % px, (rank // px) % py, rank // (px * py)
    lower, upper, fdir = block_addressing(lx, ly, lz)
    n = lx * ly * lz
    up = np.ones(lower.size)
    diag = np.zeros(n)
    np.subtract.at(diag, lower, up)
    np.subtract.at(diag, upper, up)
    c = np.arange(n, dtype=np.int64)
    i, j, k = c % lx, (c // lx) % ly, c // (lx * ly)
    # global cell index of each local cell (for the right-hand side and the reference cell)
    gi, gj, gk = i + gx * lx, j + gy * ly, k + gz * lz
    gcell = gi + nx * (gj + ny * gk)

    def patch(mask, nbr_rank):
        cells = np.nonzero(mask)[0]          # ascending local index == ascending global face order
        np.subtract.at(diag, cells, 1.0)     # boundary contribution already in the diagonal
        return Interface(neighb_rank=nbr_rank, face_cells=cells.astype(np.int32),
                         bou_coeffs=-np.ones(cells.size), int_coeffs=-np.ones(cells.size))

    def rid(a, b, c_):
        return a + px * (b + py * c_)

    cand = []
    if gz > 0:
        cand.append((rid(gx, gy, gz - 1), k == 0))
    if gy > 0:
        cand.append((rid(gx, gy - 1, gz), j == 0))
    if gx > 0:
        cand.append((rid(gx - 1, gy, gz), i == 0))
    if gx < px - 1:
        cand.append((rid(gx + 1, gy, gz), i == lx - 1))
    if gy < py - 1:
        cand.append((rid(gx, gy + 1, gz), j == ly - 1))
    if gz < pz - 1:
        cand.append((rid(gx, gy, gz + 1), k == lz - 1))
    ifaces = [patch(mask, r) for r, mask in sorted(cand, key=lambda t: t[0])]
    if rank == 0:
        diag[0] += diag[0]                   # setReference(0, 0) on global cell 0
    src = np.sin(0.37 * gcell.astype(np.float64))
    return LduSystem(n_cells=n, lower=lower, upper=upper, diag=diag, upper_coeffs=up, source=src,
                     face_weights=face_area_pair_weights(nx, ny, nz, fdir), face_dir=fdir, interfaces=ifaces,
                     shape=(lx, ly, lz))


def as_cyclic_blocks(parts):
    """The decomposed system as ONE block-diagonal LDU system: rank r's cells follow rank r-1's, every rank keeps its
    own internal faces, and each processor-patch pair becomes a cyclic patch pair between the two blocks.

    Run through the serial reference this executes the reference's *decomposed* algorithm: interface terms applied
    after the face loop in patch order, block-local (= rank-local) DIC/DILU and Gauss-Seidel with lagged interface
    values, per-block pair agglomeration with the lower rank as the master of each coarse patch
    (cyclicGAMGInterface.C:85-157 has the same pair logic as processorGAMGInterface.C:75-147; the owner half is the
    one with the lower patch index = the lower rank here), block-local DIC on the coarsest level.  Only the global
    sums differ: one sequential pass over all cells instead of per-rank partial sums added in rank order.
    GAMG needs nCellsInCoarsestLevel = 10*len(parts) to reproduce the decomposed stop criterion
    (GAMGAgglomeration.C:205-230).  Returns (LduSystem, cell offsets)."""
    offs = np.concatenate([[0], np.cumsum([p.n_cells for p in parts])]).astype(np.int64)
    sym = parts[0].symmetric
    first = np.concatenate([[0], np.cumsum([len(p.interfaces) for p in parts])])
    ifaces = []
    for r, p in enumerate(parts):
        for itf in p.interfaces:
            s = itf.neighb_rank
            back = [k for k, o in enumerate(parts[s].interfaces) if o.neighb_rank == r]
            if len(back) != 1:
                raise ValueError("as_cyclic_blocks needs exactly one processor patch per rank pair")
            ifaces.append(Interface(neighb_rank=-1, face_cells=(itf.face_cells + offs[r]).astype(np.int32),
                                    bou_coeffs=itf.bou_coeffs.copy(), int_coeffs=itf.int_coeffs.copy(),
                                    nbr_patch=int(first[s] + back[0])))
    cat = np.concatenate
    return LduSystem(
        n_cells=int(offs[-1]),
        lower=cat([p.lower + offs[r] for r, p in enumerate(parts)]).astype(np.int32),
        upper=cat([p.upper + offs[r] for r, p in enumerate(parts)]).astype(np.int32),
        diag=cat([p.diag for p in parts]), upper_coeffs=cat([p.upper_coeffs for p in parts]),
        lower_coeffs=None if sym else cat([p.lower_coeffs for p in parts]),
        source=cat([p.source for p in parts]), face_weights=cat([p.face_weights for p in parts]),
        interfaces=ifaces), offs
