// Host-side mesh analysis (integer data, built once per mesh).  See mesh.hpp.
//
// Reference behaviour restated here (paths under /root/reference/src/OpenFOAM/matrices/lduMatrix):
//   lduAddressing/lduAddressing.C:32-170          losort / ownerStart / losortStart
//   solvers/GAMG/GAMGAgglomerations/pairGAMGAgglomeration/pairGAMGAgglomerate.C:31-301
//   solvers/GAMG/GAMGAgglomerations/GAMGAgglomeration/GAMGAgglomerateLduAddressing.C:32-353
//   solvers/GAMG/GAMGAgglomerations/GAMGAgglomeration/GAMGAgglomeration.C:205-230
//   solvers/GAMG/GAMGAgglomerations/GAMGAgglomeration/GAMGAgglomerationTemplates.C:137-167
#include "mesh.hpp"

#include <algorithm>
#include <array>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <stdexcept>
#include <unordered_map>

namespace b200ls {

namespace {

// Stable bucket sort of items 0..n-1 by key[i] in [0, nKeys): offsets[nKeys+1], order[n]
void bucketSort(const std::vector<int32_t>& key, int32_t nKeys, std::vector<int32_t>& offsets,
                std::vector<int32_t>& order) {
    const size_t n = key.size();
    offsets.assign(size_t(nKeys) + 1, 0);
    for (size_t i = 0; i < n; i++) offsets[size_t(key[i]) + 1]++;
    for (int32_t k = 0; k < nKeys; k++) offsets[k + 1] += offsets[k];
    order.resize(n);
    std::vector<int32_t> cursor(offsets.begin(), offsets.end() - 1);
    for (size_t i = 0; i < n; i++) order[cursor[key[i]]++] = int32_t(i);
}

void makeTasks(const std::vector<int32_t>& offsets, std::vector<SweepTask>& tasks) {
    tasks.clear();
    for (size_t k = 0; k + 1 < offsets.size(); k++) {
        for (int32_t s = offsets[k]; s < offsets[k + 1]; s += 32) {
            tasks.push_back({s, std::min<int32_t>(32, offsets[k + 1] - s)});
        }
    }
}

}  // namespace

void buildLevel(LevelHost& L, int32_t nCells, int32_t nFaces, const int32_t* lower, const int32_t* upper,
                std::vector<HostInterface> interfaces, bool allowPencil) {
    if (nCells < 0 || nFaces < 0) throw std::runtime_error("negative mesh size");
    L.nCells = nCells;
    L.nFaces = nFaces;
    L.lower.assign(lower, lower + nFaces);
    L.upper.assign(upper, upper + nFaces);
    for (int32_t f = 0; f < nFaces; f++) {
        const int32_t l = lower[f], u = upper[f];
        if (l < 0 || u >= nCells || l >= u) {
            throw std::runtime_error("face " + std::to_string(f) + ": need 0 <= lower < upper < nCells");
        }
        if (f > 0 && lower[f - 1] > l) {
            throw std::runtime_error("faces are not in upper-triangular (owner-sorted) order at face " +
                                     std::to_string(f));
        }
    }

    // losort: faces grouped by upper cell, ascending face index inside a group
    std::vector<int32_t> nbrStart;   // proper CSR start of each losort group
    bucketSort(L.upper, nCells, nbrStart, L.losort);

    // ownerStart: number of faces whose owner is below the cell (faces are owner-sorted)
    L.ownerStart.assign(size_t(nCells) + 1, 0);
    for (int32_t f = 0; f < nFaces; f++) L.ownerStart[size_t(lower[f]) + 1]++;
    for (int32_t c = 0; c < nCells; c++) L.ownerStart[c + 1] += L.ownerStart[c];

    // losortStart as the reference leaves it: entries above the largest neighbour label are never
    // written and stay 0 (lduAddressing.C:139-169), the last entry is nFaces
    L.losortStart = nbrStart;
    {
        const int32_t maxNbr = nFaces ? L.upper[L.losort[nFaces - 1]] : -1;
        for (int32_t c = std::max(maxNbr + 1, 0); c < nCells; c++) L.losortStart[c] = 0;
        if (nFaces == 0) std::fill(L.losortStart.begin(), L.losortStart.end(), 0);
        L.losortStart[nCells] = nFaces;
    }

    // canonical wavefront levels
    std::vector<int32_t> lf(nCells, 0), lb(nCells, 0);
    int32_t nFwd = nCells ? 1 : 0, nBwd = nCells ? 1 : 0;
    for (int32_t f = 0; f < nFaces; f++) {
        const int32_t v = lf[lower[f]] + 1;
        if (v > lf[upper[f]]) lf[upper[f]] = v;
    }
    for (int32_t f = nFaces - 1; f >= 0; f--) {
        const int32_t v = lb[upper[f]] + 1;
        if (v > lb[lower[f]]) lb[lower[f]] = v;
    }
    for (int32_t c = 0; c < nCells; c++) {
        nFwd = std::max(nFwd, lf[c] + 1);
        nBwd = std::max(nBwd, lb[c] + 1);
    }
    L.maxFwdSpan = 1;
    for (int32_t f = 0; f < nFaces; f++) L.maxFwdSpan = std::max(L.maxFwdSpan, lf[upper[f]] - lf[lower[f]]);
    bucketSort(lf, nFwd, L.fwdOffsets, L.fwdRows);
    bucketSort(lb, nBwd, L.bwdOffsets, L.bwdRows);

    // native layout: structured blocks are laid out tile-major for the pencil sweeps, everything else in
    // forward-wavefront-major order (position = rank in that order)
    L.pencil = PencilPlan();
    L.fwdPos.clear();
    {
        const char* on = getenv("B200LS_PENCIL");
        const char* mc = getenv("B200LS_PENCIL_MIN_CELLS");
        const int32_t minCells = mc ? int32_t(atoi(mc)) : 16384;
        if (allowPencil && !(on && on[0] == '0') && nCells >= minCells) buildPencilPlan(L, L.pencil);
    }
    if (L.pencil.valid) {
        const PencilPlan& P = L.pencil;
        L.perm.resize(nCells);
        for (const PencilTile& t : P.tiles)
            for (int32_t i = 0; i < P.nx; i++)
                for (int32_t kk = 0; kk < t.wk; kk++)
                    for (int32_t jj = 0; jj < t.wj; jj++)
                        L.perm[size_t(t.base) + size_t(i) * t.w + jj + t.wj * kk] =
                            int32_t(i + int64_t(P.nx) * ((t.j0 + jj) + int64_t(P.ny) * (t.k0 + kk)));
    } else {
        L.perm = L.fwdRows;
    }
    L.ipos.resize(nCells);
    for (int32_t p = 0; p < nCells; p++) L.ipos[L.perm[p]] = p;
    if (L.pencil.valid) {
        // the wavefront kernels remain usable on this layout through the processing-order -> position map
        L.fwdPos.resize(nCells);
        for (int32_t k = 0; k < nCells; k++) L.fwdPos[k] = L.ipos[L.fwdRows[k]];
    }

    L.Lptr.assign(size_t(nCells) + 1, 0);
    L.Uptr.assign(size_t(nCells) + 1, 0);
    for (int32_t p = 0; p < nCells; p++) {
        const int32_t c = L.perm[p];
        L.Lptr[p + 1] = L.Lptr[p] + (nbrStart[c + 1] - nbrStart[c]);
        L.Uptr[p + 1] = L.Uptr[p] + (L.ownerStart[c + 1] - L.ownerStart[c]);
    }
    L.Lcol.resize(nFaces);
    L.Lface.resize(nFaces);
    L.Lidx.resize(nFaces);
    L.Ucol.resize(nFaces);
    L.Uface.resize(nFaces);
    L.Uidx.resize(nFaces);
    for (int32_t p = 0; p < nCells; p++) {
        const int32_t c = L.perm[p];
        int32_t j = L.Lptr[p];
        for (int32_t k = nbrStart[c]; k < nbrStart[c + 1]; k++, j++) {
            const int32_t f = L.losort[k];
            L.Lcol[j] = L.ipos[lower[f]];
            L.Lface[j] = f;
            L.Lidx[f] = j;
        }
        j = L.Uptr[p];
        for (int32_t f = L.ownerStart[c]; f < L.ownerStart[c + 1]; f++, j++) {
            L.Ucol[j] = L.ipos[upper[f]];
            L.Uface[j] = f;
            L.Uidx[f] = j;
        }
    }

    makeTasks(L.fwdOffsets, L.fwdTasks);

    // backward processing order: by backward level, ascending position inside a level
    {
        std::vector<int32_t> lbByPos(nCells);
        for (int32_t p = 0; p < nCells; p++) lbByPos[p] = lb[L.perm[p]];
        std::vector<int32_t> offs;
        bucketSort(lbByPos, nBwd, offs, L.bwdPos);
        makeTasks(offs, L.bwdTasks);
    }

    // interfaces and the rows they touch
    L.interfaces = std::move(interfaces);
    {
        std::vector<int32_t> cellOf, ifaceOf, faceOf;
        for (size_t i = 0; i < L.interfaces.size(); i++) {
            const auto& fc = L.interfaces[i].faceCells;
            for (size_t k = 0; k < fc.size(); k++) {
                if (fc[k] < 0 || fc[k] >= nCells) throw std::runtime_error("interface faceCells out of range");
                cellOf.push_back(fc[k]);
                ifaceOf.push_back(int32_t(i));
                faceOf.push_back(int32_t(k));
            }
        }
        std::vector<int32_t> order(cellOf.size());
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(),
                         [&](int32_t a, int32_t b) { return L.ipos[cellOf[a]] < L.ipos[cellOf[b]]; });
        L.bRowPos.clear();
        L.bRowPtr.assign(1, 0);
        L.bEntryIface.clear();
        L.bEntryFace.clear();
        for (size_t k = 0; k < order.size(); k++) {
            const int32_t e = order[k];
            const int32_t pos = L.ipos[cellOf[e]];
            if (L.bRowPos.empty() || L.bRowPos.back() != pos) {
                if (!L.bRowPos.empty()) L.bRowPtr.push_back(int32_t(L.bEntryIface.size()));
                L.bRowPos.push_back(pos);
            }
            L.bEntryIface.push_back(ifaceOf[e]);
            L.bEntryFace.push_back(faceOf[e]);
        }
        if (!L.bRowPos.empty()) L.bRowPtr.push_back(int32_t(L.bEntryIface.size()));
    }
}

// ------------------------------------------------------------------------------------------------------------
// structured blocks: tile-major layout for the pencil sweeps
// ------------------------------------------------------------------------------------------------------------

static bool detectBlock(const LevelHost& L, int32_t dims[3]) {
    const int64_t n = L.nCells;
    if (n < 2 || L.nFaces == 0) return false;
    // strides of the owner-side faces of cell 0: {1, nx, nx*ny} (those that exist)
    std::vector<int64_t> st;
    for (int32_t f = L.ownerStart[0]; f < L.ownerStart[1]; f++) st.push_back(L.upper[f]);
    if (st.empty() || st[0] != 1 || st.size() > 3) return false;
    int64_t nx, ny = 1, nz = 1;
    if (st.size() == 1) {
        nx = n;
    } else {
        nx = st[1];
        if (nx < 2 || n % nx) return false;
        if (st.size() == 2) {
            ny = n / nx;
        } else {
            if (st[2] % nx) return false;
            ny = st[2] / nx;
            if (ny < 2 || n % (nx * ny)) return false;
            nz = n / (nx * ny);
        }
    }
    if (nx * ny * nz != n) return false;
    const int64_t expectFaces = (nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1);
    if (expectFaces != L.nFaces) return false;
    int64_t f = 0;
    for (int64_t k = 0; k < nz; k++)
        for (int64_t j = 0; j < ny; j++)
            for (int64_t i = 0; i < nx; i++) {
                const int64_t c = i + nx * (j + ny * k);
                if (i < nx - 1) { if (L.lower[f] != c || L.upper[f] != c + 1) return false; f++; }
                if (j < ny - 1) { if (L.lower[f] != c || L.upper[f] != c + nx) return false; f++; }
                if (k < nz - 1) { if (L.lower[f] != c || L.upper[f] != c + nx * ny) return false; f++; }
            }
    dims[0] = int32_t(nx);
    dims[1] = int32_t(ny);
    dims[2] = int32_t(nz);
    return true;
}

void buildPencilPlan(const LevelHost& L, PencilPlan& plan) {
    plan = PencilPlan();
    int32_t d[3];
    if (!detectBlock(L, d)) return;
    const int32_t nx = d[0], ny = d[1], nz = d[2];
    // pencils of a full tile: 8 x 4 in 3-D, 32 x 1 in 2-D; thin blocks take all of j and fill up along k.  WK is kept
    // even whenever there are several tiles along k, and WJ is even whenever there are several along j: then every
    // tile base is even (16-byte aligned rows of doubles for the bulk copies) -- only the very last tile can be odd.
    int32_t WJ, WK;
    if (nz == 1) {
        WJ = std::min<int32_t>(ny, 32);
        WK = 1;
    } else {
        WJ = std::min<int32_t>(ny, 8);
        WK = std::min<int32_t>(nz, 32 / WJ);
        if (WK < nz && WK > 1 && (WK & 1)) WK--;
    }
    const int32_t nJ = (ny + WJ - 1) / WJ, nK = (nz + WK - 1) / WK;
    plan.nx = nx;
    plan.ny = ny;
    plan.nz = nz;
    plan.WJ = WJ;
    plan.WK = WK;
    plan.nJ = nJ;
    plan.nK = nK;
    plan.tiles.resize(size_t(nJ) * nK);
    int64_t base = 0;
    for (int32_t K = 0; K < nK; K++)
        for (int32_t J = 0; J < nJ; J++) {
            PencilTile& t = plan.tiles[size_t(J) + size_t(nJ) * K];
            t.j0 = J * WJ;
            t.k0 = K * WK;
            t.wj = std::min(WJ, ny - t.j0);
            t.wk = std::min(WK, nz - t.k0);
            t.w = t.wj * t.wk;
            if (base & 1) return;   // an odd-sized tile before the last one: keep the wavefront layout
            t.base = int32_t(base);
            base += int64_t(nx) * t.w;
            t.nbr[0] = J > 0 ? (J - 1) + nJ * K : -1;
            t.nbr[1] = K > 0 ? J + nJ * (K - 1) : -1;
            t.nbr[2] = J + 1 < nJ ? (J + 1) + nJ * K : -1;
            t.nbr[3] = K + 1 < nK ? J + nJ * (K + 1) : -1;
        }
    // launch order: tile wavefronts J + K (every tile only depends on (J-1, K) and (J, K-1))
    plan.fwdOrder.reserve(plan.tiles.size());
    for (int32_t wv = 0; wv <= nJ + nK - 2; wv++)
        for (int32_t K = 0; K < nK; K++) {
            const int32_t J = wv - K;
            if (J >= 0 && J < nJ) plan.fwdOrder.push_back(J + nJ * K);
        }
    plan.valid = true;
}

std::vector<int32_t> pairAgglomerate(int32_t& nCoarseCells, const LevelHost& fine,
                                     const std::vector<double>& w, bool& forward) {
    const int32_t n = fine.nCells;
    const int32_t nF = fine.nFaces;
    const int32_t* up = fine.upper.data();
    const int32_t* lo = fine.lower.data();

    // faces of a cell in the reference's visiting order: neighbour-side faces (ascending), then
    // owner-side faces (ascending)  (pairGAMGAgglomerate.C:139-181)
    std::vector<int32_t> nbrStart(size_t(n) + 1, 0);
    for (int32_t f = 0; f < nF; f++) nbrStart[size_t(up[f]) + 1]++;
    for (int32_t c = 0; c < n; c++) nbrStart[c + 1] += nbrStart[c];

    std::vector<int32_t> map(n, -1);
    nCoarseCells = 0;
    const double minusGreat = -1.0e+15;   // -great (reference primitives/Scalar/doubleScalar: great = 1e+15)

    auto visit = [&](int32_t c) { return forward ? c : n - 1 - c; };

    for (int32_t ci = 0; ci < n; ci++) {
        const int32_t c = visit(ci);
        if (map[c] >= 0) continue;

        int32_t match = -1;
        double best = minusGreat;
        for (int32_t k = nbrStart[c]; k < nbrStart[c + 1]; k++) {
            const int32_t f = fine.losort[k];
            if (map[up[f]] < 0 && map[lo[f]] < 0 && w[f] > best) {
                match = f;
                best = w[f];
            }
        }
        for (int32_t f = fine.ownerStart[c]; f < fine.ownerStart[c + 1]; f++) {
            if (map[up[f]] < 0 && map[lo[f]] < 0 && w[f] > best) {
                match = f;
                best = w[f];
            }
        }

        if (match >= 0) {
            map[up[match]] = nCoarseCells;
            map[lo[match]] = nCoarseCells;
            nCoarseCells++;
            continue;
        }

        // no free neighbour: join the cluster across the heaviest face
        int32_t join = -1;
        best = minusGreat;
        for (int32_t k = nbrStart[c]; k < nbrStart[c + 1]; k++) {
            const int32_t f = fine.losort[k];
            if (w[f] > best) {
                join = f;
                best = w[f];
            }
        }
        for (int32_t f = fine.ownerStart[c]; f < fine.ownerStart[c + 1]; f++) {
            if (w[f] > best) {
                join = f;
                best = w[f];
            }
        }
        if (join >= 0) map[c] = std::max(map[up[join]], map[lo[join]]);
    }

    // isolated cells become single-cell clusters, same visiting order
    for (int32_t ci = 0; ci < n; ci++) {
        const int32_t c = visit(ci);
        if (map[c] < 0) map[c] = nCoarseCells++;
    }

    if (!forward) {
        const int32_t last = nCoarseCells - 1;
        for (int32_t c = 0; c < n; c++) map[c] = last - map[c];
    }
    forward = !forward;
    return map;
}

namespace {

// GAMGAgglomeration::agglomerateLduAddressing for one level: fills fine.faceRestrictAddr/faceFlip and
// returns the coarse lower/upper addressing.
void agglomerateAddressing(LevelHost& fine, std::vector<int32_t>& cLower, std::vector<int32_t>& cUpper) {
    const int32_t nF = fine.nFaces;
    const int32_t nC = fine.nCoarseCells;
    const std::vector<int32_t>& rm = fine.restrictAddr;

    fine.faceRestrictAddr.resize(nF);
    fine.faceFlip.assign(nF, 0);

    // per coarse owner: singly linked list of its provisional coarse faces, in creation order
    std::vector<int32_t> head(nC, -1), tail(nC, -1);
    std::vector<int32_t> next, nei, own;
    next.reserve(nF / 2);
    nei.reserve(nF / 2);
    own.reserve(nF / 2);

    for (int32_t f = 0; f < nF; f++) {
        const int32_t ru = rm[fine.upper[f]];
        const int32_t rl = rm[fine.lower[f]];
        if (ru == rl) {
            fine.faceRestrictAddr[f] = -(ru + 1);
            continue;
        }
        const int32_t cOwn = std::min(ru, rl), cNei = std::max(ru, rl);
        int32_t found = -1;
        for (int32_t e = head[cOwn]; e >= 0; e = next[e]) {
            if (nei[e] == cNei) {
                found = e;
                break;
            }
        }
        if (found < 0) {
            found = int32_t(nei.size());
            nei.push_back(cNei);
            own.push_back(cOwn);
            next.push_back(-1);
            if (tail[cOwn] >= 0) next[tail[cOwn]] = found;
            else head[cOwn] = found;
            tail[cOwn] = found;
        }
        fine.faceRestrictAddr[f] = found;
    }

    // renumber into owner-major order, creation order inside an owner
    const int32_t nCF = int32_t(nei.size());
    fine.nCoarseFaces = nCF;
    std::vector<int32_t> renum(nCF);
    cLower.resize(nCF);
    cUpper.resize(nCF);
    int32_t k = 0;
    for (int32_t c = 0; c < nC; c++) {
        for (int32_t e = head[c]; e >= 0; e = next[e]) {
            cLower[k] = c;
            cUpper[k] = nei[e];
            renum[e] = k++;
        }
    }
    for (int32_t f = 0; f < nF; f++) {
        int32_t& cf = fine.faceRestrictAddr[f];
        if (cf >= 0) {
            cf = renum[cf];
            // flipped when the fine neighbour side maps to the coarse owner
            fine.faceFlip[f] = (rm[fine.upper[f]] == cLower[cf]) ? 1 : 0;
        }
    }
}

void buildMaps(LevelHost& fine, const LevelHost& coarse) {
    AgglomMaps& M = fine.maps;
    const int32_t nCf = fine.nCells, nCc = coarse.nCells;
    const int32_t nFf = fine.nFaces, nFc = coarse.nFaces;

    // restrict: coarse position -> fine positions in ascending fine cell order
    M.rPtr.assign(size_t(nCc) + 1, 0);
    for (int32_t c = 0; c < nCf; c++) M.rPtr[size_t(coarse.ipos[fine.restrictAddr[c]]) + 1]++;
    for (int32_t p = 0; p < nCc; p++) M.rPtr[p + 1] += M.rPtr[p];
    M.rFine.resize(nCf);
    {
        std::vector<int32_t> cur(M.rPtr.begin(), M.rPtr.end() - 1);
        for (int32_t c = 0; c < nCf; c++) M.rFine[cur[coarse.ipos[fine.restrictAddr[c]]]++] = fine.ipos[c];
    }
    M.pMap.resize(nCf);
    for (int32_t p = 0; p < nCf; p++) M.pMap[p] = coarse.ipos[fine.restrictAddr[fine.perm[p]]];

    // coarse off-diagonals: for each coarse U / L entry the fine value refs, ascending fine face
    M.uPtr.assign(size_t(nFc) + 1, 0);
    M.lPtr.assign(size_t(nFc) + 1, 0);
    M.dPtr.assign(size_t(nCc) + 1, 0);
    for (int32_t f = 0; f < nFf; f++) {
        const int32_t cf = fine.faceRestrictAddr[f];
        if (cf >= 0) {
            M.uPtr[size_t(coarse.Uidx[cf]) + 1]++;
            M.lPtr[size_t(coarse.Lidx[cf]) + 1]++;
        } else {
            M.dPtr[size_t(coarse.ipos[-1 - cf]) + 1]++;
        }
    }
    for (int32_t j = 0; j < nFc; j++) {
        M.uPtr[j + 1] += M.uPtr[j];
        M.lPtr[j + 1] += M.lPtr[j];
    }
    for (int32_t p = 0; p < nCc; p++) M.dPtr[p + 1] += M.dPtr[p];
    M.uSrc.resize(M.uPtr[nFc]);
    M.lSrc.resize(M.lPtr[nFc]);
    M.dU.resize(M.dPtr[nCc]);
    M.dL.resize(M.dPtr[nCc]);
    {
        std::vector<int32_t> cu(M.uPtr.begin(), M.uPtr.end() - 1), cl(M.lPtr.begin(), M.lPtr.end() - 1),
            cd(M.dPtr.begin(), M.dPtr.end() - 1);
        for (int32_t f = 0; f < nFf; f++) {
            const int32_t cf = fine.faceRestrictAddr[f];
            const int32_t uRef = fine.Uidx[f];            // fine upper[f]
            const int32_t lRef = nFf + fine.Lidx[f];      // fine lower[f]
            if (cf >= 0) {
                const bool flip = fine.faceFlip[f] != 0;
                M.uSrc[cu[coarse.Uidx[cf]]++] = flip ? lRef : uRef;
                M.lSrc[cl[coarse.Lidx[cf]]++] = flip ? uRef : lRef;
            } else {
                const int32_t p = coarse.ipos[-1 - cf];
                M.dU[cd[p]] = fine.Uidx[f];
                M.dL[cd[p]] = fine.Lidx[f];
                cd[p]++;
            }
        }
    }

    // interface coefficients: coarse patch face -> fine patch faces (ascending)
    const size_t nI = fine.interfaces.size();
    M.iPtr.assign(nI, {});
    M.iSrc.assign(nI, {});
    for (size_t i = 0; i < nI; i++) {
        const auto& pr = fine.patchFaceRestrictAddr[i];
        const int32_t nCoarsePatch = int32_t(coarse.interfaces[i].faceCells.size());
        std::vector<int32_t> key(pr.begin(), pr.end());
        bucketSort(key, nCoarsePatch, M.iPtr[i], M.iSrc[i]);
    }
}

}  // namespace

// Coarse processor interfaces (processorGAMGInterface ctor, processorGAMGInterface.C:53-140): coarse patch faces
// are the distinct (local coarse cell, neighbour coarse cell) pairs in first-seen order; the pair is ordered
// (master, slave) by rank so both sides number the coarse faces identically.
static void agglomerateInterfaces(LevelHost& fine, std::vector<HostInterface>& coarseIfaces, int myRank,
                                  const HostComm* comm) {
    const size_t nI = fine.interfaces.size();
    fine.patchFaceRestrictAddr.assign(nI, {});
    coarseIfaces.assign(nI, {});
    if (nI == 0) return;
    // interfaceInternalField(restrictMap) of every patch; the neighbour's values come from the partner patch for a
    // cyclic pair (cyclicGAMGInterface::internalFieldTransfer, cyclicGAMGInterface.C:180-197) and over the
    // communicator for processor patches (processorGAMGInterface.C:195-214)
    std::vector<std::vector<int32_t>> local(nI), remote(nI);
    for (size_t i = 0; i < nI; i++) {
        const auto& fc = fine.interfaces[i].faceCells;
        local[i].resize(fc.size());
        for (size_t k = 0; k < fc.size(); k++) local[i][k] = fine.restrictAddr[fc[k]];
    }
    std::vector<int32_t> procIdx, nbr;
    for (size_t i = 0; i < nI; i++) {
        const int32_t partner = fine.interfaces[i].partner;
        if (partner >= 0) {
            remote[i] = local[size_t(partner)];
        } else {
            procIdx.push_back(int32_t(i));
            nbr.push_back(fine.interfaces[i].neighbRank);
        }
    }
    if (!procIdx.empty()) {
        if (!comm || !comm->exchange) throw std::runtime_error("agglomerating processor interfaces needs a communicator");
        std::vector<std::vector<int32_t>> send(procIdx.size()), recv(procIdx.size());
        for (size_t j = 0; j < procIdx.size(); j++) {
            send[j] = local[size_t(procIdx[j])];
            recv[j].assign(send[j].size(), -1);
        }
        comm->exchange(nbr, send, recv);
        for (size_t j = 0; j < procIdx.size(); j++) remote[size_t(procIdx[j])] = std::move(recv[j]);
    }
    // coarse patch faces = distinct (master coarse cell, slave coarse cell) pairs in first-seen order
    // (processorGAMGInterface.C:75-147, cyclicGAMGInterface.C:85-157)
    for (size_t i = 0; i < nI; i++) {
        const HostInterface& fi = fine.interfaces[i];
        const bool master = fi.partner >= 0 ? int32_t(i) < fi.partner : myRank < fi.neighbRank;
        std::unordered_map<uint64_t, int32_t> seen;
        seen.reserve(local[i].size() * 2);
        HostInterface& ci = coarseIfaces[i];
        ci.neighbRank = fi.neighbRank;
        ci.partner = fi.partner;
        auto& pr = fine.patchFaceRestrictAddr[i];
        pr.resize(local[i].size());
        for (size_t k = 0; k < local[i].size(); k++) {
            const uint32_t a = uint32_t(master ? local[i][k] : remote[i][k]);
            const uint32_t b = uint32_t(master ? remote[i][k] : local[i][k]);
            const uint64_t key = (uint64_t(a) << 32) | b;
            auto it = seen.find(key);
            if (it == seen.end()) {
                const int32_t cf = int32_t(ci.faceCells.size());
                seen.emplace(key, cf);
                ci.faceCells.push_back(local[i][k]);
                pr[k] = cf;
            } else {
                pr[k] = it->second;
            }
        }
    }
    for (size_t i = 0; i < nI; i++) {
        const int32_t partner = coarseIfaces[i].partner;
        if (partner >= 0 && coarseIfaces[size_t(partner)].faceCells.size() != coarseIfaces[i].faceCells.size())
            throw std::runtime_error("cyclic halves agglomerated to different sizes");
    }
}

// Append the coarse level defined by `map` (fine cell -> coarse cell) below the current coarsest level:
// agglomerateLduAddressing + coarse interfaces + native layout + transfer maps.
static void appendCoarseLevel(HostMesh& mesh, std::vector<int32_t> map, int32_t nCoarse, const HostComm* comm) {
    const size_t k = mesh.levels.size() - 1;
    std::vector<HostInterface> coarseIfaces;
    {
        LevelHost& fine = mesh.levels[k];
        if (int32_t(map.size()) != fine.nCells) throw std::runtime_error("restrict map size != fine level size");
        for (int32_t v : map)
            if (v < 0 || v >= nCoarse) throw std::runtime_error("restrict map entry out of range");
        fine.hasCoarse = true;
        fine.nCoarseCells = nCoarse;
        fine.restrictAddr = std::move(map);
        agglomerateInterfaces(fine, coarseIfaces, mesh.rank, comm);
    }
    std::vector<int32_t> cLower, cUpper;
    agglomerateAddressing(mesh.levels[k], cLower, cUpper);
    mesh.levels.emplace_back();   // may move the storage: re-take references
    LevelHost& fine = mesh.levels[k];
    LevelHost& coarse = mesh.levels[k + 1];
    buildLevel(coarse, nCoarse, int32_t(cLower.size()), cLower.data(), cUpper.data(), std::move(coarseIfaces));
    buildMaps(fine, coarse);
}

// A mesh that gets a GAMG hierarchy keeps every level, the finest included, in the wavefront-major layout (the pencil
// layout only pays for the Krylov solvers' DIC/DILU sweeps; B200LS_PENCIL_GAMG=1 keeps it).
static void finestLevelForGamg(HostMesh& mesh) {
    LevelHost& L0 = mesh.levels[0];
    const char* keep = getenv("B200LS_PENCIL_GAMG");
    if (!L0.pencil.valid || (keep && keep[0] == '1')) return;
    const std::vector<int32_t> lower = L0.lower, upper = L0.upper;
    std::vector<HostInterface> ifaces = L0.interfaces;
    LevelHost fresh;
    buildLevel(fresh, L0.nCells, L0.nFaces, lower.data(), upper.data(), std::move(ifaces), false);
    L0 = std::move(fresh);
}

int agglomerate(HostMesh& mesh, const double* faceWeights, int32_t minCellsPerProcessor, int32_t mergeLevels,
                bool& forward, const HostComm* comm) {
    if (mesh.nRanks > 1 && (!comm || !comm->sum)) throw std::runtime_error("multi-rank agglomeration needs a communicator");
    if (mergeLevels != 1) throw std::runtime_error("mergeLevels != 1 is not supported");
    mesh.levels.resize(1);
    finestLevelForGamg(mesh);
    mesh.levels[0].hasCoarse = false;
    mesh.agglomerated = false;

    const int32_t maxLevels = 50;   // GAMGAgglomeration.C:248
    std::vector<double> w(faceWeights, faceWeights + mesh.levels[0].nFaces);

    int nCreated = 0;
    while (nCreated < maxLevels - 1) {
        int32_t nCoarse = -1;
        std::vector<int32_t> map = pairAgglomerate(nCoarse, mesh.levels[nCreated], w, forward);

        // continueAgglomerating: global sums over ranks
        const int64_t totalCoarse = mesh.nRanks > 1 ? comm->sum(nCoarse) : nCoarse;
        const int64_t totalFine = mesh.nRanks > 1 ? comm->sum(mesh.levels[nCreated].nCells) : mesh.levels[nCreated].nCells;
        if (totalCoarse < int64_t(mesh.nRanks) * minCellsPerProcessor || !(totalCoarse < totalFine)) break;

        appendCoarseLevel(mesh, std::move(map), nCoarse, comm);

        // restrictFaceField of the weights for the next level (sequential adds in fine-face order)
        const LevelHost& fine = mesh.levels[nCreated];
        std::vector<double> cw(mesh.levels[nCreated + 1].nFaces, 0.0);
        for (int32_t f = 0; f < fine.nFaces; f++) {
            const int32_t cf = fine.faceRestrictAddr[f];
            if (cf >= 0) cw[cf] += w[f];
        }
        w.swap(cw);
        nCreated++;
    }
    mesh.agglomerated = true;
    return nCreated;
}

int agglomerateFromMaps(HostMesh& mesh, int32_t nCoarseLevels, const int32_t* const* restrictAddr,
                        const int32_t* nCoarseCells, const HostComm* comm) {
    mesh.levels.resize(1);
    finestLevelForGamg(mesh);
    mesh.levels[0].hasCoarse = false;
    mesh.agglomerated = false;
    for (int32_t k = 0; k < nCoarseLevels; k++) {
        const int32_t nFine = mesh.levels[k].nCells;
        appendCoarseLevel(mesh, std::vector<int32_t>(restrictAddr[k], restrictAddr[k] + nFine), nCoarseCells[k], comm);
    }
    mesh.agglomerated = true;
    return nCoarseLevels;
}

}  // namespace b200ls
