// Host-side mesh analysis for libb200ls: everything that is integer data and built once per
// mesh (SURVEY.md 8(a) rows a5, a17, a18 + the canonical wavefront definition).
//
// No CUDA in this translation unit: it is exercised on CPU-only machines through
// b200ls_mesh_create / b200ls_agglomerate / b200ls_mesh_get_i32.
#pragma once

#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace b200ls {

// One coupled patch.  partner < 0: processor patch to rank neighbRank (processorLduInterface).  partner >= 0: one half
// of a cyclic pair on this rank (cyclicLduInterface): face i is coupled to face i of patch `partner`; the half with
// the lower patch index is the owner (cyclicLduInterface::owner()).
struct HostInterface {
    int32_t neighbRank = -1;
    int32_t partner = -1;
    std::vector<int32_t> faceCells;      // reference cell index of each patch face
};

// One (start, count<=32) slice of a wavefront: the unit of work of one warp in the sweep kernels.
struct SweepTask {
    int32_t start;
    int32_t count;
};

// Native-space maps from one level to the next coarser one (device gathers, all order-preserving).
struct AgglomMaps {
    // restrictField: coarse position -> fine positions, ascending FINE CELL index
    std::vector<int32_t> rPtr, rFine;
    // prolongField: fine position -> coarse position
    std::vector<int32_t> pMap;
    // agglomerateMatrix: coarse U/L entry -> fine value refs (ascending fine face index);
    // ref r < nFineFaces selects fine Uval[r], otherwise fine Lval[r - nFineFaces]
    std::vector<int32_t> uPtr, uSrc, lPtr, lSrc;
    // fine faces interior to a coarse cell, by coarse position (ascending fine face): Uval + Lval entry
    std::vector<int32_t> dPtr, dU, dL;
    // interface coefficient restriction, per interface: coarse patch face -> fine patch faces
    std::vector<std::vector<int32_t>> iPtr, iSrc;
};

// Pencil plan (structured hex blocks numbered i-fastest, e.g. a blockMesh single block or a `simple` subdomain of one).
// The rows of the level are laid out TILE-MAJOR instead of wavefront-major: a tile is a bundle of wj x wk <= 32
// "pencils" (lines of cells along i); position = tile.base + i*tile.w + (jj + wj*kk).  One warp owns a tile and walks
// it along i with lane = pencil; lane (jj, kk) is skewed by jj + kk steps so that the three lower neighbours of a row
// are the lane's own previous row and the previous rows of two neighbouring lanes (registers + shuffles).  Only lanes
// on the low-j / low-k faces of a tile read values of other tiles (tile (J-1, K) and (J, K-1)), which were launched
// earlier (tile-wavefront order J + K).  Every per-row array is then contiguous per (tile, i-range): the kernels
// stream it with bulk async copies and need no per-row records.  (csrc/pencil.cuh)
struct PencilTile {
    int32_t base;       // first position of the tile
    int32_t w;          // lanes in use = wj*wk (row stride of the tile)
    int32_t wj, wk;     // pencils along j / along k
    int32_t j0, k0;     // first pencil
    int32_t nbr[4];     // tile index of (J-1,K), (J,K-1), (J+1,K), (J,K+1) or -1
};
struct PencilPlan {
    bool valid = false;
    int32_t nx = 0, ny = 0, nz = 0;
    int32_t WJ = 0, WK = 0;             // pencils of a full tile
    int32_t nJ = 0, nK = 0;             // tiles along j / k
    std::vector<PencilTile> tiles;      // memory order: J fastest
    std::vector<int32_t> fwdOrder;      // launch order of the forward sweeps (tile wavefronts); backward = reversed
};

struct LevelHost {
    int32_t nCells = 0, nFaces = 0;
    std::vector<int32_t> lower, upper;                      // reference face order
    std::vector<int32_t> losort, ownerStart, losortStart;   // reference-exact (lduAddressing.C:32-170)
    std::vector<int32_t> fwdOffsets, fwdRows, bwdOffsets, bwdRows;   // canonical wavefronts (cells)
    int32_t maxFwdSpan = 1;         // max over faces of Lf[upper] - Lf[lower] (1 on structured blocks)

    // ---- native layout: rows at "positions": forward-wavefront-major order, or tile-major when the level has a
    // pencil plan ----
    std::vector<int32_t> perm;      // position -> cell
    std::vector<int32_t> ipos;      // cell -> position
    std::vector<int32_t> Lptr, Lcol, Lface;   // neighbour-side entries of each row (ascending face)
    std::vector<int32_t> Uptr, Ucol, Uface;   // owner-side entries of each row (ascending face)
    std::vector<int32_t> Lidx, Uidx;          // face -> entry index in the L / U arrays
    std::vector<SweepTask> fwdTasks, bwdTasks;
    std::vector<int32_t> bwdPos;    // backward processing order -> position
    std::vector<int32_t> fwdPos;    // forward processing order -> position (empty = identity: wavefront-major layout)

    // ---- structured block: tile-major layout + pencil sweeps (invalid otherwise) ----
    PencilPlan pencil;

    std::vector<HostInterface> interfaces;
    // rows touched by interfaces: boundary-row CSR in (patch, face) order
    std::vector<int32_t> bRowPos, bRowPtr, bEntryIface, bEntryFace;

    // ---- agglomeration to the next level (empty on the coarsest) ----
    bool hasCoarse = false;
    int32_t nCoarseCells = 0, nCoarseFaces = 0;
    std::vector<int32_t> restrictAddr, faceRestrictAddr, faceFlip;
    std::vector<std::vector<int32_t>> patchFaceRestrictAddr;   // per interface
    AgglomMaps maps;
};

// Host-side communication needed while agglomerating a decomposed mesh (once per mesh): the neighbour exchange
// of restrictMap over every processor patch (GAMGAgglomerateLduAddressing.C:268-283,
// processorGAMGInterface.C:195-214) and the global sums of continueAgglomerating (GAMGAgglomeration.C:211-229).
struct HostComm {
    // send[i] goes to nbr[i]; recv[i] (same length) comes from nbr[i]; patches in this rank's patch order
    std::function<void(const std::vector<int32_t>& nbr, const std::vector<std::vector<int32_t>>& send,
                       std::vector<std::vector<int32_t>>& recv)> exchange;
    std::function<int64_t(int64_t)> sum;
};

struct HostMesh {
    std::vector<LevelHost> levels;      // [0] = finest
    bool agglomerated = false;
    int nRanks = 1, rank = 0;
};

// Throws std::runtime_error on invalid input (not upper-triangular ordered, out-of-range labels).
// allowPencil = false keeps the wavefront-major layout even on a structured block (levels of a GAMG hierarchy: their
// fused multi-sweep Gauss-Seidel kernel works on wavefront-major rows).
void buildLevel(LevelHost& L, int32_t nCells, int32_t nFaces, const int32_t* lower, const int32_t* upper,
                std::vector<HostInterface> interfaces, bool allowPencil = true);

// Detects an nx*ny*nz hex block numbered i-fastest with faces in upper-triangular order (blockMesh single block)
// and chooses the pencil tiling; plan.valid stays false otherwise.  Called by buildLevel (B200LS_PENCIL=0 disables,
// B200LS_PENCIL_MIN_CELLS sets the size below which the wavefront layout is kept).
void buildPencilPlan(const LevelHost& L, PencilPlan& plan);

// pairGAMGAgglomeration::agglomerate(nCoarseCells, addressing, weights) (pairGAMGAgglomerate.C:123-301).
// `forward` is the reference's static forward_ flag: read, used, and toggled.
std::vector<int32_t> pairAgglomerate(int32_t& nCoarseCells, const LevelHost& fine,
                                     const std::vector<double>& faceWeights, bool& forward);

// Whole level loop of pairGAMGAgglomeration::agglomerate(mesh, weights) (pairGAMGAgglomerate.C:31-118)
// with continueAgglomerating (GAMGAgglomeration.C:205-230).  Returns number of coarse levels.
int agglomerate(HostMesh& mesh, const double* faceWeights, int32_t minCellsPerProcessor, int32_t mergeLevels,
                bool& forward, const HostComm* comm = nullptr);

// Levels supplied by the caller (the plugin passes GAMGAgglomeration::restrictAddressing(level) of the reference's
// own cached agglomeration object); everything derived from them is rebuilt here.
int agglomerateFromMaps(HostMesh& mesh, int32_t nCoarseLevels, const int32_t* const* restrictAddr,
                        const int32_t* nCoarseCells, const HostComm* comm = nullptr);

}  // namespace b200ls
