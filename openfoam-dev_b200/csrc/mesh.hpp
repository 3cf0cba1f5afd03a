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

// Streamed sweep plan (structured hex blocks): the rows are partitioned into PARTS of <=32 "pencils" (lines of cells
// along the fastest index); one warp owns a part and walks it step by step, lane = pencil.  A dependency produced
// by the same warp in the previous step is read from a shared-memory ring (no L2 round trip); every other
// dependency is "external": its producer belongs to a part earlier in the launch order (or to an earlier step of
// the same part) and is polled from global memory -- prefetched several steps ahead, which works because a part
// naturally lags the parts it depends on.  Arithmetic and its order per row are those of the wavefront kernels.
struct StreamRec {
    int32_t pos;      // row position handled by this lane in this step, -1 = idle
    int32_t ebase;    // first entry of the row in the triangle's CSR value array
    int32_t ext0;     // positions of up to two external dependencies (-1 = none)
    int32_t ext1;
};
struct StreamPlan {
    bool valid = false;
    int32_t nParts = 0;
    std::vector<int32_t> partStart;   // [nParts + 1], in steps; record index = step*32 + lane
    std::vector<StreamRec> rec;
    // bits 0-2: number of dependencies nd (<= 3), in processing order n = 0..nd-1 (forward: ascending CSR
    // entries; backward: descending).  Dependency n: bit 3+6n = 1 if external, bits 4+6n .. 8+6n = source lane
    // (internal) or external slot 0/1.
    std::vector<uint32_t> meta;
};

struct LevelHost {
    int32_t nCells = 0, nFaces = 0;
    std::vector<int32_t> lower, upper;                      // reference face order
    std::vector<int32_t> losort, ownerStart, losortStart;   // reference-exact (lduAddressing.C:32-170)
    std::vector<int32_t> fwdOffsets, fwdRows, bwdOffsets, bwdRows;   // canonical wavefronts (cells)
    int32_t maxFwdSpan = 1;         // max over faces of Lf[upper] - Lf[lower] (1 on structured blocks)

    // ---- native layout: rows in forward-wavefront-major order ("positions") ----
    std::vector<int32_t> perm;      // position -> cell
    std::vector<int32_t> ipos;      // cell -> position
    std::vector<int32_t> Lptr, Lcol, Lface;   // neighbour-side entries of each row (ascending face)
    std::vector<int32_t> Uptr, Ucol, Uface;   // owner-side entries of each row (ascending face)
    std::vector<int32_t> Lidx, Uidx;          // face -> entry index in the L / U arrays
    std::vector<SweepTask> fwdTasks, bwdTasks;
    std::vector<int32_t> bwdPos;    // backward processing order -> position

    // ---- streamed sweeps (empty unless the addressing is a structured block, see buildStreamPlans) ----
    int32_t blockDims[3] = {0, 0, 0};
    StreamPlan fwdStream, bwdStream;

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
void buildLevel(LevelHost& L, int32_t nCells, int32_t nFaces, const int32_t* lower, const int32_t* upper,
                std::vector<HostInterface> interfaces);

// Detects an nx*ny*nz hex block numbered i-fastest (blockMesh single block) and builds the streamed sweep plans
// for it; leaves them invalid otherwise.  minCells: do not bother below this size.
void buildStreamPlans(LevelHost& L, int32_t minCells);

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
