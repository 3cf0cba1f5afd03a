// Device-side containers of libb200ls: context (stream, NCCL), per-level addressing in HBM, per-matrix
// coefficient storage and the work-vector pool.
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdint>
#include <cstdio>
#include <array>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "mesh.hpp"

namespace b200ls {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define B2_CUDA(call)                                                                                    \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            throw ::b200ls::CudaError(std::string(#call) + " failed: " + cudaGetErrorString(e_) + " (" + \
                                      __FILE__ + ":" + std::to_string(__LINE__) + ")");                  \
        }                                                                                                \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p;
            n = o.n;
            o.p = nullptr;
            o.n = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count) {
        if (count == n && p) return;
        release();
        n = count;
        // 16 bytes of slack: the streamed sweeps prefetch aligned 16-byte pairs of doubles
        if (count) B2_CUDA(cudaMalloc(&p, count * sizeof(T) + 16));
    }
    void upload(const std::vector<T>& h, cudaStream_t s) {
        alloc(h.size());
        if (!h.empty()) B2_CUDA(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    }
};

// NCCL entry points resolved with dlopen at b200ls_init (no link-time dependency: the library must load on
// machines without NCCL/GPU for the host-logic tests)
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

// Peer-to-peer communication over NVLink (one process per GPU): every rank exports one arena with CUDA IPC and maps
// the arenas of all peers.  Small all-reduces and halo exchanges then are plain remote stores + release/acquire
// flags issued from our own kernels instead of NCCL calls.
static constexpr int kMaxRanks = 16;
static constexpr size_t kP2PReduceBytes = 4096;          // [2 parities][kMaxRanks][4] doubles + [2][kMaxRanks] epochs
struct P2PView {                                          // passed by value to kernels
    int rank, nRanks;
    char* peer[kMaxRanks];                                // arena base of every rank as mapped in this process
};
struct P2PState {
    bool enabled = false;
    char* arena = nullptr;
    size_t arenaBytes = 0, bump = 0;
    P2PView view{};
    unsigned long long reduceEpoch = 0;
    std::vector<void*> opened;                            // cudaIpcOpenMemHandle results
    std::multimap<size_t, char*> freeBlocks;              // arena blocks handed back by destroyed matrices, by size
};

struct Context {
    bool initialised = false;
    P2PState p2p;
    int device = 0;
    int numSMs = 148;
    int rank = 0, nRanks = 1;
    cudaStream_t stream = nullptr;
    NcclApi nccl;
    ncclComm_t comm = nullptr;
    // reduction scratch
    DevBuf<double> partials;
    DevBuf<unsigned int> ticket;
    DevBuf<int> errFlag;
    DevBuf<int> stopFlag;                      // device-side loop condition of speculatively enqueued Krylov iterations
    double* pinned = nullptr;       // pinned host scratch for scalar read-back (64 doubles)
    int64_t launches = 0;           // kernels launched since the counter was last reset
    int sweepBlocksPerSM = 0;       // 0 = occupancy maximum
};

Context& ctx();
void ensureInit();

// device copy of mesh.hpp PencilTile plus what the helper warp needs of its four neighbours (pencil.cuh)
struct PencilTileDev {
    int base, w, wj, wk;
    int j0, k0, pad0, pad1;
    int nbrBase[4];     // (J-1,K), (J,K-1), (J+1,K), (J,K+1): first position or -1
    int nbrW[4];        // row stride of that tile
    int nbrWj[4];       // its pencils along j
};

// Addressing of one level in HBM (built once per mesh)
// Coarsest GAMG level with coupled patches (several ranks and/or cyclic halves): the blocks of all ranks concatenated
// in rank order, so that one single-thread kernel per rank can replay the reference's distributed PCG/PBiCGStab without
// a launch or a collective per operation (solver.cu: ensureCoarsestGather / solveCoarsest).
struct CoarsestGather {
    bool tried = false, ok = false;
    int nCells = 0, nFaces = 0, nCouple = 0;        // totals over ranks
    int maxCells = 0, maxFaces = 0, maxCouple = 0;  // per-rank maxima: padded block sizes of the all-gathers
    int blockLen = 0;                               // maxCells + 2*maxFaces + maxCouple doubles per rank
    int myCouple = 0;
    std::vector<int> ifaceCoupleOff;                // start of every local interface inside this rank's coupling block
    DevBuf<int> lower, upper;                       // global cell numbers (block offset added), face blocks in rank order
    DevBuf<int> cRow, cCol;                         // couplings in (rank, patch, face) order: Apsi[cRow] -= coef*psi[cCol]
    DevBuf<int> cellOff, faceOff, coupleOff;        // [nRanks+1]
};

struct DevLevel {
    int nCells = 0, nFaces = 0;
    DevBuf<int> perm, ipos;
    DevBuf<int> Lptr, Lcol, Lface, Uptr, Ucol, Uface, LtoU;
    DevBuf<unsigned char> Lslot;    // slot of each L entry inside its owner's U row (symmetric SpMV); empty if > 255
    bool hasLslot = false;
    DevBuf<int2> fwdTasks, bwdTasks;
    std::vector<int2> hostFwdTasks;                 // host copy + wavefront index of every task, for fused task lists
    std::vector<int> hostFwdTaskLevel;
    int nFwdLevels = 0;
    int maxFwdSpan = 1;                             // max wavefront distance between the two cells of a face
    std::map<int, DevBuf<int4>> multiSweepTasks;    // nSweeps -> tasks ordered by tau = level + 2*sweep
    DevBuf<int> bwdPos;
    int nFwdTasks = 0, nBwdTasks = 0;
    DevBuf<int> fwdPos;                             // forward processing order -> position (empty: identity)
    // pencil sweeps (structured blocks only; pencil.cuh)
    bool hasPencil = false;
    int pNx = 0, pNy = 0, pNz = 0, pWJ = 0, pWK = 0;
    int pCols = 0;                                  // neighbour values an interior tile needs per step (one side)
    int pSkewUnits = 0;                             // (WJ-1) + (WK-1): skew of the last lane of a full tile, in units of SKEW
    DevBuf<PencilTileDev> pTiles;
    DevBuf<int> pOrder;
    int nPencilTiles = 0;
    DevBuf<int> pGroupTiles;                        // pairs of tiles consecutive along k, in group-wavefront order (k_pencil G = 2)
    int nPencilGroups = 0;
    // interfaces
    int nIfaces = 0;
    std::vector<int> ifaceSize, ifaceNbr;
    std::vector<int> ifacePartner;      // >= 0: cyclic half coupled to that patch of this rank; -1: processor patch
    bool anyProcIface = false;
    std::vector<DevBuf<int>> ifaceCellsPos;
    DevBuf<int> bRowPos, bRowPtr, bEntIface, bEntFace;
    int nBRows = 0;
    // maps to the next coarser level
    bool hasCoarse = false;
    DevBuf<int> rPtr, rFine, pMap, auPtr, auSrc, alPtr, alSrc, adPtr, adU, adL;
    std::vector<DevBuf<int>> aiPtr, aiSrc;
    // coarsest-level direct data (reference order) for the single-thread coarsest solve
    DevBuf<int> refLower, refUpper, Uidx, Lidx;
    std::vector<int> hostRefLower, hostRefUpper;
    std::vector<std::vector<int>> ifaceCellsRef;    // per interface: faceCells in reference cell numbering (host)
    std::vector<std::vector<int>> hostIfaceCellsPos;   // per interface: positions of the patch cells (host copy)
    DevBuf<int> bRowOf;                             // position -> boundary row or -1 (coupled fused Gauss-Seidel)
    CoarsestGather gather;
};

struct DeviceMesh {
    std::vector<std::unique_ptr<DevLevel>> levels;
};

}  // namespace b200ls

struct b200ls_mesh_s {
    b200ls::HostMesh host;
    std::unique_ptr<b200ls::DeviceMesh> dev;   // created lazily at first device use
    int generation = 0;                        // bumped by re-agglomeration
};
