// Solver drivers of libb200ls: matrix upload, SpMV family, DIC/DILU, smoothers, PCG, PBiCGStab, GAMG.
//
// Reference algorithms (paths under /root/reference/src/OpenFOAM/matrices/lduMatrix):
//   solvers/PCG/PCG.C:65-193, solvers/PBiCGStab/PBiCGStab.C:68-254, solvers/smoothSolver/smoothSolver.C:77-193
//   solvers/GAMG/GAMGSolverSolve.C:31-552, GAMGSolverScale.C:31-76, GAMGSolverAgglomerateMatrix.C:33-270
//   lduMatrix/lduMatrixSolver.C:174-197 (normFactor), LduMatrix/LduMatrix/SolverPerformance.C:32-92
#include "solver.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "kernels.cuh"
#include "pencil.cuh"

namespace b200ls {

// ------------------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------------------

static constexpr int kMaxPartials = 2 * 65536;   // up to 65536 blocks x 2 values

Context& ctx() {
    static Context c;
    return c;
}

void ensureInit() {
    Context& c = ctx();
    if (c.initialised) return;
    int nDev = 0;
    cudaError_t e = cudaGetDeviceCount(&nDev);
    if (e != cudaSuccess || nDev == 0) {
        throw CudaError("no CUDA device available: libb200ls has no CPU fallback (" +
                        std::string(cudaGetErrorString(e)) + ")");
    }
    B2_CUDA(cudaSetDevice(c.device));
    cudaDeviceProp prop;
    B2_CUDA(cudaGetDeviceProperties(&prop, c.device));
    c.numSMs = prop.multiProcessorCount;
    B2_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.partials.alloc(kMaxPartials);
    c.ticket.alloc(1);
    c.errFlag.alloc(1);
    c.stopFlag.alloc(1);
    B2_CUDA(cudaMemsetAsync(c.stopFlag.p, 0, sizeof(int), c.stream));
    B2_CUDA(cudaMemsetAsync(c.ticket.p, 0, sizeof(unsigned int), c.stream));
    B2_CUDA(cudaMemsetAsync(c.errFlag.p, 0, sizeof(int), c.stream));
    B2_CUDA(cudaMallocHost(&c.pinned, 64 * sizeof(double)));
    if (const char* s = getenv("B200LS_SWEEP_BLOCKS_PER_SM")) c.sweepBlocksPerSM = atoi(s);
    c.initialised = true;
}

static inline cudaStream_t S() { return ctx().stream; }

static inline int gridStride(int n) {
    const int b = (n + 255) / 256;
    return std::max(1, std::min(b, ctx().numSMs * 8));
}
static inline int gridRows(int n) { return std::max(1, (n + 255) / 256); }

#define LAUNCH(kernel, grid, block, ...)                  \
    do {                                                  \
        kernel<<<(grid), (block), 0, S()>>>(__VA_ARGS__); \
        ctx().launches++;                                 \
    } while (0)

// Opt-in phase timer for the V-cycle (B200LS_VPROF=<file prefix>): events on the solver stream at phase boundaries,
// summed per label when the solve ends.  Off: one branch per mark.
struct PhaseProf {
    bool on = false;
    std::string prefix;
    std::vector<std::pair<std::string, cudaEvent_t>> marks;
    PhaseProf() {
        const char* e = getenv("B200LS_VPROF");
        on = e && *e;
        if (on) prefix = e;
    }
};
static PhaseProf g_vprof;
static inline void vmark(const char* phase, int level) {
    if (!g_vprof.on) return;
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    cudaEventRecord(ev, S());
    char label[64];
    snprintf(label, sizeof(label), "L%02d %s", level, phase);
    g_vprof.marks.emplace_back(label, ev);
}
static void vprofDump(int nCycles) {
    if (!g_vprof.on || g_vprof.marks.size() < 2) return;
    cudaStreamSynchronize(S());
    std::map<std::string, std::pair<double, int>> sum;
    double total = 0;
    for (size_t i = 0; i + 1 < g_vprof.marks.size(); i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, g_vprof.marks[i].second, g_vprof.marks[i + 1].second);
        auto& e = sum[g_vprof.marks[i + 1].first];   // a mark closes the phase it names
        e.first += ms;
        e.second++;
        total += ms;
    }
    FILE* f = fopen((g_vprof.prefix + ".rank" + std::to_string(ctx().rank)).c_str(), "a");
    if (f) {
        fprintf(f, "# V-cycle phases, %d cycles, %.3f ms in marked phases\n", nCycles, total);
        for (auto& kv : sum) fprintf(f, "%-28s %9.4f ms  x%d\n", kv.first.c_str(), kv.second.first, kv.second.second);
        fclose(f);
    }
    for (auto& mk : g_vprof.marks) cudaEventDestroy(mk.second);
    g_vprof.marks.clear();
}

static void checkLaunch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw CudaError(std::string(what) + ": " + cudaGetErrorString(e));
}

template <class K>
static void launchSweep(K kernel, SweepArgs& a) {
    if (a.nTasks == 0) return;
    static std::map<const void*, int> occCache;
    Context& c = ctx();
    int occ;
    auto it = occCache.find((const void*)kernel);
    if (it == occCache.end()) {
        B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, 0));
        occCache[(const void*)kernel] = occ;
    } else {
        occ = it->second;
    }
    // look-ahead of ~4 CTAs per SM hides the static-load latency; more only adds polling traffic (measured)
    occ = std::min(occ, c.sweepBlocksPerSM > 0 ? c.sweepBlocksPerSM : 4);
    const int blocks = std::max(1, std::min(occ * c.numSMs, (a.nTasks + 7) / 8));
    a.err = c.errFlag.p;
    a.partials = c.partials.p;
    a.ticket = c.ticket.p;
    void* args[] = {&a};
    B2_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3(blocks), dim3(256), args, 0, c.stream));
    c.launches++;
}

// Pencil sweeps (pencil.cuh): one four-warp CTA per tile, persistent over the tile list.  Cooperative launch: a tile
// polls values of tiles earlier in the launch order, which must be resident or finished.
template <int MODE, int SKEW, int NS, int G>
static void launchPencilCfg(PencilArgs& a, int cols) {
    if (a.nTiles == 0) return;
    Context& c = ctx();
    constexpr bool GS = PencilTraits<MODE>::GS;
    auto kernel = k_pencil<MODE, SKEW, NS, G>;
    if (cols * (GS ? 2 : 1) > 32 * kPencilCPL) throw CudaError("pencil sweep: too many neighbour columns for the helper warp");
    const size_t smem = G * ((size_t(pencilSmemBytes<MODE, NS>(a.extW, a.dotOut != nullptr)) + 127) & ~size_t(127));
    // per instantiation: the dynamic shared-memory limit only ever grows; occupancy per size
    static size_t smemLimit = 0;
    static std::map<size_t, int> occCache;
    if (smem > smemLimit) {
        B2_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        smemLimit = smem;
    }
    int occ;
    auto it = occCache.find(smem);
    if (it == occCache.end()) {
        B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kPencilThreads * G, smem));
        occCache[smem] = occ;
    } else {
        occ = it->second;
    }
    if (occ < 1) throw CudaError("pencil sweep kernel does not fit on an SM");
    const int cap = getenv("B200LS_PENCIL_CTAS_PER_SM") ? atoi(getenv("B200LS_PENCIL_CTAS_PER_SM")) : 0;
    if (cap > 0) occ = std::min(occ, cap);
    const int maxCtas = getenv("B200LS_PENCIL_MAX_CTAS") ? atoi(getenv("B200LS_PENCIL_MAX_CTAS")) : 0;   // debugging
    int blocks = std::max(1, std::min(occ * c.numSMs, G == 1 ? a.nTiles : a.nGroups));
    if (maxCtas > 0) blocks = std::min(blocks, maxCtas);
    a.err = c.errFlag.p;
    a.partials = c.partials.p;
    a.ticket = c.ticket.p;
    const int dbg = getenv("B200LS_PENCIL_DEBUG") ? atoi(getenv("B200LS_PENCIL_DEBUG")) : 0;
    a.debug = dbg;
    // debugging aid: B200LS_PENCIL_PROF=<file> dumps 16 counters per tile of every pencil launch (pencil.cuh PencilArgs::prof)
    static const char* profFile = getenv("B200LS_PENCIL_PROF");
    static DevBuf<unsigned long long> profBuf;
    if (profFile) {
        profBuf.alloc(size_t(a.nTiles) * kPencilProfWords);
        B2_CUDA(cudaMemsetAsync(profBuf.p, 0, size_t(a.nTiles) * kPencilProfWords * 8, c.stream));
        a.prof = profBuf.p;
    }
    void* args[] = {&a};
    cudaError_t le = cudaLaunchCooperativeKernel((const void*)kernel, dim3(blocks), dim3(kPencilThreads * G), args, smem, c.stream);
    if (le != cudaSuccess)
        throw CudaError(std::string("pencil sweep launch failed: ") + cudaGetErrorString(le) + " (mode " +
                        std::to_string(MODE) + ", " + std::to_string(blocks) + " CTAs, " + std::to_string(smem) +
                        " B shared memory, occupancy " + std::to_string(occ) + "/SM)");
    c.launches++;
    if (profFile) {
        std::vector<unsigned long long> h(size_t(a.nTiles) * kPencilProfWords);
        B2_CUDA(cudaMemcpyAsync(h.data(), profBuf.p, h.size() * 8, cudaMemcpyDeviceToHost, c.stream));
        B2_CUDA(cudaStreamSynchronize(c.stream));
        if (FILE* f = fopen(profFile, "a")) {
            fprintf(f, "# mode %d skew %d stages %d tiles %d ctas %d group %d\n", MODE, SKEW, NS, a.nTiles, blocks, G);
            for (int t = 0; t < a.nTiles; t++) {
                for (int q = 0; q < kPencilProfWords; q++) fprintf(f, "%llu ", h[size_t(t) * kPencilProfWords + q]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
}

// (skew, ring stages) of a pencil launch: the first configuration of the mode's list whose raw ring holds the rows a
// step can touch (pencilFits); B200LS_PENCIL_CFG=<skew><stages> (14, 18, 24, 28) moves one to the front.
static bool pencilConfig(int skewUnits, bool gs, int& skew, int& stages) {
    const int forced = getenv("B200LS_PENCIL_CFG") ? atoi(getenv("B200LS_PENCIL_CFG")) : 0;
    const int order[5] = {forced, 14, 18, 24, 28};
    for (int q = forced ? 0 : 1; q < 5; q++) {
        const int sk = order[q] / 10, ns = order[q] % 10;
        if ((sk == 1 || sk == 2) && (ns == 4 || ns == 8) && pencilFits(skewUnits, sk, ns, gs)) {
            skew = sk;
            stages = ns;
            return true;
        }
    }
    return false;
}

template <int MODE>
static void launchPencil(PencilArgs& a, const DevLevel& D) {
    int skew = 0, stages = 0;
    if (!pencilConfig(D.pSkewUnits, PencilTraits<MODE>::GS, skew, stages))
        throw CudaError("pencil sweep: no ring configuration fits this tile shape");
    // EXPERIMENT, opt-in (B200LS_PENCIL_GROUP=2): two tiles per CTA (k_pencil<..., G = 2>: the K faces between them
    // are handed over in shared memory; 30 instead of 46 L2 hops at 128^3, -5 % there).  With the per-step barriers
    // initialised once and their phase carried across tiles it is bit-exact on every small multi-round case tried,
    // but one of the 128^3 checks of tests/test_gpu_scale.py still failed with it on, so it stays off by default and
    // out of the test suite.  Gauss-Seidel and the profiling aid always keep one tile per CTA.
    const bool noGroup = !(getenv("B200LS_PENCIL_GROUP") && atoi(getenv("B200LS_PENCIL_GROUP")) >= 2);
    static const bool prof = getenv("B200LS_PENCIL_PROF") != nullptr;
    const int groupModes = getenv("B200LS_PENCIL_GROUP_MODES") ? atoi(getenv("B200LS_PENCIL_GROUP_MODES")) : 7;   // debugging: bit MODE
    if (!PencilTraits<MODE>::GS && ((groupModes >> MODE) & 1) && skew == 1 && stages == 4 && !noGroup && !prof && D.pNz > D.pWK &&
        D.nPencilGroups > 0) {
        a.groupTiles = D.pGroupTiles.p;
        a.nGroups = D.nPencilGroups;
        return launchPencilCfg<MODE, 1, 4, kPencilGroup>(a, D.pCols);
    }
    if (skew == 1 && stages == 4) return launchPencilCfg<MODE, 1, 4, 1>(a, D.pCols);
    if (skew == 1) return launchPencilCfg<MODE, 1, 8, 1>(a, D.pCols);
    if (stages == 4) return launchPencilCfg<MODE, 2, 4, 1>(a, D.pCols);
    return launchPencilCfg<MODE, 2, 8, 1>(a, D.pCols);
}

void checkSweepError() {
    Context& c = ctx();
    int h = 0;
    B2_CUDA(cudaMemcpyAsync(&h, c.errFlag.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    B2_CUDA(cudaStreamSynchronize(c.stream));
    if (h) {
        B2_CUDA(cudaMemsetAsync(c.errFlag.p, 0, sizeof(int), c.stream));
        throw CudaError(h == 2 ? "timed out waiting for a peer GPU (P2P halo / all-reduce)"
                               : "wavefront sweep timed out waiting for a dependency (internal error)");
    }
}

static void readScalars(const double* dev, int n) {
    B2_CUDA(cudaMemcpyAsync(ctx().pinned, dev, n * sizeof(double), cudaMemcpyDeviceToHost, S()));
    B2_CUDA(cudaStreamSynchronize(S()));
}

// sum over ranks of `count` device doubles (gSum*/reduce(sumOp), allReduceTemplates.C:157)
static void allReduce(double* dev, int count) {
    Context& c = ctx();
    if (c.nRanks == 1) return;
    if (c.p2p.enabled && count <= 4) {
        // our own kernel over the mapped peer arenas: remote stores + release/acquire epochs, no NCCL launch
        k_allreduce_p2p<<<1, 32, 0, c.stream>>>(dev, count, c.p2p.view, ++c.p2p.reduceEpoch, c.errFlag.p);
        c.launches++;
        return;
    }
    int r = c.nccl.AllReduce(dev, dev, count, ncclDouble, ncclSum, c.comm, c.stream);
    if (r != 0) throw CudaError(std::string("ncclAllReduce: ") + c.nccl.GetErrorString((ncclResult_t)r));
}

// ------------------------------------------------------------------------------------------------------------
// device mesh
// ------------------------------------------------------------------------------------------------------------

static void uploadLevel(DevLevel& D, const LevelHost& H, bool coarsest) {
    cudaStream_t s = S();
    D.nCells = H.nCells;
    D.nFaces = H.nFaces;
    D.perm.upload(H.perm, s);
    D.ipos.upload(H.ipos, s);
    D.Lptr.upload(H.Lptr, s);
    D.Lcol.upload(H.Lcol, s);
    D.Lface.upload(H.Lface, s);
    D.Uptr.upload(H.Uptr, s);
    D.Ucol.upload(H.Ucol, s);
    D.Uface.upload(H.Uface, s);
    {
        std::vector<int32_t> ltou(H.nFaces);
        for (int32_t j = 0; j < H.nFaces; j++) ltou[j] = H.Uidx[H.Lface[j]];
        D.LtoU.upload(ltou, s);
        // slot of the matching entry inside the owner's U row
        std::vector<unsigned char> slot(H.nFaces);
        bool fits = true;
        for (int32_t j = 0; j < H.nFaces && fits; j++) {
            const int32_t d = ltou[j] - H.Uptr[H.Lcol[j]];
            if (d < 0 || d > 255) fits = false;
            else slot[j] = (unsigned char)d;
        }
        D.hasLslot = fits && H.nFaces > 0;
        if (D.hasLslot) D.Lslot.upload(slot, s);
        B2_CUDA(cudaStreamSynchronize(s));   // ltou/slot are temporaries
    }
    // structured blocks: tile table of the pencil sweeps
    D.hasPencil = H.pencil.valid;
    D.fwdPos.upload(H.fwdPos, s);
    D.nPencilTiles = 0;
    if (D.hasPencil) {
        const PencilPlan& P = H.pencil;
        D.pNx = P.nx;
        D.pNy = P.ny;
        D.pNz = P.nz;
        D.pWJ = P.WJ;
        D.pWK = P.WK;
        D.pCols = (P.nJ > 1 ? P.WK : 0) + (P.nK > 1 ? P.WJ : 0);
        D.pSkewUnits = (P.WJ - 1) + (P.WK - 1);
        std::vector<PencilTileDev> tiles(P.tiles.size());
        for (size_t t = 0; t < P.tiles.size(); t++) {
            const PencilTile& h = P.tiles[t];
            PencilTileDev& d = tiles[t];
            d.base = h.base;
            d.w = h.w;
            d.wj = h.wj;
            d.wk = h.wk;
            d.j0 = h.j0;
            d.k0 = h.k0;
            d.pad0 = d.pad1 = 0;
            for (int q = 0; q < 4; q++) {
                const int nb = h.nbr[q];
                d.nbrBase[q] = nb >= 0 ? P.tiles[nb].base : -1;
                d.nbrW[q] = nb >= 0 ? P.tiles[nb].w : 1;
                d.nbrWj[q] = nb >= 0 ? P.tiles[nb].wj : 1;
            }
        }
        D.pTiles.upload(tiles, s);
        D.pOrder.upload(P.fwdOrder, s);
        D.nPencilTiles = int(tiles.size());
        // groups of kPencilGroup tiles that follow each other along k (tile index = J + nJ*K), in wavefront order J + M
        {
            const int nM = (P.nK + kPencilGroup - 1) / kPencilGroup;
            std::vector<int> groups;
            for (int d = 0; d <= P.nJ - 1 + nM - 1; d++)
                for (int J = 0; J < P.nJ; J++) {
                    const int Mg = d - J;
                    if (Mg < 0 || Mg >= nM) continue;
                    for (int q = 0; q < kPencilGroup; q++) {
                        const int K = Mg * kPencilGroup + q;
                        groups.push_back(K < P.nK ? J + P.nJ * K : -1);
                    }
                }
            D.nPencilGroups = int(groups.size()) / kPencilGroup;
            D.pGroupTiles.upload(groups, s);
        }
        B2_CUDA(cudaStreamSynchronize(s));   // `tiles` is a temporary
    }
    static_assert(sizeof(SweepTask) == sizeof(int2), "task layout");
    D.nFwdTasks = int(H.fwdTasks.size());
    D.nBwdTasks = int(H.bwdTasks.size());
    D.fwdTasks.alloc(H.fwdTasks.size());
    D.bwdTasks.alloc(H.bwdTasks.size());
    if (D.nFwdTasks)
        B2_CUDA(cudaMemcpyAsync(D.fwdTasks.p, H.fwdTasks.data(), H.fwdTasks.size() * sizeof(int2),
                                cudaMemcpyHostToDevice, s));
    if (D.nBwdTasks)
        B2_CUDA(cudaMemcpyAsync(D.bwdTasks.p, H.bwdTasks.data(), H.bwdTasks.size() * sizeof(int2),
                                cudaMemcpyHostToDevice, s));
    D.bwdPos.upload(H.bwdPos, s);
    D.hostFwdTasks.assign(reinterpret_cast<const int2*>(H.fwdTasks.data()),
                          reinterpret_cast<const int2*>(H.fwdTasks.data()) + H.fwdTasks.size());
    D.nFwdLevels = int(H.fwdOffsets.size()) - 1;
    D.maxFwdSpan = H.maxFwdSpan;
    D.hostFwdTaskLevel.resize(H.fwdTasks.size());
    {
        int lev = 0;
        for (size_t t = 0; t < H.fwdTasks.size(); t++) {
            while (lev + 1 < D.nFwdLevels && H.fwdTasks[t].start >= H.fwdOffsets[lev + 1]) lev++;
            D.hostFwdTaskLevel[t] = lev;
        }
    }
    D.multiSweepTasks.clear();

    D.nIfaces = int(H.interfaces.size());
    D.ifaceSize.clear();
    D.ifaceNbr.clear();
    D.ifacePartner.clear();
    D.anyProcIface = false;
    D.ifaceCellsPos.clear();
    D.ifaceCellsPos.resize(D.nIfaces);
    D.hostIfaceCellsPos.clear();
    D.bRowOf.release();
    for (int i = 0; i < D.nIfaces; i++) {
        const auto& fc = H.interfaces[i].faceCells;
        D.ifaceSize.push_back(int(fc.size()));
        D.ifaceNbr.push_back(H.interfaces[i].neighbRank);
        D.ifacePartner.push_back(H.interfaces[i].partner);
        D.anyProcIface = D.anyProcIface || H.interfaces[i].partner < 0;
        std::vector<int32_t> pos(fc.size());
        for (size_t k = 0; k < fc.size(); k++) pos[k] = H.ipos[fc[k]];
        D.ifaceCellsPos[i].upload(pos, s);
        B2_CUDA(cudaStreamSynchronize(s));
        D.hostIfaceCellsPos.emplace_back(pos.begin(), pos.end());
    }
    D.ifaceCellsRef.clear();
    if (coarsest)
        for (int i = 0; i < D.nIfaces; i++)
            D.ifaceCellsRef.emplace_back(H.interfaces[i].faceCells.begin(), H.interfaces[i].faceCells.end());
    D.gather = CoarsestGather();
    D.nBRows = int(H.bRowPos.size());
    D.bRowPos.upload(H.bRowPos, s);
    D.bRowPtr.upload(H.bRowPtr, s);
    D.bEntIface.upload(H.bEntryIface, s);
    D.bEntFace.upload(H.bEntryFace, s);

    D.hasCoarse = H.hasCoarse;
    if (H.hasCoarse) {
        const AgglomMaps& M = H.maps;
        D.rPtr.upload(M.rPtr, s);
        D.rFine.upload(M.rFine, s);
        D.pMap.upload(M.pMap, s);
        D.auPtr.upload(M.uPtr, s);
        D.auSrc.upload(M.uSrc, s);
        D.alPtr.upload(M.lPtr, s);
        D.alSrc.upload(M.lSrc, s);
        D.adPtr.upload(M.dPtr, s);
        D.adU.upload(M.dU, s);
        D.adL.upload(M.dL, s);
        D.aiPtr.clear();
        D.aiSrc.clear();
        D.aiPtr.resize(M.iPtr.size());
        D.aiSrc.resize(M.iSrc.size());
        for (size_t i = 0; i < M.iPtr.size(); i++) {
            D.aiPtr[i].upload(M.iPtr[i], s);
            D.aiSrc[i].upload(M.iSrc[i], s);
        }
    }
    if (coarsest) {
        D.refLower.upload(H.lower, s);
        D.refUpper.upload(H.upper, s);
        D.Uidx.upload(H.Uidx, s);
        D.Lidx.upload(H.Lidx, s);
        D.hostRefLower.assign(H.lower.begin(), H.lower.end());
        D.hostRefUpper.assign(H.upper.begin(), H.upper.end());
    }
    B2_CUDA(cudaStreamSynchronize(s));
}

void ensureDeviceMesh(b200ls_mesh_s* mesh) {
    ensureInit();
    if (mesh->dev) return;
    mesh->dev.reset(new DeviceMesh);
    const size_t nL = mesh->host.levels.size();
    for (size_t k = 0; k < nL; k++) {
        mesh->dev->levels.emplace_back(new DevLevel);
        uploadLevel(*mesh->dev->levels.back(), mesh->host.levels[k], k + 1 == nL);
    }
}

}  // namespace b200ls

double* b200ls_matrix_s::vec(const std::string& name) {
    b200ls::Vec& v = vecs[name];
    const size_t n = mesh->host.levels[0].nCells;
    if (v.buf.n != n || !v.buf.p) v.buf.alloc(n);
    return v.buf.p;
}

namespace b200ls {

static DevLevel& DL(b200ls_matrix_s* m, int level) { return *m->mesh->dev->levels[level]; }

void toPositions(b200ls_matrix_s* m, int level, double* outPos, const double* inCell) {
    DevLevel& D = DL(m, level);
    if (D.nCells) LAUNCH(k_gather, gridStride(D.nCells), 256, outPos, inCell, D.perm.p, D.nCells);
}
void toCells(b200ls_matrix_s* m, int level, double* outCell, const double* inPos) {
    DevLevel& D = DL(m, level);
    if (D.nCells) LAUNCH(k_gather, gridStride(D.nCells), 256, outCell, inPos, D.ipos.p, D.nCells);
}

static void syncMatrixWithMesh(b200ls_matrix_s* m) {
    ensureDeviceMesh(m->mesh);
    if (m->meshGeneration != m->mesh->generation || m->levels.size() != m->mesh->dev->levels.size()) {
        m->levels.clear();
        m->levels.resize(m->mesh->dev->levels.size());
        m->meshGeneration = m->mesh->generation;
        m->coarseValid = false;
        m->valuesSet = false;
    }
}

static void setupCoupledGS(b200ls_matrix_s* m, int level);
static void fillSentinel(double* p, int n);
// Allocation inside this rank's IPC arena for one matrix level: a block of the same size handed back by a destroyed
// matrix if there is one (zeroed again: epoch flags restart at 0), else bump allocation; nullptr when the arena is
// exhausted.  Matrices are created and destroyed in the same order on every rank, so the choice is the same everywhere;
// the neighbours learn the new offsets in the exchange that follows, which is stream-ordered after the memset.
static char* arenaAlloc(MatLevel& M, size_t bytes) {
    P2PState& P = ctx().p2p;
    const size_t aligned = (bytes + 255) & ~size_t(255);
    if (!P.enabled) return nullptr;
    char* p = nullptr;
    auto it = P.freeBlocks.find(aligned);
    if (it != P.freeBlocks.end()) {
        p = it->second;
        P.freeBlocks.erase(it);
        B2_CUDA(cudaMemsetAsync(p, 0, aligned, S()));
    } else {
        if (P.bump + aligned > P.arenaBytes) return nullptr;
        p = P.arena + P.bump;
        P.bump += aligned;
    }
    M.arenaBlocks.emplace_back(p, aligned);
    return p;
}
void arenaRelease(std::vector<std::pair<char*, size_t>>& blocks) {
    if (blocks.empty()) return;
    P2PState& P = ctx().p2p;
    if (P.enabled)
        for (auto& b : blocks) P.freeBlocks.emplace(b.second, b.first);
    blocks.clear();
}

static void allGatherInts(const std::vector<int>& mine, std::vector<int>& all) {
    Context& c = ctx();
    if (c.nRanks == 1) {
        all = mine;
        return;
    }
    DevBuf<int> snd, rcv;
    snd.upload(mine, c.stream);
    rcv.alloc(mine.size() * c.nRanks);
    int r = c.nccl.AllGather(snd.p, rcv.p, mine.size(), ncclInt32, c.comm, c.stream);
    if (r != 0) throw CudaError(std::string("ncclAllGather: ") + c.nccl.GetErrorString((ncclResult_t)r));
    all.resize(mine.size() * c.nRanks);
    B2_CUDA(cudaMemcpyAsync(all.data(), rcv.p, all.size() * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    B2_CUDA(cudaStreamSynchronize(c.stream));
}

// P2P halos: place the receive buffers (two parities) and epoch flags of this level in the arena and swap their
// arena offsets with the neighbours (once per matrix level).  Falls back to the NCCL path if anything is missing.
static void setupP2PHalos(b200ls_matrix_s* m, int level) {
    DevLevel& D = DL(m, level);
    MatLevel& M = m->levels[level];
    Context& c = ctx();
    // Every rank calls this once per (matrix, level) -- also ranks without a processor patch on the level -- and the
    // decisions below are taken COLLECTIVELY: p2p.enabled is all-or-none by construction (setupP2P), the environment
    // switch is process-wide, and anything rank-local goes through an all-reduce before any rank leaves.
    if (M.p2pTried || !c.p2p.enabled) return;
    M.p2pTried = true;
    static const bool off = getenv("B200LS_NO_P2P_HALO") != nullptr;
    if (off) return;
    auto isProc = [&](int i) { return D.ifacePartner[i] < 0; };   // cyclic halves never leave this GPU
    auto anyRankSays = [&](bool mine) {   // global OR
        DevBuf<double> flag;
        flag.alloc(1);
        const double v = mine ? 1.0 : 0.0;
        B2_CUDA(cudaMemcpyAsync(flag.p, &v, sizeof(double), cudaMemcpyHostToDevice, c.stream));
        if (c.nccl.AllReduce(flag.p, flag.p, 1, ncclDouble, ncclSum, c.comm, c.stream) != 0)
            throw CudaError("ncclAllReduce failed");
        double sum = 0;
        B2_CUDA(cudaMemcpyAsync(&sum, flag.p, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        B2_CUDA(cudaStreamSynchronize(c.stream));
        return sum != 0.0;
    };
    bool duplicate = false;   // several patches to one neighbour (e.g. processor + processorCyclic): keep NCCL
    for (int i = 0; i < D.nIfaces; i++)
        for (int j = 0; j < i; j++)
            if (isProc(i) && isProc(j) && D.ifaceNbr[i] == D.ifaceNbr[j]) duplicate = true;
    if (anyRankSays(duplicate)) return;
    std::vector<long long> mine(2 * D.nIfaces), theirs(2 * D.nIfaces, -1);
    M.p2pLocalRecv.assign(D.nIfaces, nullptr);
    M.p2pLocalFlag.assign(D.nIfaces, nullptr);
    bool ok = true;
    for (int i = 0; i < D.nIfaces; i++) {
        if (!isProc(i)) {
            mine[2 * i] = mine[2 * i + 1] = theirs[2 * i] = theirs[2 * i + 1] = 0;
            continue;
        }
        char* r = arenaAlloc(M, size_t(2) * std::max(D.ifaceSize[i], 1) * sizeof(double));
        char* f = arenaAlloc(M, sizeof(unsigned long long));
        if (!r || !f) ok = false;
        M.p2pLocalRecv[i] = reinterpret_cast<double*>(r);
        M.p2pLocalFlag[i] = reinterpret_cast<unsigned long long*>(f);
        mine[2 * i] = r ? r - c.p2p.arena : -1;
        mine[2 * i + 1] = f ? f - c.p2p.arena : -1;
    }
    // exchange the offsets pairwise (also tells the neighbour if we ran out of arena)
    DevBuf<long long> dMine, dTheirs;
    dMine.upload(mine, c.stream);
    dTheirs.alloc(std::max<size_t>(theirs.size(), 1));
    B2_CUDA(cudaStreamSynchronize(c.stream));
    c.nccl.GroupStart();
    for (int i = 0; i < D.nIfaces; i++) {
        if (!isProc(i)) continue;
        c.nccl.Send(dMine.p + 2 * i, 2, ncclInt64, D.ifaceNbr[i], c.comm, c.stream);
        c.nccl.Recv(dTheirs.p + 2 * i, 2, ncclInt64, D.ifaceNbr[i], c.comm, c.stream);
    }
    if (c.nccl.GroupEnd() != 0) throw CudaError("nccl exchange of P2P halo offsets failed");
    {
        std::vector<long long> got(theirs.size());
        B2_CUDA(cudaMemcpyAsync(got.data(), dTheirs.p, sizeof(long long) * got.size(), cudaMemcpyDeviceToHost, c.stream));
        B2_CUDA(cudaStreamSynchronize(c.stream));
        for (int i = 0; i < D.nIfaces; i++)
            if (isProc(i)) { theirs[2 * i] = got[2 * i]; theirs[2 * i + 1] = got[2 * i + 1]; }
    }
    for (long long v : theirs) ok = ok && v >= 0;
    // every rank must take the same decision for a given interface pair; a global AND keeps it simple
    if (anyRankSays(!ok)) return;
    M.p2pRemoteRecv.resize(D.nIfaces);
    M.p2pRemoteFlag.resize(D.nIfaces);
    for (int i = 0; i < D.nIfaces; i++) {
        if (!isProc(i)) { M.p2pRemoteRecv[i] = nullptr; M.p2pRemoteFlag[i] = nullptr; continue; }
        char* base = c.p2p.view.peer[D.ifaceNbr[i]];
        M.p2pRemoteRecv[i] = reinterpret_cast<double*>(base + theirs[2 * i]);
        M.p2pRemoteFlag[i] = reinterpret_cast<unsigned long long*>(base + theirs[2 * i + 1]);
    }
    M.p2pTickets.alloc(D.nIfaces);
    B2_CUDA(cudaMemsetAsync(M.p2pTickets.p, 0, sizeof(unsigned int) * D.nIfaces, c.stream));
    M.haloEpoch = 0;
    M.p2pReady = true;
    setupCoupledGS(m, level);
}

// Fused Gauss-Seidel sweeps across processor patches (k_gs_multi<true>): sweep s of a patch row needs the neighbour
// rank's value after sweep s-1.  Ordering every (row, sweep) of every rank by tau = wavefront + lag*sweep with ONE lag
// for all ranks, lag > (neighbour's wavefront - my wavefront) for every coupled pair, makes every dependency --
// local or across ranks -- point to a smaller tau, so the pipelined kernels of all ranks cannot deadlock and the chain
// is nLevels + lag*(nSweeps-1) hops instead of nSweeps*nLevels plus a halo exchange per sweep.
// Collective (called by every rank at the end of setupP2PHalos).
static void setupCoupledGS(b200ls_matrix_s* m, int level) {
    DevLevel& D = DL(m, level);
    MatLevel& M = m->levels[level];
    Context& c = ctx();
    M.gsCoupled = false;
    static const bool off = getenv("B200LS_NO_COUPLED_FUSED_GS") != nullptr;
    if (off) return;
    bool okLocal = D.fwdPos.n == 0 && !D.hasPencil && int(D.hostIfaceCellsPos.size()) == D.nIfaces;
    for (int i = 0; i < D.nIfaces; i++) okLocal = okLocal && D.ifacePartner[i] < 0;
    // wavefront of every patch cell
    std::vector<std::vector<int>> myLevel(D.nIfaces), nbLevel(D.nIfaces);
    if (okLocal && D.nIfaces) {
        std::vector<int> lvl(D.nCells, 0);
        for (size_t t = 0; t < D.hostFwdTasks.size(); t++)
            for (int r = 0; r < D.hostFwdTasks[t].y; r++) lvl[D.hostFwdTasks[t].x + r] = D.hostFwdTaskLevel[t];
        for (int i = 0; i < D.nIfaces; i++) {
            myLevel[i].resize(D.ifaceSize[i]);
            for (int f = 0; f < D.ifaceSize[i]; f++) myLevel[i][f] = lvl[D.hostIfaceCellsPos[i][f]];
        }
    }
    for (int i = 0; i < D.nIfaces; i++) {
        myLevel[i].resize(D.ifaceSize[i], 0);
        nbLevel[i].assign(D.ifaceSize[i], 0);
    }
    // slots in my arena: [2 parities][kCoupledSlotSweeps][size], all-sentinel
    M.gsSlotLocal.assign(D.nIfaces, nullptr);
    M.gsSlotRemote.assign(D.nIfaces, nullptr);
    std::vector<long long> mine(std::max(D.nIfaces, 1), -1), theirs(std::max(D.nIfaces, 1), -1);
    for (int i = 0; i < D.nIfaces && okLocal; i++) {
        const size_t cnt = size_t(2) * kCoupledSlotSweeps * std::max(D.ifaceSize[i], 1);
        char* q = arenaAlloc(M, cnt * sizeof(double));
        if (!q) {
            okLocal = false;
            break;
        }
        M.gsSlotLocal[i] = reinterpret_cast<double*>(q);
        fillSentinel(M.gsSlotLocal[i], int(cnt));
        mine[i] = q - c.p2p.arena;
    }
    if (!okLocal) std::fill(mine.begin(), mine.end(), -1);
    // pairwise: slot offsets and wavefronts of the patch cells
    DevBuf<long long> dMine, dTheirs;
    dMine.upload(mine, c.stream);
    dTheirs.alloc(mine.size());
    std::vector<DevBuf<int>> dLm(D.nIfaces), dLn(D.nIfaces);
    for (int i = 0; i < D.nIfaces; i++) {
        dLm[i].upload(myLevel[i], c.stream);
        dLn[i].alloc(std::max(D.ifaceSize[i], 1));
    }
    B2_CUDA(cudaStreamSynchronize(c.stream));
    c.nccl.GroupStart();
    for (int i = 0; i < D.nIfaces; i++) {
        if (D.ifacePartner[i] >= 0) continue;
        c.nccl.Send(dMine.p + i, 1, ncclInt64, D.ifaceNbr[i], c.comm, c.stream);
        c.nccl.Recv(dTheirs.p + i, 1, ncclInt64, D.ifaceNbr[i], c.comm, c.stream);
        if (D.ifaceSize[i]) {
            c.nccl.Send(dLm[i].p, D.ifaceSize[i], ncclInt32, D.ifaceNbr[i], c.comm, c.stream);
            c.nccl.Recv(dLn[i].p, D.ifaceSize[i], ncclInt32, D.ifaceNbr[i], c.comm, c.stream);
        }
    }
    if (c.nccl.GroupEnd() != 0) throw CudaError("nccl exchange of the coupled Gauss-Seidel set-up failed");
    if (D.nIfaces)
        B2_CUDA(cudaMemcpyAsync(theirs.data(), dTheirs.p, sizeof(long long) * D.nIfaces, cudaMemcpyDeviceToHost, c.stream));
    for (int i = 0; i < D.nIfaces; i++)
        if (D.ifaceSize[i])
            B2_CUDA(cudaMemcpyAsync(nbLevel[i].data(), dLn[i].p, sizeof(int) * D.ifaceSize[i], cudaMemcpyDeviceToHost,
                                    c.stream));
    B2_CUDA(cudaStreamSynchronize(c.stream));
    int lag = D.maxFwdSpan + 1;
    for (int i = 0; i < D.nIfaces; i++) {
        if (D.ifacePartner[i] < 0 && theirs[i] < 0) okLocal = false;
        for (int f = 0; f < D.ifaceSize[i]; f++) lag = std::max(lag, nbLevel[i][f] - myLevel[i][f] + 1);
    }
    std::vector<int> all;
    // rows per wavefront: wide wavefronts are bandwidth-bound, and with sweep s+1 trailing by `lag` wavefronts the
    // iterate has left L2 before it is read again -- measured at 384^3 on 2 GPUs: 29.5k rows per wavefront 15 % slower
    // fused than sweep by sweep, 12k rows 15 % faster, 3k rows 45 % faster.  Below the limit the level is latency-bound.
    static const int maxRows = getenv("B200LS_COUPLED_FUSED_MAX_ROWS") ? atoi(getenv("B200LS_COUPLED_FUSED_MAX_ROWS")) : 16000;
    allGatherInts({lag, D.nFwdLevels, okLocal ? 1 : 0, D.nCells / std::max(1, D.nFwdLevels)}, all);
    int maxLevels = 0, rowsPerLevel = 0;
    bool ok = true;
    for (int r = 0; r < c.nRanks; r++) {
        lag = std::max(lag, all[4 * r]);
        maxLevels = std::max(maxLevels, all[4 * r + 1]);
        ok = ok && all[4 * r + 2] != 0;
        rowsPerLevel = std::max(rowsPerLevel, all[4 * r + 3]);
    }
    if (!ok || lag >= maxLevels || rowsPerLevel > maxRows) return;
    M.gsLag = lag;
    M.gsEpoch = 0;
    if (D.nIfaces == 0) return;   // nothing coupled here: this rank keeps its local fused sweeps
    for (int par = 0; par < 2; par++) {
        std::vector<CoupledView> v(D.nIfaces);
        for (int i = 0; i < D.nIfaces; i++) {
            const size_t parOff = size_t(par) * kCoupledSlotSweeps * std::max(D.ifaceSize[i], 1);
            M.gsSlotRemote[i] = reinterpret_cast<double*>(c.p2p.view.peer[D.ifaceNbr[i]] + theirs[i]);
            v[i].coeffs = M.bou[i].p;   // allocated before setupIfaceViews; the size never changes for a mesh level
            v[i].mine = M.gsSlotLocal[i] + parOff;
            v[i].theirs = M.gsSlotRemote[i] + parOff;
            v[i].size = D.ifaceSize[i];
            v[i].pad = 0;
        }
        M.gsViews[par].alloc(v.size() * sizeof(CoupledView));
        B2_CUDA(cudaMemcpyAsync(M.gsViews[par].p, v.data(), v.size() * sizeof(CoupledView), cudaMemcpyHostToDevice,
                                c.stream));
        B2_CUDA(cudaStreamSynchronize(c.stream));
    }
    if (D.bRowOf.n != size_t(D.nCells)) {
        D.bRowOf.alloc(D.nCells);
        B2_CUDA(cudaMemsetAsync(D.bRowOf.p, 0xFF, sizeof(int) * D.nCells, c.stream));
        if (D.nBRows) LAUNCH(k_scatter_index, gridRows(D.nBRows), 256, D.bRowOf.p, D.bRowPos.p, D.nBRows);
    }
    M.gsCoupled = true;
}

static void setupIfaceViews(b200ls_matrix_s* m, int level) {
    DevLevel& D = DL(m, level);
    MatLevel& M = m->levels[level];
    if (ctx().nRanks > 1) setupP2PHalos(m, level);   // collective: before any rank-local return
    if (D.nIfaces == 0) return;
    std::vector<IfaceView> v(D.nIfaces);
    for (int i = 0; i < D.nIfaces; i++) {
        M.sendBuf[i].alloc(D.ifaceSize[i]);
        M.recvBuf[i].alloc(D.ifaceSize[i]);
    }
    for (int i = 0; i < D.nIfaces; i++) {
        v[i].coeffs = M.bou[i].p;
        if (D.ifacePartner[i] >= 0) {
            // cyclic half: the neighbour values are the partner patch's packed boundary cells, same GPU, same stream
            // (cyclicGAMGInterfaceField.C:124-146: pnf = nbrPatch().interfaceInternalField(psi))
            v[i].recv = M.sendBuf[D.ifacePartner[i]].p;
            v[i].flag = nullptr;
        } else {
            v[i].recv = M.p2pReady ? M.p2pLocalRecv[i] : M.recvBuf[i].p;
            v[i].flag = M.p2pReady ? M.p2pLocalFlag[i] : nullptr;
        }
        v[i].size = D.ifaceSize[i];
    }
    M.ifaceViews.alloc(v.size() * sizeof(IfaceView));
    B2_CUDA(cudaMemcpyAsync(M.ifaceViews.p, v.data(), v.size() * sizeof(IfaceView), cudaMemcpyHostToDevice, S()));
    B2_CUDA(cudaStreamSynchronize(S()));
}

void matrixSet(b200ls_matrix_s* m, const double* diag, const double* upper, const double* lower,
               const double* const* bou, const double* const* inn, bool devicePointers) {
    syncMatrixWithMesh(m);
    DevLevel& D = DL(m, 0);
    MatLevel& M = m->levels[0];
    const int nC = D.nCells, nF = D.nFaces;
    m->symmetric = (lower == nullptr);
    m->hasFingerprint = false;
    M.diag.alloc(nC);
    M.vals.alloc(size_t(2) * nF);
    cudaStream_t s = S();
    // reference order -> native layout; host arrays pass through a staging buffer, device arrays are gathered in place
    const double *dDiag = diag, *dUpper = upper, *dLower = lower;
    if (!devicePointers) {
        m->stageA.alloc(std::max(nC, nF));
        if (lower) m->stageB.alloc(std::max(nC, nF));
    }
    if (nC) {
        if (!devicePointers) {
            h2dBytes(m->stageA.p, diag, sizeof(double) * nC);
            dDiag = m->stageA.p;
        }
        LAUNCH(k_gather, gridStride(nC), 256, M.diag.p, dDiag, D.perm.p, nC);
    }
    if (nF) {
        if (!devicePointers) {
            h2dBytes(m->stageA.p, upper, sizeof(double) * nF);
            dUpper = m->stageA.p;
        }
        LAUNCH(k_gather, gridStride(nF), 256, M.Uval(), dUpper, D.Uface.p, nF);
        if (lower) {
            if (!devicePointers) {
                h2dBytes(m->stageB.p, lower, sizeof(double) * nF);
                dLower = m->stageB.p;
            }
            LAUNCH(k_gather, gridStride(nF), 256, M.Lval(nF), dLower, D.Lface.p, nF);
        } else {
            LAUNCH(k_gather, gridStride(nF), 256, M.Lval(nF), dUpper, D.Lface.p, nF);
        }
    }
    M.bou.resize(D.nIfaces);
    M.inn.resize(D.nIfaces);
    M.sendBuf.resize(D.nIfaces);
    M.recvBuf.resize(D.nIfaces);
    for (int i = 0; i < D.nIfaces; i++) {
        M.bou[i].alloc(D.ifaceSize[i]);
        M.inn[i].alloc(D.ifaceSize[i]);
        if (D.ifaceSize[i]) {
            const cudaMemcpyKind kind = devicePointers ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
            B2_CUDA(cudaMemcpyAsync(M.bou[i].p, bou[i], sizeof(double) * D.ifaceSize[i], kind, s));
            B2_CUDA(cudaMemcpyAsync(M.inn[i].p, inn[i], sizeof(double) * D.ifaceSize[i], kind, s));
        }
    }
    setupIfaceViews(m, 0);
    B2_CUDA(cudaStreamSynchronize(s));
    checkLaunch("matrixSet");
    for (auto& L : m->levels) {
        L.rDValid = false;
        L.pPlanesValid = false;
    }
    m->coarseValid = false;
    m->valuesSet = true;
}

// ------------------------------------------------------------------------------------------------------------
// halo exchange (initMatrixInterfaces/updateMatrixInterfaces, lduMatrixUpdateMatrixInterfaces.C:30-266;
// processorFvPatchScalarField.C:36-152): pack -> ncclSend/ncclRecv group -> ordered apply
// ------------------------------------------------------------------------------------------------------------

static void haloExchange(b200ls_matrix_s* m, int level, const double* psi) {
    DevLevel& D = DL(m, level);
    MatLevel& M = m->levels[level];
    Context& c = ctx();
    if (D.nIfaces == 0) return;
    if (D.anyProcIface && c.nRanks == 1)
        throw CudaError("matrix has processor interfaces but b200ls_init was called with nRanks=1");
    if (M.p2pReady) {
        // remote stores into the neighbours' receive buffers + release of the epoch flag, from our own kernel
        const unsigned long long epoch = ++M.haloEpoch;
        for (int i = 0; i < D.nIfaces; i++) {
            const int n = D.ifaceSize[i];
            if (D.ifacePartner[i] >= 0) {
                if (n) LAUNCH(k_iface_pack, gridRows(n), 256, M.sendBuf[i].p, psi, D.ifaceCellsPos[i].p, n);
                continue;
            }
            LAUNCH(k_iface_pack_p2p, gridRows(std::max(n, 1)), 256, M.p2pRemoteRecv[i] + (epoch & 1) * n, psi,
                   D.ifaceCellsPos[i].p, n, M.p2pTickets.p + i, M.p2pRemoteFlag[i], epoch);
        }
        return;
    }
    for (int i = 0; i < D.nIfaces; i++) {
        if (D.ifaceSize[i])
            LAUNCH(k_iface_pack, gridRows(D.ifaceSize[i]), 256, M.sendBuf[i].p, psi, D.ifaceCellsPos[i].p,
                   D.ifaceSize[i]);
    }
    if (!D.anyProcIface) return;
    c.nccl.GroupStart();
    for (int i = 0; i < D.nIfaces; i++) {
        if (D.ifacePartner[i] >= 0) continue;
        c.nccl.Send(M.sendBuf[i].p, D.ifaceSize[i], ncclDouble, D.ifaceNbr[i], c.comm, c.stream);
        c.nccl.Recv(M.recvBuf[i].p, D.ifaceSize[i], ncclDouble, D.ifaceNbr[i], c.comm, c.stream);
    }
    int r = c.nccl.GroupEnd();
    if (r != 0) throw CudaError(std::string("nccl halo exchange: ") + c.nccl.GetErrorString((ncclResult_t)r));
}

static void ifaceApply(b200ls_matrix_s* m, int level, double* result, double sign) {
    DevLevel& D = DL(m, level);
    MatLevel& M = m->levels[level];
    if (D.nBRows == 0) return;
    if (D.nIfaces > 256) throw CudaError("more than 256 coupled patches on one rank");
    LAUNCH(k_iface_apply, gridRows(D.nBRows), 256, result, D.bRowPos.p, D.bRowPtr.p, D.bEntIface.p, D.bEntFace.p,
           reinterpret_cast<const IfaceView*>(M.ifaceViews.p), D.nIfaces, M.haloEpoch, sign, D.nBRows,
           ctx().errFlag.p);
}

// ------------------------------------------------------------------------------------------------------------
// SpMV family
// ------------------------------------------------------------------------------------------------------------

// Non-null while a Krylov loop enqueues iterations ahead of the host's convergence read-back: the kernels of such an
// iteration return at once when the word it points to is set (solvePCG).
static const int* g_stop = nullptr;

static bool usePencil(const DevLevel& D);
static void ensurePencilPlanes(b200ls_matrix_s* m, int level);
static size_t planeStride(const DevLevel& D);

// Amul on a pencil level (tile-major structured block): the 7-point stencil kernel on the coefficient planes, optionally
// fused with the dot product wA.x (dotOut != nullptr)
static void pencilSpmv(b200ls_matrix_s* m, int level, double* out, const double* x, double* dotOut) {
    DevLevel& D = DL(m, level);
    MatLevel& M = m->levels[level];
    Context& c = ctx();
    ensurePencilPlanes(m, level);
    const size_t np = planeStride(D);
    const int chunks = std::max(1, (D.pNx * 32 + 255) / 256);
    const int grid = std::max(1, std::min(D.nPencilTiles * chunks, kMaxPartials / 2));
#define B2_PSPMV(SYM, DOT)                                                                                            \
    LAUNCH((k_pencil_spmv<SYM, DOT>), grid, 256, out, x, M.diag.p, M.pcL.p, M.pcU.p, np, D.pTiles.p, D.nPencilTiles, \
           D.pNx, chunks, dotOut, c.partials.p, c.ticket.p, g_stop)
    if (m->symmetric) {
        if (dotOut) B2_PSPMV(true, true);
        else B2_PSPMV(true, false);
    } else {
        if (dotOut) B2_PSPMV(false, true);
        else B2_PSPMV(false, false);
    }
#undef B2_PSPMV
}

template <int MODE>
static void spmv(b200ls_matrix_s* m, int level, double* out, double* out2, const double* x, const double* b) {
    DevLevel& D = DL(m, level);
    MatLevel& M = m->levels[level];
    if (D.nCells == 0) return;
    static const bool noPencilSpmv = getenv("B200LS_NO_PENCIL_SPMV") != nullptr;
    if (MODE == SPMV_AMUL && !noPencilSpmv && usePencil(D)) {
        pencilSpmv(m, level, out, x, nullptr);
        return;
    }
    static const bool noSym = getenv("B200LS_NO_SYM_SPMV") != nullptr;
    static const bool x2 = getenv("B200LS_NO_SPMV_X2") == nullptr;   // two rows per thread: +7 % at 256^3
    if (MODE == SPMV_AMUL && m->symmetric && D.hasLslot && !noSym && x2) {
        LAUNCH(k_spmv_sym_x2, std::max(1, (D.nCells + 511) / 512), 256, out, x, M.diag.p, D.Lptr.p, D.Lcol.p, D.Lslot.p,
               D.Uptr.p, D.Ucol.p, M.Uval(), D.nCells);
        return;
    }
    if (MODE != SPMV_SUMA && m->symmetric && D.hasLslot && !noSym) {
        LAUNCH(k_spmv_sym<MODE>, gridRows(D.nCells), 256, out, out2, x, b, M.diag.p, D.Lptr.p, D.Lcol.p, D.Lslot.p,
               D.Uptr.p, D.Ucol.p, M.Uval(), D.nCells);
        return;
    }
    LAUNCH(k_spmv<MODE>, gridRows(D.nCells), 256, out, out2, x, b, M.diag.p, D.Lptr.p, D.Lcol.p, M.Lval(D.nFaces),
           D.Uptr.p, D.Ucol.p, M.Uval(), D.nCells);
}

void opAmul(b200ls_matrix_s* m, int level, double* out, const double* x) {
    haloExchange(m, level, x);
    spmv<SPMV_AMUL>(m, level, out, nullptr, x, nullptr);
    ifaceApply(m, level, out, 1.0);
}

void opResidual(b200ls_matrix_s* m, int level, double* out, const double* x, const double* b) {
    haloExchange(m, level, x);
    spmv<SPMV_RESIDUAL>(m, level, out, nullptr, x, b);
    ifaceApply(m, level, out, -1.0);
}

// wA = A psi ; rA = source - wA   (PCG.C:93-96)
static void opAmulAndResidual(b200ls_matrix_s* m, int level, double* wA, double* rA, const double* x,
                              const double* b) {
    DevLevel& D = DL(m, level);
    if (D.nIfaces == 0) {
        spmv<SPMV_AMUL_AND_RESIDUAL>(m, level, wA, rA, x, b);
    } else {
        opAmul(m, level, wA, x);
        LAUNCH(k_sub, gridStride(D.nCells), 256, rA, b, wA, D.nCells);
    }
}

void opSumA(b200ls_matrix_s* m, int level, double* out) {
    DevLevel& D = DL(m, level);
    MatLevel& M = m->levels[level];
    spmv<SPMV_SUMA>(m, level, out, nullptr, nullptr, nullptr);
    if (D.nBRows)
        LAUNCH(k_iface_suma, gridRows(D.nBRows), 256, out, D.bRowPos.p, D.bRowPtr.p, D.bEntIface.p, D.bEntFace.p,
               reinterpret_cast<const IfaceView*>(M.ifaceViews.p), D.nBRows);
}

// ------------------------------------------------------------------------------------------------------------
// DIC / DILU
// ------------------------------------------------------------------------------------------------------------

static void fillSentinel(double* p, int n) {
    static const bool noWait = getenv("B200LS_DEBUG_NO_WAIT") != nullptr;   // profiling aid: dependencies pre-satisfied
    if (noWait) {
        if (n) LAUNCH(k_fill, gridStride(n), 256, p, 1.0, n);
        return;
    }
    if (n) LAUNCH(k_fill_sentinel, gridStride(n), 256, p, n);
}

static void ensureLevelScratch(b200ls_matrix_s* m, int level) {
    DevLevel& D = DL(m, level);
    MatLevel& M = m->levels[level];
    if (M.tmpA.n != size_t(D.nCells) || (!M.tmpA.p && D.nCells)) {
        M.tmpA.alloc(D.nCells);
        M.tmpB.alloc(D.nCells);
        M.tmpC.alloc(D.nCells);
        M.tmpASentinel = false;
    }
}

// ---- pencil levels (structured blocks, pencil.cuh) ----

static bool usePencil(const DevLevel& D) {
    if (!D.hasPencil) return false;
    const char* e = getenv("B200LS_PENCIL_SWEEPS");   // "0": run the wavefront kernels on the tile-major layout
    if (e && e[0] == '0') return false;
    int skew, stages;
    return pencilConfig(D.pSkewUnits, true, skew, stages);
}
// planes are [slot][position] with an even stride (16-byte aligned rows for the bulk copies)
static size_t planeStride(const DevLevel& D) { return (size_t(D.nCells) + 1) & ~size_t(1); }
static dim3 pencilGrid(const DevLevel& D) { return dim3(unsigned(std::max(1, (D.pNx * 32 + 255) / 256)), unsigned(D.nPencilTiles)); }

static PencilArgs pencilArgs(const DevLevel& D, bool gs) {
    PencilArgs a{};
    a.tiles = D.pTiles.p;
    a.order = D.pOrder.p;
    a.nTiles = D.nPencilTiles;
    a.nx = D.pNx;
    a.ny = D.pNy;
    a.nz = D.pNz;
    a.extW = (D.pWJ + D.pWK) * (gs ? 2 : 1);
    return a;
}

static void ensurePencilPlanes(b200ls_matrix_s* m, int level) {
    DevLevel& D = DL(m, level);
    MatLevel& M = m->levels[level];
    if (M.pPlanesValid) return;
    const size_t np = planeStride(D);
    M.pcL.alloc(3 * np);
    M.pcU.alloc(3 * np);
    if (!m->symmetric) M.pcLu.alloc(3 * np);
    LAUNCH(k_pencil_planes, pencilGrid(D), 256, M.pcL.p, m->symmetric ? nullptr : M.pcLu.p, M.pcU.p, D.pTiles.p, D.pNx,
           D.pNy, D.pNz, np, D.Lptr.p, M.Lval(D.nFaces), D.LtoU.p, D.Uptr.p, M.Uval());
    M.pPlanesValid = true;
}

void ensureFactor(b200ls_matrix_s* m, int level, int precond) {
    DevLevel& D = DL(m, level);
    MatLevel& M = m->levels[level];
    const int kind = (precond == B200LS_DIAGONAL) ? B200LS_DIAGONAL : B200LS_DIC;   // DIC and DILU share rD
    if (M.rDValid && M.rDKind == kind) return;
    M.rDKind = kind;
    M.rD.alloc(D.nCells);
    if (precond == B200LS_DIAGONAL) {
        // rD = 1/diag (diagonalPreconditioner.C:59-62)
        if (D.nCells) {
            LAUNCH(k_fill, gridStride(D.nCells), 256, M.rD.p, 1.0, D.nCells);
            LAUNCH(k_div, gridStride(D.nCells), 256, M.rD.p, M.rD.p, M.diag.p, D.nCells);
        }
        M.rDValid = true;
        return;
    }
    M.dWork.alloc(D.nCells);
    fillSentinel(M.dWork.p, D.nCells);
    if (usePencil(D)) {
        // structured block: pencil factorisation, then the pre-multiplied substitution coefficients
        ensurePencilPlanes(m, level);
        const size_t np = planeStride(D);
        PencilArgs a = pencilArgs(D, false);
        a.plane[0] = M.diag.p;
        for (int q = 0; q < 3; q++) {
            a.plane[1 + q] = M.pcL.p + q * np;
            a.plane[4 + q] = m->symmetric ? nullptr : M.pcLu.p + q * np;
        }
        a.out = M.dWork.p;
        a.out2 = M.rD.p;
        launchPencil<PM_FACTOR>(a, D);
        M.ptL.alloc(3 * np);
        M.ptU.alloc(3 * np);
        LAUNCH(k_pencil_pack, pencilGrid(D), 256, M.ptL.p, M.ptU.p, M.pcL.p, M.pcU.p, M.rD.p, D.pTiles.p, D.pNx, D.pNy,
               D.pNz, np);
        M.rDValid = true;
        return;
    }
    SweepArgs a{};
    a.tasks = D.fwdTasks.p;
    a.nTasks = D.nFwdTasks;
    a.rowOf = D.fwdPos.p;
    a.ptr = D.Lptr.p;
    a.col = D.Lcol.p;
    a.val = M.Lval(D.nFaces);
    a.col2 = m->symmetric ? nullptr : D.LtoU.p;
    a.val2 = M.Uval();
    a.diag = M.diag.p;
    a.out = M.dWork.p;
    a.out2 = M.rD.p;
    launchSweep(k_factor, a);
    M.rDValid = true;
}

// wA = M^-1 rA.  DIC/DILU: forward sweep into the level's sentinel scratch, backward sweep into wA; each sweep
// re-arms the other's output buffer, so no fill kernels run in steady state.
void opPrecondition(b200ls_matrix_s* m, int level, int precond, double* wA, const double* rA, double* dotOut) {
    DevLevel& D = DL(m, level);
    MatLevel& M = m->levels[level];
    const int n = D.nCells;
    if (n == 0) {
        // a rank without rows on this level still owns a term of the global sum
        if (dotOut) B2_CUDA(cudaMemsetAsync(dotOut, 0, sizeof(double), S()));
        return;
    }
    if (precond == B200LS_NONE) {
        B2_CUDA(cudaMemcpyAsync(wA, rA, sizeof(double) * n, cudaMemcpyDeviceToDevice, S()));
        return;
    }
    ensureFactor(m, level, precond);
    if (precond == B200LS_DIAGONAL) {
        // wA = rD*rA  (diagonalPreconditioner.C:80-83); rD holds 1/diag
        LAUNCH(k_mul, gridStride(n), 256, wA, M.rD.p, rA, n);
        return;
    }
    if (precond != B200LS_DIC && precond != B200LS_DILU) throw CudaError("unknown preconditioner");
    ensureLevelScratch(m, level);
    if (!M.tmpASentinel) {
        fillSentinel(M.tmpA.p, n);
        M.tmpASentinel = true;
    }
    if (usePencil(D)) {
        const size_t np = planeStride(D);
        PencilArgs f = pencilArgs(D, false);
        f.plane[0] = rA;
        f.plane[1] = M.rD.p;
        for (int q = 0; q < 3; q++) f.plane[2 + q] = M.ptL.p + q * np;
        f.out = M.tmpA.p;
        f.clear = wA;
        f.stop = g_stop;
        launchPencil<PM_FWD>(f, D);
        static const bool fwdOnly = getenv("B200LS_DEBUG_FWD_ONLY") != nullptr;
        if (fwdOnly) {   // debugging aid (benchmarks/pencil_group_fwd.py): return the forward sweep's result
            B2_CUDA(cudaMemcpyAsync(wA, M.tmpA.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, S()));
            M.tmpASentinel = false;
            return;
        }
        PencilArgs b = pencilArgs(D, false);
        b.plane[0] = M.tmpA.p;
        for (int q = 0; q < 3; q++) b.plane[1 + q] = M.ptU.p + q * np;
        b.out = wA;
        b.clear = M.tmpA.p;
        b.stop = g_stop;
        if (dotOut) {   // fused wA.rA
            b.plane[4] = rA;
            b.dotOut = dotOut;
        }
        launchPencil<PM_BWD>(b, D);
        return;
    }
    SweepArgs f{};
    f.tasks = D.fwdTasks.p;
    f.nTasks = D.nFwdTasks;
    f.rowOf = D.fwdPos.p;
    f.ptr = D.Lptr.p;
    f.col = D.Lcol.p;
    f.val = M.Lval(D.nFaces);
    f.rD = M.rD.p;
    f.in = rA;
    f.out = M.tmpA.p;
    f.clear = wA;
    f.stop = g_stop;
    launchSweep(k_sweep_fwd, f);

    SweepArgs b{};
    b.tasks = D.bwdTasks.p;
    b.nTasks = D.nBwdTasks;
    b.rowOf = D.bwdPos.p;
    b.ptr = D.Uptr.p;
    b.col = D.Ucol.p;
    b.val = M.Uval();
    b.rD = M.rD.p;
    b.in = M.tmpA.p;
    b.out = wA;
    b.clear = M.tmpA.p;
    b.stop = g_stop;
    if (dotOut) {   // fused wA.rA
        b.dotWith = rA;
        b.dotOut = dotOut;
    }
    launchSweep(k_sweep_bwd, b);
}

// ------------------------------------------------------------------------------------------------------------
// smoothers
// ------------------------------------------------------------------------------------------------------------

// psi/spare are swapped by Gauss-Seidel sweeps: on return `psi` holds the result.
void opSmooth(b200ls_matrix_s* m, int level, int smoother, double*& psi, double*& spare, const double* source,
              int nSweeps) {
    DevLevel& D = DL(m, level);
    MatLevel& M = m->levels[level];
    const int n = D.nCells;
    if (n == 0) return;
    ensureLevelScratch(m, level);
    static const bool noFusedGS = getenv("B200LS_NO_FUSED_GS") != nullptr;
    // lag between consecutive sweeps: an upper neighbour may sit up to maxFwdSpan wavefronts ahead, and its previous
    // sweep must come earlier in the task order (deadlock freedom), so sweep s starts lag = maxFwdSpan + 1
    // wavefronts after sweep s-1 (2 on structured blocks)
    const int lag = D.maxFwdSpan + 1;
    // symGaussSeidel stays on the general wavefront kernels (they run on the tile-major layout through the
    // processing-order map): its pencil variant (k_pencil<PM_GS_REV>) gave a result that was not bit-identical in 2 of
    // ~15 test-suite runs on one shape ((150, 9, 5)) and in none of 1,400 stress repetitions of the same calls -- an
    // unexplained, context-dependent race; B200LS_PENCIL_SYMGS=1 turns it back on.
    const bool symGsPencil = getenv("B200LS_PENCIL_SYMGS") && atoi(getenv("B200LS_PENCIL_SYMGS")) > 0;
    const bool pencil = usePencil(D) && (smoother != B200LS_SYM_GAUSS_SEIDEL || symGsPencil);
    if (pencil) ensurePencilPlanes(m, level);
    const bool coupledFused = !pencil && smoother == B200LS_GAUSS_SEIDEL && D.nIfaces > 0 && M.gsCoupled &&
                              nSweeps >= 2 && nSweeps <= 1 + kCoupledSlotSweeps && !noFusedGS;
    const bool localFused = !pencil && smoother == B200LS_GAUSS_SEIDEL && D.nIfaces == 0 && nSweeps >= 2 &&
                            nSweeps <= kMaxFusedSweeps && !noFusedGS && 2 * lag <= D.nFwdLevels;
    if (coupledFused || localFused) {
        // all sweeps in one pipelined launch (k_gs_multi): chain of nLevels + lag*(nSweeps-1) hops instead of
        // nSweeps*nLevels.  Levels without processor patches need no communication between sweeps; with patches the
        // neighbours' intermediate values travel through P2P slots inside the kernel (setupCoupledGS).
        const int useLag = coupledFused ? M.gsLag : lag;
        const int key = coupledFused ? 1000 + nSweeps : nSweeps;
        auto it = D.multiSweepTasks.find(key);
        if (it == D.multiSweepTasks.end()) {
            // bucket the (task, sweep) pairs by tau = level + lag*sweep
            const int nT = int(D.hostFwdTasks.size());
            std::vector<int> levelStart(D.nFwdLevels + 1, nT);
            for (int t = nT - 1; t >= 0; t--) levelStart[D.hostFwdTaskLevel[t]] = t;
            for (int l = D.nFwdLevels - 1; l >= 0; l--)
                if (levelStart[l] == nT) levelStart[l] = levelStart[l + 1];
            std::vector<int4> fused;
            fused.reserve(size_t(nT) * nSweeps);
            const int maxTau = D.nFwdLevels - 1 + useLag * (nSweeps - 1);
            for (int tau = 0; tau <= maxTau; tau++) {
                for (int sw = 0; sw < nSweeps; sw++) {
                    const int l = tau - useLag * sw;
                    if (l < 0 || l >= D.nFwdLevels) continue;
                    for (int t = levelStart[l]; t < levelStart[l + 1]; t++)
                        fused.push_back(make_int4(D.hostFwdTasks[t].x, D.hostFwdTasks[t].y, sw, 0));
                }
            }
            DevBuf<int4>& buf = D.multiSweepTasks[key];
            buf.alloc(fused.size());
            if (!fused.empty())
                B2_CUDA(cudaMemcpyAsync(buf.p, fused.data(), fused.size() * sizeof(int4), cudaMemcpyHostToDevice, S()));
            B2_CUDA(cudaStreamSynchronize(S()));
            it = D.multiSweepTasks.find(key);
        }
        if (M.gsBufs.n < size_t(nSweeps - 1) * n) M.gsBufs.alloc(size_t(nSweeps - 1) * n);
        fillSentinel(M.gsBufs.p, (nSweeps - 1) * n);
        fillSentinel(spare, n);
        MultiSweepArgs a{};
        a.tasks = it->second.p;
        a.nTasks = int(it->second.n);
        a.rowOf = D.fwdPos.p;
        a.Lptr = D.Lptr.p;
        a.Lcol = D.Lcol.p;
        a.Lval = M.Lval(D.nFaces);
        a.Uptr = D.Uptr.p;
        a.Ucol = D.Ucol.p;
        a.Uval = M.Uval();
        a.diag = M.diag.p;
        a.b = source;
        a.X[0] = psi;
        for (int sw = 1; sw < nSweeps; sw++) a.X[sw] = M.gsBufs.p + size_t(sw - 1) * n;
        a.X[nSweeps] = spare;
        a.err = ctx().errFlag.p;
        a.nSweeps = nSweeps;
        if (coupledFused) {
            // sweep 0: bPrime = source + bouCoeffs*psi_nbr through the regular halo exchange (which also orders this
            // call after the neighbours' previous one: the slot parity protocol relies on it)
            B2_CUDA(cudaMemcpyAsync(M.tmpB.p, source, sizeof(double) * n, cudaMemcpyDeviceToDevice, S()));
            haloExchange(m, level, psi);
            ifaceApply(m, level, M.tmpB.p, -1.0);
            a.b0 = M.tmpB.p;
            a.bRowOf = D.bRowOf.p;
            a.bRowPtr = D.bRowPtr.p;
            a.bEntIface = D.bEntIface.p;
            a.bEntFace = D.bEntFace.p;
            a.views = reinterpret_cast<const CoupledView*>(M.gsViews[(M.gsEpoch++) & 1].p);
        }
        {
            Context& c = ctx();
            static int occ[2] = {0, 0};
            const void* fn = coupledFused ? (const void*)k_gs_multi<true> : (const void*)k_gs_multi<false>;
            int& oc = occ[coupledFused ? 1 : 0];
            if (!oc) B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc, fn, 256, 0));
            const int perSM = std::min(oc, c.sweepBlocksPerSM > 0 ? c.sweepBlocksPerSM : 4);
            const int blocks = std::max(1, std::min(perSM * c.numSMs, (a.nTasks + 7) / 8));
            void* args[] = {&a};
            B2_CUDA(cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(256), args, 0, c.stream));
            c.launches++;
        }
        std::swap(psi, spare);
        return;
    }
    if (smoother == B200LS_GAUSS_SEIDEL || smoother == B200LS_SYM_GAUSS_SEIDEL) {
        for (int sweep = 0; sweep < nSweeps; sweep++) {
            const double* bPrime = source;
            if (D.nIfaces) {
                // bPrime = source + bouCoeffs*psi_nbr (negated coefficients, GaussSeidelSmoother.C:110-145)
                B2_CUDA(cudaMemcpyAsync(M.tmpB.p, source, sizeof(double) * n, cudaMemcpyDeviceToDevice, S()));
                haloExchange(m, level, psi);
                ifaceApply(m, level, M.tmpB.p, -1.0);
                bPrime = M.tmpB.p;
            }
            // plain Gauss-Seidel re-arms the old iterate while it sweeps (every reader of old[p] sits in an
            // earlier wavefront than row p), so only the first sweep of a call needs a fill kernel
            const bool rearm = (smoother == B200LS_GAUSS_SEIDEL);
            if (!(rearm && sweep > 0)) fillSentinel(spare, n);
            if (pencil) {
                const size_t np = planeStride(D);
                PencilArgs a = pencilArgs(D, true);
                a.plane[0] = bPrime;
                a.plane[1] = M.diag.p;
                for (int q = 0; q < 3; q++) {
                    a.plane[2 + q] = M.pcL.p + q * np;
                    a.plane[5 + q] = M.pcU.p + q * np;
                }
                a.plane[8] = psi;
                a.out = spare;
                a.clear = (rearm && sweep + 1 < nSweeps) ? psi : nullptr;
                launchPencil<PM_GS_FWD>(a, D);
                if (smoother == B200LS_GAUSS_SEIDEL) {
                    std::swap(psi, spare);
                    continue;
                }
                // symGaussSeidel: reverse loop over the forward result (symGaussSeidelSmoother.C:178-205)
                fillSentinel(psi, n);
                PencilArgs r = a;
                r.plane[8] = spare;
                r.out = psi;
                r.clear = nullptr;
                launchPencil<PM_GS_REV>(r, D);
                continue;
            }
            SweepArgs a{};
            a.tasks = D.fwdTasks.p;
            a.nTasks = D.nFwdTasks;
            a.rowOf = D.fwdPos.p;
            a.ptr = D.Lptr.p;
            a.col = D.Lcol.p;
            a.val = M.Lval(D.nFaces);
            a.ptr2 = D.Uptr.p;
            a.col2 = D.Ucol.p;
            a.val2 = M.Uval();
            a.diag = M.diag.p;
            a.in = bPrime;
            a.old = psi;
            a.out = spare;
            a.clear = (rearm && sweep + 1 < nSweeps) ? psi : nullptr;
            launchSweep(k_gs_sweep, a);
            if (smoother == B200LS_GAUSS_SEIDEL) {
                std::swap(psi, spare);
                continue;
            }
            // symGaussSeidel: reverse loop over the forward result (symGaussSeidelSmoother.C:178-205)
            fillSentinel(psi, n);
            SweepArgs r{};
            r.tasks = D.bwdTasks.p;
            r.nTasks = D.nBwdTasks;
            r.rowOf = D.bwdPos.p;
            r.ptr = D.Uptr.p;
            r.col = D.Ucol.p;
            r.val = M.Uval();
            r.ptr2 = D.Lptr.p;
            r.col2 = D.Lcol.p;
            r.val2 = M.Lval(D.nFaces);
            r.diag = M.diag.p;
            r.in = bPrime;
            r.old = spare;
            r.out = psi;
            launchSweep(k_gs_sweep_rev, r);
        }
        return;
    }
    if (smoother == B200LS_DIC || smoother == B200LS_DILU) {
        // DICSmoother.C:84-115 / DILUSmoother.C: rA = residual ; rA = M^-1 rA ; psi += rA
        ensureFactor(m, level, smoother);
        for (int sweep = 0; sweep < nSweeps; sweep++) {
            opResidual(m, level, M.tmpB.p, psi, source);
            opPrecondition(m, level, smoother, M.tmpC.p, M.tmpB.p);
            LAUNCH(k_add_inplace, gridStride(n), 256, psi, M.tmpC.p, n);
        }
        return;
    }
    if (smoother == B200LS_DIC_GAUSS_SEIDEL || smoother == B200LS_DILU_GAUSS_SEIDEL) {
        // DICGaussSeidelSmoother.C:79-89: nSweeps of DIC (DILU) followed by nSweeps of Gauss-Seidel
        opSmooth(m, level, smoother == B200LS_DIC_GAUSS_SEIDEL ? B200LS_DIC : B200LS_DILU, psi, spare, source, nSweeps);
        opSmooth(m, level, B200LS_GAUSS_SEIDEL, psi, spare, source, nSweeps);
        return;
    }
    throw CudaError("unknown smoother");
}

// ------------------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------------------

template <int OP>
static void reduce(double* out, const double* x, const double* y, int n) {
    Context& c = ctx();
    LAUNCH(k_reduce<OP>, kReduceBlocks, kReduceThreads, out, x, y, n, c.partials.p, c.ticket.p);
}

enum {
    S_SUMPSI = 0, S_NORM = 1, S_RES = 2, S_WARA0 = 3, S_WARA1 = 4, S_WAPA = 5,
    S_RHO0 = 6, S_RHO1 = 7, S_RA0AYA = 8, S_ALPHA = 9, S_OMEGA = 10, S_TASA = 11, S_TATA = 12,
    S_NUM = 13, S_DEN = 14, S_SINGULAR = 15, S_COUNT = 32
};

// Two banks: the coarsest-level Krylov solve nested inside a V-cycle (bank 1) must not clobber the scalars the
// enclosing solver keeps on the device across its iterations (e.g. PCG's previous wArA when GAMG is its preconditioner).
static int g_scalarBank = 0;
static double* scalar(b200ls_matrix_s* m, int i) {
    if (m->scalars.n != 2 * S_COUNT) {
        m->scalars.alloc(2 * S_COUNT);
        B2_CUDA(cudaMemsetAsync(m->scalars.p, 0, 2 * S_COUNT * sizeof(double), S()));
    }
    return m->scalars.p + g_scalarBank * S_COUNT + i;
}

// normFactor (lduMatrixSolver.C:174-197). tmp is clobbered.
static double normFactor(b200ls_matrix_s* m, int lv, const double* psi, const double* source, const double* Apsi,
                         double* tmp, int64_t nGlobalCells) {
    Context& c = ctx();
    DevLevel& D = DL(m, lv);
    const int n = D.nCells;
    opSumA(m, lv, tmp);
    reduce<RED_SUM>(scalar(m, S_SUMPSI), psi, nullptr, n);
    allReduce(scalar(m, S_SUMPSI), 1);
    LAUNCH(k_norm_factor, kReduceBlocks, kReduceThreads, scalar(m, S_NORM), Apsi, source, tmp, scalar(m, S_SUMPSI),
           double(nGlobalCells), n, c.partials.p, c.ticket.p);
    allReduce(scalar(m, S_NORM), 1);
    readScalars(scalar(m, S_NORM), 1);
    return c.pinned[0] + 1e-20;   // solverPerformance::small_
}

static bool converged(double finalRes, double initRes, double tol, double relTol) {
    // SolverPerformance.C:75-82
    return finalRes < tol || (relTol > 1e-20 && finalRes < relTol * initRes);
}

// SolverPerformance::checkConvergence (SolverPerformance.C:60-92) stores its result: the flag a solve reports is the
// one of the last call the loop conditions actually made (at maxIter the `&&` skips the call).
static bool checkConvergence(b200ls_perf* perf, const b200ls_controls& c) {
    perf->converged = converged(perf->finalResidual, perf->initialResidual, c.tolerance, c.relTol) ? 1 : 0;
    return perf->converged != 0;
}

// named work vector of a level (level 0 vectors are the solver's finest-level vectors)
static double* lvec(b200ls_matrix_s* m, int lv, const char* name) {
    if (lv == 0) return m->vec(name);
    Vec& v = m->vecs[std::string(name) + "@" + std::to_string(lv)];
    const size_t n = DL(m, lv).nCells;
    if (v.buf.n != n || !v.buf.p) v.buf.alloc(n);
    return v.buf.p;
}

static int64_t globalCells(b200ls_matrix_s* m, int lv = 0) {
    Context& c = ctx();
    const int64_t n = DL(m, lv).nCells;
    if (c.nRanks == 1) return n;
    double* s = scalar(m, S_NUM);
    const double h = double(n);
    B2_CUDA(cudaMemcpyAsync(s, &h, sizeof(double), cudaMemcpyHostToDevice, S()));
    B2_CUDA(cudaStreamSynchronize(S()));
    allReduce(s, 1);
    readScalars(s, 1);
    return int64_t(c.pinned[0] + 0.5);
}

static const double kVSmall = 2.2250738585072014e-308;   // vSmall = DBL_MIN (doubleScalar.H:57)

static void record(b200ls_perf* perf, const b200ls_controls& c, double r) {
    if (c.recordHistory && perf->nHistory < B200LS_MAX_HISTORY) perf->history[perf->nHistory++] = r;
}

// preconditioner dispatch of the Krylov solvers: DIC/DILU/diagonal/none per level, or GAMG V-cycles (finest level)
static void gamgPrecondition(b200ls_matrix_s* m, const b200ls_controls& c, double* wA, const double* rA);
static void buildCoarseMatrices(b200ls_matrix_s* m);
static void gatherCoarsestCoefs(b200ls_matrix_s* m);

static void preparePrecond(b200ls_matrix_s* m, const b200ls_controls& c, int lv) {
    if (c.precond == B200LS_GAMG_PRECOND) {
        if (lv != 0) throw CudaError("GAMG preconditioning is only available on the finest level");
        if (m->levels.size() < 2) throw CudaError("preconditioner GAMG needs an agglomerated mesh");
        buildCoarseMatrices(m);
        return;
    }
    ensureFactor(m, lv, c.precond);
}

static void applyPrecond(b200ls_matrix_s* m, const b200ls_controls& c, int lv, double* wA, const double* rA) {
    if (c.precond == B200LS_GAMG_PRECOND) gamgPrecondition(m, c, wA, rA);
    else opPrecondition(m, lv, c.precond, wA, rA);
}

// ------------------------------------------------------------------------------------------------------------
// PCG (PCG.C:65-193)
// ------------------------------------------------------------------------------------------------------------

static void solvePCG(b200ls_matrix_s* m, const b200ls_controls& c, int lv, double* psi, const double* source,
                     b200ls_perf* perf, cudaEvent_t evLoopStart) {
    Context& cx = ctx();
    DevLevel& D = DL(m, lv);
    MatLevel& M = m->levels[lv];
    const int n = D.nCells;
    double* pA = lvec(m, lv, "pA");
    double* wA = lvec(m, lv, "wA");
    double* rA = lvec(m, lv, "rA");

    opAmulAndResidual(m, lv, wA, rA, psi, source);
    const double nf = normFactor(m, lv, psi, source, wA, pA, globalCells(m, lv));
    perf->normFactor = nf;
    reduce<RED_SUMMAG>(scalar(m, S_RES), rA, nullptr, n);
    allReduce(scalar(m, S_RES), 1);
    readScalars(scalar(m, S_RES), 1);
    perf->initialResidual = cx.pinned[0] / nf;
    perf->finalResidual = perf->initialResidual;
    perf->nIterations = 0;

    if (evLoopStart) B2_CUDA(cudaEventRecord(evLoopStart, S()));
    if (c.minIter > 0 || !checkConvergence(perf, c)) {
        preparePrecond(m, c, lv);
        B2_CUDA(cudaMemsetAsync(scalar(m, S_SINGULAR), 0, sizeof(double), S()));
        // the dot products ride on the kernels that produce their operands when nothing sits in between
        const bool fuseSweepDot = (c.precond == B200LS_DIC || c.precond == B200LS_DILU);
        const bool fuseSpmvDot = (D.nIfaces == 0);
        const int spmvGrid = std::max(1, std::min(gridRows(n), kMaxPartials / 2));   // one row per thread when it fits
        static const bool noPencilSpmv = getenv("B200LS_NO_PENCIL_SPMV") != nullptr;
        static const bool noPipeline = getenv("B200LS_NO_PCG_PIPELINE") != nullptr;
        // One rank, no coupled patches: iteration k+1 is enqueued BEFORE the host has read the residual of iteration
        // k, so the read-back round trip and the launch latencies hide behind the kernels.  The loop condition is
        // evaluated on the device too (k_pcg_stop_flag) and every kernel of an iteration enqueued too far returns at
        // once; the host applies the same test to the same numbers one step later, so the iteration count, the
        // convergence flag and all fields are those of the synchronous loop.
        const bool pipelined = cx.nRanks == 1 && D.nIfaces == 0 && fuseSweepDot && lv == 0 && !noPipeline;
        const int* stop = pipelined ? cx.stopFlag.p : nullptr;
        auto enqueue = [&](int it) {
            const int cur = S_WARA0 + (it & 1);
            const int old = S_WARA0 + ((it + 1) & 1);
            if (fuseSweepDot) {
                opPrecondition(m, lv, c.precond, wA, rA, scalar(m, cur));
            } else {
                applyPrecond(m, c, lv, wA, rA);
                reduce<RED_DOT>(scalar(m, cur), wA, rA, n);
            }
            allReduce(scalar(m, cur), 1);
            LAUNCH(k_pcg_update_p, gridStride(n), 256, pA, wA, scalar(m, cur), scalar(m, old), it == 0 ? 1 : 0, n, stop);
            if (fuseSpmvDot && !noPencilSpmv && usePencil(D)) {
                pencilSpmv(m, lv, wA, pA, scalar(m, S_WAPA));
            } else if (fuseSpmvDot) {
                if (m->symmetric && D.hasLslot) {
                    LAUNCH(k_spmv_dot<true>, spmvGrid, 256, wA, pA, M.diag.p, D.Lptr.p, D.Lcol.p, M.Lval(D.nFaces),
                           D.Lslot.p, D.Uptr.p, D.Ucol.p, M.Uval(), n, scalar(m, S_WAPA), cx.partials.p, cx.ticket.p,
                           stop);
                } else {
                    LAUNCH(k_spmv_dot<false>, spmvGrid, 256, wA, pA, M.diag.p, D.Lptr.p, D.Lcol.p, M.Lval(D.nFaces),
                           D.Lslot.p, D.Uptr.p, D.Ucol.p, M.Uval(), n, scalar(m, S_WAPA), cx.partials.p, cx.ticket.p,
                           stop);
                }
            } else {
                opAmul(m, lv, wA, pA);
                reduce<RED_DOT>(scalar(m, S_WAPA), wA, pA, n);
            }
            allReduce(scalar(m, S_WAPA), 1);
            // the singularity test (PCG.C:165) is evaluated on the device by every thread of the update kernel
            LAUNCH(k_pcg_update_xr, kReduceBlocks, kReduceThreads, psi, rA, pA, wA, scalar(m, cur),
                   scalar(m, S_WAPA), nf, scalar(m, S_SINGULAR), scalar(m, S_RES), n, cx.partials.p, cx.ticket.p, stop);
            allReduce(scalar(m, S_RES), 1);
        };
        static_assert(S_SINGULAR - S_NUM == 2 && S_RES < S_NUM, "scalar layout");
        if (!pipelined) {
            do {
                enqueue(perf->nIterations);
                readScalars(scalar(m, 0), S_SINGULAR + 1);   // one read-back per iteration
                if (cx.pinned[S_SINGULAR] != 0.0) {
                    perf->singular = 1;
                    break;
                }
                perf->finalResidual = cx.pinned[S_RES] / nf;
                record(perf, c, perf->finalResidual);
            } while ((++perf->nIterations < c.maxIter &&
                      !checkConvergence(perf, c)) ||
                     perf->nIterations < c.minIter);
        } else {
            constexpr int kSlot = S_SINGULAR + 1;   // doubles per read-back slot of the pinned block
            static_assert(2 * kSlot <= 64, "pinned scalar block");
            struct StopGuard {
                cudaEvent_t ev[2] = {nullptr, nullptr};
                StopGuard() {
                    cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
                    cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
                }
                ~StopGuard() {
                    g_stop = nullptr;
                    cudaEventDestroy(ev[0]);
                    cudaEventDestroy(ev[1]);
                }
            } guard;
            B2_CUDA(cudaMemsetAsync(cx.stopFlag.p, 0, sizeof(int), S()));
            g_stop = stop;
            auto enqueueAll = [&](int it) {
                enqueue(it);
                LAUNCH(k_pcg_stop_flag, 1, 1, cx.stopFlag.p, scalar(m, S_RES), scalar(m, S_SINGULAR), nf, c.tolerance,
                       c.relTol, perf->initialResidual, it + 1 >= c.minIter ? 1 : 0);
                B2_CUDA(cudaMemcpyAsync(cx.pinned + kSlot * (it & 1), scalar(m, 0), kSlot * sizeof(double),
                                        cudaMemcpyDeviceToHost, S()));
                B2_CUDA(cudaEventRecord(guard.ev[it & 1], S()));
            };
            enqueueAll(0);
            for (;;) {
                const int it = perf->nIterations;
                if (it + 1 < c.maxIter || it + 1 < c.minIter) enqueueAll(it + 1);   // ahead of the read-back of `it`
                B2_CUDA(cudaEventSynchronize(guard.ev[it & 1]));
                const double* h = cx.pinned + kSlot * (it & 1);
                if (h[S_SINGULAR] != 0.0) {
                    perf->singular = 1;
                    break;
                }
                perf->finalResidual = h[S_RES] / nf;
                record(perf, c, perf->finalResidual);
                if (!((++perf->nIterations < c.maxIter && !checkConvergence(perf, c)) ||
                      perf->nIterations < c.minIter))
                    break;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// PBiCGStab (PBiCGStab.C:68-254)
// ------------------------------------------------------------------------------------------------------------

static void solvePBiCGStab(b200ls_matrix_s* m, const b200ls_controls& c, int lv, double* psi, const double* source,
                           b200ls_perf* perf, cudaEvent_t evLoopStart) {
    Context& cx = ctx();
    const int n = DL(m, lv).nCells;
    double* pA = lvec(m, lv, "pA");
    double* yA = lvec(m, lv, "yA");
    double* rA = lvec(m, lv, "rA");

    opAmulAndResidual(m, lv, yA, rA, psi, source);
    const double nf = normFactor(m, lv, psi, source, yA, pA, globalCells(m, lv));
    perf->normFactor = nf;
    reduce<RED_SUMMAG>(scalar(m, S_RES), rA, nullptr, n);
    allReduce(scalar(m, S_RES), 1);
    readScalars(scalar(m, S_RES), 1);
    perf->initialResidual = cx.pinned[0] / nf;
    perf->finalResidual = perf->initialResidual;
    perf->nIterations = 0;

    if (evLoopStart) B2_CUDA(cudaEventRecord(evLoopStart, S()));
    if (c.minIter > 0 || !checkConvergence(perf, c)) {
        double* AyA = lvec(m, lv, "AyA");
        double* sA = lvec(m, lv, "sA");
        double* zA = lvec(m, lv, "zA");
        double* tA = lvec(m, lv, "tA");
        double* rA0 = lvec(m, lv, "rA0");
        B2_CUDA(cudaMemcpyAsync(rA0, rA, sizeof(double) * n, cudaMemcpyDeviceToDevice, S()));
        preparePrecond(m, c, lv);
        double omega = 0;
        do {
            const int cur = S_RHO0 + (perf->nIterations & 1);
            const int old = S_RHO0 + ((perf->nIterations + 1) & 1);
            reduce<RED_DOT>(scalar(m, cur), rA0, rA, n);
            allReduce(scalar(m, cur), 1);
            readScalars(scalar(m, cur), 1);
            if (std::fabs(cx.pinned[0]) < kVSmall) {
                perf->singular = 1;
                break;
            }
            if (perf->nIterations > 0 && std::fabs(omega) < kVSmall) {
                perf->singular = 1;
                break;
            }
            LAUNCH(k_bicg_update_p, gridStride(n), 256, pA, rA, AyA, scalar(m, 0), cur, old, S_ALPHA, S_OMEGA,
                   perf->nIterations == 0 ? 1 : 0, n);
            applyPrecond(m, c, lv, yA, pA);
            opAmul(m, lv, AyA, yA);
            reduce<RED_DOT>(scalar(m, S_RA0AYA), rA0, AyA, n);
            allReduce(scalar(m, S_RA0AYA), 1);
            LAUNCH(k_bicg_update_s, kReduceBlocks, kReduceThreads, sA, rA, AyA, scalar(m, cur), scalar(m, S_RA0AYA),
                   scalar(m, S_ALPHA), scalar(m, S_RES), n, cx.partials.p, cx.ticket.p);
            allReduce(scalar(m, S_RES), 1);
            readScalars(scalar(m, S_RES), 1);
            perf->finalResidual = cx.pinned[0] / nf;
            if (++perf->nIterations >= c.minIter &&
                checkConvergence(perf, c)) {
                LAUNCH(k_axpy_s, gridStride(n), 256, psi, yA, scalar(m, S_ALPHA), n);
                record(perf, c, perf->finalResidual);
                perf->converged = 1;
                return;
            }
            applyPrecond(m, c, lv, zA, sA);
            opAmul(m, lv, tA, zA);
            // tAsA and tAtA in one pass (and one 2-element allreduce)
            LAUNCH(k_dot2, kReduceBlocks, kReduceThreads, scalar(m, S_TASA), sA, tA, tA, n, cx.partials.p,
                   cx.ticket.p);
            allReduce(scalar(m, S_TASA), 2);
            LAUNCH(k_bicg_update_xr, kReduceBlocks, kReduceThreads, psi, rA, yA, zA, sA, tA, scalar(m, S_ALPHA),
                   scalar(m, S_TASA), scalar(m, S_OMEGA), scalar(m, S_RES), n, cx.partials.p, cx.ticket.p);
            allReduce(scalar(m, S_RES), 1);
            readScalars(scalar(m, S_OMEGA), 1);
            omega = cx.pinned[0];
            readScalars(scalar(m, S_RES), 1);
            perf->finalResidual = cx.pinned[0] / nf;
            record(perf, c, perf->finalResidual);
        } while ((perf->nIterations < c.maxIter &&
                  !checkConvergence(perf, c)) ||
                 perf->nIterations < c.minIter);
    }
}

// ------------------------------------------------------------------------------------------------------------
// smoothSolver (smoothSolver.C:77-193)
// ------------------------------------------------------------------------------------------------------------

static void solveSmooth(b200ls_matrix_s* m, const b200ls_controls& c, double*& psi, double*& spare,
                        const double* source, b200ls_perf* perf, cudaEvent_t evLoopStart) {
    Context& cx = ctx();
    const int n = DL(m, 0).nCells;
    perf->nIterations = 0;
    if (c.nSweeps < 0) {
        B2_CUDA(cudaEventRecord(evLoopStart, S()));
        opSmooth(m, 0, c.precond, psi, spare, source, -c.nSweeps);
        perf->nIterations -= c.nSweeps;
        return;
    }
    double* Apsi = m->vec("wA");
    double* tmp = m->vec("pA");
    double* rA = m->vec("rA");
    opAmulAndResidual(m, 0, Apsi, rA, psi, source);
    const double nf = normFactor(m, 0, psi, source, Apsi, tmp, globalCells(m));
    perf->normFactor = nf;
    reduce<RED_SUMMAG>(scalar(m, S_RES), rA, nullptr, n);
    allReduce(scalar(m, S_RES), 1);
    readScalars(scalar(m, S_RES), 1);
    perf->initialResidual = cx.pinned[0] / nf;
    perf->finalResidual = perf->initialResidual;
    B2_CUDA(cudaEventRecord(evLoopStart, S()));
    if (c.minIter > 0 || !checkConvergence(perf, c)) {
        do {
            opSmooth(m, 0, c.precond, psi, spare, source, c.nSweeps);
            opResidual(m, 0, rA, psi, source);
            reduce<RED_SUMMAG>(scalar(m, S_RES), rA, nullptr, n);
            allReduce(scalar(m, S_RES), 1);
            readScalars(scalar(m, S_RES), 1);
            perf->finalResidual = cx.pinned[0] / nf;
            record(perf, c, perf->finalResidual);
        } while (((perf->nIterations += c.nSweeps) < c.maxIter &&
                  !checkConvergence(perf, c)) ||
                 perf->nIterations < c.minIter);
    }
}

// ------------------------------------------------------------------------------------------------------------
// GAMG
// ------------------------------------------------------------------------------------------------------------

// agglomerateMatrix for every level (GAMGSolver.C:196-208 -> GAMGSolverAgglomerateMatrix.C:33-193)
static void buildCoarseMatrices(b200ls_matrix_s* m) {
    if (m->coarseValid) return;
    const int nL = int(m->levels.size());
    for (int k = 0; k + 1 < nL; k++) {
        DevLevel& F = DL(m, k);
        DevLevel& C = DL(m, k + 1);
        MatLevel& MF = m->levels[k];
        MatLevel& MC = m->levels[k + 1];
        MC.diag.alloc(C.nCells);
        MC.vals.alloc(size_t(2) * C.nFaces);
        MC.rDValid = false;
        if (C.nCells)
            LAUNCH(k_agglomerate_diag, gridRows(C.nCells), 256, MC.diag.p, MF.diag.p, MF.Uval(), MF.Lval(F.nFaces),
                   F.rPtr.p, F.rFine.p, F.adPtr.p, F.adU.p, F.adL.p, C.nCells);
        if (C.nFaces) {
            LAUNCH(k_agglomerate_offdiag, gridRows(C.nFaces), 256, MC.Uval(), MF.vals.p, F.auPtr.p, F.auSrc.p,
                   C.nFaces);
            LAUNCH(k_agglomerate_offdiag, gridRows(C.nFaces), 256, MC.Lval(C.nFaces), MF.vals.p, F.alPtr.p,
                   F.alSrc.p, C.nFaces);
        }
        // interface coefficients: restrictField over patchFaceRestrictAddressing (:253-267)
        MC.bou.resize(C.nIfaces);
        MC.inn.resize(C.nIfaces);
        MC.sendBuf.resize(C.nIfaces);
        MC.recvBuf.resize(C.nIfaces);
        for (int i = 0; i < C.nIfaces; i++) {
            MC.bou[i].alloc(C.ifaceSize[i]);
            MC.inn[i].alloc(C.ifaceSize[i]);
            if (C.ifaceSize[i]) {
                LAUNCH(k_agglomerate_offdiag, gridRows(C.ifaceSize[i]), 256, MC.bou[i].p, MF.bou[i].p, F.aiPtr[i].p,
                       F.aiSrc[i].p, C.ifaceSize[i]);
                LAUNCH(k_agglomerate_offdiag, gridRows(C.ifaceSize[i]), 256, MC.inn[i].p, MF.inn[i].p, F.aiPtr[i].p,
                       F.aiSrc[i].p, C.ifaceSize[i]);
            }
        }
        setupIfaceViews(m, k + 1);
        MC.corr.alloc(C.nCells);
        MC.src.alloc(C.nCells);
    }
    gatherCoarsestCoefs(m);
    m->coarseValid = true;
}

// Coarsest-level solve on one thread, in the reference's exact sequential order (PCG+DIC for symmetric,
// PBiCGStab+DILU for asymmetric, diagonalSolver when the level has no faces; GAMGSolver.C:269-319,
// GAMGSolverSolve.C:520-552).  The level has ~10-20 cells: latency, not throughput.
struct CoarsestArgs {
    int nCells, nFaces, symmetric;
    const int *lower, *upper, *Uidx, *Lidx, *ipos;
    const double *diag, *Uval, *Lval;
    const double* source;   // position order
    double* psi;            // position order, overwritten (initial guess 0)
    double* work;           // 10*nCells + 2*nFaces doubles
    double tolerance, relTol;
    int maxIter;
    // gathered mode: the level of every rank, concatenated in rank order (CoarsestGather).  lower/upper then hold global
    // cell numbers for nFaces = all faces, nCells = all cells; coefficients and sources come as padded per-rank blocks.
    int gathered, nRanks, myRank, nCouple;
    int maxCells, maxFaces, blockLen;
    const int *cellOff, *faceOff, *coupleOff, *cRow, *cCol;
    const double *gCoef, *gSrc;
    const double* cc;       // coupling coefficients (filled by the kernel from gCoef)
};

// The kernel runs on one warp.  Lane r replays rank r of the distributed reference on block r of the gathered level
// (cells, faces and coupled-patch entries of that rank); a level that is not gathered is a single block on lane 0.
// Everything inside a block is sequential in the reference's order; blocks only meet in the sums (rank partials added
// in rank order, like a linear reduce) and in the coupled-patch update of Amul.
struct CoarsestBlock {
    int c0, c1, f0, f1, q0, q1;
};

template <class F>
__device__ static double c_sum(const CoarsestArgs& a, const CoarsestBlock& B, F term) {
    double s = 0.0;
    for (int c = B.c0; c < B.c1; c++) s += term(c);
    const int nB = a.gathered ? a.nRanks : 1;
    double total = __shfl_sync(0xffffffffu, s, 0);
    for (int r = 1; r < nB; r++) total += __shfl_sync(0xffffffffu, s, r);
    return total;
}

__device__ static void c_amul(const CoarsestArgs& a, const CoarsestBlock& B, const double* up, const double* lo,
                              const double* dg, double* out, const double* x) {
    for (int c = B.c0; c < B.c1; c++) out[c] = dg[c] * x[c];
    for (int f = B.f0; f < B.f1; f++) {
        out[a.upper[f]] += lo[f] * x[a.lower[f]];
        out[a.lower[f]] += up[f] * x[a.upper[f]];
    }
    // coupled patches after the faces, patch by patch (lduMatrixUpdateMatrixInterfaces.C, processorFvPatchField /
    // cyclicFvPatchField::updateInterfaceMatrix: result[faceCells] -= coeffs*psiNeighbour); x of the other blocks
    // was written by their lanes
    __syncwarp();
    for (int q = B.q0; q < B.q1; q++) out[a.cRow[q]] -= a.cc[q] * x[a.cCol[q]];
    __syncwarp();   // x may be overwritten by its owners from here on
}

__device__ static void c_precondition(const CoarsestArgs& a, const CoarsestBlock& B, const double* up,
                                      const double* lo, const double* rD, double* w, const double* r) {
    for (int c = B.c0; c < B.c1; c++) w[c] = rD[c] * r[c];
    // DILU walks the lower triangle in losort order, DIC in face order: both visit the faces of a row in ascending
    // order and every row only depends on finished rows (owner-sorted), so plain face order reproduces both
    for (int f = B.f0; f < B.f1; f++) w[a.upper[f]] -= rD[a.upper[f]] * lo[f] * w[a.lower[f]];
    for (int f = B.f1 - 1; f >= B.f0; f--) w[a.lower[f]] -= rD[a.lower[f]] * up[f] * w[a.upper[f]];
}

__device__ static bool c_converged(double fin, double ini, double tol, double relTol) {
    return fin < tol || (relTol > 1e-20 && fin < relTol * ini);
}

__global__ void __launch_bounds__(32) k_coarsest_solve(CoarsestArgs a_, int useSmem) {
    // When the level fits, its vectors, coefficients and addressing are staged in shared memory first (the loops are
    // chains of dependent loads).
    extern __shared__ double csm[];
    CoarsestArgs a = a_;
    const int n = a.nCells, nF = a.nFaces;
    const int nQ = a.nCouple;
    const int lane = threadIdx.x;
    if (useSmem) {
        int* sl = reinterpret_cast<int*>(csm + 2 * nF + 12 * n + nQ);
        int* su = sl + nF;
        int* sr = su + nF;
        int* sc = sr + nQ;
        int* so = sc + nQ;
        for (int f = lane; f < nF; f += 32) {
            sl[f] = a.lower[f];
            su[f] = a.upper[f];
        }
        for (int q = lane; q < nQ; q += 32) {
            sr[q] = a.cRow[q];
            sc[q] = a.cCol[q];
        }
        if (a.gathered)
            for (int r = lane; r <= a.nRanks; r += 32) {
                so[r] = a.cellOff[r];
                so[a.nRanks + 1 + r] = a.faceOff[r];
                so[2 * (a.nRanks + 1) + r] = a.coupleOff[r];
            }
        __syncwarp();
        a.lower = sl;
        a.upper = su;
        a.cRow = sr;
        a.cCol = sc;
        if (a.gathered) {
            a.cellOff = so;
            a.faceOff = so + a.nRanks + 1;
            a.coupleOff = so + 2 * (a.nRanks + 1);
        }
    }
    CoarsestBlock B = {0, 0, 0, 0, 0, 0};
    if (a.gathered) {
        if (lane < a.nRanks)
            B = {a.cellOff[lane], a.cellOff[lane + 1], a.faceOff[lane], a.faceOff[lane + 1], a.coupleOff[lane],
                 a.coupleOff[lane + 1]};
    } else if (lane == 0) {
        B = {0, n, 0, nF, 0, 0};
    }
    double* w = useSmem ? csm : a.work;
    double* up = w;            w += nF;
    double* lo = w;            w += nF;
    double* dg = w;            w += n;
    double* b = w;             w += n;
    double* x = w;             w += n;
    double* v0 = w;            w += n;
    double* v1 = w;            w += n;
    double* v2 = w;            w += n;
    double* rD = w;            w += n;
    double* v3 = w;            w += n;
    double* v4 = w;            w += n;
    double* v5 = w;            w += n;
    double* v6 = w;            w += n;
    double* v7 = w;            w += n;
    double* cc = w;            w += nQ;
    a.cc = cc;
    if (a.gathered) {
        // per-rank block of gCoef: [diag (maxCells) | upper (maxFaces) | lower (maxFaces) | coupling (maxCouple)]
        const double* blk = a.gCoef + size_t(lane) * a.blockLen;
        const double* sb = a.gSrc + size_t(lane) * a.maxCells;
        for (int c = B.c0; c < B.c1; c++) {
            dg[c] = blk[c - B.c0];
            b[c] = sb[c - B.c0];
            x[c] = 0.0;
        }
        for (int f = B.f0; f < B.f1; f++) {
            up[f] = blk[a.maxCells + f - B.f0];
            lo[f] = blk[a.maxCells + a.maxFaces + f - B.f0];
        }
        for (int q = B.q0; q < B.q1; q++) cc[q] = blk[a.maxCells + 2 * a.maxFaces + q - B.q0];
    } else {
        for (int f = B.f0; f < B.f1; f++) {
            up[f] = a.Uval[a.Uidx[f]];
            lo[f] = a.Lval[a.Lidx[f]];
        }
        for (int c = B.c0; c < B.c1; c++) {
            dg[c] = a.diag[a.ipos[c]];
            b[c] = a.source[a.ipos[c]];
            x[c] = 0.0;
        }
    }
    __syncwarp();
    const double vSmall = 2.2250738585072014e-308;

    if (nF == 0 && !a.gathered) {
        for (int c = B.c0; c < B.c1; c++) a.psi[a.ipos[c]] = b[c] / dg[c];   // diagonalSolver.C:66
        return;
    }

    // common prologue: Amul, residual, normFactor
    double* Ax = v0;
    double* r = v1;
    double* tmp = v2;
    c_amul(a, B, up, lo, dg, Ax, x);
    for (int c = B.c0; c < B.c1; c++) r[c] = b[c] - Ax[c];
    for (int c = B.c0; c < B.c1; c++) tmp[c] = dg[c];
    for (int f = B.f0; f < B.f1; f++) {
        tmp[a.upper[f]] += lo[f];
        tmp[a.lower[f]] += up[f];
    }
    for (int q = B.q0; q < B.q1; q++) tmp[a.cRow[q]] -= cc[q];   // lduMatrix::sumA, coupled patches (lduMatrixATmul.C:187-198)
    const double sumX = c_sum(a, B, [&](int c) { return x[c]; });
    const double xbar = sumX / n;
    double nfac = c_sum(a, B, [&](int c) {
        const double t = tmp[c] * xbar;
        return fabs(Ax[c] - t) + fabs(b[c] - t);
    });
    nfac += 1e-20;
    double sm = c_sum(a, B, [&](int c) { return fabs(r[c]); });
    const double ini = sm / nfac;
    double fin = ini;

    if (!c_converged(fin, ini, a.tolerance, a.relTol)) {
        // reciprocal preconditioned diagonal
        for (int c = B.c0; c < B.c1; c++) rD[c] = dg[c];
        for (int f = B.f0; f < B.f1; f++) rD[a.upper[f]] -= up[f] * lo[f] / rD[a.lower[f]];
        for (int c = B.c0; c < B.c1; c++) rD[c] = 1.0 / rD[c];

        if (a.symmetric) {
            double* p = v3;
            double* wA = Ax;
            double wArA = 1e20, wArAold;
            int it = 0;
            do {
                wArAold = wArA;
                c_precondition(a, B, up, lo, rD, wA, r);
                wArA = c_sum(a, B, [&](int c) { return wA[c] * r[c]; });
                if (it == 0) {
                    for (int c = B.c0; c < B.c1; c++) p[c] = wA[c];
                } else {
                    const double beta = wArA / wArAold;
                    for (int c = B.c0; c < B.c1; c++) p[c] = wA[c] + beta * p[c];
                }
                c_amul(a, B, up, lo, dg, wA, p);
                const double wApA = c_sum(a, B, [&](int c) { return wA[c] * p[c]; });
                if (fabs(wApA) / nfac < vSmall) break;
                const double alpha = wArA / wApA;
                for (int c = B.c0; c < B.c1; c++) {
                    x[c] += alpha * p[c];
                    r[c] -= alpha * wA[c];
                }
                sm = c_sum(a, B, [&](int c) { return fabs(r[c]); });
                fin = sm / nfac;
            } while (++it < a.maxIter && !c_converged(fin, ini, a.tolerance, a.relTol));
        } else {
            double* p = v3;
            double* y = Ax;
            double* AyA = v2;
            double* s = v4;
            double* z = v5;
            double* t = v6;
            double* r0 = v7;
            for (int c = B.c0; c < B.c1; c++) r0[c] = r[c];
            double rho = 0, alpha = 0, omega = 0;
            int it = 0;
            do {
                const double rhoOld = rho;
                rho = c_sum(a, B, [&](int c) { return r0[c] * r[c]; });
                if (fabs(rho) < vSmall) break;
                if (it == 0) {
                    for (int c = B.c0; c < B.c1; c++) p[c] = r[c];
                } else {
                    if (fabs(omega) < vSmall) break;
                    const double beta = (rho / rhoOld) * (alpha / omega);
                    for (int c = B.c0; c < B.c1; c++) p[c] = r[c] + beta * (p[c] - omega * AyA[c]);
                }
                c_precondition(a, B, up, lo, rD, y, p);
                c_amul(a, B, up, lo, dg, AyA, y);
                const double r0AyA = c_sum(a, B, [&](int c) { return r0[c] * AyA[c]; });
                alpha = rho / r0AyA;
                for (int c = B.c0; c < B.c1; c++) s[c] = r[c] - alpha * AyA[c];
                sm = c_sum(a, B, [&](int c) { return fabs(s[c]); });
                fin = sm / nfac;
                if (++it >= 0 && c_converged(fin, ini, a.tolerance, a.relTol)) {
                    for (int c = B.c0; c < B.c1; c++) x[c] += alpha * y[c];
                    break;
                }
                c_precondition(a, B, up, lo, rD, z, s);
                c_amul(a, B, up, lo, dg, t, z);
                const double tt = c_sum(a, B, [&](int c) { return t[c] * t[c]; });
                const double ts = c_sum(a, B, [&](int c) { return t[c] * s[c]; });
                omega = ts / tt;
                for (int c = B.c0; c < B.c1; c++) {
                    x[c] += alpha * y[c] + omega * z[c];
                    r[c] = s[c] - omega * t[c];
                }
                sm = c_sum(a, B, [&](int c) { return fabs(r[c]); });
                fin = sm / nfac;
            } while (it < a.maxIter && !c_converged(fin, ini, a.tolerance, a.relTol));
        }
    }
    if (a.gathered) {
        if (lane == a.myRank)
            for (int c = B.c0; c < B.c1; c++) a.psi[a.ipos[c - B.c0]] = x[c];
    } else {
        for (int c = B.c0; c < B.c1; c++) a.psi[a.ipos[c]] = x[c];
    }
}

// ---- gathered coarsest level ------------------------------------------------------------------------------
// With coupled patches on the coarsest level (processor patches of a decomposed case, cyclic halves) the reference
// runs its regular distributed PCG+DIC / PBiCGStab+DILU there (GAMGSolver.C:286-319).  Launch by launch that is a few
// thousand tiny kernels and all-reduces per V-cycle.  Instead every rank receives the coarsest blocks of all ranks
// (topology once per mesh, coefficients once per matrix, the source once per cycle: one all-gather) and replays the
// distributed iteration in k_coarsest_solve: block-local DIC/DILU (faces never cross ranks), couplings as explicit
// off-diagonal entries, sums rank by rank.  All ranks compute the same numbers; each keeps its own slice.

static constexpr int kMaxGatheredCells = 4096;

// Collective: every rank calls it at the same point (first GAMG solve on a mesh) and all ranks reach the same
// decision, because it is taken from the all-gathered data only.
static void ensureCoarsestGather(b200ls_matrix_s* m, int k) {
    DevLevel& D = DL(m, k);
    CoarsestGather& G = D.gather;
    Context& c = ctx();
    if (G.tried) return;
    G.tried = true;
    static const bool off = getenv("B200LS_NO_COARSEST_GATHER") != nullptr;
    if (off || int(D.ifaceCellsRef.size()) != D.nIfaces) return;
    const int R = c.nRanks;
    G.ifaceCoupleOff.assign(D.nIfaces + 1, 0);
    for (int i = 0; i < D.nIfaces; i++) G.ifaceCoupleOff[i + 1] = G.ifaceCoupleOff[i] + D.ifaceSize[i];
    G.myCouple = G.ifaceCoupleOff[D.nIfaces];
    int unsupported = 0;   // several processor patches to one neighbour: the pairing by rank is ambiguous
    for (int i = 0; i < D.nIfaces; i++)
        for (int j = 0; j < i; j++)
            if (D.ifacePartner[i] < 0 && D.ifacePartner[j] < 0 && D.ifaceNbr[i] == D.ifaceNbr[j]) unsupported = 1;
    std::vector<int> sizes;
    allGatherInts({D.nCells, D.nFaces, G.myCouple, unsupported}, sizes);
    std::vector<int> cellOff(R + 1, 0), faceOff(R + 1, 0), coupleOff(R + 1, 0);
    G.maxCells = G.maxFaces = G.maxCouple = 0;
    for (int r = 0; r < R; r++) {
        cellOff[r + 1] = cellOff[r] + sizes[4 * r];
        faceOff[r + 1] = faceOff[r] + sizes[4 * r + 1];
        coupleOff[r + 1] = coupleOff[r] + sizes[4 * r + 2];
        G.maxCells = std::max(G.maxCells, sizes[4 * r]);
        G.maxFaces = std::max(G.maxFaces, sizes[4 * r + 1]);
        G.maxCouple = std::max(G.maxCouple, sizes[4 * r + 2]);
        if (sizes[4 * r + 3]) unsupported = 1;
    }
    if (unsupported || cellOff[R] > kMaxGatheredCells) return;
    G.nCells = cellOff[R];
    G.nFaces = faceOff[R];
    G.nCouple = coupleOff[R];
    G.blockLen = G.maxCells + 2 * G.maxFaces + G.maxCouple;

    // topology block of this rank: [lower | upper] (maxFaces each), then per coupling entry
    // [cell | peer rank | peer patch (cyclic partner on the same rank, -1: the peer's patch towards me) | patch | face]
    const int mf = G.maxFaces, mq = G.maxCouple;
    std::vector<int> topo(size_t(2) * mf + size_t(5) * mq, -1), all;
    for (int f = 0; f < D.nFaces; f++) {
        topo[f] = D.hostRefLower[f];
        topo[mf + f] = D.hostRefUpper[f];
    }
    for (int i = 0; i < D.nIfaces; i++)
        for (int e = 0; e < D.ifaceSize[i]; e++) {
            int* t = topo.data() + 2 * mf;
            const int q = G.ifaceCoupleOff[i] + e;
            t[q] = D.ifaceCellsRef[i][e];
            t[mq + q] = D.ifacePartner[i] >= 0 ? c.rank : D.ifaceNbr[i];
            t[2 * mq + q] = D.ifacePartner[i];
            t[3 * mq + q] = i;
            t[4 * mq + q] = e;
        }
    allGatherInts(topo, all);
    const size_t tb = topo.size();
    std::vector<int> lower(G.nFaces), upper(G.nFaces), cRow(G.nCouple), cCol(G.nCouple, -1);
    // (rank, patch, face) -> global cell, and for processor patches (rank, neighbour rank) -> patch
    std::map<std::array<int, 3>, int> cellOf;
    std::map<std::array<int, 2>, int> patchTo;
    for (int r = 0; r < R; r++) {
        const int* t = all.data() + tb * r;
        for (int f = 0; f < sizes[4 * r + 1]; f++) {
            lower[faceOff[r] + f] = t[f] + cellOff[r];
            upper[faceOff[r] + f] = t[mf + f] + cellOff[r];
        }
        t += 2 * mf;
        for (int q = 0; q < sizes[4 * r + 2]; q++) {
            cellOf[{r, t[3 * mq + q], t[4 * mq + q]}] = t[q] + cellOff[r];
            if (t[2 * mq + q] < 0) patchTo[{r, t[mq + q]}] = t[3 * mq + q];
        }
    }
    for (int r = 0; r < R; r++) {
        const int* t = all.data() + tb * r + 2 * mf;
        for (int q = 0; q < sizes[4 * r + 2]; q++) {
            const int peer = t[mq + q];
            int peerPatch = t[2 * mq + q];
            if (peerPatch < 0) {
                auto it = patchTo.find({peer, r});
                if (it == patchTo.end()) return;
                peerPatch = it->second;
            }
            auto it = cellOf.find({peer, peerPatch, t[4 * mq + q]});
            if (it == cellOf.end()) return;
            cRow[coupleOff[r] + q] = t[q] + cellOff[r];
            cCol[coupleOff[r] + q] = it->second;
        }
    }
    G.lower.upload(lower, c.stream);
    G.upper.upload(upper, c.stream);
    G.cRow.upload(cRow, c.stream);
    G.cCol.upload(cCol, c.stream);
    G.cellOff.upload(cellOff, c.stream);
    G.faceOff.upload(faceOff, c.stream);
    G.coupleOff.upload(coupleOff, c.stream);
    B2_CUDA(cudaStreamSynchronize(c.stream));
    G.ok = true;
}

static void allGatherDoubles(const double* mine, double* all, size_t count) {
    Context& c = ctx();
    int r = c.nccl.AllGather(mine, all, count, ncclDouble, c.comm, c.stream);
    if (r != 0) throw CudaError(std::string("ncclAllGather: ") + c.nccl.GetErrorString((ncclResult_t)r));
}

// coefficients of the coarsest level of every rank, after the coarse matrices were (re)built
static void gatherCoarsestCoefs(b200ls_matrix_s* m) {
    const int k = int(m->levels.size()) - 1;
    DevLevel& D = DL(m, k);
    MatLevel& M = m->levels[k];
    Context& c = ctx();
    if (!(c.nRanks > 1 || D.nIfaces > 0)) return;
    ensureCoarsestGather(m, k);
    const CoarsestGather& G = D.gather;
    if (!G.ok) return;
    m->gSendCoef.alloc(G.blockLen);
    m->gSendSrc.alloc(G.maxCells);
    if (G.blockLen) B2_CUDA(cudaMemsetAsync(m->gSendCoef.p, 0, sizeof(double) * G.blockLen, S()));
    if (G.maxCells) B2_CUDA(cudaMemsetAsync(m->gSendSrc.p, 0, sizeof(double) * G.maxCells, S()));
    double* blk = m->gSendCoef.p;
    if (D.nCells) LAUNCH(k_gather, gridStride(D.nCells), 256, blk, M.diag.p, D.ipos.p, D.nCells);
    if (D.nFaces) {
        LAUNCH(k_gather, gridStride(D.nFaces), 256, blk + G.maxCells, M.Uval(), D.Uidx.p, D.nFaces);
        LAUNCH(k_gather, gridStride(D.nFaces), 256, blk + G.maxCells + G.maxFaces, M.Lval(D.nFaces), D.Lidx.p,
               D.nFaces);
    }
    for (int i = 0; i < D.nIfaces; i++)
        if (D.ifaceSize[i])
            B2_CUDA(cudaMemcpyAsync(blk + G.maxCells + 2 * G.maxFaces + G.ifaceCoupleOff[i], M.bou[i].p,
                                    sizeof(double) * D.ifaceSize[i], cudaMemcpyDeviceToDevice, S()));
    if (c.nRanks > 1) {
        m->gCoef.alloc(size_t(G.blockLen) * c.nRanks);
        m->gSrc.alloc(size_t(G.maxCells) * c.nRanks);
        allGatherDoubles(m->gSendCoef.p, m->gCoef.p, G.blockLen);
    }
}

static void launchCoarsest(const CoarsestArgs& a, size_t smemBytes) {
    constexpr size_t kMax = 200 * 1024;
    const int useSmem = smemBytes <= kMax ? 1 : 0;
    static bool attr = false;
    if (!attr) {
        B2_CUDA(cudaFuncSetAttribute(k_coarsest_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMax)));
        attr = true;
    }
    k_coarsest_solve<<<1, 32, useSmem ? smemBytes : 0, S()>>>(a, useSmem);
    ctx().launches++;
}

static void solveCoarsest(b200ls_matrix_s* m, const b200ls_controls& c) {
    const int k = int(m->levels.size()) - 1;
    DevLevel& D = DL(m, k);
    MatLevel& M = m->levels[k];
    if (ctx().nRanks == 1 && D.nIfaces > 0 && D.nFaces == 0) {
        // lduMatrix::diagonal() -> diagonalSolver, which ignores the interfaces (GAMGSolver.C:271-284, diagonalSolver.C:66)
        LAUNCH(k_div, gridStride(D.nCells), 256, M.corr.p, M.src.p, M.diag.p, D.nCells);
        return;
    }
    if ((ctx().nRanks > 1 || D.nIfaces > 0) && D.gather.ok) {
        const CoarsestGather& G = D.gather;
        Context& cx = ctx();
        const double *gCoef = m->gSendCoef.p, *gSrc = m->gSendSrc.p;
        if (D.nCells) LAUNCH(k_gather, gridStride(D.nCells), 256, m->gSendSrc.p, M.src.p, D.ipos.p, D.nCells);
        if (cx.nRanks > 1) {
            allGatherDoubles(m->gSendSrc.p, m->gSrc.p, G.maxCells);
            gCoef = m->gCoef.p;
            gSrc = m->gSrc.p;
        }
        const size_t nD = size_t(2) * G.nFaces + size_t(12) * G.nCells + G.nCouple;
        m->coarsestWork.alloc(nD + 16);
        CoarsestArgs a;
        memset(&a, 0, sizeof(a));
        a.nCells = G.nCells;
        a.nFaces = G.nFaces;
        a.symmetric = m->symmetric ? 1 : 0;
        a.lower = G.lower.p;
        a.upper = G.upper.p;
        a.ipos = D.ipos.p;
        a.psi = M.corr.p;
        a.work = m->coarsestWork.p;
        a.tolerance = c.tolerance;
        a.relTol = c.relTol;
        a.maxIter = 1000;   // lduMatrix::solver::defaultMaxIter_
        a.gathered = 1;
        a.nRanks = cx.nRanks;
        a.myRank = cx.rank;
        a.nCouple = G.nCouple;
        a.maxCells = G.maxCells;
        a.maxFaces = G.maxFaces;
        a.blockLen = G.blockLen;
        a.cellOff = G.cellOff.p;
        a.faceOff = G.faceOff.p;
        a.coupleOff = G.coupleOff.p;
        a.cRow = G.cRow.p;
        a.cCol = G.cCol.p;
        a.gCoef = gCoef;
        a.gSrc = gSrc;
        launchCoarsest(a, nD * sizeof(double) +
                              (size_t(2) * G.nFaces + size_t(2) * G.nCouple + size_t(3) * (cx.nRanks + 1)) * sizeof(int));
        return;
    }
    if (ctx().nRanks > 1 || D.nIfaces > 0) {
        // coarsest level with coupled patches (distributed, or cyclic on one rank): the regular PCG+DIC / PBiCGStab+DILU on that level with the parent's
        // tolerance and relTol, zero initial guess (GAMGSolver.C:286-319, GAMGSolverSolve.C:538-545)
        b200ls_controls cc;
        memset(&cc, 0, sizeof(cc));
        cc.tolerance = c.tolerance;
        cc.relTol = c.relTol;
        cc.maxIter = 1000;
        cc.minIter = 0;
        cc.precond = m->symmetric ? B200LS_DIC : B200LS_DILU;
        static b200ls_perf cperf;
        memset(&cperf, 0, offsetof(b200ls_perf, history));
        B2_CUDA(cudaMemsetAsync(M.corr.p, 0, sizeof(double) * D.nCells, S()));
        g_scalarBank = 1;
        try {
            if (m->symmetric) solvePCG(m, cc, k, M.corr.p, M.src.p, &cperf, nullptr);
            else solvePBiCGStab(m, cc, k, M.corr.p, M.src.p, &cperf, nullptr);
        } catch (...) {
            g_scalarBank = 0;
            throw;
        }
        g_scalarBank = 0;
        return;
    }
    m->coarsestWork.alloc(size_t(12) * D.nCells + size_t(2) * D.nFaces + 16);
    CoarsestArgs a;
    a.nCells = D.nCells;
    a.nFaces = D.nFaces;
    a.symmetric = m->symmetric ? 1 : 0;
    a.lower = D.refLower.p;
    a.upper = D.refUpper.p;
    a.Uidx = D.Uidx.p;
    a.Lidx = D.Lidx.p;
    a.ipos = D.ipos.p;
    a.diag = M.diag.p;
    a.Uval = M.Uval();
    a.Lval = M.Lval(D.nFaces);
    a.source = M.src.p;
    a.psi = M.corr.p;
    a.work = m->coarsestWork.p;
    a.tolerance = c.tolerance;
    a.relTol = c.relTol;
    a.maxIter = 1000;   // lduMatrix::solver::defaultMaxIter_
    a.gathered = 0;
    a.nRanks = 1;
    a.myRank = 0;
    a.nCouple = 0;
    a.cRow = a.cCol = nullptr;
    launchCoarsest(a, (size_t(2) * D.nFaces + size_t(12) * D.nCells) * sizeof(double) +
                          size_t(2) * D.nFaces * sizeof(int));
}

// GAMGSolver::scale (GAMGSolverScale.C:31-76)
static void gamgScale(b200ls_matrix_s* m, int level, double* field, double* Acf, const double* source) {
    Context& cx = ctx();
    DevLevel& D = DL(m, level);
    MatLevel& M = m->levels[level];
    const int n = D.nCells;
    opAmul(m, level, Acf, field);
    LAUNCH(k_dot2, kReduceBlocks, kReduceThreads, scalar(m, S_NUM), source, Acf, field, n, cx.partials.p,
           cx.ticket.p);
    allReduce(scalar(m, S_NUM), 2);
    LAUNCH(k_gamg_scale, gridStride(n), 256, field, source, Acf, M.diag.p, scalar(m, S_NUM), n);
}

static void restrictTo(b200ls_matrix_s* m, int fineLevel, double* coarse, const double* fine) {
    DevLevel& F = DL(m, fineLevel);
    DevLevel& C = DL(m, fineLevel + 1);
    if (C.nCells) LAUNCH(k_restrict, gridRows(C.nCells), 256, coarse, fine, F.rPtr.p, F.rFine.p, C.nCells);
}
static void prolongTo(b200ls_matrix_s* m, int fineLevel, double* fine, const double* coarse) {
    DevLevel& F = DL(m, fineLevel);
    if (F.nCells) LAUNCH(k_prolong, gridRows(F.nCells), 256, fine, coarse, F.pMap.p, F.nCells);
}

// one V-cycle (GAMGSolverSolve.C:148-443).  My level k>0 is the reference's matrixLevels_[k-1];
// coarseCorrFields[l] / coarseSources[l] live on level l+1 as levels[l+1].corr / .src.
static void vcycle(b200ls_matrix_s* m, const b200ls_controls& c, double*& psi, double*& psiSpare,
                   const double* source, double* Apsi, double* finestCorrection, double* finestResidual,
                   bool scaleCorrection) {
    const int nL = int(m->levels.size());
    const int coarsest = nL - 2;   // reference coarsestLevel index into matrixLevels_
    const int n0 = DL(m, 0).nCells;

    vmark("start", 0);
    restrictTo(m, 0, m->levels[1].src.p, finestResidual);
    vmark("restrict", 0);
    for (int l = 0; l < coarsest; l++) {
        MatLevel& ML = m->levels[l + 1];
        if (c.nPreSweeps) {
            DevLevel& D = DL(m, l + 1);
            B2_CUDA(cudaMemsetAsync(ML.corr.p, 0, sizeof(double) * D.nCells, S()));
            ensureLevelScratch(m, l + 1);
            double* corr = ML.corr.p;
            double* spare = ML.tmpC.p;
            opSmooth(m, l + 1, c.precond, corr, spare, ML.src.p,
                     std::min(c.nPreSweeps + c.preSweepsLevelMultiplier * l, c.maxPreSweeps));
            vmark("pre-smooth", l + 1);
            if (corr != ML.corr.p) std::swap(ML.corr.p, ML.tmpC.p);
            double* ACf = ML.tmpB.p;
            if (scaleCorrection && l < coarsest - 1) gamgScale(m, l + 1, ML.corr.p, ACf, ML.src.p);
            opAmul(m, l + 1, ACf, ML.corr.p);
            // coarseSources[l] -= ACf
            LAUNCH(k_sub_inplace, gridStride(D.nCells), 256, ML.src.p, ACf, D.nCells);
            vmark("scale+residual", l + 1);
        }
        restrictTo(m, l + 1, m->levels[l + 2].src.p, ML.src.p);
        vmark("restrict", l + 1);
    }

    solveCoarsest(m, c);
    vmark("coarsest solve", nL - 1);

    for (int l = coarsest - 1; l >= 0; l--) {
        DevLevel& D = DL(m, l + 1);
        MatLevel& ML = m->levels[l + 1];
        ensureLevelScratch(m, l + 1);
        double* pre = nullptr;
        if (c.nPreSweeps) {
            // preSmoothedCoarseCorrField = coarseCorrFields[l]
            ML.dWork.alloc(D.nCells);
            pre = ML.dWork.p;
            B2_CUDA(cudaMemcpyAsync(pre, ML.corr.p, sizeof(double) * D.nCells, cudaMemcpyDeviceToDevice, S()));
        }
        prolongTo(m, l + 1, ML.corr.p, m->levels[l + 2].corr.p);
        if (scaleCorrection && l < coarsest - 1) gamgScale(m, l + 1, ML.corr.p, ML.tmpB.p, ML.src.p);
        if (pre) LAUNCH(k_add_inplace, gridStride(D.nCells), 256, ML.corr.p, pre, D.nCells);
        vmark("prolong+scale", l + 1);
        double* corr = ML.corr.p;
        double* spare = ML.tmpC.p;
        // Gauss-Seidel sweeps swap corr/spare; the other smoothers leave corr in place
        opSmooth(m, l + 1, c.precond, corr, spare, ML.src.p,
                 std::min(c.nPostSweeps + c.postSweepsLevelMultiplier * l, c.maxPostSweeps));
        vmark("post-smooth", l + 1);
        if (corr != ML.corr.p) std::swap(ML.corr.p, ML.tmpC.p);
    }

    prolongTo(m, 0, finestCorrection, m->levels[1].corr.p);
    if (scaleCorrection) gamgScale(m, 0, finestCorrection, Apsi, finestResidual);
    LAUNCH(k_add_inplace, gridStride(n0), 256, psi, finestCorrection, n0);
    vmark("prolong+scale", 0);
    opSmooth(m, 0, c.precond, psi, psiSpare, source, c.nFinestSweeps);
    vmark("finest smooth", 0);
}

// GAMGPreconditioner::precondition (GAMGPreconditioner.C:81-148): nVcycles V-cycles on A wA = rA from wA = 0
static void gamgPrecondition(b200ls_matrix_s* m, const b200ls_controls& c, double* wA, const double* rA) {
    const int n = DL(m, 0).nCells;
    b200ls_controls g = c;
    g.precond = c.precSmoother;
    g.tolerance = c.precTolerance;       // inherited by the coarsest-level solver
    g.relTol = c.precRelTol;
    const bool scaleCorrection = g.scaleCorrection < 0 ? m->symmetric : (g.scaleCorrection != 0);
    double* AwA = m->vec("gAwA");
    double* finestCorrection = m->vec("gCorr");
    double* finestResidual = m->vec("gRes");
    double* psi = wA;
    double* spare = m->vec("gSpare");
    B2_CUDA(cudaMemsetAsync(psi, 0, sizeof(double) * n, S()));
    B2_CUDA(cudaMemcpyAsync(finestResidual, rA, sizeof(double) * n, cudaMemcpyDeviceToDevice, S()));
    for (int cycle = 0; cycle < g.nVcycles; cycle++) {
        vcycle(m, g, psi, spare, rA, AwA, finestCorrection, finestResidual, scaleCorrection);
        if (cycle < g.nVcycles - 1) {
            opAmul(m, 0, AwA, psi);
            LAUNCH(k_sub, gridStride(n), 256, finestResidual, rA, AwA, n);
        }
    }
    if (psi != wA) {
        // Gauss-Seidel sweeps ping-pong between the two buffers: bring the result home
        B2_CUDA(cudaMemcpyAsync(wA, psi, sizeof(double) * n, cudaMemcpyDeviceToDevice, S()));
        // keep the named scratch vector pointing at its own storage (psi/spare were only swapped locally)
    }
}

static void solveGAMG(b200ls_matrix_s* m, const b200ls_controls& c, double*& psi, double*& psiSpare,
                      const double* source, b200ls_perf* perf, cudaEvent_t evLoopStart) {
    Context& cx = ctx();
    if (m->levels.size() < 2) {
        throw CudaError("No coarse levels created, either matrix too small for GAMG or minCellsPerProcessor too "
                        "large (call b200ls_agglomerate first)");
    }
    const int n = DL(m, 0).nCells;
    double* Apsi = m->vec("wA");
    double* finestCorrection = m->vec("pA");
    double* finestResidual = m->vec("rA");
    const bool scaleCorrection = c.scaleCorrection < 0 ? m->symmetric : (c.scaleCorrection != 0);

    // solver construction: coarse matrices (+ smoother factorisations) -- timed as setup
    buildCoarseMatrices(m);
    // (DIC/DILU smoothers factorise lazily inside opSmooth, once per level per solve)

    opAmul(m, 0, Apsi, psi);
    const double nf = normFactor(m, 0, psi, source, Apsi, finestCorrection, globalCells(m));
    perf->normFactor = nf;
    LAUNCH(k_sub, gridStride(n), 256, finestResidual, source, Apsi, n);
    reduce<RED_SUMMAG>(scalar(m, S_RES), finestResidual, nullptr, n);
    allReduce(scalar(m, S_RES), 1);
    readScalars(scalar(m, S_RES), 1);
    perf->initialResidual = cx.pinned[0] / nf;
    perf->finalResidual = perf->initialResidual;
    perf->nIterations = 0;

    B2_CUDA(cudaEventRecord(evLoopStart, S()));
    if (c.minIter > 0 || !checkConvergence(perf, c)) {
        do {
            vcycle(m, c, psi, psiSpare, source, Apsi, finestCorrection, finestResidual, scaleCorrection);
            opAmul(m, 0, Apsi, psi);
            LAUNCH(k_sub, gridStride(n), 256, finestResidual, source, Apsi, n);
            reduce<RED_SUMMAG>(scalar(m, S_RES), finestResidual, nullptr, n);
            allReduce(scalar(m, S_RES), 1);
            readScalars(scalar(m, S_RES), 1);
            perf->finalResidual = cx.pinned[0] / nf;
            record(perf, c, perf->finalResidual);
        } while ((++perf->nIterations < c.maxIter &&
                  !checkConvergence(perf, c)) ||
                 perf->nIterations < c.minIter);
    }
}

// ------------------------------------------------------------------------------------------------------------
// entry
// ------------------------------------------------------------------------------------------------------------

void solveDev(b200ls_matrix_s* m, const b200ls_controls& c, double* psiCell, const double* sourceCell,
              b200ls_perf* perf) {
    if (!m->valuesSet) throw CudaError("b200ls_solve: matrix coefficients not set");
    Context& cx = ctx();
    const int64_t launches0 = cx.launches;
    memset(perf, 0, offsetof(b200ls_perf, history));

    cudaEvent_t ev0, ev1, ev2;
    B2_CUDA(cudaEventCreate(&ev0));
    B2_CUDA(cudaEventCreate(&ev1));
    B2_CUDA(cudaEventCreate(&ev2));
    B2_CUDA(cudaEventRecord(ev0, S()));

    // named vectors are looked up before taking raw pointers (std::map nodes are stable)
    b200ls::Vec& vPsi = m->vecs["psi"];
    b200ls::Vec& vSpare = m->vecs["psiSpare"];
    m->vec("psi");
    m->vec("psiSpare");
    double* source = m->vec("source");
    toPositions(m, 0, vPsi.buf.p, psiCell);
    toPositions(m, 0, source, sourceCell);

    try {
        switch (c.solver) {
            case B200LS_PCG:
                if (!m->symmetric) throw CudaError("PCG requires a symmetric matrix");
                solvePCG(m, c, 0, vPsi.buf.p, source, perf, ev1);
                break;
            case B200LS_PBICGSTAB:
                solvePBiCGStab(m, c, 0, vPsi.buf.p, source, perf, ev1);
                break;
            case B200LS_GAMG:
                solveGAMG(m, c, vPsi.buf.p, vSpare.buf.p, source, perf, ev1);
                break;
            case B200LS_SMOOTH_SOLVER:
                solveSmooth(m, c, vPsi.buf.p, vSpare.buf.p, source, perf, ev1);
                break;
            case B200LS_DIAGONAL_SOLVER: {
                // diagonalSolver.C:62-79: psi = source/diag; reports zero residuals, zero iterations, converged
                const int n = DL(m, 0).nCells;
                B2_CUDA(cudaEventRecord(ev1, S()));
                if (n) LAUNCH(k_div, gridStride(n), 256, vPsi.buf.p, source, m->levels[0].diag.p, n);
                perf->converged = 1;
                break;
            }
            default:
                throw CudaError("unknown solver");
        }
        toCells(m, 0, psiCell, vPsi.buf.p);
        vprofDump(perf->nIterations);
        B2_CUDA(cudaEventRecord(ev2, S()));
        B2_CUDA(cudaStreamSynchronize(S()));
        checkLaunch("solve");
        checkSweepError();
    } catch (...) {
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
        cudaEventDestroy(ev2);
        throw;
    }
    float msSetup = 0, msLoop = 0;
    cudaEventElapsedTime(&msSetup, ev0, ev1);
    cudaEventElapsedTime(&msLoop, ev1, ev2);
    perf->setupMs = msSetup;
    perf->solveMs = msLoop;
    perf->kernelLaunches = cx.launches - launches0;
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    cudaEventDestroy(ev2);
}

}  // namespace b200ls
