// Pencil sweeps: DIC/DILU factorisation + substitution and Gauss-Seidel on structured blocks (mesh.hpp PencilPlan).
//
// The wavefront kernels (kernels.cuh) pay one L2 round trip per wavefront because every dependency crosses SMs:
// 382 hops at 128^3.  Here ONE WARP owns a tile of wj x wk <= 32 pencils (lines of cells along i) and walks it along
// i, lane = pencil, lane (jj, kk) skewed by SKEW*(jj + kk) steps.  The three lower neighbours of a row then are
//   * the lane's own previous row                       -> a register,
//   * the row of lane-1 / lane-wj computed SKEW steps ago -> a warp shuffle (issued a step early when SKEW = 2),
//   * for lanes on the low-j / low-k face of the tile: a row of tile (J-1,K) / (J,K-1), launched earlier.
// so the dependency chain of a step is one multiply and one subtract (the own-row term is the last one in the
// reference's accumulation order), and only tile-to-tile hand-overs go through L2: nJ + nK hops (48 at 128^3).
//
// A CTA is a pipeline of five warps around one tile:
//   * the HELPER warp (a) streams the per-row operands, which are contiguous per (tile, i-range) in the tile-major
//     layout, into a shared-memory raw ring with bulk async copies (cp.async.bulk + mbarrier complete_tx, one elected
//     lane), and (b) fetches the values the face lanes need from the neighbouring tiles a window of steps ahead:
//     polls them in L2 until none is the sentinel and deposits them in a small ring, one mbarrier per step;
//   * two PREP warps turn raw operands + neighbour values into per-step records (2-4 16-byte vectors per lane:
//     pre-multiplied coefficients, right-hand side), so that
//   * the CHAIN warp does nothing but: load record, two shuffles, the reference's multiply-subtract sequence in
//     exactly the order of kernels.cuh / the reference, store the result to the result ring;
//   * the WRITER warp drains the result ring: face lanes publish at once with one 8-byte L2 store (sentinel protocol,
//     as the wavefront kernels), complete rows are stored coalesced, the other buffer is re-armed, the fused dot
//     product accumulated.
// Hand-overs inside the CTA are mbarrier full/empty pairs (one elected lane arrives after __syncwarp()).
// Backward sweeps run the same code on reflected coordinates.
//
// Reference order per row (bit-exact, -fmad=false):
//   DIC/DILU forward   wA[c] = rD*rA - (rD*l_K)*wA[c-K] - (rD*l_J)*wA[c-J] - (rD*l_I)*wA[c-1]     DICPreconditioner.C:109-117
//   backward           wA[c] = wA[c] - (rD*u_K)*wA[c+K] - (rD*u_J)*wA[c+J] - (rD*u_I)*wA[c+1]     DICPreconditioner.C:119-122
//   calcReciprocalD    d[c]  = diag - u_K*l_K/d[c-K] - u_J*l_J/d[c-J] - u_I*l_I/d[c-1]            DICPreconditioner.C:71-83
//   GaussSeidel        psi[c] = (b - l_K*psi[c-K] - l_J*psi[c-J] - l_I*psi[c-1]
//                                  - u_I*old[c+1] - u_J*old[c+J] - u_K*old[c+K]) / diag             GaussSeidelSmoother.C:151-176
//   symGaussSeidel reverse half: lower terms with the forward values first, then u_I, u_J, u_K with the new ones
//                                                                                                   symGaussSeidelSmoother.C:178-205
// (faces of a cell ascend with the neighbour's index: K-, J-, I- on the lower side, I+, J+, K+ on the upper side).
#pragma once

#include "kernels.cuh"

namespace b200ls {

enum { PM_FWD = 0, PM_BWD = 1, PM_FACTOR = 2, PM_GS_FWD = 3, PM_GS_REV = 4 };

static constexpr int kPencilMaxPlanes = 9;
static constexpr int kPencilR = 8;                          // rows per bulk copy / ring stage
static constexpr int kPencilPlaneBytes = kPencilR * 32 * 8; // one plane of one stage
static constexpr int kPencilE = 32;                         // steps held by the neighbour-value ring
static constexpr int kPencilCPL = 3;                        // ring columns per helper lane (up to 96 columns)
static constexpr int kPencilLook = 6;                       // steps per column a helper round looks ahead

struct PencilArgs {
    const PencilTileDev* tiles;
    const int* order;       // launch order of the forward sweeps; backward sweeps walk it from the end
    int nTiles;
    int nx, ny, nz;
    int extW;               // doubles per step of the neighbour-value ring (full tile)
    // operand planes in position order (see k_pencil for the meaning per mode); unused entries are null
    const double* plane[kPencilMaxPlanes];
    double* out;            // sentinel-armed result
    double* out2;           // PM_FACTOR: reciprocal
    double* clear;          // optional: entry p is re-armed with the sentinel once row p is done
    // optional fused dot product of the result with plane[4] (PM_BWD; plane[4] must be set iff dotOut is)
    double* dotOut;
    double* partials;
    unsigned int* ticket;
    int* err;
    // debugging aid (B200LS_PENCIL_PROF): 16 counters per tile in launch order --
    // chain: start, end (globaltimer ns), cycles waiting for records, for result-ring space;
    // prep: cycles waiting for operands, neighbour values, record slots; writer: cycles waiting for results;
    // helper: polling rounds, cycles waiting for ring capacity
    unsigned long long* prof;
    int debug;              // B200LS_PENCIL_DEBUG bits: 2 = slow helper rounds
    const int* stop;        // optional: a non-zero word makes the launch a no-op (speculatively enqueued iterations)
    // G > 1 tiles per CTA (consecutive along k; template parameter G of k_pencil): groups in launch order, G tile indices
    // each (-1: the group has no such tile)
    const int* groupTiles;
    int nGroups;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity, int* err) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > kMaxSpins) {
            *err = 1;
            break;
        }
    }
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void st_release_cta(unsigned addr, int v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta(unsigned addr) {
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_cg(const double* p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ double2 lds_v2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v2(unsigned addr, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}

template <int MODE>
struct PencilTraits {
    static constexpr int DIR = (MODE == PM_BWD || MODE == PM_GS_REV) ? -1 : 1;
    static constexpr bool GS = (MODE == PM_GS_FWD || MODE == PM_GS_REV);
    // operand planes streamed into the raw ring
    static constexpr int NP = MODE == PM_FWD ? 5 : MODE == PM_BWD ? 5 : MODE == PM_FACTOR ? 7 : 9;
    static constexpr int STAGE_BYTES = NP * kPencilPlaneBytes;
    // 16-byte vectors of a step record (what the chain warp reads per lane and step)
    static constexpr int NV = MODE == PM_GS_FWD ? 4 : MODE == PM_GS_REV ? 3 : 2;
    // lane stride: an odd number of vectors keeps the 16-byte accesses of a quarter warp on distinct banks
    static constexpr int REC_LANE_BYTES = (NV | 1) * 16;
    static constexpr int REC_STEP_BYTES = 32 * REC_LANE_BYTES;
};

// shared-memory layout of a CTA (dynamic):
// [full[8] | empty[8] | progress words | neighbour ring | record ring | result ring | dot ring | raw ring]
static constexpr int kPencilProfWords = 64;  // debugging counters per tile (B200LS_PENCIL_PROF)
static constexpr int kPencilD = 8;          // steps held by the record ring
static constexpr int kPencilPrep = 2;       // prep warps (step s is prepared by warp s % kPencilPrep)
static constexpr int kPencilThreads = 32 * (3 + kPencilPrep);   // chain, helper, writer + the prep warps
static constexpr int kPencilGroup = 2;      // tiles per CTA of the grouped substitution sweeps (k_pencil template parameter G)
static constexpr int kPencilOutRows = 64;   // steps held by the result ring (power of two, > largest skew + 26)
static constexpr int kPencilHeaderBytes = 768;
__host__ __device__ inline int pencilExtBytes(int extW) { return (kPencilE * extW * 8 + 127) / 128 * 128; }
__host__ __device__ constexpr int pencilOutBytes() { return kPencilOutRows * 32 * 8; }
template <int MODE, int NS>
__host__ __device__ inline int pencilSmemBytes(int extW, bool dot) {
    return kPencilHeaderBytes + pencilExtBytes(extW) + kPencilD * PencilTraits<MODE>::REC_STEP_BYTES +
           pencilOutBytes() * (dot ? 2 : 1) + NS * PencilTraits<MODE>::STAGE_BYTES;
}
// rows of the raw ring a step can touch at once: the skew of the tile, the chunk being loaded and the row read ahead
__host__ __device__ constexpr bool pencilFits(int skewUnits, int SKEW, int NS, bool gs) {
    return SKEW * skewUnits + kPencilR + (gs ? 2 : 1) <= NS * kPencilR && SKEW * skewUnits + 26 <= kPencilOutRows;
}

// One CTA = four warps working on one tile at a time:
//   warp 0  CHAIN   the recurrence itself: per step one record (3-5 LDS.128), two shuffles, the arithmetic, one STS
//   warp 1  HELPER  bulk copies of the operand planes into the raw ring; polls the neighbour tiles' face values in L2
//   warp 2  PREP    reads the raw ring with the lane skew, does everything that does not depend on new values
//                   (pre-multiplications, old-value terms of Gauss-Seidel, zero records for idle lanes) and writes the
//                   step records
//   warp 3  WRITER  takes the results from the result ring: face lanes are published at once (a neighbour tile waits
//                   for them), whole rows are stored coalesced once the most skewed lane has finished them; re-arms
//                   the other buffer with the sentinel; carries the fused dot product
// Hand-overs inside the CTA are sentinel words in shared memory (the value is the flag), mbarriers for the bulk
// copies, and three progress words.
// G tiles per CTA: G such pipelines side by side (warp w: pipeline w / 5), working on G tiles that follow each other along
// k.  The K-face values then never leave the SM: the writer warp of the producing tile deposits them straight into the
// neighbour-value ring of the consuming pipeline (second arrival on its per-step barrier), so only the J faces and
// the K faces between groups are handed over through L2.
template <int MODE, int SKEW, int NS, int G>
__global__ void __launch_bounds__(kPencilThreads * G, 1) k_pencil(PencilArgs a) {
    using T = PencilTraits<MODE>;
    constexpr int DIR = T::DIR;
    constexpr bool GS = T::GS;
    constexpr int NP = T::NP;
    constexpr int NV = T::NV;
    constexpr int R = kPencilR;
    constexpr int LA = GS ? 1 : 0;   // Gauss-Seidel reads the old value of the next row
    constexpr int kNever = 0x7fffffff;
    static_assert((NS & (NS - 1)) == 0 && NS <= 8, "ring stages: a power of two, at most 8");
    extern __shared__ __align__(128) unsigned char pencilSmem[];
    if (a.stop && *a.stop) return;   // (every CTA reads the same word: it only changes between launches)
    const bool doDot = MODE == PM_BWD && a.dotOut != nullptr;
    constexpr int kWarpsPerSub = kPencilThreads / 32;
    const int sub = G == 1 ? 0 : int(threadIdx.x >> 5) / kWarpsPerSub;
    const unsigned subBytes = unsigned(pencilSmemBytes<MODE, NS>(a.extW, doDot) + 127) & ~127u;
    const unsigned smCta = smem_u32(pencilSmem);
    unsigned smBase = smCta + unsigned(sub) * subBytes;
    asm volatile("mov.u32 %0, %0;" : "+r"(smBase));   // opaque: keep the base in a register (not re-derived from SR_CgaCtaId)
    // mbarriers: raw stages [8], record full / empty [8 each], result written [16], neighbour values of a step [32]
    const unsigned barFull = smBase, recFull = smBase + 64, recEmpty = smBase + 128, outFull = smBase + 192;
    const unsigned extFull = smBase + 320, rawDone = smBase + 576;   // + raw stage read by every prep warp [8]
    const unsigned wChainProg = smBase + 640, wWriterProg = smBase + 656;
    const unsigned extRing = smBase + kPencilHeaderBytes;
    const unsigned recRing = extRing + pencilExtBytes(a.extW);
    const unsigned outRing = recRing + kPencilD * T::REC_STEP_BYTES;
    const unsigned dotRing = outRing + pencilOutBytes();
    const unsigned dataRing = dotRing + (doDot ? pencilOutBytes() : 0);
    const int lane = threadIdx.x & 31;
    const int wid = int(threadIdx.x >> 5) - sub * kWarpsPerSub;
    // warp roles: 0 chain, 1 helper, 2 .. 1+kPencilPrep prep, last writer
    const int role = wid == 0 ? 0 : wid == 1 ? 1 : wid == 2 + kPencilPrep ? 3 : 2;
    const int prepId = wid - 2;
    const double sent = sentinel();
    const double NEUTRAL = MODE == PM_FACTOR ? 1.0 : 0.0;
    const int nx = a.nx;
    const int nChunks = (nx + R - 1) / R;
    const int pad = DIR > 0 ? 0 : (R - nx % R) % R;   // backward: the first (top) chunk is the partial one
    double dsum[1] = {0.0};

    if (wid == 0 && lane == 0) {
        for (int q = 0; q < NS; q++) {
            mbar_init(barFull + q * 8, 1);
            mbar_init(rawDone + q * 8, kPencilPrep);
        }
        for (int q = 0; q < kPencilD; q++) {
            // one arrival each, by lane 0 after __syncwarp() (a 32-lane arrive is 32 serialised shared-memory atomics)
            mbar_init(recFull + q * 8, 1);     // prep warp: record stored
            mbar_init(recEmpty + q * 8, 1);    // chain warp: record loaded
        }
        for (int q = 0; q < 16; q++) mbar_init(outFull + q * 8, 1);   // chain warp: result stored
        // G > 1: the per-step "neighbour values" barriers receive arrivals from another pipeline's writer as well, so
        // they are initialised ONCE (two arrivals per step, always) and their phase parity is carried across tiles;
        // re-initialising them at tile starts, as the one-tile-per-CTA path does, lost or duplicated the other warp's
        // last arrivals of the previous tile
        if (G > 1)
            for (int q = 0; q < kPencilE; q++) mbar_init(extFull + q * 8, 2);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    unsigned gchunk0 = 0;   // chunks / record groups handed over by earlier tiles of this CTA (same count in every warp)
    unsigned ggroup0 = 0;
    unsigned gext0 = 0;     // G > 1: laps of the neighbour-value ring (kPencilE steps each) completed by earlier tiles
    const int nWork = G == 1 ? a.nTiles : a.nGroups;
    for (int wi = blockIdx.x; wi < nWork; wi += gridDim.x) {
        // the tile of this pipeline; with G > 1 also who produces its chain-side K-face values (prodTile) and who
        // consumes its own (consTile) inside the CTA
        int ti, prodTile = -1, consTile = -1;
        if (G == 1) {
            ti = a.order[DIR > 0 ? wi : a.nTiles - 1 - wi];
        } else {
            const int* gt = a.groupTiles + size_t(DIR > 0 ? wi : a.nGroups - 1 - wi) * G;
            ti = gt[sub];
            if (sub - DIR >= 0 && sub - DIR < G) prodTile = gt[sub - DIR];
            if (sub + DIR >= 0 && sub + DIR < G) consTile = gt[sub + DIR];
            if (ti < 0) {          // no such tile in this group: keep the block barriers of the others company
                __syncthreads();   // tile start
                __syncthreads();   // tile end
                continue;
            }
        }
        const PencilTileDev* tp = a.tiles + ti;
        const int4 t0 = *reinterpret_cast<const int4*>(&tp->base);      // base, w, wj, wk
        const int4 tB = *reinterpret_cast<const int4*>(tp->nbrBase);
        const int tbase = t0.x, w = t0.y, wj = t0.z, wk = t0.w;
        const int skewMax = SKEW * ((wj - 1) + (wk - 1));
        const int S = nx + skewMax;
        const int S8 = (S + kPencilD - 1) & ~(kPencilD - 1);   // the pipeline runs whole record groups; the extra steps are idle
        const unsigned group0 = ggroup0;
        ggroup0 += unsigned(S8 / kPencilD);
        // G > 1: every tile completes whole laps of the neighbour-value barriers (arrivals padded to E32 steps)
        const int E32 = (S + kPencilE - 1) & ~(kPencilE - 1);
        const unsigned ext0 = G > 1 ? gext0 : 0u;
        gext0 += unsigned(E32 / kPencilE);
        // neighbour tiles: chain side = where the new values come from, static side = old values (Gauss-Seidel) and
        // the tiles that wait for our results
        const int baseCJ = DIR > 0 ? tB.x : tB.z, baseCK = DIR > 0 ? tB.y : tB.w;
        const int baseSJ = DIR > 0 ? tB.z : tB.x, baseSK = DIR > 0 ? tB.w : tB.y;
        const bool hasCJ = baseCJ >= 0, hasCK = baseCK >= 0;
        const bool hasSJ = GS && baseSJ >= 0, hasSK = GS && baseSK >= 0;
        const bool anyExt = hasCJ || hasCK || hasSJ || hasSK;
        // K faces handed over inside the CTA (G > 1): the chain-side K neighbour is the tile of the pipeline next door
        const bool ctaCK = G > 1 && hasCK && prodTile >= 0 && !(a.debug & 4);   // (debug bit 4: K faces through L2 after all)
        const bool ctaCons = G > 1 && consTile >= 0 && !(a.debug & 4);
        // ring columns of a step: [chain J (wk) | chain K (wj) | static J (wk) | static K (wj)]
        const int colCK = wk, colSJ = wk + wj, colSK = 2 * wk + wj;
        const unsigned extRowB = unsigned(a.extW) * 8;
        // lane geometry (chain, prep and writer warps)
        const bool laneOn = lane < w;
        const int jj = laneOn ? lane % wj : 0, kk = laneOn ? lane / wj : 0;
        const int jr = DIR > 0 ? jj : wj - 1 - jj, kr = DIR > 0 ? kk : wk - 1 - kk;
        const int skew = SKEW * (jr + kr);
        const bool extJ = (jr == 0), extK = (kr == 0);   // chain-side values come from the neighbour ring

        if (role == 1) {
            // ------------------------------------------------------------------------------------------------
            // helper warp: neighbour values
            // ------------------------------------------------------------------------------------------------
            if (lane == 0) {
                // the per-step "neighbour values are in the ring" barriers start every tile in phase 0
                // (arrivals per step: this warp if it fetches any column, the producing pipeline's writer if the K face
                //  comes from inside the CTA)
                if (G == 1) {
                    for (int q = 0; q < kPencilE; q++) mbar_init(extFull + q * 8, 1);
                    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                }
                st_release_cta(wChainProg, 0);
                st_release_cta(wWriterProg, 0);
            }
            // columns without a source tile hold a constant for the whole tile
            for (int e = lane; e < kPencilE * a.extW; e += 32)
                sts_f64(extRing + unsigned(e) * 8, (e % a.extW) < colSJ ? NEUTRAL : 0.0);
            __syncthreads();   // tile start: rings initialised, progress words reset

            // ---- every lane streams up to kPencilCPL columns of the ring (one source pencil of a neighbour tile each) ----
            const int4 tW = *reinterpret_cast<const int4*>(tp->nbrW);
            const int4 tJ = *reinterpret_cast<const int4*>(tp->nbrWj);
            const bool hasCKh = hasCK && !ctaCK;   // K-face columns this warp has to fetch from L2
            const int nCols = (hasCJ ? wk : 0) + (hasCKh ? wj : 0) + (hasSJ ? wk : 0) + (hasSK ? wj : 0);
            const double* cSrc[kPencilCPL];   // element (row 0) of the source pencil; nullptr: no column
            int cStride[kPencilCPL];          // row stride of the source tile
            int cSkew[kPencilCPL];            // skew of the lane of this tile that consumes the column
            unsigned cDst[kPencilCPL];        // ring address of the column in row 0
            bool cChain[kPencilCPL];          // new values (polled) or old values (plain loads)
            int cProg[kPencilCPL];            // steps of the column already in the ring
#pragma unroll
            for (int c = 0; c < kPencilCPL; c++) {
                int q = c * 32 + lane;
                cSrc[c] = nullptr;
                cStride[c] = cSkew[c] = 0;
                cDst[c] = 0;
                cChain[c] = false;
                cProg[c] = S;   // a lane without column never holds anyone back
                if (q < nCols) {
                    // which group does column q of the compacted list belong to: 0 chain J, 1 chain K, 2 static J, 3 static K
                    int grp = -1, idx = 0;
                    if (hasCJ) { if (grp < 0 && q < wk) { grp = 0; idx = q; } q -= wk; }
                    if (hasCKh) { if (grp < 0 && q >= 0 && q < wj) { grp = 1; idx = q; } q -= wj; }
                    if (hasSJ) { if (grp < 0 && q >= 0 && q < wk) { grp = 2; idx = q; } q -= wk; }
                    if (hasSK) { if (grp < 0 && q >= 0 && q < wj) { grp = 3; idx = q; } q -= wj; }
                    const bool isJ = (grp == 0 || grp == 2), isChain = grp < 2;
                    // lower-side neighbours ((J-1,K), (J,K-1)) touch our low face with their high face; upper side reversed
                    const bool lowSide = (DIR > 0) == isChain;
                    const int nBase = grp == 0 ? baseCJ : grp == 1 ? baseCK : grp == 2 ? baseSJ : baseSK;
                    const int nW = isJ ? (lowSide ? tW.x : tW.z) : (lowSide ? tW.y : tW.w);
                    const int nWj = isJ ? (lowSide ? tJ.x : tJ.z) : (lowSide ? tJ.y : tJ.w);
                    const int nWk = nW / nWj;
                    int ej, ek, srcLane;
                    if (isJ) {
                        ek = idx;
                        ej = lowSide ? 0 : wj - 1;
                        srcLane = (lowSide ? nWj - 1 : 0) + nWj * ek;
                    } else {
                        ej = idx;
                        ek = lowSide ? 0 : wk - 1;
                        srcLane = ej + nWj * (lowSide ? nWk - 1 : 0);
                    }
                    const int ejr = DIR > 0 ? ej : wj - 1 - ej, ekr = DIR > 0 ? ek : wk - 1 - ek;
                    const int col = (grp == 0 ? 0 : grp == 1 ? colCK : grp == 2 ? colSJ : colSK) + idx;
                    cSrc[c] = (isChain ? a.out : a.plane[8]) + nBase + srcLane;
                    cStride[c] = nW;
                    cSkew[c] = SKEW * (ejr + ekr);
                    cDst[c] = extRing + unsigned(col) * 8;
                    cChain[c] = isChain;
                    cProg[c] = 0;
                }
            }

            // Rounds: every lane loads the next kPencilLook steps of its columns (all loads in flight together), deposits
            // the values that have arrived -- a producer publishes a pencil in order, so they form a prefix -- and the
            // steps complete in EVERY column are handed over (one barrier per step).
            unsigned long long pRounds = 0, pCap = 0;
            if (nCols > 0 || G > 1) {
                int published = 0, chainProg = 0;
                unsigned spins = 0;
                const int target = G > 1 ? E32 : S;   // G > 1: also the idle steps that complete the last ring lap
                while (published < target) {
                    pRounds++;
                    if (a.debug & 2) __nanosleep(1000);
                    // ring capacity: rows of steps the chain warp (the last reader) has finished may be overwritten
                    const int limit = min(S, chainProg + kPencilE);
                    double v[kPencilCPL][kPencilLook];
#pragma unroll
                    for (int c = 0; c < kPencilCPL; c++) {
                        if (c * 32 >= nCols) break;   // (uniform) columns per lane actually in use
#pragma unroll
                        for (int d = 0; d < kPencilLook; d++) {
                            const int st = cProg[c] + d, r = st - cSkew[c];
                            v[c][d] = cChain[c] ? NEUTRAL : 0.0;   // rows outside the block: a constant
                            if (cSrc[c] && st < limit && unsigned(r) < unsigned(nx)) {
                                const double* p = cSrc[c] + size_t(DIR > 0 ? r : nx - 1 - r) * size_t(cStride[c]);
                                v[c][d] = cChain[c] ? ld_l2(p) : ld_cg(p);
                            }
                        }
                    }
                    int myMin = S;
#pragma unroll
                    for (int c = 0; c < kPencilCPL; c++) {
                        if (c * 32 >= nCols) break;
                        if (cSrc[c]) {
                            int cnt = 0;
#pragma unroll
                            for (int d = 0; d < kPencilLook; d++) {
                                const int st = cProg[c] + d;
                                if (cnt == d && st < limit && !(cChain[c] && is_sentinel(v[c][d]))) {
                                    sts_f64(cDst[c] + unsigned(st & (kPencilE - 1)) * extRowB, v[c][d]);
                                    cnt++;
                                }
                            }
                            cProg[c] += cnt;
                            myMin = min(myMin, cProg[c]);
                        }
                    }
                    int ready = __reduce_min_sync(0xffffffffu, myMin);
                    if (G > 1) {
                        // lanes without a column report S: steps are still only handed over inside the ring window,
                        // and past S (no data) the window is the only condition
                        ready = min(ready, limit);
                        if (ready >= S) ready = min(E32, chainProg + kPencilE);
                    }
                    __syncwarp();
                    if (ready > published) {
                        // arrive = release of the ring stores above; the chain / prep warps sleep on these barriers
                        if (lane == 0)
                            for (int q = published; q < ready; q++) {
                                mbar_arrive(extFull + unsigned(q & (kPencilE - 1)) * 8);
                                // G > 1: two arrivals per step -- the second one is the in-CTA producer's for the
                                // steps it deposits (q < S), ours otherwise
                                if (G > 1 && !(ctaCK && q < S)) mbar_arrive(extFull + unsigned(q & (kPencilE - 1)) * 8);
                            }
                        published = ready;
                        spins = 0;
                    } else if (++spins > kMaxSpins) {
                        *a.err = 1;
                        break;
                    }
                    if (published + kPencilLook > chainProg + kPencilE) {
                        const long long c0 = a.prof ? clock64() : 0;
                        chainProg = ld_acquire_cta(wChainProg);
                        if (a.prof) pCap += clock64() - c0;
                    }
                }
            }
            if (a.prof && lane == 0) {
                unsigned long long* q = a.prof + size_t(ti) * kPencilProfWords;
                q[8] = pRounds;
                q[9] = pCap;
            }
            __syncthreads();   // tile end
        } else if (role == 2) {
            // ------------------------------------------------------------------------------------------------
            // prep warp: raw ring (skewed rows) + neighbour ring -> step records
            // ------------------------------------------------------------------------------------------------
            // Gauss-Seidel: the old values of the other side live in the raw ring (same tile) or the neighbour ring
            const bool sExtJ = (jr == wj - 1), sExtK = (kr == wk - 1);
            const int oJoff = sExtJ ? 0 : DIR * 8, oKoff = sExtK ? 0 : DIR * wj * 8;
            const unsigned wB = unsigned(w) * 8;
            const unsigned ringLane = dataRing + unsigned(lane) * 8;
            const unsigned sJaddr = extRing + unsigned(colSJ + kk) * 8, sKaddr = extRing + unsigned(colSK + jj) * 8;
            const unsigned recLane = recRing + unsigned(lane) * T::REC_LANE_BYTES;
            const unsigned dotLane = dotRing + unsigned(lane) * 8;
            const bool asym = MODE == PM_FACTOR && a.plane[4] != nullptr;
            __syncthreads();   // tile start

            // raw-ring address of the operands of processing row r (any r: an idle lane's record is zeroed anyway)
            auto rowOff = [&](int r) -> unsigned {
                const int i = DIR > 0 ? r : nx - 1 - r;
                const unsigned st = (gchunk0 + unsigned((r + pad) >> 3)) & (NS - 1);
                return ringLane + st * T::STAGE_BYTES + unsigned(i & (R - 1)) * wB;
            };
            // bulk copies of one chunk of every operand plane into its ring stage (one lane; the stage is free: either it
            // has never been used by this tile or this warp has just finished reading its previous occupant)
            auto issueChunk = [&](int cL) {
                const unsigned gc = gchunk0 + unsigned(cL);
                const unsigned st = gc & (NS - 1);
                const int i0 = (DIR > 0 ? cL : nChunks - 1 - cL) * R;   // rows of the chunk in memory order
                const int rows = min(R, nx - i0);
                const unsigned bytes = (unsigned(rows) * unsigned(w) * 8u + 15u) & ~15u;
                int np = 0;
#pragma unroll
                for (int p = 0; p < NP; p++) np += a.plane[p] ? 1 : 0;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // our reads of the stage precede the copy
                mbar_arrive_expect_tx(barFull + st * 8, bytes * unsigned(np));
                const size_t e0 = size_t(tbase) + size_t(i0) * size_t(w);
#pragma unroll
                for (int p = 0; p < NP; p++)
                    if (a.plane[p])
                        bulk_g2s(dataRing + st * T::STAGE_BYTES + p * kPencilPlaneBytes, a.plane[p] + e0, bytes,
                                 barFull + st * 8);
            };
            if (prepId == 0 && lane == 0)
                for (int cL = 0; cL < min(NS, nChunks); cL++) issueChunk(cL);
            int waitRow = 0;                          // first processing row of the next chunk to wait for
            int waitStep = 0;                         // step at which that wait falls due
            int relChunk = 0;
            int relStep = min(nx - 1, R - pad - 1) + skewMax;   // last step that reads the oldest chunk
            unsigned long long pData = 0, pExt = 0, pSlot = 0;
            for (int s = prepId; s < S8; s += kPencilPrep) {
                const int r = s - skew;
                const bool act = laneOn && unsigned(r) < unsigned(nx);
                // barrier tests first (their latency hides behind the loads): neighbour values of this step in the ring,
                // record slot taken by the chain warp
                const unsigned eb = extFull + unsigned(s & (kPencilE - 1)) * 8, ep = unsigned(s >> 5) & 1;
                const unsigned grp = group0 + unsigned(s >> 3), slot = unsigned(s & 7);
                const bool needExt = GS && (hasSJ || hasSK) && s < S;   // old values of the neighbour tiles
                const bool extOk = !needExt || mbar_test_wait(eb, ep);
                const bool slotOk = grp == 0 || mbar_test_wait(recEmpty + slot * 8, (grp - 1) & 1);
                if (s >= waitStep) {                  // chunks holding processing rows <= s + LA
                    const long long c0 = a.prof ? clock64() : 0;
                    while (waitRow < nx && waitRow <= s + LA) {
                        const unsigned gc = gchunk0 + unsigned((waitRow + pad) >> 3);
                        mbar_wait(barFull + (gc & (NS - 1)) * 8, (gc / NS) & 1, a.err);
                        waitRow = (((waitRow + pad) >> 3) + 1) * R - pad;
                    }
                    waitStep = waitRow < nx ? waitRow - LA : kNever;
                    if (a.prof) pData += clock64() - c0;
                }
                const unsigned off = rowOff(r);
                double c[NP];
#pragma unroll
                for (int p = 0; p < NP; p++) c[p] = lds_f64(off + p * kPencilPlaneBytes);
                if (!extOk) {
                    const long long c0 = a.prof ? clock64() : 0;
                    mbar_wait(eb, ep, a.err);
                    if (a.prof) pExt += clock64() - c0;
                }
                const unsigned er = unsigned(s & (kPencilE - 1)) * extRowB;
                double oI = 0.0, oJ = 0.0, oK = 0.0;
                if (GS) {
                    const double esJ = lds_f64(sJaddr + er), esK = lds_f64(sKaddr + er);
                    oJ = lds_f64(off + 8 * kPencilPlaneBytes + oJoff);
                    oK = lds_f64(off + 8 * kPencilPlaneBytes + oKoff);
                    oI = lds_f64(rowOff(r + 1) + 8 * kPencilPlaneBytes);
                    if (sExtJ) oJ = esJ;
                    if (sExtK) oK = esK;
                    if (!(unsigned(r + 1) < unsigned(nx))) oI = 0.0;   // no such row: exact +0
                }
                // the record: everything of the step that does not depend on new values
                double rec[2 * NV];
                if (MODE == PM_FWD) {
                    // planes: in, rD, t_K, t_J, t_I (t = rD*lower, +0 where there is no face)
                    rec[0] = c[1] * c[0];
                    rec[1] = c[2];
                    rec[2] = c[3];
                    rec[3] = c[4];
                } else if (MODE == PM_BWD) {
                    // planes: in, t_K, t_J, t_I (t = rD*upper), dotWith
                    rec[0] = c[0];
                    rec[1] = c[1];
                    rec[2] = c[2];
                    rec[3] = c[3];
                } else if (MODE == PM_FACTOR) {
                    // planes: diag, l_K, l_J, l_I, u_K, u_J, u_I of the lower-side faces (symmetric: no u planes)
                    rec[0] = c[0];
                    rec[1] = (asym ? c[4] : c[1]) * c[1];
                    rec[2] = (asym ? c[5] : c[2]) * c[2];
                    rec[3] = (asym ? c[6] : c[3]) * c[3];
                } else if (MODE == PM_GS_FWD) {
                    // planes: b, diag, l_K, l_J, l_I, u_I, u_J, u_K, old
                    rec[0] = c[0];
                    rec[1] = c[2];
                    rec[2] = c[3];
                    rec[3] = c[4];
                    rec[4] = c[5] * oI;
                    rec[5] = c[6] * oJ;
                    rec[6] = c[7] * oK;
                    rec[7] = c[1];
                } else {
                    // reverse half of symGaussSeidel: the lower terms use the forward values, summed here in order
                    double acc = c[0];
                    acc -= c[2] * oK;
                    acc -= c[3] * oJ;
                    acc -= c[4] * oI;
                    rec[0] = acc;
                    rec[1] = c[5];
                    rec[2] = c[6];
                    rec[3] = c[7];
                    rec[4] = c[1];
                    rec[5] = 0.0;
                }
                if (!act) {
                    // idle lane: a record whose result is NEUTRAL whatever (finite) values the neighbours hold
#pragma unroll
                    for (int q = 0; q < 2 * NV; q++) rec[q] = 0.0;
                    if (MODE == PM_FACTOR) rec[0] = 1.0;
                    if (MODE == PM_GS_FWD) rec[7] = 1.0;
                    if (MODE == PM_GS_REV) rec[4] = 1.0;
                }
                // wait until the chain warp has taken the previous occupant of the slot, write, hand over (release)
                const unsigned ra = recLane + slot * T::REC_STEP_BYTES;
                {
                    if (!slotOk) {
                        const long long c0 = a.prof ? clock64() : 0;
                        mbar_wait(recEmpty + slot * 8, (grp - 1) & 1, a.err);
                        if (a.prof) pSlot += clock64() - c0;
                    }
#pragma unroll
                    for (int q = 0; q < NV; q++) sts_v2(ra + q * 16, rec[2 * q], rec[2 * q + 1]);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(recFull + slot * 8);
                }
                if (doDot) sts_f64(dotLane + unsigned(s & (kPencilOutRows - 1)) * 256u, act ? c[4] : 0.0);
                // Once the most skewed lane has read the last row of the oldest chunk (step relStep) no step of this warp
                // touches the chunk any more; when that holds for every prep warp its stage takes the next chunk.
                while (s + kPencilPrep > relStep) {
                    const unsigned gc = gchunk0 + unsigned(relChunk);
                    const unsigned db = rawDone + (gc & (NS - 1)) * 8;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(db);
                    if (prepId == 0 && relChunk + NS < nChunks) {
                        mbar_wait(db, (gc / NS) & 1, a.err);
                        if (lane == 0) issueChunk(relChunk + NS);
                    }
                    relChunk++;
                    const int nextFirst = relChunk * R - pad;
                    relStep = nextFirst < nx ? min(nx - 1, nextFirst + R - 1) + skewMax : kNever;
                }
            }
            if (a.prof && lane == 0 && prepId == 0) {
                unsigned long long* q = a.prof + size_t(ti) * kPencilProfWords;
                q[4] = pData;
                q[6] = pSlot;
                q[11] = pExt;
            }
            __syncthreads();   // tile end
        } else if (role == 0) {
            // ------------------------------------------------------------------------------------------------
            // chain warp
            // ------------------------------------------------------------------------------------------------
            const int srcJ = (extJ || !laneOn) ? lane : lane - DIR;   // shuffle sources
            const int srcK = (extK || !laneOn) ? lane : lane - DIR * wj;
            const unsigned recLane = recRing + unsigned(lane) * T::REC_LANE_BYTES;
            const unsigned outLane = outRing + unsigned(lane) * 8;
            __syncthreads();   // tile start
            const unsigned long long pStart = a.prof ? globaltimer_ns() : 0;
            const long long pClk0 = a.prof ? clock64() : 0;
            unsigned long long pRec = 0, pOut = 0, pExt = 0;
            const unsigned eJaddr = extRing + unsigned(kk) * 8, eKaddr = extRing + unsigned(colCK + jj) * 8;
            double y1 = NEUTRAL, sJ = NEUTRAL, sK = NEUTRAL;
            int writerProg = 0;
            // Records are taken one step ahead: the wait and the loads of step s+1 are in flight during step s.  The
            // loop is unrolled over the eight slots of the record ring, so slot and barrier addresses are immediates.
            // (both barrier tests are issued first and the loads go out speculatively, so the test latencies overlap each
            // other and the arithmetic of the current step; only when a test fails is the warp put to sleep and the
            // loads repeated)
            const bool chainExt = hasCJ || hasCK;
            auto fetch = [&](double2 (&v)[NV], double2& e, int slot, unsigned parity, int s) {
                const unsigned rb = recFull + slot * 8;
                const unsigned eb = extFull + unsigned(s & (kPencilE - 1)) * 8, ep = (ext0 + unsigned(s >> 5)) & 1;
                const unsigned er = unsigned(s & (kPencilE - 1)) * extRowB;
                const bool needExt = chainExt && s < S;
                const bool recOk = mbar_test_wait(rb, parity);
                const bool extOk = !needExt || mbar_test_wait(eb, ep);
#pragma unroll
                for (int q = 0; q < NV; q++) v[q] = lds_v2(recLane + slot * T::REC_STEP_BYTES + q * 16);
                e.x = lds_f64(eKaddr + er);
                e.y = lds_f64(eJaddr + er);
                return recOk && extOk;
            };
            auto refetch = [&](double2 (&v)[NV], double2& e, int slot, unsigned parity, int s) {   // slow path
                const long long c0 = a.prof ? clock64() : 0;
                mbar_wait(recFull + slot * 8, parity, a.err);
                const long long c1 = a.prof ? clock64() : 0;
                if (chainExt && s < S) mbar_wait(extFull + unsigned(s & (kPencilE - 1)) * 8, (ext0 + unsigned(s >> 5)) & 1, a.err);
                if (a.prof) {
                    pRec += c1 - c0;
                    pExt += clock64() - c1;
                }
                const unsigned er = unsigned(s & (kPencilE - 1)) * extRowB;
#pragma unroll
                for (int q = 0; q < NV; q++) v[q] = lds_v2(recLane + slot * T::REC_STEP_BYTES + q * 16);
                e.x = lds_f64(eKaddr + er);
                e.y = lds_f64(eJaddr + er);
            };
            auto step = [&](const double2 (&v)[NV], const double2& e, int slot, unsigned outBase, int o16) {
                if (SKEW == 1) {
                    sJ = __shfl_sync(0xffffffffu, y1, srcJ);
                    sK = __shfl_sync(0xffffffffu, y1, srcK);
                }
                const double vK = extK ? e.x : sK, vJ = extJ ? e.y : sJ;
                double acc, y;
                if (MODE == PM_FWD || MODE == PM_BWD) {
                    // record: {rD*in | in, t_K} {t_J, t_I}
                    acc = v[0].x;
                    acc -= v[0].y * vK;
                    acc -= v[1].x * vJ;
                    acc -= v[1].y * y1;
                    y = acc;
                } else if (MODE == PM_FACTOR) {
                    // record: {diag, u_K*l_K} {u_J*l_J, u_I*l_I}
                    acc = v[0].x;
                    acc -= v[0].y / vK;
                    acc -= v[1].x / vJ;
                    acc -= v[1].y / y1;
                    y = acc;
                } else if (MODE == PM_GS_FWD) {
                    // record: {b, l_K} {l_J, l_I} {u_I*old_I, u_J*old_J} {u_K*old_K, diag}
                    acc = v[0].x;
                    acc -= v[0].y * vK;
                    acc -= v[1].x * vJ;
                    acc -= v[1].y * y1;
                    acc -= v[2].x;
                    acc -= v[2].y;
                    acc -= v[3].x;
                    y = acc / v[3].y;
                } else {
                    // record: {b - lower terms, u_I} {u_J, u_K} {diag, -}
                    acc = v[0].x;
                    acc -= v[0].y * y1;
                    acc -= v[1].x * vJ;
                    acc -= v[1].y * vK;
                    y = acc / v[2].x;
                }
                if (SKEW == 2) {
                    sJ = __shfl_sync(0xffffffffu, y1, srcJ);
                    sK = __shfl_sync(0xffffffffu, y1, srcK);
                }
                y1 = y;
                sts_f64(outBase + slot * 256, y);
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(outFull + (o16 + slot) * 8);   // result written (release): the writer sleeps on this barrier
                    mbar_arrive(recEmpty + slot * 8);           // the loads of this record have long completed
                }
            };
            double2 vA[NV], vB[NV], eA, eB;
            if (!fetch(vA, eA, 0, group0 & 1, 0)) refetch(vA, eA, 0, group0 & 1, 0);
            for (int s0 = 0; s0 < S8; s0 += kPencilD) {
                const unsigned par = (group0 + unsigned(s0 >> 3)) & 1;
                const unsigned outBase = outLane + unsigned(s0 & (kPencilOutRows - 1)) * 256u;
                const int o16 = int((group0 + unsigned(s0 >> 3)) & 1) * 8;   // result barriers of this group
                if (a.prof && lane == 0 && (s0 >> 3) < 48) a.prof[size_t(ti) * kPencilProfWords + 16 + (s0 >> 3)] = globaltimer_ns();
                // the writer waits on 16 result barriers by phase parity: it must have finished the group before last
                if (writerProg < s0 - 8) {
                    const long long c0 = a.prof ? clock64() : 0;
                    unsigned spins = 0;
                    while (writerProg < s0 - 8) {
                        __nanosleep(40);
                        writerProg = ld_acquire_cta(wWriterProg);
                        if (++spins > kMaxSpins) {
                            *a.err = 1;
                            break;
                        }
                    }
                    if (a.prof) pOut += clock64() - c0;
                }
#pragma unroll
                for (int k = 0; k < kPencilD; k += 2) {
                    const bool okB = fetch(vB, eB, k + 1, par, s0 + k + 1);
                    step(vA, eA, k, outBase, o16);
                    if (!okB) refetch(vB, eB, k + 1, par, s0 + k + 1);
                    const int kn = (k + 2) & (kPencilD - 1);
                    const unsigned pn = k + 2 < kPencilD ? par : par ^ 1;
                    const bool okA = fetch(vA, eA, kn, pn, s0 + k + 2);   // (past the last group: never consumed)
                    step(vB, eB, k + 1, outBase, o16);
                    if (!okA && s0 + k + 2 < S8) refetch(vA, eA, kn, pn, s0 + k + 2);
                }
                // tell the helper how far the neighbour ring has been consumed
                if ((anyExt || G > 1) && lane == 0) st_release_cta(wChainProg, s0 + kPencilD - 1);
            }
            if (a.prof && lane == 0) {
                unsigned long long* q = a.prof + size_t(ti) * kPencilProfWords;
                q[0] = pStart;
                q[1] = globaltimer_ns();
                q[2] = pRec;
                q[3] = pOut;
                q[10] = (unsigned long long)(clock64() - pClk0);
                q[5] = pExt;
            }
            __syncthreads();   // tile end
        } else {
            // ------------------------------------------------------------------------------------------------
            // writer warp
            // ------------------------------------------------------------------------------------------------
            // the lanes on the high faces of the tile, whose values a neighbour tile is waiting for, publish at once
            const bool faceLane = laneOn && ((jr == wj - 1 && baseSJ >= 0) || (kr == wk - 1 && baseSK >= 0 && !ctaCons));
            // in-CTA consumer of the K face: ring, per-step barriers and progress word of the pipeline next door
            const unsigned consBase = smCta + unsigned(sub + DIR) * subBytes;
            const unsigned consExtFull = consBase + 320, consProgW = consBase + 640;
            const int consWk = ctaCons ? a.tiles[consTile].wk : 1;
            const int consS = nx + SKEW * ((wj - 1) + (consWk - 1));
            const unsigned consCol = consBase + kPencilHeaderBytes + unsigned(consWk + jj) * 8;   // column colCK + jj
            const int consShift = SKEW * (wk - 1);   // consumer step = producer step - shift
            const bool kFace = laneOn && kr == wk - 1;
            int consProg = 0;
            auto deposit = [&](int sc, double val) {   // K-face values of consumer step sc (uniform over the warp)
                if (sc >= consProg + kPencilE) {       // ring capacity: wait for the consumer's chain warp
                    unsigned spins = 0;
                    while (sc >= consProg + kPencilE) {
                        consProg = ld_acquire_cta(consProgW);
                        if (++spins > kMaxSpins) {
                            *a.err = 1;
                            break;
                        }
                    }
                }
                if (kFace) sts_f64(consCol + unsigned(sc & (kPencilE - 1)) * extRowB, val);
                __syncwarp();
                if (lane == 0) mbar_arrive(consExtFull + unsigned(sc & (kPencilE - 1)) * 8);
            };
            const int delay = skewMax - skew;   // steps between this lane's result of a row and the row being complete
            const unsigned outLane = outRing + unsigned(lane) * 8;
            const unsigned dotLane = dotRing + unsigned(lane) * 8;
            const bool doClear = a.clear != nullptr;
            int elem = tbase + lane + (DIR > 0 ? -skew : nx - 1 + skew) * w;            // position of this lane's row of step 0
            int elemRow = tbase + lane + (DIR > 0 ? -skewMax : nx - 1 + skewMax) * w;   // ... of the row completed at step 0
            __syncthreads();   // tile start
            unsigned long long pWait = 0;
            for (int s = 0; s < S8; s++) {
                const unsigned slotS = outLane + unsigned(s & (kPencilOutRows - 1)) * 256u;
                {
                    // sleep until the chain warp has stored the results of this step
                    const unsigned gs = group0 * kPencilD + unsigned(s);
                    const unsigned ob = outFull + (gs & 15) * 8, op = (gs >> 4) & 1;
                    if (!mbar_test_wait(ob, op)) {
                        const long long c0 = a.prof ? clock64() : 0;
                        mbar_wait(ob, op, a.err);
                        if (a.prof) pWait += clock64() - c0;
                    }
                }
                const double v = lds_f64(slotS);
                const int r = s - skew;
                const bool act = laneOn && unsigned(r) < unsigned(nx);
                if (act && faceLane) st_l2(a.out + elem, v);
                if (ctaCons && s >= consShift && s - consShift < consS) deposit(s - consShift, act ? v : NEUTRAL);
                const int q = s - skewMax;   // processing row every lane has finished now
                if (laneOn && unsigned(q) < unsigned(nx)) {
                    const unsigned slotQ = outLane + unsigned((s - delay) & (kPencilOutRows - 1)) * 256u;
                    const double vq = lds_f64(slotQ);   // this lane's own result `delay` steps ago (== v when delay is 0)
                    st_l2(a.out + elemRow, vq);
                    if (MODE == PM_FACTOR) a.out2[elemRow] = 1.0 / vq;
                    if (doClear) a.clear[elemRow] = sent;
                    if (doDot) dsum[0] += vq * lds_f64(dotLane + unsigned((s - delay) & (kPencilOutRows - 1)) * 256u);
                }
                elem += DIR * w;
                elemRow += DIR * w;
                if ((s & 7) == 7 && lane == 0) st_release_cta(wWriterProg, s + 1);
            }
            // consumer steps past our own last one only see rows outside the block
            if (ctaCons)
                for (int sc = max(0, S8 - consShift); sc < consS; sc++) deposit(sc, NEUTRAL);
            if (a.prof && lane == 0) a.prof[size_t(ti) * kPencilProfWords + 7] = pWait;
            __syncthreads();   // tile end
        }
        gchunk0 += nChunks;
    }
    if (MODE == PM_BWD && a.dotOut) {
        if (grid_reduce<1>(dsum, a.partials, a.ticket)) a.dotOut[0] = dsum[0];
    }
}

// ------------------------------------------------------------------------------------------------------------
// Amul on a pencil level: a 7-point stencil on the tile-major layout
// ------------------------------------------------------------------------------------------------------------

// wA = A x (lduMatrixATmul.C:34-92) for a structured block in tile-major positions, optionally fused with wA.x
// (PCG.C:159-161).  Neighbour positions follow from the tile geometry (same tile: +-1 along j, +-wj along k, +-w along
// i; across a tile face: the facing pencil of the neighbour tile), so no row pointers or column indices are read: per
// row the diagonal, x, the result and six coefficients -- the upper-side planes cU (I+, J+, K+) of the row and, for the
// lower side, either the planes cL (K-, J-, I-) or (SYM) the neighbours' own cU entries of the shared faces, which the
// neighbouring threads load anyway.  Absent faces carry +0 coefficients and point at the row itself.  Accumulation
// order = k_spmv: diagonal, K-, J-, I-, I+, J+, K+.  One work item = 256 consecutive rows of one tile; grid-stride
// over the items so that the block partials of the fused dot product fit the reduction scratch.
template <bool SYM, bool DOT>
__global__ void __launch_bounds__(256)
k_pencil_spmv(double* __restrict__ out, const double* __restrict__ x, const double* __restrict__ diag,
              const double* __restrict__ cL, const double* __restrict__ cU, size_t np,
              const PencilTileDev* __restrict__ tiles, int nTiles, int nx, int chunksPerTile,
              double* __restrict__ dotOut, double* __restrict__ partials, unsigned int* __restrict__ ticket,
              const int* __restrict__ stop) {
    if (stop && *stop) return;
    double v[1] = {0.0};
    const int nItems = nTiles * chunksPerTile;
    for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
        const int ti = item / chunksPerTile;
        const int local = (item - ti * chunksPerTile) * 256 + threadIdx.x;
        const PencilTileDev* tp = tiles + ti;
        const int4 t0 = *reinterpret_cast<const int4*>(&tp->base);   // base, w, wj, wk
        const int w = t0.y, wj = t0.z, wk = t0.w;
        if (local >= nx * w) continue;
        const int4 nB = *reinterpret_cast<const int4*>(tp->nbrBase);
        const int4 nW = *reinterpret_cast<const int4*>(tp->nbrW);
        const int4 nJ = *reinterpret_cast<const int4*>(tp->nbrWj);
        const int i = local / w, ln = local - i * w;
        const int kk = ln / wj, jj = ln - kk * wj;
        const int p = t0.x + local;
        // neighbour positions (p itself where there is no cell: the coefficient is +0 there)
        const int qIm = i > 0 ? p - w : p;
        const int qIp = i < nx - 1 ? p + w : p;
        const int qJm = jj > 0 ? p - 1 : (nB.x >= 0 ? nB.x + i * nW.x + (nJ.x - 1) + nJ.x * kk : p);
        const int qJp = jj < wj - 1 ? p + 1 : (nB.z >= 0 ? nB.z + i * nW.z + nJ.z * kk : p);
        const int qKm = kk > 0 ? p - wj : (nB.y >= 0 ? nB.y + i * nW.y + jj + nJ.y * (nW.y / nJ.y - 1) : p);
        const int qKp = kk < wk - 1 ? p + wj : (nB.w >= 0 ? nB.w + i * nW.w + jj : p);
        const double xp = x[p];
        double lK, lJ, lI;
        if (SYM) {
            // the lower coefficient of a face equals its upper coefficient, stored with the owner (the lower-side
            // neighbour): slot K+ of q_K-, J+ of q_J-, I+ of q_I-
            lK = qKm != p ? cU[2 * np + qKm] : 0.0;
            lJ = qJm != p ? cU[np + qJm] : 0.0;
            lI = qIm != p ? cU[qIm] : 0.0;
        } else {
            lK = cL[p];
            lJ = cL[np + p];
            lI = cL[2 * np + p];
        }
        double acc = diag[p] * xp;
        acc += lK * x[qKm];
        acc += lJ * x[qJm];
        acc += lI * x[qIm];
        acc += cU[p] * x[qIp];
        acc += cU[np + p] * x[qJp];
        acc += cU[2 * np + p] * x[qKp];
        out[p] = acc;
        if (DOT) v[0] += acc * xp;
    }
    if (DOT) {
        if (grid_reduce<1>(v, partials, ticket)) dotOut[0] = v[0];
    }
}

// ------------------------------------------------------------------------------------------------------------
// coefficient planes of a pencil level
// ------------------------------------------------------------------------------------------------------------

// From the native CSR triangles: cL[s][p] = lower-side coefficient of row p in slot s = (K-, J-, I-), cLu = the upper
// coefficient of the same faces, cU[s][p] = upper-side coefficient in slot s = (I+, J+, K+); +0 where the face does not
// exist.  grid = (blocks over the rows of a tile, tiles).
__global__ void k_pencil_planes(double* __restrict__ cL, double* __restrict__ cLu, double* __restrict__ cU,
                                const PencilTileDev* __restrict__ tiles, int nx, int ny, int nz, size_t n,
                                const int* __restrict__ Lptr, const double* __restrict__ Lval,
                                const int* __restrict__ LtoU, const int* __restrict__ Uptr,
                                const double* __restrict__ Uval) {
    const PencilTileDev t = tiles[blockIdx.y];
    const int rows = nx * t.w;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < rows; q += gridDim.x * blockDim.x) {
        const int i = q / t.w, l = q - i * t.w;
        const int j = t.j0 + l % t.wj, k = t.k0 + l / t.wj;
        const size_t p = size_t(t.base) + q;
        int e = Lptr[p];
        const bool has[3] = {k > 0, j > 0, i > 0};
#pragma unroll
        for (int s = 0; s < 3; s++) {
            double lo = 0.0, up = 0.0;
            if (has[s]) {
                lo = Lval[e];
                up = Uval[LtoU[e]];
                e++;
            }
            cL[s * n + p] = lo;
            if (cLu) cLu[s * n + p] = up;
        }
        e = Uptr[p];
        const bool hasU[3] = {i < nx - 1, j < ny - 1, k < nz - 1};
#pragma unroll
        for (int s = 0; s < 3; s++) {
            double up = 0.0;
            if (hasU[s]) up = Uval[e++];
            cU[s * n + p] = up;
        }
    }
}

// Per factorisation: tL[s] = rD*cL[s] (forward substitution, slots K-, J-, I-) and tU[s] = rD*cU[2-s] (backward
// substitution, slots in processing order K+, J+, I+); +0 where the face does not exist.
__global__ void k_pencil_pack(double* __restrict__ tL, double* __restrict__ tU, const double* __restrict__ cL,
                              const double* __restrict__ cU, const double* __restrict__ rD,
                              const PencilTileDev* __restrict__ tiles, int nx, int ny, int nz, size_t n) {
    const PencilTileDev t = tiles[blockIdx.y];
    const int rows = nx * t.w;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < rows; q += gridDim.x * blockDim.x) {
        const int i = q / t.w, l = q - i * t.w;
        const int j = t.j0 + l % t.wj, k = t.k0 + l / t.wj;
        const size_t p = size_t(t.base) + q;
        const double rd = rD[p];
        const bool hasL[3] = {k > 0, j > 0, i > 0};
        const bool hasU[3] = {k < nz - 1, j < ny - 1, i < nx - 1};
#pragma unroll
        for (int s = 0; s < 3; s++) {
            tL[s * n + p] = hasL[s] ? rd * cL[s * n + p] : 0.0;
            tU[s * n + p] = hasU[s] ? rd * cU[(2 - s) * n + p] : 0.0;
        }
    }
}

}  // namespace b200ls
