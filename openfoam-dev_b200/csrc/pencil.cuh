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
// A CTA is a warp pair:
//   * the CHAIN warp does the arithmetic, in exactly the order of kernels.cuh / the reference, and publishes every
//     result with one 8-byte L2 store (sentinel protocol, as the wavefront kernels);
//   * the HELPER warp (a) streams the per-row operands, which are contiguous per (tile, i-range) in the tile-major
//     layout, into a shared-memory ring with bulk async copies (cp.async.bulk + mbarrier complete_tx, one elected
//     lane), and (b) fetches the values the face lanes need from the neighbouring tiles a window of steps ahead:
//     polls them in L2 until none is the sentinel, deposits them in a small shared-memory ring indexed by step and
//     publishes its progress (release/acquire on a shared-memory word).  The chain warp never touches global memory
//     except for its stores.
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
static constexpr int kPencilMaxE = 6;                       // neighbour values per helper lane and window

struct PencilArgs {
    const PencilTileDev* tiles;
    const int* order;       // launch order of the forward sweeps; backward sweeps walk it from the end
    int nTiles;
    int nx, ny, nz;
    int extW;               // doubles per step of the neighbour-value ring (full tile)
    int window;             // steps per helper window
    // operand planes in position order (see k_pencil for the meaning per mode); unused entries are null
    const double* plane[kPencilMaxPlanes];
    double* out;            // sentinel-armed result
    double* out2;           // PM_FACTOR: reciprocal
    double* clear;          // optional: entry p is re-armed with the sentinel once row p is done
    // optional fused dot product of the result with plane[4] (PM_BWD)
    double* dotOut;
    double* partials;
    unsigned int* ticket;
    int* err;
};

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

template <int MODE>
struct PencilTraits {
    static constexpr int DIR = (MODE == PM_BWD || MODE == PM_GS_REV) ? -1 : 1;
    static constexpr bool GS = (MODE == PM_GS_FWD || MODE == PM_GS_REV);
    static constexpr int NP = MODE == PM_FWD ? 5 : MODE == PM_BWD ? 5 : MODE == PM_FACTOR ? 7 : 9;
    static constexpr int STAGE_BYTES = NP * kPencilPlaneBytes;
};

// shared-memory layout of a CTA (dynamic): [full[NS] | empty[NS] | extReady, chainProg | neighbour ring | data ring]
__host__ __device__ constexpr int pencilHeaderBytes(int NS) { return ((2 * NS * 8 + 16) + 127) / 128 * 128; }
__host__ __device__ inline int pencilExtBytes(int extW) { return (kPencilE * extW * 8 + 127) / 128 * 128; }
template <int MODE, int NS>
__host__ __device__ inline int pencilSmemBytes(int extW) {
    return pencilHeaderBytes(NS) + pencilExtBytes(extW) + NS * PencilTraits<MODE>::STAGE_BYTES;
}

template <int MODE, int SKEW, int NS>
__global__ void __launch_bounds__(64, 2) k_pencil(PencilArgs a) {
    using T = PencilTraits<MODE>;
    constexpr int DIR = T::DIR;
    constexpr bool GS = T::GS;
    constexpr int NP = T::NP;
    constexpr int R = kPencilR;
    constexpr int LA = GS ? 2 : 1;   // the leading lane reads LA rows ahead (operand prefetch; + the old value of row r+1)
    constexpr int kNever = 0x7fffffff;
    static_assert((NS & (NS - 1)) == 0, "ring stages must be a power of two");
    extern __shared__ __align__(128) unsigned char pencilSmem[];
    const unsigned smBase = smem_u32(pencilSmem);
    const unsigned barFull = smBase, barEmpty = smBase + NS * 8;
    const unsigned wExtReady = smBase + 2 * NS * 8, wChainProg = wExtReady + 4;
    const unsigned extRing = smBase + pencilHeaderBytes(NS);
    const unsigned dataRing = extRing + pencilExtBytes(a.extW);
    const int lane = threadIdx.x & 31;
    const bool helper = threadIdx.x >= 32;
    const double NEUTRAL = MODE == PM_FACTOR ? 1.0 : 0.0;
    const int nx = a.nx;
    const int nChunks = (nx + R - 1) / R;
    const int pad = DIR > 0 ? 0 : (R - nx % R) % R;   // backward: the first (top) chunk is the partial one
    double dsum[1] = {0.0};

    if (threadIdx.x == 0) {
        for (int q = 0; q < NS; q++) {
            mbar_init(barFull + q * 8, 1);
            mbar_init(barEmpty + q * 8, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    unsigned gchunk0 = 0;   // chunks handed over by earlier tiles of this CTA (same count in both warps)
    for (int ti = blockIdx.x; ti < a.nTiles; ti += gridDim.x, gchunk0 += nChunks) {
        const PencilTileDev* tp = a.tiles + a.order[DIR > 0 ? ti : a.nTiles - 1 - ti];
        const int4 t0 = *reinterpret_cast<const int4*>(&tp->base);      // base, w, wj, wk
        const int4 tB = *reinterpret_cast<const int4*>(tp->nbrBase);
        const int tbase = t0.x, w = t0.y, wj = t0.z, wk = t0.w;
        const int skewMax = SKEW * ((wj - 1) + (wk - 1));
        const int S = nx + skewMax;
        // neighbour tiles: chain side = where the new values come from, static side = old values (Gauss-Seidel)
        const int baseCJ = DIR > 0 ? tB.x : tB.z, baseCK = DIR > 0 ? tB.y : tB.w;
        const int baseSJ = DIR > 0 ? tB.z : tB.x, baseSK = DIR > 0 ? tB.w : tB.y;
        const bool hasCJ = baseCJ >= 0, hasCK = baseCK >= 0;
        const bool hasSJ = GS && baseSJ >= 0, hasSK = GS && baseSK >= 0;
        const bool anyExt = hasCJ || hasCK || hasSJ || hasSK;
        // ring columns of a step: [chain J (wk) | chain K (wj) | static J (wk) | static K (wj)]
        const int colCK = wk, colSJ = wk + wj, colSK = 2 * wk + wj;
        const unsigned extRowB = unsigned(a.extW) * 8;

        if (helper) {
            if (lane == 0) {
                st_release_cta(wExtReady, anyExt ? 0 : kNever);
                st_release_cta(wChainProg, 0);
            }
            // columns without a source tile hold a constant for the whole tile
            for (int e = lane; e < kPencilE * a.extW; e += 32)
                sts_f64(extRing + unsigned(e) * 8, (e % a.extW) < colSJ ? NEUTRAL : 0.0);
            __syncthreads();   // tile start: ring initialised, progress words reset

            // ---- neighbour-value entries of this lane: (step offset, column) pairs over the present columns ----
            const int4 tW = *reinterpret_cast<const int4*>(tp->nbrW);
            const int4 tJ = *reinterpret_cast<const int4*>(tp->nbrWj);
            const int nCols = (hasCJ ? wk : 0) + (hasCK ? wj : 0) + (hasSJ ? wk : 0) + (hasSK ? wj : 0);
            const int W = a.window;
            const int nE = W * nCols;
            int eOff[kPencilMaxE];      // element offset of row 0 of the source pencil, -1: no entry
            int eStride[kPencilMaxE];   // row stride of the source tile
            int eInfo[kPencilMaxE];     // step offset | skew << 8 | column << 16 | chain << 24
#pragma unroll
            for (int k = 0; k < kPencilMaxE; k++) {
                const int e = k * 32 + lane;
                eOff[k] = -1;
                eStride[k] = 0;
                eInfo[k] = 0;
                if (nCols > 0 && e < nE) {
                    const int stepOff = e / nCols;
                    int q = e - stepOff * nCols;
                    // which group does column q of the compacted list belong to: 0 chain J, 1 chain K, 2 static J, 3 static K
                    int grp = -1, idx = 0;
                    if (hasCJ) { if (grp < 0 && q < wk) { grp = 0; idx = q; } q -= wk; }
                    if (hasCK) { if (grp < 0 && q >= 0 && q < wj) { grp = 1; idx = q; } q -= wj; }
                    if (hasSJ) { if (grp < 0 && q >= 0 && q < wk) { grp = 2; idx = q; } q -= wk; }
                    if (hasSK) { if (grp < 0 && q >= 0 && q < wj) { grp = 3; idx = q; } q -= wj; }
                    const bool isJ = (grp == 0 || grp == 2), isChain = grp < 2;
                    // lower-side neighbours ((J-1,K), (J,K-1)) touch our low face with their high face; upper side reversed
                    const bool lowSide = (DIR > 0) == isChain;
                    const int nBase = grp == 0 ? baseCJ : grp == 1 ? baseCK : grp == 2 ? baseSJ : baseSK;
                    const int nW = isJ ? (lowSide ? tW.x : tW.z) : (lowSide ? tW.y : tW.w);
                    const int nWj = isJ ? (lowSide ? tJ.x : tJ.z) : (lowSide ? tJ.y : tJ.w);
                    const int nWk = nW / nWj;
                    int jj, kk, srcLane;
                    if (isJ) {
                        kk = idx;
                        jj = lowSide ? 0 : wj - 1;
                        srcLane = (lowSide ? nWj - 1 : 0) + nWj * kk;
                    } else {
                        jj = idx;
                        kk = lowSide ? 0 : wk - 1;
                        srcLane = jj + nWj * (lowSide ? nWk - 1 : 0);
                    }
                    const int jr = DIR > 0 ? jj : wj - 1 - jj, kr = DIR > 0 ? kk : wk - 1 - kk;
                    const int col = (grp == 0 ? 0 : grp == 1 ? colCK : grp == 2 ? colSJ : colSK) + idx;
                    eOff[k] = nBase + srcLane;
                    eStride[k] = nW;
                    eInfo[k] = stepOff | ((SKEW * (jr + kr)) << 8) | (col << 16) | (isChain ? (1 << 24) : 0);
                }
            }
            const double* chainSrc = a.out;
            const double* staticSrc = a.plane[8];

            int cL = 0;   // chunks issued
            auto serviceLoader = [&](bool block) {
                while (cL < nChunks) {
                    const unsigned gc = gchunk0 + unsigned(cL);
                    const unsigned st = gc & (NS - 1), use = gc / NS;
                    if (use > 0) {
                        if (block) mbar_wait(barEmpty + st * 8, (use - 1) & 1, a.err);
                        else if (!mbar_test_wait(barEmpty + st * 8, (use - 1) & 1)) return;
                    }
                    if (lane == 0) {
                        // rows of this chunk in memory order
                        const int i0 = (DIR > 0 ? cL : nChunks - 1 - cL) * R;
                        const int rows = min(R, nx - i0);
                        const unsigned bytes = (unsigned(rows) * unsigned(w) * 8u + 15u) & ~15u;
                        int np = 0;
#pragma unroll
                        for (int p = 0; p < NP; p++) np += a.plane[p] ? 1 : 0;
                        mbar_arrive_expect_tx(barFull + st * 8, bytes * unsigned(np));
                        const size_t e0 = size_t(tbase) + size_t(i0) * size_t(w);
#pragma unroll
                        for (int p = 0; p < NP; p++)
                            if (a.plane[p])
                                bulk_g2s(dataRing + st * T::STAGE_BYTES + p * kPencilPlaneBytes, a.plane[p] + e0, bytes,
                                         barFull + st * 8);
                    }
                    cL++;
                }
            };

            serviceLoader(false);
            if (nCols > 0) {
                int chainProg = 0;
                for (int s0 = 0; s0 < S; s0 += W) {
                    // ring capacity: the window may only overwrite steps the chain warp has left behind
                    unsigned spins = 0;
                    while (s0 + W - kPencilE > chainProg) {
                        serviceLoader(false);
                        chainProg = ld_acquire_cta(wChainProg);
                        if (++spins > kMaxSpins) {
                            *a.err = 1;
                            break;
                        }
                    }
                    // entries of this window: constants are deposited at once, the others are polled
                    unsigned pend = 0;
#pragma unroll
                    for (int k = 0; k < kPencilMaxE; k++) {
                        const int step = s0 + (eInfo[k] & 0xff);
                        if (eOff[k] >= 0 && step < S) {
                            const int r = step - ((eInfo[k] >> 8) & 0xff);
                            if (r >= 0 && r < nx) {
                                pend |= 1u << k;
                            } else {
                                sts_f64(extRing + unsigned(step & (kPencilE - 1)) * extRowB + unsigned((eInfo[k] >> 16) & 0xff) * 8,
                                        (eInfo[k] >> 24) ? NEUTRAL : 0.0);
                            }
                        }
                    }
                    int published = 0;
                    spins = 0;
                    while (true) {
                        // one polling round: every value still missing, all loads in flight together
                        double v[kPencilMaxE];
#pragma unroll
                        for (int k = 0; k < kPencilMaxE; k++) {
                            if (pend & (1u << k)) {
                                const int r = s0 + (eInfo[k] & 0xff) - ((eInfo[k] >> 8) & 0xff);
                                const int i = DIR > 0 ? r : nx - 1 - r;
                                const size_t el = size_t(eOff[k]) + size_t(i) * size_t(eStride[k]);
                                v[k] = (eInfo[k] >> 24) ? ld_l2(chainSrc + el) : ld_cg(staticSrc + el);
                            }
                        }
                        int firstMissing = nE;
#pragma unroll
                        for (int k = 0; k < kPencilMaxE; k++) {
                            if ((pend & (1u << k)) && !((eInfo[k] >> 24) && is_sentinel(v[k]))) {
                                const int step = s0 + (eInfo[k] & 0xff);
                                sts_f64(extRing + unsigned(step & (kPencilE - 1)) * extRowB + unsigned((eInfo[k] >> 16) & 0xff) * 8,
                                        v[k]);
                                pend &= ~(1u << k);
                            }
                            const unsigned m = __ballot_sync(0xffffffffu, (pend >> k) & 1u);
                            if (m && firstMissing == nE) firstMissing = k * 32 + (__ffs(m) - 1);
                        }
                        // steps whose values are all in the ring
                        const int ready = firstMissing == nE ? min(W, S - s0) : firstMissing / nCols;
                        __syncwarp();
                        if (ready > published) {
                            published = ready;
                            if (lane == 0) st_release_cta(wExtReady, s0 + published);
                        }
                        if (firstMissing == nE) break;
                        serviceLoader(false);
                        if (++spins > kMaxSpins) {
                            *a.err = 1;
                            break;
                        }
                    }
                    serviceLoader(false);
                }
            }
            serviceLoader(true);
            __syncthreads();   // tile end
        } else {
            // ------------------------------------------------------------------------------------------------
            // chain warp
            // ------------------------------------------------------------------------------------------------
            const bool laneOn = lane < w;
            const int jj = laneOn ? lane % wj : 0, kk = laneOn ? lane / wj : 0;
            const int jr = DIR > 0 ? jj : wj - 1 - jj, kr = DIR > 0 ? kk : wk - 1 - kk;
            const int skew = SKEW * (jr + kr);
            const bool extJ = (jr == 0), extK = (kr == 0);            // chain-side values come from the ring
            const int srcJ = (extJ || !laneOn) ? lane : lane - DIR;   // shuffle sources
            const int srcK = (extK || !laneOn) ? lane : lane - DIR * wj;
            // Gauss-Seidel: the old values of the other side live in the data ring (same tile) or the neighbour ring
            const bool sExtJ = (jr == wj - 1), sExtK = (kr == wk - 1);
            const int oJoff = sExtJ ? 0 : DIR * 8, oKoff = sExtK ? 0 : DIR * wj * 8;
            const unsigned wB = unsigned(w) * 8;
            const unsigned ringLane = dataRing + unsigned(lane) * 8;
            const unsigned eJaddr = extRing + unsigned(kk) * 8, eKaddr = extRing + unsigned(colCK + jj) * 8;
            const unsigned sJaddr = extRing + unsigned(colSJ + kk) * 8, sKaddr = extRing + unsigned(colSK + jj) * 8;
            const bool doClear = a.clear != nullptr;
            const bool doDot = MODE == PM_BWD && a.plane[4] != nullptr;
            const bool asym = MODE == PM_FACTOR && a.plane[4] != nullptr;
            const double sent = sentinel();

            __syncthreads();   // tile start (helper has reset the progress words and the ring)

            // ring address of the operands of processing row r (any r: the result of an inactive step is discarded)
            auto rowOff = [&](int r) -> unsigned {
                const int i = DIR > 0 ? r : nx - 1 - r;
                const unsigned st = (gchunk0 + unsigned((r + pad) >> 3)) & (NS - 1);
                return ringLane + st * T::STAGE_BYTES + unsigned(i & (R - 1)) * wB;
            };
            struct Ops {
                double c[NP];
                double eJ, eK, esJ, esK, oI, oJ, oK;
            };
            auto loadOps = [&](Ops& o, int s, unsigned off, unsigned offNext) {
#pragma unroll
                for (int p = 0; p < NP; p++) o.c[p] = lds_f64(off + p * kPencilPlaneBytes);
                const unsigned er = unsigned(s & (kPencilE - 1)) * extRowB;
                o.eJ = lds_f64(eJaddr + er);
                o.eK = lds_f64(eKaddr + er);
                if (GS) {
                    o.esJ = lds_f64(sJaddr + er);
                    o.esK = lds_f64(sKaddr + er);
                    o.oJ = lds_f64(off + 8 * kPencilPlaneBytes + oJoff);
                    o.oK = lds_f64(off + 8 * kPencilPlaneBytes + oKoff);
                    o.oI = lds_f64(offNext + 8 * kPencilPlaneBytes);
                }
            };

            // waits, as step numbers at which they fall due
            int waitRow = 0;                          // first processing row of the next chunk to wait for
            int waitStep = -1;                        // ... needed by the loads issued at this step
            int relChunk = 0;
            int relStep = min(nx - 1, R - pad - 1) + skewMax;   // last step that reads the chunk
            int extAvail = anyExt ? 0 : kNever;
            auto waitData = [&](int s) {              // chunks holding processing rows <= s + LA
                while (waitRow < nx && waitRow <= s + LA) {
                    const unsigned gc = gchunk0 + unsigned((waitRow + pad) >> 3);
                    mbar_wait(barFull + (gc & (NS - 1)) * 8, (gc / NS) & 1, a.err);
                    waitRow = (((waitRow + pad) >> 3) + 1) * R - pad;
                }
                waitStep = waitRow < nx ? waitRow - LA : kNever;
            };
            auto waitExt = [&](int step) {            // neighbour values of `step` are in the ring
                unsigned spins = 0;
                while (extAvail <= step && step < S) {
                    extAvail = ld_acquire_cta(wExtReady);
                    if (++spins > kMaxSpins) {
                        *a.err = 1;
                        break;
                    }
                }
            };

            double y1 = NEUTRAL, sJ = NEUTRAL, sK = NEUTRAL;
            int elem = tbase + lane + (DIR > 0 ? -skew : nx - 1 + skew) * w;   // position of the row of step 0
            int r = -skew;

            // one step: operands in `cur`, prefetch of the next step's into `nxt`
            auto step = [&](int s, Ops& cur, Ops& nxt, unsigned& offNext) {
                const bool act = laneOn && unsigned(r) < unsigned(nx);
                if (s >= waitStep) waitData(s);
                if (s + 1 >= extAvail) waitExt(s + 1);
                const unsigned off = offNext;
                offNext = rowOff(r + 2);
                loadOps(nxt, s + 1, off, offNext);

                if (SKEW == 1) {
                    sJ = __shfl_sync(0xffffffffu, y1, srcJ);
                    sK = __shfl_sync(0xffffffffu, y1, srcK);
                }
                const double vJ = extJ ? cur.eJ : sJ, vK = extK ? cur.eK : sK;
                const double* c = cur.c;
                double acc, y;
                if (MODE == PM_FWD) {
                    // planes: in, rD, t_K, t_J, t_I  (t = rD*lower, +0 where there is no face)
                    acc = c[1] * c[0];
                    acc -= c[2] * vK;
                    acc -= c[3] * vJ;
                    acc -= c[4] * y1;
                    y = acc;
                } else if (MODE == PM_BWD) {
                    // planes: in, t_K, t_J, t_I (t = rD*upper), dotWith
                    acc = c[0];
                    acc -= c[1] * vK;
                    acc -= c[2] * vJ;
                    acc -= c[3] * y1;
                    y = acc;
                } else if (MODE == PM_FACTOR) {
                    // planes: diag, l_K, l_J, l_I, u_K, u_J, u_I (coefficients of the lower-side faces; symmetric
                    // matrices pass no u planes)
                    acc = c[0];
                    acc -= ((asym ? c[4] : c[1]) * c[1]) / vK;
                    acc -= ((asym ? c[5] : c[2]) * c[2]) / vJ;
                    acc -= ((asym ? c[6] : c[3]) * c[3]) / y1;
                    y = acc;
                } else {
                    // Gauss-Seidel planes: b, diag, l_K, l_J, l_I, u_I, u_J, u_K, old
                    const double oI = (unsigned(r + 1) < unsigned(nx)) ? cur.oI : 0.0;   // no such row: exact +0
                    const double oJ = sExtJ ? cur.esJ : cur.oJ, oK = sExtK ? cur.esK : cur.oK;
                    acc = c[0];
                    if (MODE == PM_GS_FWD) {
                        acc -= c[2] * vK;
                        acc -= c[3] * vJ;
                        acc -= c[4] * y1;
                        acc -= c[5] * oI;
                        acc -= c[6] * oJ;
                        acc -= c[7] * oK;
                    } else {
                        // reverse half of symGaussSeidel: lower terms with the forward values, then the new upper ones
                        acc -= c[2] * oK;
                        acc -= c[3] * oJ;
                        acc -= c[4] * oI;
                        acc -= c[5] * y1;
                        acc -= c[6] * vJ;
                        acc -= c[7] * vK;
                    }
                    y = acc / c[1];
                }
                if (SKEW == 2) {
                    sJ = __shfl_sync(0xffffffffu, y1, srcJ);
                    sK = __shfl_sync(0xffffffffu, y1, srcK);
                }
                y1 = act ? y : NEUTRAL;
                if (act) {
                    st_l2(a.out + elem, y);
                    if (MODE == PM_FACTOR) a.out2[elem] = 1.0 / y;
                    if (doClear) a.clear[elem] = sent;
                    if (doDot) dsum[0] += y * c[4];
                }
                elem += DIR * w;
                r++;
                // hand the chunk back to the helper once its last row has been read by the most skewed lane
                if (s == relStep) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(barEmpty + ((gchunk0 + unsigned(relChunk)) & (NS - 1)) * 8);
                    relChunk++;
                    const int nextFirst = relChunk * R - pad;
                    relStep = nextFirst < nx ? min(nx - 1, nextFirst + R - 1) + skewMax : kNever;
                }
                // tell the helper how far the neighbour ring has been consumed
                if ((s & 7) == 7 && anyExt && lane == 0) st_release_cta(wChainProg, s);
            };

            Ops A, B;
            waitData(-1);
            if (anyExt) waitExt(0);
            unsigned offNext = rowOff(1 - skew);
            loadOps(A, 0, rowOff(-skew), offNext);
            int s = 0;
            for (; s + 1 < S; s += 2) {
                step(s, A, B, offNext);
                step(s + 1, B, A, offNext);
            }
            if (s < S) step(s, A, B, offNext);
            __syncthreads();   // tile end
        }
    }
    if (MODE == PM_BWD && a.dotOut) {
        if (grid_reduce<1>(dsum, a.partials, a.ticket)) a.dotOut[0] = dsum[0];
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
