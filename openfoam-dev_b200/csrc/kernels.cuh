// sm_100a kernels of libb200ls.  Every kernel on this path is HBM/L2-latency bound fp64 + int32 work:
// plain CUDA cores, no tensor cores (nothing here is a dense contraction).
//
// Numerical contract (SURVEY.md 8(a) "canonical derived integer data"): each row accumulates in the
// reference's order -- diagonal term, neighbour-side faces ascending, owner-side faces ascending (descending
// for backward sweeps) -- and the translation unit is compiled with -fmad=false, so per-row results are
// bit-identical to the reference's sequential face loops.  Only the global reductions (dot products) differ
// in summation order.
//
// Rows live in "positions" = forward-wavefront-major order (mesh.hpp).  L* arrays hold the neighbour-side
// entries of each row, U* arrays the owner-side entries.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "device.cuh"

namespace b200ls {

// Signalling-NaN bit pattern that fp64 arithmetic can never produce (arithmetic yields the canonical quiet
// NaN): used as "row not finished yet" marker by the sync-free sweeps.
static constexpr unsigned long long kSentinelBits = 0x7FF4B2005E471AE1ull;

__device__ __forceinline__ double sentinel() { return __longlong_as_double((long long)kSentinelBits); }

// L2-coherent load/store (sweeps communicate between SMs through L2; L1 must be bypassed)
__device__ __forceinline__ double ld_l2(const double* p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_l2(double* p, double v) {
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

__device__ __forceinline__ bool is_sentinel(double v) { return __double_as_longlong(v) == (long long)kSentinelBits; }

// Poll budget of one wait: ~2^22 L2 round trips (seconds).  A logic error sets *err instead of hanging the GPU.
static constexpr unsigned kMaxSpins = 1u << 22;

// system-scope release/acquire on 64-bit flags (peer GPUs over NVLink) and system-scope relaxed data stores
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(double* p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// ------------------------------------------------------------------------------------------------------------
// elementwise / permutation
// ------------------------------------------------------------------------------------------------------------

__global__ void k_gather(double* __restrict__ out, const double* __restrict__ in, const int* __restrict__ idx,
                         int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[idx[i]];
}

__global__ void k_scatter_index(int* __restrict__ out, const int* __restrict__ idx, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[idx[i]] = i;
}

__global__ void k_fill(double* __restrict__ out, double v, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = v;
}

__global__ void k_fill_sentinel(double* __restrict__ out, int n) {
    const double s = sentinel();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = s;
}

// out = a - b
__global__ void k_sub(double* __restrict__ out, const double* __restrict__ a, const double* __restrict__ b, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = a[i] - b[i];
}

// x += y
__global__ void k_add_inplace(double* __restrict__ x, const double* __restrict__ y, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] += y[i];
}

// x -= y
__global__ void k_sub_inplace(double* __restrict__ x, const double* __restrict__ y, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] -= y[i];
}

// out = a * b   (diagonalPreconditioner::precondition: wA = rD*rA)
__global__ void k_mul(double* __restrict__ out, const double* __restrict__ a, const double* __restrict__ b, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = a[i] * b[i];
}

// out = x / d   (diagonal preconditioner / diagonalSolver)
__global__ void k_div(double* __restrict__ out, const double* __restrict__ x, const double* __restrict__ d, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = x[i] / d[i];
}

// ------------------------------------------------------------------------------------------------------------
// grid-wide reduction helper (used by the reductions below and by the fused sweep / SpMV epilogues)
// ------------------------------------------------------------------------------------------------------------

static constexpr int kReduceBlocks = 592;    // 4 per SM on 148 SMs
static constexpr int kReduceThreads = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-level sum of up to 2 values; returns true in the single thread that holds the grid totals.
template <int NV>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double* __restrict__ partials,
                                            unsigned int* __restrict__ ticket) {
    __shared__ double sm[NV][32];   // one entry per warp of the block (any block size)
    __shared__ bool isLast;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        v[k] = warp_sum(v[k]);
        if (lane == 0) sm[k][w] = v[k];
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double s = lane < (blockDim.x >> 5) ? sm[k][lane] : 0.0;
            s = warp_sum(s);
            if (lane == 0) partials[k * gridDim.x + blockIdx.x] = s;
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicInc(ticket, gridDim.x - 1);
        isLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!isLast) return false;
    __threadfence();
    // fixed-order fold of the block partials by the whole last block: strided per-thread sums, then the same
    // warp/shared-memory tree as above
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double s = 0.0;
        for (int i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += ld_l2(partials + k * gridDim.x + i);
        s = warp_sum(s);
        if (lane == 0) sm[k][w] = s;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double s = lane < (blockDim.x >> 5) ? sm[k][lane] : 0.0;
            v[k] = warp_sum(s);
        }
        return lane == 0;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------------------
// SpMV family: lduMatrix::Amul / residual / sumA  (lduMatrixATmul.C:34-92, 203-280, 154-200)
// ------------------------------------------------------------------------------------------------------------

enum { SPMV_AMUL = 0, SPMV_RESIDUAL = 1, SPMV_AMUL_AND_RESIDUAL = 2, SPMV_SUMA = 3 };

template <int MODE>
__global__ void __launch_bounds__(256)
k_spmv(double* __restrict__ out, double* __restrict__ out2, const double* __restrict__ x,
       const double* __restrict__ b, const double* __restrict__ diag, const int* __restrict__ Lptr,
       const int* __restrict__ Lcol, const double* __restrict__ Lval, const int* __restrict__ Uptr,
       const int* __restrict__ Ucol, const double* __restrict__ Uval, int n) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int l0 = Lptr[p], l1 = Lptr[p + 1];
    const int u0 = Uptr[p], u1 = Uptr[p + 1];
    double acc;
    if (MODE == SPMV_SUMA) {
        acc = diag[p];
        for (int j = l0; j < l1; j++) acc += Lval[j];
        for (int j = u0; j < u1; j++) acc += Uval[j];
        out[p] = acc;
        return;
    }
    if (MODE == SPMV_RESIDUAL) {
        acc = b[p] - diag[p] * x[p];
        for (int j = l0; j < l1; j++) acc -= Lval[j] * x[Lcol[j]];
        for (int j = u0; j < u1; j++) acc -= Uval[j] * x[Ucol[j]];
        out[p] = acc;
        return;
    }
    acc = diag[p] * x[p];
    for (int j = l0; j < l1; j++) acc += Lval[j] * x[Lcol[j]];
    for (int j = u0; j < u1; j++) acc += Uval[j] * x[Ucol[j]];
    out[p] = acc;
    if (MODE == SPMV_AMUL_AND_RESIDUAL) out2[p] = b[p] - acc;
}

// Symmetric matrices: the neighbour-side coefficient of row p for neighbour q is the owner-side coefficient that
// row q stores for p, so it is gathered from Uval[Uptr[q] + slot] (L2 hit: row q streamed it a wavefront earlier)
// instead of being read from a duplicated Lval array: 83 B/row of DRAM traffic instead of 104 on a hex mesh.
template <int MODE>
__global__ void __launch_bounds__(256)
k_spmv_sym(double* __restrict__ out, double* __restrict__ out2, const double* __restrict__ x,
           const double* __restrict__ b, const double* __restrict__ diag, const int* __restrict__ Lptr,
           const int* __restrict__ Lcol, const unsigned char* __restrict__ Lslot, const int* __restrict__ Uptr,
           const int* __restrict__ Ucol, const double* __restrict__ Uval, int n) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int l0 = Lptr[p], l1 = Lptr[p + 1];
    const int u0 = Uptr[p], u1 = Uptr[p + 1];
    double acc;
    if (MODE == SPMV_RESIDUAL) {
        acc = b[p] - diag[p] * x[p];
        for (int j = l0; j < l1; j++) {
            const int q = Lcol[j];
            acc -= Uval[Uptr[q] + Lslot[j]] * x[q];
        }
        for (int j = u0; j < u1; j++) acc -= Uval[j] * x[Ucol[j]];
        out[p] = acc;
        return;
    }
    acc = diag[p] * x[p];
    for (int j = l0; j < l1; j++) {
        const int q = Lcol[j];
        acc += Uval[Uptr[q] + Lslot[j]] * x[q];
    }
    for (int j = u0; j < u1; j++) acc += Uval[j] * x[Ucol[j]];
    out[p] = acc;
    if (MODE == SPMV_AMUL_AND_RESIDUAL) out2[p] = b[p] - acc;
}

// Two rows per thread (rows p and p + blockDim.x of a 2*blockDim.x tile): the gathers of the symmetric kernel form a
// four-deep dependent chain (Lptr -> Lcol/Lslot -> Uptr[q] -> Uval), so two independent chains per thread double the
// loads in flight.  Same arithmetic order per row as k_spmv_sym<SPMV_AMUL>.
__global__ void __launch_bounds__(256)
k_spmv_sym_x2(double* __restrict__ out, const double* __restrict__ x, const double* __restrict__ diag,
              const int* __restrict__ Lptr, const int* __restrict__ Lcol, const unsigned char* __restrict__ Lslot,
              const int* __restrict__ Uptr, const int* __restrict__ Ucol, const double* __restrict__ Uval, int n) {
    const int pa = blockIdx.x * (2 * blockDim.x) + threadIdx.x;
    const int pb = pa + blockDim.x;
    const bool va = pa < n, vb = pb < n;
    int la0 = 0, la1 = 0, ua0 = 0, ua1 = 0, lb0 = 0, lb1 = 0, ub0 = 0, ub1 = 0;
    double acca = 0.0, accb = 0.0;
    if (va) {
        la0 = Lptr[pa];
        la1 = Lptr[pa + 1];
        ua0 = Uptr[pa];
        ua1 = Uptr[pa + 1];
    }
    if (vb) {
        lb0 = Lptr[pb];
        lb1 = Lptr[pb + 1];
        ub0 = Uptr[pb];
        ub1 = Uptr[pb + 1];
    }
    if (va) acca = diag[pa] * x[pa];
    if (vb) accb = diag[pb] * x[pb];
    const int nl = max(la1 - la0, lb1 - lb0);
    for (int k = 0; k < nl; k++) {
        const bool ka = la0 + k < la1, kb = lb0 + k < lb1;
        int qa = 0, qb = 0, sa = 0, sb = 0;
        if (ka) {
            qa = Lcol[la0 + k];
            sa = Lslot[la0 + k];
        }
        if (kb) {
            qb = Lcol[lb0 + k];
            sb = Lslot[lb0 + k];
        }
        int ba = 0, bb = 0;
        if (ka) ba = Uptr[qa];
        if (kb) bb = Uptr[qb];
        double ca = 0.0, cb = 0.0, xa = 0.0, xb = 0.0;
        if (ka) {
            ca = Uval[ba + sa];
            xa = x[qa];
        }
        if (kb) {
            cb = Uval[bb + sb];
            xb = x[qb];
        }
        if (ka) acca += ca * xa;
        if (kb) accb += cb * xb;
    }
    const int nu = max(ua1 - ua0, ub1 - ub0);
    for (int k = 0; k < nu; k++) {
        const bool ka = ua0 + k < ua1, kb = ub0 + k < ub1;
        double ca = 0.0, cb = 0.0, xa = 0.0, xb = 0.0;
        if (ka) {
            ca = Uval[ua0 + k];
            xa = x[Ucol[ua0 + k]];
        }
        if (kb) {
            cb = Uval[ub0 + k];
            xb = x[Ucol[ub0 + k]];
        }
        if (ka) acca += ca * xa;
        if (kb) accb += cb * xb;
    }
    if (va) out[pa] = acca;
    if (vb) out[pb] = accb;
}

// wA = A pA fused with wApA = wA.pA (PCG.C:159-161).  Grid-stride over rows with a bounded grid so that the
// per-block partials fit the reduction scratch.
template <bool SYM>
__global__ void __launch_bounds__(256)
k_spmv_dot(double* __restrict__ out, const double* __restrict__ x, const double* __restrict__ diag,
           const int* __restrict__ Lptr, const int* __restrict__ Lcol, const double* __restrict__ Lval,
           const unsigned char* __restrict__ Lslot, const int* __restrict__ Uptr, const int* __restrict__ Ucol,
           const double* __restrict__ Uval, int n, double* __restrict__ dotOut, double* __restrict__ partials,
           unsigned int* __restrict__ ticket, const int* __restrict__ stop) {
    if (stop && *stop) return;
    double v[1] = {0.0};
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int l0 = Lptr[p], l1 = Lptr[p + 1];
        const int u0 = Uptr[p], u1 = Uptr[p + 1];
        const double xp = x[p];
        double acc = diag[p] * xp;
        for (int j = l0; j < l1; j++) {
            const int q = Lcol[j];
            acc += (SYM ? Uval[Uptr[q] + Lslot[j]] : Lval[j]) * x[q];
        }
        for (int j = u0; j < u1; j++) acc += Uval[j] * x[Ucol[j]];
        out[p] = acc;
        v[0] += acc * xp;
    }
    if (grid_reduce<1>(v, partials, ticket)) dotOut[0] = v[0];
}

// Coupled-interface epilogue: result[faceCells[i]] -= coeffs[i]*recv[i] in (patch, face) order
// (processorFvPatchScalarField.C:133-136).  One thread per boundary row.  sign = +1 for Amul, -1 when the
// caller negated the coefficients (residual, GaussSeidel).
struct IfaceView {
    const double* coeffs;           // per patch face
    const double* recv;             // per patch face, neighbour values (P2P: two parity buffers back to back)
    const unsigned long long* flag; // P2P: epoch written by the neighbour once its values have landed; else null
    int size;
};
__global__ void k_iface_apply(double* __restrict__ result, const int* __restrict__ rowPos,
                              const int* __restrict__ rowPtr, const int* __restrict__ entIface,
                              const int* __restrict__ entFace, const IfaceView* __restrict__ views, int nIfaces,
                              unsigned long long epoch, double sign, int nRows, int* err) {
    // P2P halos: wait until every neighbour has published this exchange (release/acquire over NVLink)
    if (threadIdx.x < nIfaces) {
        const unsigned long long* f = views[threadIdx.x].flag;
        if (f) {
            unsigned spins = 0;
            while (ld_acquire_sys(f) < epoch) {
                if (++spins > (1u << 26)) {
                    *err = 2;
                    break;
                }
            }
        }
    }
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nRows) return;
    const int p = rowPos[r];
    double acc = result[p];
    for (int e = rowPtr[r]; e < rowPtr[r + 1]; e++) {
        const IfaceView v = views[entIface[e]];
        const int f = entFace[e];
        const double* rv = v.flag ? v.recv + (epoch & 1) * v.size : v.recv;
        acc -= (sign * v.coeffs[f]) * __ldcg(rv + f);
    }
    result[p] = acc;
}

// P2P halo send: the gathered boundary values are stored straight into the neighbour's receive buffer; the last
// block to finish publishes the epoch flag (release, system scope).
__global__ void k_iface_pack_p2p(double* __restrict__ remoteRecv, const double* __restrict__ psi,
                                 const int* __restrict__ faceCellsPos, int n, unsigned int* __restrict__ ticket,
                                 unsigned long long* remoteFlag, unsigned long long epoch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_sys(remoteRecv + i, psi[faceCellsPos[i]]);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicInc(ticket, gridDim.x - 1);
        if (t == gridDim.x - 1) {
            __threadfence_system();
            st_release_sys(remoteFlag, epoch);
        }
    }
}

// All-reduce (sum) of <= 4 doubles across ranks through the mapped peer arenas: every rank stores its values into
// slot [parity][rank] of every arena and releases an epoch; every rank then acquires the nRanks epochs of its own
// arena and adds the values in rank order (identical, deterministic result on all ranks).
__global__ void k_allreduce_p2p(double* __restrict__ data, int count, P2PView v, unsigned long long epoch, int* err) {
    const int lane = threadIdx.x;
    const int par = int(epoch & 1);
    const size_t valOff = (size_t(par) * kMaxRanks) * 4 * sizeof(double);
    const size_t epoOff = 2 * kMaxRanks * 4 * sizeof(double) + size_t(par) * kMaxRanks * sizeof(unsigned long long);
    if (lane < v.nRanks) {
        double* dst = reinterpret_cast<double*>(v.peer[lane] + valOff) + v.rank * 4;
        for (int k = 0; k < count; k++) st_sys(dst + k, data[k]);
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long*>(v.peer[lane] + epoOff) + v.rank, epoch);
    }
    __syncwarp();
    double vals[4] = {0.0, 0.0, 0.0, 0.0};
    if (lane < v.nRanks) {
        const unsigned long long* e = reinterpret_cast<const unsigned long long*>(v.peer[v.rank] + epoOff) + lane;
        unsigned spins = 0;
        while (ld_acquire_sys(e) < epoch) {
            if (++spins > (1u << 26)) {
                *err = 2;
                break;
            }
        }
        const double* src = reinterpret_cast<const double*>(v.peer[v.rank] + valOff) + lane * 4;
        for (int k = 0; k < count; k++) vals[k] = ld_sys(src + k);
    }
    for (int k = 0; k < count; k++) {
        double s = 0.0;
        for (int r = 0; r < v.nRanks; r++) s += __shfl_sync(0xffffffffu, vals[k], r);
        if (lane == 0) data[k] = s;
    }
}

// send[i] = psi[faceCellsPos[i]]  (patchInternalField, processorFvPatchScalarField.C:45)
__global__ void k_iface_pack(double* __restrict__ send, const double* __restrict__ psi,
                             const int* __restrict__ faceCellsPos, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) send[i] = psi[faceCellsPos[i]];
}

// sumA -= bouCoeffs on interface rows (lduMatrixATmul.C:185-199)
__global__ void k_iface_suma(double* __restrict__ sumA, const int* __restrict__ rowPos,
                             const int* __restrict__ rowPtr, const int* __restrict__ entIface,
                             const int* __restrict__ entFace, const IfaceView* __restrict__ views, int nRows) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nRows) return;
    const int p = rowPos[r];
    double acc = sumA[p];
    for (int e = rowPtr[r]; e < rowPtr[r + 1]; e++) acc -= views[entIface[e]].coeffs[entFace[e]];
    sumA[p] = acc;
}

// ------------------------------------------------------------------------------------------------------------
// Sync-free wavefront sweeps.  Persistent grid (every CTA co-resident, cooperative launch): warp w takes tasks
// w, w+W, ... ; a task is <=32 rows of ONE wavefront, so lanes of a warp never wait on each other, and a warp
// only ever waits on tasks with a smaller index => deadlock free.  A row publishes its result with a single
// 8-byte L2 store; consumers spin on the sentinel (wait_row).
// ------------------------------------------------------------------------------------------------------------

struct SweepArgs {
    const int2* tasks;
    int nTasks;
    const int* rowOf;       // processing slot -> position (nullptr = identity: forward sweeps on the wavefront-major layout)
    const int* ptr;         // Lptr (forward) / Uptr (backward)
    const int* col;
    const double* val;
    const int* ptr2;        // the other triangle (GaussSeidel, factor)
    const int* col2;
    const double* val2;
    const double* diag;
    const double* rD;
    const double* in;       // rA (forward) / forward result (backward) / source (GS)
    const double* old;      // GS: previous iterate
    double* out;            // sentinel-initialised result
    double* out2;           // optional second output (factor: rD)
    double* clear;          // optional: array whose entry p is reset to the sentinel once row p is done
    int* err;
    // optional fused dot product of the result with another vector (PCG: wArA = wA.rA, PCG.C:138)
    const double* dotWith;
    double* dotOut;
    double* partials;
    unsigned int* ticket;
    const int* stop;        // optional: a non-zero word makes the launch a no-op (speculatively enqueued iterations)
};

// Dependency gather of one row: acc -= (scale*val[j]) * y[col[j]] for j ascending (DESC=false) or descending
// (DESC=true), waiting for each y entry to be published.  The static loads (col, val) of a chunk of 4 entries
// are issued first, then ALL outstanding dependencies are polled together in every spin round: a row whose
// dependencies become visible at about the same time pays one L2 round trip, not one per entry.  (Measured on
// B200, 128^3: 1.8 ms per sweep polling entry by entry, 0.25 ms polling all at once + warp reconvergence.)
template <bool DESC, bool SCALE>
__device__ __forceinline__ double gather_deps(double acc, double scale, int j0, int j1, const int* __restrict__ col,
                                              const double* __restrict__ val, const double* y, int* err) {
    for (int base = 0; base < j1 - j0; base += 4) {
        const int n = min(4, j1 - j0 - base);
        int c[4];
        double v[4], w[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (k < n) {
                const int j = DESC ? (j1 - 1 - base - k) : (j0 + base + k);
                c[k] = col[j];
                v[k] = val[j];
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < n) w[k] = ld_l2(y + c[k]);
        unsigned spins = 0;
        while (true) {
            bool pending = false;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (k < n && is_sentinel(w[k])) pending = true;
            if (!pending) break;
            if (++spins > kMaxSpins) {
                *err = 1;
                break;
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (k < n && is_sentinel(w[k])) w[k] = ld_l2(y + c[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < n) acc -= (SCALE ? scale * v[k] : v[k]) * w[k];
    }
    return acc;
}

#define SWEEP_TASK_LOOP(a)                                                            \
    const int wpb = blockDim.x >> 5;                                                  \
    const int nW = gridDim.x * wpb;                                                   \
    const int lane = threadIdx.x & 31;                                                \
    int t = blockIdx.x * wpb + (threadIdx.x >> 5);                                    \
    int2 task = t < (a).nTasks ? (a).tasks[t] : make_int2(0, 0);                      \
    for (; t < (a).nTasks; t += nW)

#define SWEEP_NEXT_TASK(a) ((t + nW) < (a).nTasks ? (a).tasks[t + nW] : make_int2(0, 0))

// DIC/DILU calcReciprocalD (DICPreconditioner.C:71-83, DILUPreconditioner.C:72-84):
//   d[u] = diag[u] - sum_{f: upper(f)=u, ascending} upper[f]*lower[f]/d[lower(f)] ;  rD = 1/d
// val = Lval (lower[f]), val2 = the matching upper[f] gathered through LtoU (col2).
__global__ void __launch_bounds__(256) k_factor(SweepArgs a) {
    const double sent = sentinel();
    SWEEP_TASK_LOOP(a) {
        const int2 next = SWEEP_NEXT_TASK(a);
        if (lane < task.y) {
            const int p = a.rowOf ? a.rowOf[task.x + lane] : task.x + lane;
            double acc = a.diag[p];
            const int j0 = a.ptr[p], j1 = a.ptr[p + 1];
            for (int base = j0; base < j1; base += 4) {
                const int n = min(4, j1 - base);
                int c[4];
                double num[4], w[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k < n) {
                        const double lo = a.val[base + k];
                        const double up = a.col2 ? a.val2[a.col2[base + k]] : lo;
                        c[k] = a.col[base + k];
                        num[k] = up * lo;
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (k < n) w[k] = ld_l2(a.out + c[k]);
                unsigned spins = 0;
                while (true) {
                    bool pending = false;
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (k < n && is_sentinel(w[k])) pending = true;
                    if (!pending) break;
                    if (++spins > kMaxSpins) {
                        *a.err = 1;
                        break;
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (k < n && is_sentinel(w[k])) w[k] = ld_l2(a.out + c[k]);
                }
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (k < n) acc -= num[k] / w[k];
            }
            st_l2(a.out + p, acc);
            a.out2[p] = 1.0 / acc;
            if (a.clear) a.clear[p] = sent;
        }
        __syncwarp();
        task = next;
    }
}

// forward substitution of DIC/DILU precondition with the scaling pass fused:
//   y[u] = rD[u]*rA[u] - sum_{f ascending} (rD[u]*lower[f]) * y[lower(f)]
__global__ void __launch_bounds__(256) k_sweep_fwd(SweepArgs a) {
    if (a.stop && *a.stop) return;
    const double sent = sentinel();
    SWEEP_TASK_LOOP(a) {
        const int2 next = SWEEP_NEXT_TASK(a);
        if (lane < task.y) {
            const int p = a.rowOf ? a.rowOf[task.x + lane] : task.x + lane;
            const double rd = a.rD[p];
            const int j0 = a.ptr[p], j1 = a.ptr[p + 1];
            double acc = rd * a.in[p];
            acc = gather_deps<false, true>(acc, rd, j0, j1, a.col, a.val, a.out, a.err);
            st_l2(a.out + p, acc);
            if (a.clear) a.clear[p] = sent;
        }
        __syncwarp();
        task = next;
    }
}

// backward substitution:  z[l] = y[l] - sum_{f descending} (rD[l]*upper[f]) * z[upper(f)]
__global__ void __launch_bounds__(256) k_sweep_bwd(SweepArgs a) {
    if (a.stop && *a.stop) return;
    const double sent = sentinel();
    double dsum[1] = {0.0};
    SWEEP_TASK_LOOP(a) {
        const int2 next = SWEEP_NEXT_TASK(a);
        if (lane < task.y) {
            const int p = a.rowOf ? a.rowOf[task.x + lane] : task.x + lane;
            const double rd = a.rD[p];
            const int j0 = a.ptr[p], j1 = a.ptr[p + 1];
            double acc = a.in[p];
            acc = gather_deps<true, true>(acc, rd, j0, j1, a.col, a.val, a.out, a.err);
            st_l2(a.out + p, acc);
            if (a.clear) a.clear[p] = sent;
            if (a.dotWith) dsum[0] += acc * a.dotWith[p];
        }
        __syncwarp();
        task = next;
    }
    if (a.dotWith) {
        if (grid_reduce<1>(dsum, a.partials, a.ticket)) a.dotOut[0] = dsum[0];
    }
}

// Gauss-Seidel sweep (GaussSeidelSmoother.C:151-176) in gather form:
//   psi_new[c] = ( b'[c] - sum_{nbr faces asc} lower[f]*psi_new[l] - sum_{own faces asc} upper[f]*psi_old[u] ) / diag[c]
__global__ void __launch_bounds__(256) k_gs_sweep(SweepArgs a) {
    const double sent = sentinel();
    SWEEP_TASK_LOOP(a) {
        const int2 next = SWEEP_NEXT_TASK(a);
        if (lane < task.y) {
            const int p = a.rowOf ? a.rowOf[task.x + lane] : task.x + lane;
            const int j0 = a.ptr[p], j1 = a.ptr[p + 1];
            const int k0 = a.ptr2[p], k1 = a.ptr2[p + 1];
            const double dg = a.diag[p];
            double acc = a.in[p];
            // owner-side products only need the old iterate: fetch them before waiting on anything
            double up[4];
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (k0 + k < k1) up[k] = a.val2[k0 + k] * a.old[a.col2[k0 + k]];
            acc = gather_deps<false, false>(acc, 1.0, j0, j1, a.col, a.val, a.out, a.err);
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (k0 + k < k1) acc -= up[k];
            for (int k = k0 + 4; k < k1; k++) acc -= a.val2[k] * a.old[a.col2[k]];
            st_l2(a.out + p, acc / dg);
            if (a.clear) a.clear[p] = sent;
        }
        __syncwarp();
        task = next;
    }
}

// Several Gauss-Seidel sweeps in ONE launch, pipelined: sweep s of wavefront L only needs sweep s of wavefront L-1
// (lower neighbours) and sweep s-1 of wavefront L+1 (upper neighbours), so sweep s trails sweep s-1 by two
// wavefronts instead of starting after it has finished.  Tasks are ordered by tau = L + 2s; the dependency chain of
// nSweeps sweeps shrinks from nSweeps*nLevels to nLevels + 2*(nSweeps-1) hops.  X[0] is the input iterate, sweep s
// reads X[s] (plain loads for s = 0, polled otherwise) and publishes X[s+1]; every X[s>0] starts all-sentinel.
// Arithmetic per row is identical to k_gs_sweep.
// entries beyond the first four of a row (rare on hex meshes): one at a time, kept out of line to save registers
__device__ __noinline__ double gather_tail(double acc, int j0, int j1, const int* __restrict__ col,
                                           const double* __restrict__ val, const double* y, int* err) {
    for (int j = j0; j < j1; j++) {
        const int c = col[j];
        double w = ld_l2(y + c);
        unsigned spins = 0;
        while (is_sentinel(w)) {
            if (++spins > kMaxSpins) {
                *err = 1;
                break;
            }
            w = ld_l2(y + c);
        }
        acc -= val[j] * w;
    }
    return acc;
}

static constexpr int kMaxFusedSweeps = 8;
static constexpr int kCoupledSlotSweeps = 3;   // coupled fused Gauss-Seidel: up to 1 + 3 sweeps per launch
struct CoupledView {
    const double* coeffs;   // interfaceBouCoeffs of the patch
    double* mine;           // [kCoupledSlotSweeps][size] slots of this call's parity in my arena (neighbour writes)
    double* theirs;         // the neighbour's slots for its side of the patch (I write)
    int size, pad;
};
struct MultiSweepArgs {
    const int4* tasks;      // (start, count, sweep, -)
    int nTasks;
    const int* rowOf;       // processing slot -> position (nullptr = identity)
    const int* Lptr;
    const int* Lcol;
    const double* Lval;
    const int* Uptr;
    const int* Ucol;
    const double* Uval;
    const double* diag;
    const double* b;
    double* X[kMaxFusedSweeps + 1];
    int* err;
    // COUPLED (processor patches, P2P): rows on a patch take the neighbour rank's values of the previous sweep from
    // slots in this rank's arena (the neighbour stores them there the moment it has finished the row) and publish
    // their own the same way.  Sweep 0 uses b0 = source with the regular halo exchange already applied.
    int nSweeps;
    const double* b0;
    const int* bRowOf;      // position -> boundary row or -1
    const int* bRowPtr;
    const int* bEntIface;
    const int* bEntFace;
    const CoupledView* views;
};
template <bool COUPLED>
__global__ void __launch_bounds__(256, 2) k_gs_multi(MultiSweepArgs a) {
    const int wpb = blockDim.x >> 5;
    const int nW = gridDim.x * wpb;
    const int lane = threadIdx.x & 31;
    int t = blockIdx.x * wpb + (threadIdx.x >> 5);
    int4 task = t < a.nTasks ? a.tasks[t] : make_int4(0, 0, 0, 0);
    for (; t < a.nTasks; t += nW) {
        const int4 next = (t + nW) < a.nTasks ? a.tasks[t + nW] : make_int4(0, 0, 0, 0);
        if (lane < task.y) {
            const int p = a.rowOf ? a.rowOf[task.x + lane] : task.x + lane;
            const int s = task.z;
            const double* xo = a.X[s];
            double* xn = a.X[s + 1];
            const int j0 = a.Lptr[p], j1 = a.Lptr[p + 1];
            const int k0 = a.Uptr[p], k1 = a.Uptr[p + 1];
            const double dg = a.diag[p];
            const int br = COUPLED ? a.bRowOf[p] : -1;
            double acc = (COUPLED && s == 0) ? a.b0[p] : a.b[p];
            if (COUPLED && s > 0 && br >= 0) {
                // bPrime = source - sum (-bouCoeffs)*psiNeighbour (GaussSeidelSmoother.C:110-145), neighbour values
                // of the previous sweep; patch by patch, face by face like k_iface_apply
                for (int e = a.bRowPtr[br]; e < a.bRowPtr[br + 1]; e++) {
                    const CoupledView v = a.views[a.bEntIface[e]];
                    const int f = a.bEntFace[e];
                    double* slot = v.mine + size_t(s - 1) * v.size + f;
                    double w = ld_sys(slot);
                    unsigned spins = 0;
                    while (is_sentinel(w)) {
                        if (++spins > kMaxSpins) {
                            *a.err = 2;
                            break;
                        }
                        w = ld_sys(slot);
                    }
                    acc -= (-1.0 * v.coeffs[f]) * w;
                    *slot = sentinel();   // re-armed for the call after next (same parity)
                }
            }
            if (s == 0) {
                double up[4];
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (k0 + k < k1) up[k] = a.Uval[k0 + k] * xo[a.Ucol[k0 + k]];
                acc = gather_deps<false, false>(acc, 1.0, j0, j1, a.Lcol, a.Lval, xn, a.err);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (k0 + k < k1) acc -= up[k];
                for (int k = k0 + 4; k < k1; k++) acc -= a.Uval[k] * xo[a.Ucol[k]];
            } else {
                // both dependency sets (this sweep's lower neighbours, the previous sweep's upper neighbours)
                // become visible at about the same time: poll the first four of each together
                const int nl = min(4, j1 - j0), nu = min(4, k1 - k0);
                int cl[4], cu[4];
                double vl[4], vu[4], wl[4], wu[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k < nl) {
                        cl[k] = a.Lcol[j0 + k];
                        vl[k] = a.Lval[j0 + k];
                    }
                    if (k < nu) {
                        cu[k] = a.Ucol[k0 + k];
                        vu[k] = a.Uval[k0 + k];
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k < nl) wl[k] = ld_l2(xn + cl[k]);
                    if (k < nu) wu[k] = ld_l2(xo + cu[k]);
                }
                unsigned spins = 0;
                while (true) {
                    bool pending = false;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (k < nl && is_sentinel(wl[k])) pending = true;
                        if (k < nu && is_sentinel(wu[k])) pending = true;
                    }
                    if (!pending) break;
                    if (++spins > kMaxSpins) {
                        *a.err = 1;
                        break;
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (k < nl && is_sentinel(wl[k])) wl[k] = ld_l2(xn + cl[k]);
                        if (k < nu && is_sentinel(wu[k])) wu[k] = ld_l2(xo + cu[k]);
                    }
                }
                // accumulate in the reference's order: all lower-neighbour terms, then all upper-neighbour terms
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (k < nl) acc -= vl[k] * wl[k];
                if (j0 + nl < j1) acc = gather_tail(acc, j0 + nl, j1, a.Lcol, a.Lval, xn, a.err);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (k < nu) acc -= vu[k] * wu[k];
                if (k0 + nu < k1) acc = gather_tail(acc, k0 + nu, k1, a.Ucol, a.Uval, xo, a.err);
            }
            const double xNew = acc / dg;
            st_l2(xn + p, xNew);
            if (COUPLED && br >= 0 && s + 1 < a.nSweeps) {
                for (int e = a.bRowPtr[br]; e < a.bRowPtr[br + 1]; e++) {
                    const CoupledView v = a.views[a.bEntIface[e]];
                    st_sys(v.theirs + size_t(s) * v.size + a.bEntFace[e], xNew);
                }
            }
        }
        __syncwarp();
        task = next;
    }
}

// Reverse half of symGaussSeidel (symGaussSeidelSmoother.C:178-205) in gather form.  After the forward loop
// bPrime[c] = b'[c] - sum_{nbr faces asc} lower[f]*psi_fwd[l]; the reverse loop then subtracts the owner side with
// the already reverse-updated upper neighbours:
//   psi_rev[c] = ( b'[c] - sum_{nbr faces asc} lower[f]*psi_fwd[l] - sum_{own faces asc} upper[f]*psi_rev[u] ) / diag[c]
// ptr/col/val = U triangle (polled, rows in backward-wavefront order), ptr2/col2/val2 = L triangle read from `old`
// (the forward result).
__global__ void __launch_bounds__(256) k_gs_sweep_rev(SweepArgs a) {
    const double sent = sentinel();
    SWEEP_TASK_LOOP(a) {
        const int2 next = SWEEP_NEXT_TASK(a);
        if (lane < task.y) {
            const int p = a.rowOf ? a.rowOf[task.x + lane] : task.x + lane;
            const int j0 = a.ptr[p], j1 = a.ptr[p + 1];
            const int k0 = a.ptr2[p], k1 = a.ptr2[p + 1];
            const double dg = a.diag[p];
            double acc = a.in[p];
            for (int k = k0; k < k1; k++) acc -= a.val2[k] * a.old[a.col2[k]];
            acc = gather_deps<false, false>(acc, 1.0, j0, j1, a.col, a.val, a.out, a.err);
            st_l2(a.out + p, acc / dg);
            if (a.clear) a.clear[p] = sent;
        }
        __syncwarp();
        task = next;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Reductions: per-thread grid-stride partial -> warp shuffle -> shared memory -> one partial per block; the
// last block to finish (atomic ticket) folds the partials in a fixed order.  Deterministic for a fixed grid.
// ------------------------------------------------------------------------------------------------------------

enum { RED_DOT = 0, RED_SUMMAG = 1, RED_SUM = 2, RED_SUMSQR = 3 };

// out[0] = sum f(x, y)
template <int OP>
__global__ void __launch_bounds__(kReduceThreads)
k_reduce(double* __restrict__ out, const double* __restrict__ x, const double* __restrict__ y, int n,
         double* __restrict__ partials, unsigned int* __restrict__ ticket) {
    double v[1] = {0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (OP == RED_DOT) v[0] += x[i] * y[i];
        if (OP == RED_SUMMAG) v[0] += fabs(x[i]);
        if (OP == RED_SUM) v[0] += x[i];
        if (OP == RED_SUMSQR) v[0] += x[i] * x[i];
    }
    if (grid_reduce<1>(v, partials, ticket)) out[0] = v[0];
}

// out[0] = sum x*y, out[1] = sum z*y   (GAMG scale: source.field, Acf.field ; PBiCGStab: tA.sA, tA.tA)
__global__ void __launch_bounds__(kReduceThreads)
k_dot2(double* __restrict__ out, const double* __restrict__ x, const double* __restrict__ z,
       const double* __restrict__ y, int n, double* __restrict__ partials, unsigned int* __restrict__ ticket) {
    double v[2] = {0.0, 0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double yi = y[i];
        v[0] += x[i] * yi;
        v[1] += z[i] * yi;
    }
    if (grid_reduce<2>(v, partials, ticket)) {
        out[0] = v[0];
        out[1] = v[1];
    }
}

// normFactor (lduMatrixSolver.C:174-197): out[0] = sum |Apsi - xbar*sumA| + |source - xbar*sumA|,
// xbar = sumPsi[0]/nGlobal
__global__ void __launch_bounds__(kReduceThreads)
k_norm_factor(double* __restrict__ out, const double* __restrict__ Apsi, const double* __restrict__ source,
              const double* __restrict__ sumA, const double* __restrict__ sumPsi, double nGlobal, int n,
              double* __restrict__ partials, unsigned int* __restrict__ ticket) {
    const double xbar = sumPsi[0] / nGlobal;
    double v[1] = {0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double t = sumA[i] * xbar;
        v[0] += fabs(Apsi[i] - t) + fabs(source[i] - t);
    }
    if (grid_reduce<1>(v, partials, ticket)) out[0] = v[0];
}

// ------------------------------------------------------------------------------------------------------------
// PCG vector updates (PCG.C:142-179); scalars stay on the device
// ------------------------------------------------------------------------------------------------------------

// pA = wA + (wArA/wArAold)*pA       (first iteration: pA = wA)
__global__ void k_pcg_update_p(double* __restrict__ pA, const double* __restrict__ wA,
                               const double* __restrict__ wArA, const double* __restrict__ wArAold, int first,
                               int n, const int* __restrict__ stop) {
    if (stop && *stop) return;
    if (first) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) pA[i] = wA[i];
        return;
    }
    const double beta = wArA[0] / wArAold[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        pA[i] = wA[i] + beta * pA[i];
}

// alpha = wArA/wApA ; psi += alpha*pA ; rA -= alpha*wA ; out = sum |rA|
__global__ void __launch_bounds__(kReduceThreads)
k_pcg_update_xr(double* __restrict__ psi, double* __restrict__ rA, const double* __restrict__ pA,
                const double* __restrict__ wA, const double* __restrict__ wArA, const double* __restrict__ wApA,
                double normFactor, double* __restrict__ singularFlag, double* __restrict__ out, int n,
                double* __restrict__ partials, unsigned int* __restrict__ ticket, const int* __restrict__ stop) {
    if (stop && *stop) return;
    // checkSingularity(mag(wApA)/normFactor) (PCG.C:165): leave psi/rA untouched and raise the flag
    if (fabs(wApA[0]) / normFactor < 2.2250738585072014e-308) {
        if (blockIdx.x == 0 && threadIdx.x == 0) singularFlag[0] = 1.0;
        return;
    }
    const double alpha = wArA[0] / wApA[0];
    double v[1] = {0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        psi[i] += alpha * pA[i];
        const double r = rA[i] - alpha * wA[i];
        rA[i] = r;
        v[0] += fabs(r);
    }
    if (grid_reduce<1>(v, partials, ticket)) out[0] = v[0];
}

// Device-side copy of the loop condition of PCG.C:182-190 for iterations enqueued ahead of the host's read-back:
// stop = singular, or converged (SolverPerformance.C:75-82) once minIter allows it.  The host evaluates the same
// expression on the same numbers one iteration later.
__global__ void k_pcg_stop_flag(int* __restrict__ stop, const double* __restrict__ sumMagR,
                                const double* __restrict__ singular, double normFactor, double tolerance,
                                double relTol, double initialResidual, int mayStop) {
    if (*stop) return;
    const double fin = sumMagR[0] / normFactor;
    const bool conv = fin < tolerance || (relTol > 1e-20 && fin < relTol * initialResidual);
    if (singular[0] != 0.0 || (conv && mayStop)) *stop = 1;
}

// ------------------------------------------------------------------------------------------------------------
// PBiCGStab vector updates (PBiCGStab.C:166-242)
// ------------------------------------------------------------------------------------------------------------

// pA = rA + beta*(pA - omega*AyA), beta = (rA0rA/rA0rAold)*(alpha/omega)
__global__ void k_bicg_update_p(double* __restrict__ pA, const double* __restrict__ rA,
                                const double* __restrict__ AyA, const double* __restrict__ S, int iRho,
                                int iRhoOld, int iAlpha, int iOmega, int first, int n) {
    if (first) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) pA[i] = rA[i];
        return;
    }
    const double omega = S[iOmega];
    const double beta = (S[iRho] / S[iRhoOld]) * (S[iAlpha] / omega);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        pA[i] = rA[i] + beta * (pA[i] - omega * AyA[i]);
}

// alpha = rA0rA/rA0AyA ; sA = rA - alpha*AyA ; out = sum |sA| ; alphaOut = alpha
__global__ void __launch_bounds__(kReduceThreads)
k_bicg_update_s(double* __restrict__ sA, const double* __restrict__ rA, const double* __restrict__ AyA,
                const double* __restrict__ rho, const double* __restrict__ rA0AyA, double* __restrict__ alphaOut,
                double* __restrict__ out, int n, double* __restrict__ partials, unsigned int* __restrict__ ticket) {
    const double alpha = rho[0] / rA0AyA[0];
    double v[1] = {0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double s = rA[i] - alpha * AyA[i];
        sA[i] = s;
        v[0] += fabs(s);
    }
    if (grid_reduce<1>(v, partials, ticket)) {
        out[0] = v[0];
        alphaOut[0] = alpha;
    }
}

// psi += alpha*yA   (early exit of PBiCGStab, PBiCGStab.C:213-216)
__global__ void k_axpy_s(double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ a, int n) {
    const double alpha = a[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] += alpha * y[i];
}

// omega = tAsA/tAtA ; psi += alpha*yA + omega*zA ; rA = sA - omega*tA ; out = sum |rA|
__global__ void __launch_bounds__(kReduceThreads)
k_bicg_update_xr(double* __restrict__ psi, double* __restrict__ rA, const double* __restrict__ yA,
                 const double* __restrict__ zA, const double* __restrict__ sA, const double* __restrict__ tA,
                 const double* __restrict__ alphaP, const double* __restrict__ tAsA_tAtA,
                 double* __restrict__ omegaOut, double* __restrict__ out, int n, double* __restrict__ partials,
                 unsigned int* __restrict__ ticket) {
    const double alpha = alphaP[0];
    const double omega = tAsA_tAtA[0] / tAsA_tAtA[1];
    double v[1] = {0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        psi[i] += alpha * yA[i] + omega * zA[i];
        const double r = sA[i] - omega * tA[i];
        rA[i] = r;
        v[0] += fabs(r);
    }
    if (grid_reduce<1>(v, partials, ticket)) {
        out[0] = v[0];
        omegaOut[0] = omega;
    }
}

// ------------------------------------------------------------------------------------------------------------
// GAMG transfer operators and coarse-matrix assembly (all order-preserving gathers)
// ------------------------------------------------------------------------------------------------------------

// restrictField (GAMGAgglomerationTemplates.C:76-89): cf[c] = sum of ff over the fine cells of c, ascending
__global__ void k_restrict(double* __restrict__ cf, const double* __restrict__ ff, const int* __restrict__ rPtr,
                           const int* __restrict__ rFine, int nCoarse) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCoarse) return;
    double acc = 0.0;
    for (int k = rPtr[c]; k < rPtr[c + 1]; k++) acc += ff[rFine[k]];
    cf[c] = acc;
}

// prolongField (:214-217): ff[i] = cf[map[i]]
__global__ void k_prolong(double* __restrict__ ff, const double* __restrict__ cf, const int* __restrict__ pMap,
                          int nFine) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nFine) ff[i] = cf[pMap[i]];
}

// coarse off-diagonal entry = sum of the referenced fine coefficients, ascending fine face
// (GAMGSolverAgglomerateMatrix.C:142-190); fineVals = [Uval | Lval]
__global__ void k_agglomerate_offdiag(double* __restrict__ cv, const double* __restrict__ fineVals,
                                      const int* __restrict__ ptr, const int* __restrict__ src, int nEntries) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nEntries) return;
    double acc = 0.0;
    for (int k = ptr[j]; k < ptr[j + 1]; k++) acc += fineVals[src[k]];
    cv[j] = acc;
}

// coarse diagonal = restrict(fine diag) then += upper+lower of every fine face interior to the coarse cell
// (:61-69, :164, :188)
__global__ void k_agglomerate_diag(double* __restrict__ cd, const double* __restrict__ fd,
                                   const double* __restrict__ fU, const double* __restrict__ fL,
                                   const int* __restrict__ rPtr, const int* __restrict__ rFine,
                                   const int* __restrict__ dPtr, const int* __restrict__ dU,
                                   const int* __restrict__ dL, int nCoarse) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCoarse) return;
    double acc = 0.0;
    for (int k = rPtr[c]; k < rPtr[c + 1]; k++) acc += fd[rFine[k]];
    for (int k = dPtr[c]; k < dPtr[c + 1]; k++) acc += fU[dU[k]] + fL[dL[k]];
    cd[c] = acc;
}

// GAMGSolver::scale second half (GAMGSolverScale.C:62-75):
//   sf = num/stabilise(den, vSmall) ; field = sf*field + (source - sf*Acf)/D
__global__ void k_gamg_scale(double* __restrict__ field, const double* __restrict__ source,
                             const double* __restrict__ Acf, const double* __restrict__ D,
                             const double* __restrict__ numDen, int n) {
    const double den = numDen[1];
    const double vSmall = 2.2250738585072014e-308;
    const double sf = numDen[0] / (den >= 0 ? den + vSmall : den - vSmall);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        field[i] = sf * field[i] + (source[i] - sf * Acf[i]) / D[i];
}

}  // namespace b200ls
