// Round-2 prototype of the structure-specialised streamed forward sweep (see README.md "What the chain itself costs").
// N^3 hex block, tiles of 8x4 pencils, one CTA of three warps per tile:
//   L  loader: cp.async ring of {tK, tJ, tI, rD} packs, the gathered input value and the row position;
//   E  fetches the values of other tiles from L2 ahead of time (ring of 8 steps, refresh-all-on-miss) and hands them to
//      C through shared-memory words that hold the sentinel until written;
//   C  the chain: acc = rD*in - tK*vK - tJ*vJ - tI*x with vK / vJ from shuffles (or from E on the low faces of the tile),
//      publishes its own result with st.relaxed.gpu.
// Result checked bit for bit against the sequential recurrence.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false stream2_dbg.cu -o stream2_dbg ; ./stream2_dbg 128
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
static constexpr unsigned long long SENT = 0x7FF4B2005E471AE1ull;
static constexpr int R = 32, G = 4, NG = R / G, E = 8, TJ = 8, TK = 4;
__device__ __forceinline__ double ld_l2(const double* p) { double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_l2(double* p, double v) { asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ bool isS(double v) { return __double_as_longlong(v) == (long long)SENT; }
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(void* s, const void* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(s)), "l"(g) : "memory"); }
__device__ __forceinline__ void cp8(void* s, const void* g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s32(s)), "l"(g) : "memory"); }
__device__ __forceinline__ void cp4(void* s, const void* g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s32(s)), "l"(g) : "memory"); }
__device__ __forceinline__ void mb_init(unsigned long long* b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mb_arrive_cp(unsigned long long* b) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_arrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ bool mb_try(unsigned long long* b, unsigned par) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    return ok != 0;
}
__device__ __noinline__ void mb_wait(unsigned long long* b, unsigned par, int* err) { unsigned n = 0; while (!mb_try(b, par)) if (++n > (1u << 24)) { *err = 1; break; } }
__device__ __forceinline__ double ldsv(const double* p) { double v; asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(s32(p)) : "memory"); return v; }
__device__ __forceinline__ void stsv(double* p, double v) { asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(s32(p)), "d"(v) : "memory"); }
__device__ __noinline__ double ldsv_wait(const double* p, int* err) { unsigned n = 0; double v = ldsv(p); while (isS(v)) { if (++n > (1u << 26)) { *err = 2; break; } v = ldsv(p); } return v; }

struct Stage { double2 a[32]; double2 b[32]; double in[32]; double eK[32]; double eJ[32]; int pos[32]; };
struct Smem { Stage st[R]; unsigned long long full[NG], empty[NG]; };
struct Args { int nTiles, S; const double2* pack; const int* pos; const int2* ext; const double* in; double* out; int* err; long long* prof; };

__global__ void __launch_bounds__(96) k_chain_sweep(Args a) {
    extern __shared__ __align__(16) unsigned char raw[];
    Smem& sm = *reinterpret_cast<Smem*>(raw);
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    const double sent = __longlong_as_double((long long)SENT);
    if (threadIdx.x == 0) {
        for (int q = 0; q < NG; q++) { mb_init(&sm.full[q], 32); mb_init(&sm.empty[q], 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < R * 32; i += blockDim.x) { sm.st[i >> 5].eK[i & 31] = sent; sm.st[i >> 5].eJ[i & 31] = sent; }
    __syncthreads();
    const bool bK = lane < TJ, bJ = (lane % TJ) == 0;       // pencils on the low-k / low-j face of the tile
    const int S = a.S;
    unsigned g = 0;
    if (role == 0) {
        for (int T = blockIdx.x; T < a.nTiles; T += gridDim.x) {
            const size_t base = size_t(T) * S;
            for (int u0 = 0; u0 < S; u0 += 8) {
                int pb[8];
#pragma unroll
                for (int j = 0; j < 8; j++) pb[j] = (u0 + j < S) ? __ldg(a.pos + (base + u0 + j) * 32 + lane) : -1;
#pragma unroll
                for (int j = 0; j < 8; j++) if (u0 + j < S) {
                    const unsigned s = g & (R - 1), q = s / G, use = g / R;
                    if ((g & (G - 1)) == 0 && use > 0) mb_wait(&sm.empty[q], (use - 1) & 1, a.err);
                    const size_t idx = (base + u0 + j) * 32 + lane;
                    Stage& st = sm.st[s];
                    cp4(&st.pos[lane], a.pos + idx);
                    if (pb[j] >= 0) { cp16(&st.a[lane], a.pack + idx * 2); cp16(&st.b[lane], a.pack + idx * 2 + 1); cp8(&st.in[lane], a.in + pb[j]); }
                    g++;
                    if ((g & (G - 1)) == 0 || u0 + j == S - 1) mb_arrive_cp(&sm.full[q]);
                }
            }
            g = (g + G - 1) & ~unsigned(G - 1);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (role == 1) {
        if (!(bK || bJ)) return;                            // only the boundary pencils have external dependencies
        for (int T = blockIdx.x; T < a.nTiles; T += gridDim.x) {
            const size_t base = size_t(T) * S;
            // positions are fetched 2E steps ahead, values E steps ahead: neither load is waited for when it is issued
            int pK[E], pJ[E], qK[E], qJ[E]; double vK[E], vJ[E];
#pragma unroll
            for (int j = 0; j < E; j++) {
                pK[j] = pJ[j] = qK[j] = qJ[j] = -1; vK[j] = vJ[j] = 0.0;
                if (j < S) { const int2 e = __ldg(a.ext + (base + j) * 32 + lane); pK[j] = e.x; pJ[j] = e.y; }
                if (j + E < S) { const int2 e = __ldg(a.ext + (base + j + E) * 32 + lane); qK[j] = e.x; qJ[j] = e.y; }
            }
#pragma unroll
            for (int j = 0; j < E; j++) {
                if (pK[j] >= 0) vK[j] = ld_l2(a.out + pK[j]);
                if (pJ[j] >= 0) vJ[j] = ld_l2(a.out + pJ[j]);
            }
            for (int t0 = 0; t0 < S; t0 += E) {
#pragma unroll
                for (int j = 0; j < E; j++) {
                    const int t = t0 + j;
                    if (t < S) {
                        const unsigned s = g & (R - 1), use = g / R;
                        // the stage's hand-over words are free again once C has given the group back
                        if (use > 0 && (g & (G - 1)) == 0) mb_wait(&sm.empty[s / G], (use - 1) & 1, a.err);
                        if ((pK[j] >= 0 && isS(vK[j])) || (pJ[j] >= 0 && isS(vJ[j]))) {
                            unsigned n = 0;
                            while (true) {
#pragma unroll
                                for (int k = 0; k < E; k++) {
                                    if (pK[k] >= 0 && isS(vK[k])) vK[k] = ld_l2(a.out + pK[k]);
                                    if (pJ[k] >= 0 && isS(vJ[k])) vJ[k] = ld_l2(a.out + pJ[k]);
                                }
                                if (!((pK[j] >= 0 && isS(vK[j])) || (pJ[j] >= 0 && isS(vJ[j])))) break;
                                if (++n > (1u << 22)) { *a.err = 3; break; }
                            }
                        }
                        Stage& st = sm.st[s];
                        if (bK) stsv(&st.eK[lane], pK[j] >= 0 ? vK[j] : 0.0);
                        if (bJ) stsv(&st.eJ[lane], pJ[j] >= 0 ? vJ[j] : 0.0);
                        // slot j: values of step t+E (position fetched E steps ago), position of step t+2E
                        pK[j] = qK[j]; pJ[j] = qJ[j];
                        if (pK[j] >= 0) vK[j] = ld_l2(a.out + pK[j]);
                        if (pJ[j] >= 0) vJ[j] = ld_l2(a.out + pJ[j]);
                        qK[j] = qJ[j] = -1;
                        if (t + 2 * E < S) { const int2 e = __ldg(a.ext + (base + t + 2 * E) * 32 + lane); qK[j] = e.x; qJ[j] = e.y; }
                        g++;
                    }
                }
            }
            g = (g + G - 1) & ~unsigned(G - 1);
        }
    } else {
        const int lK = lane >= TJ ? lane - TJ : lane, lJ = (lane % TJ) ? lane - 1 : lane;
        const long long c0 = clock64();
        long long wE = 0, wL = 0;
        for (int T = blockIdx.x; T < a.nTiles; T += gridDim.x) {
            double x = 0.0;
            for (int t = 0; t < S; t++) {
                const unsigned s = g & (R - 1);
                Stage& st = sm.st[s];
                if ((g & (G - 1)) == 0) { const long long w0 = clock64(); mb_wait(&sm.full[s / G], (g / R) & 1, a.err); wL += clock64() - w0; }
                const int pos = st.pos[lane];
                const double2 pa = st.a[lane], pb = st.b[lane];
                const double in = st.in[lane];
                double eK = 0.0, eJ = 0.0;
                if (bK | bJ) {
                    const long long w0 = clock64();
                    if (bK) { eK = ldsv(&st.eK[lane]); if (isS(eK)) eK = ldsv_wait(&st.eK[lane], a.err); stsv(&st.eK[lane], sent); }
                    if (bJ) { eJ = ldsv(&st.eJ[lane]); if (isS(eJ)) eJ = ldsv_wait(&st.eJ[lane], a.err); stsv(&st.eJ[lane], sent); }
                    wE += clock64() - w0;
                }
                const double sK = __shfl_sync(0xffffffffu, x, lK), sJ = __shfl_sync(0xffffffffu, x, lJ);
                const double vK = bK ? eK : sK, vJ = bJ ? eJ : sJ;
                double acc = pb.y * in;          // rD * in
                acc -= pa.x * vK;
                acc -= pa.y * vJ;
                acc -= pb.x * x;
                const bool act = pos >= 0;
                x = act ? acc : 0.0;
                if (act) st_l2(a.out + pos, acc);
                g++;
                if ((g & (G - 1)) == 0 || t == S - 1) mb_arrive(&sm.empty[s / G]);
            }
            g = (g + G - 1) & ~unsigned(G - 1);
        }
        if (a.prof) {
            for (int o = 16; o > 0; o >>= 1) wE = max(wE, __shfl_xor_sync(0xffffffffu, wE, o));
            if (lane == 0) { a.prof[blockIdx.x * 3] = clock64() - c0; a.prof[blockIdx.x * 3 + 1] = wE; a.prof[blockIdx.x * 3 + 2] = wL; }
        }
    }
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 128, reps = argc > 2 ? atoi(argv[2]) : 8;
    const int nx = N, ny = N, nz = N;
    if (ny % TJ || nz % TK) { printf("N must be a multiple of 8\n"); return 1; }
    const int nJ = ny / TJ, nK = nz / TK, S = nx + TJ - 1 + TK - 1;
    const size_t n = size_t(nx) * ny * nz;
    // wavefront-major positions as in the library: by level i+j+k, ascending cell inside a level
    std::vector<int> ipos(n);
    { std::vector<int> cnt(nx + ny + nz, 0);
      for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) cnt[i + j + k + 1]++;
      for (size_t l = 1; l < cnt.size(); l++) cnt[l] += cnt[l - 1];
      for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) ipos[i + size_t(nx) * (j + size_t(ny) * k)] = cnt[i + j + k]++; }
    auto cell = [&](int i, int j, int k) { return i + size_t(nx) * (j + size_t(ny) * k); };
    std::vector<double> rD(n), in(n), cK(n), cJ(n), cI(n), ref(n);
    srand(1);
    for (size_t c = 0; c < n; c++) { rD[c] = 0.15 + 0.05 * (rand() / double(RAND_MAX)); in[c] = rand() / double(RAND_MAX) - 0.5;
        cK[c] = -(0.5 + rand() / double(RAND_MAX)); cJ[c] = -(0.5 + rand() / double(RAND_MAX)); cI[c] = -(0.5 + rand() / double(RAND_MAX)); }
    // sequential reference (cell order is a valid order): deps k-1, j-1, i-1 in ascending face order
    for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) {
        const size_t c = cell(i, j, k);
        double acc = rD[c] * in[c];
        if (k > 0) acc -= (rD[c] * cK[c]) * ref[cell(i, j, k - 1)];
        if (j > 0) acc -= (rD[c] * cJ[c]) * ref[cell(i, j - 1, k)];
        if (i > 0) acc -= (rD[c] * cI[c]) * ref[cell(i - 1, j, k)];
        ref[c] = acc;
    }
    std::vector<std::pair<int, int>> tiles;
    for (int w = 0; w <= nJ + nK - 2; w++) for (int K = 0; K < nK; K++) { const int J = w - K; if (J >= 0 && J < nJ) tiles.emplace_back(J, K); }
    const size_t nRec = tiles.size() * size_t(S) * 32;
    std::vector<double> pack(nRec * 4, 0.0), inPos(n);
    std::vector<int> pos(nRec, -1);
    std::vector<int> ext(nRec * 2, -1);
    for (size_t c = 0; c < n; c++) inPos[ipos[c]] = in[c];
    for (size_t t = 0; t < tiles.size(); t++) {
        const int J = tiles[t].first, K = tiles[t].second;
        for (int s = 0; s < S; s++) for (int kk = 0; kk < TK; kk++) for (int jj = 0; jj < TJ; jj++) {
            const int i = s - jj - kk; if (i < 0 || i >= nx) continue;
            const int j = J * TJ + jj, k = K * TK + kk, lane = jj + TJ * kk;
            const size_t c = cell(i, j, k), r = (t * S + s) * 32 + lane;
            pos[r] = ipos[c];
            pack[r * 4 + 0] = k > 0 ? rD[c] * cK[c] : 0.0;
            pack[r * 4 + 1] = j > 0 ? rD[c] * cJ[c] : 0.0;
            pack[r * 4 + 2] = i > 0 ? rD[c] * cI[c] : 0.0;
            pack[r * 4 + 3] = rD[c];
            if (kk == 0 && k > 0) ext[r * 2] = ipos[cell(i, j, k - 1)];
            if (jj == 0 && j > 0) ext[r * 2 + 1] = ipos[cell(i, j - 1, k)];
        }
    }
    double *dPack, *dIn, *dOut; int *dPos, *dExt, *dErr; long long* dProf;
    CK(cudaMalloc(&dPack, nRec * 32)); CK(cudaMalloc(&dIn, n * 8)); CK(cudaMalloc(&dOut, n * 8)); CK(cudaMalloc(&dPos, nRec * 4)); CK(cudaMalloc(&dExt, nRec * 8));
    CK(cudaMalloc(&dErr, 4)); CK(cudaMemset(dErr, 0, 4));
    CK(cudaMemcpy(dPack, pack.data(), nRec * 32, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dIn, inPos.data(), n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dPos, pos.data(), nRec * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dExt, ext.data(), nRec * 8, cudaMemcpyHostToDevice));
    const size_t smem = sizeof(Smem);
    CK(cudaFuncSetAttribute((const void*)k_chain_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int occ = 0, sms = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_chain_sweep, 96, smem));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int blocks = std::max(1, std::min(occ * sms, int(tiles.size())));
    CK(cudaMalloc(&dProf, size_t(blocks) * 24)); CK(cudaMemset(dProf, 0, size_t(blocks) * 24));
    Args a{int(tiles.size()), S, reinterpret_cast<const double2*>(dPack), dPos, reinterpret_cast<const int2*>(dExt), dIn, dOut, dErr, dProf};
    std::vector<unsigned long long> sent(n, SENT);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaMemcpy(dOut, sent.data(), n * 8, cudaMemcpyHostToDevice));
        void* args[] = {&a};
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((const void*)k_chain_sweep, dim3(blocks), dim3(96), args, smem, 0));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
    }
    std::vector<double> out(n); CK(cudaMemcpy(out.data(), dOut, n * 8, cudaMemcpyDeviceToHost));
    int err; CK(cudaMemcpy(&err, dErr, 4, cudaMemcpyDeviceToHost));
    size_t bad = 0; for (size_t c = 0; c < n; c++) if (out[ipos[c]] != ref[c]) bad++;
    std::vector<long long> prof(size_t(blocks) * 3); CK(cudaMemcpy(prof.data(), dProf, prof.size() * 8, cudaMemcpyDeviceToHost));
    printf("N=%d tiles=%zu S=%d blocks=%d occ=%d smem=%zu  best %.3f ms  mismatches %zu err %d\n", N, tiles.size(), S, blocks, occ, smem, best, bad, err);
    for (int b = 0; b < std::min(blocks, 4); b++) { const int st_ = S * ((int(tiles.size()) - b + blocks - 1) / blocks);
        printf("  CTA %d: C total %lld cycles over %d steps: waiting for E %.0f/step, for L %.0f/step, rest %.0f/step\n", b, prof[b * 3], st_,
               double(prof[b * 3 + 1]) / st_, double(prof[b * 3 + 2]) / st_, double(prof[b * 3] - prof[b * 3 + 1] - prof[b * 3 + 2]) / st_); }
    return (bad || err) ? 1 : 0;
}
