// How fast can ONE warp walk the dependency chain of a pencil tile if everything else is taken off its hands?
// Per step: 2 x LDS.128 (pre-multiplied coefficients t0,t1,t2 and b), 2 shuffles of the previous result (neighbouring
// pencils), selects, acc = b - t0*v0 - t1*v1 - t2*x, STS of the result.  Variants add, cumulatively: (1) a predicated
// shared-memory "external value" read + sentinel test, (2) an L2-coherent global store of the result,
// (3) an L2-coherent global load issued 8 steps ahead (external prefetch).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false chain_step.cu -o chain_step
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
constexpr int R = 32;          // ring stages
struct Stage { double2 a[32]; double2 b[32]; double ext[32]; double res[32]; };
__device__ __forceinline__ double ld_l2(const double* p) { double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_l2(double* p, double v) { asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }

template <int V>
__global__ void __launch_bounds__(32) k(int steps, double* out, long long* cycles, double* sink) {
    __shared__ Stage st[R];
    const int lane = threadIdx.x;
    for (int s = 0; s < R; s++) {
        st[s].a[lane] = make_double2(0.11 + 0.001 * lane, 0.07);
        st[s].b[lane] = make_double2(0.05, 1.0 + 0.01 * s);
        st[s].ext[lane] = 0.5;
        st[s].res[lane] = 0.0;
    }
    __syncwarp();
    const int lj = lane > 0 ? lane - 1 : 0, lk = lane >= 8 ? lane - 8 : 0;
    const bool extJ = (lane & 7) == 0, extK = lane < 8;
    double x = 0.3, pf[8];
#pragma unroll
    for (int j = 0; j < 8; j++) pf[j] = 0.0;
    const long long t0 = clock64();
    for (int t = 0; t < steps; t += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const Stage& S = st[(t + j) & (R - 1)];
            const double2 a = S.a[lane], b = S.b[lane];
            double e = 0.5;
            if (V >= 1) {
                if (extJ || extK) e = S.ext[lane];
                if (__double_as_longlong(e) == 0x7FF4B2005E471AE1ll) e = 0.25;   // never taken
            }
            if (V >= 3) e += 1e-300 * pf[j];
            const double sj = __shfl_sync(0xffffffffu, x, lj), sk = __shfl_sync(0xffffffffu, x, lk);
            const double v0 = extK ? e : sk, v1 = extJ ? e : sj;
            double acc = b.y;
            acc -= a.x * v0;
            acc -= a.y * v1;
            acc -= b.x * x;
            x = acc;
            const_cast<Stage&>(S).res[lane] = acc;
            if (V >= 2) st_l2(out + (size_t(t + j) * 32 + lane), acc);
            if (V >= 3) pf[j] = ld_l2(out + ((size_t(t + j) * 32 + lane) & 0xffff));
        }
    }
    const long long t1 = clock64();
    if (lane == 0) cycles[0] = t1 - t0;
    sink[lane] = x;
}

int main() {
    const int steps = 4096;
    double *out, *sink; long long* cyc;
    CK(cudaMalloc(&out, size_t(steps) * 32 * 8 + 1024)); CK(cudaMalloc(&sink, 32 * 8)); CK(cudaMalloc(&cyc, 8));
    CK(cudaMemset(out, 0, size_t(steps) * 32 * 8));
    long long h;
    for (int rep = 0; rep < 2; rep++) {
        k<0><<<1, 32>>>(steps, out, cyc, sink); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        if (rep) printf("chain only (LDS, shuffles, 6 FP64 ops, STS):        %.0f cycles/step\n", double(h) / steps);
        k<1><<<1, 32>>>(steps, out, cyc, sink); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        if (rep) printf("+ shared-memory external value and sentinel test:   %.0f cycles/step\n", double(h) / steps);
        k<2><<<1, 32>>>(steps, out, cyc, sink); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        if (rep) printf("+ st.relaxed.gpu of the result:                     %.0f cycles/step\n", double(h) / steps);
        k<3><<<1, 32>>>(steps, out, cyc, sink); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        if (rep) printf("+ ld.relaxed.gpu issued 8 steps ahead of its use:   %.0f cycles/step\n", double(h) / steps);
    }
    return 0;
}
