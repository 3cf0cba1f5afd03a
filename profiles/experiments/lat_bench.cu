// Single-warp latencies that bound a step of the pencil chain warp (B200, sm_100a), built like the library with
// -fmad=false:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false lat_bench.cu -o lat_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int WHICH>
__global__ void k(double* out, long long* cyc, int n, double seed) {
    __shared__ __align__(16) double sm[1024];
    __shared__ unsigned long long bar[4];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seed + i * 1e-9;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(smem_u32(bar) + 8));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar) + 16));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar) + 24));
    }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    double x = seed + lane, y = seed * 0.5, c1 = 1.0000001, c2 = 1e-9;
    unsigned a = smem_u32(sm) + lane * 8;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        if (WHICH == 0) x = x - c2;                                  // DADD chain
        if (WHICH == 1) x = x * c1;                                  // DMUL chain
        if (WHICH == 2) { x = x * c1; x = x - c2; }                  // DMUL + DADD
        if (WHICH == 3) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);   // 64-bit shuffle chain
        if (WHICH == 4) {                                            // dependent LDS.64 chain (address from the loaded value)
            double v;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
            a = smem_u32(sm) + ((__double2loint(v) & 0) + lane) * 8;
            x += v;
        }
        if (WHICH == 5) {                                            // chain step: shuffle x2, select, 3 mul, 3 sub
            const double sJ = __shfl_sync(0xffffffffu, x, (lane + 31) & 31);
            const double sK = __shfl_sync(0xffffffffu, x, (lane + 24) & 31);
            const double vJ = (lane & 7) ? sJ : c2, vK = (lane >> 3) ? sK : c2;
            double acc = y;
            acc -= c2 * vK;
            acc -= c2 * vJ;
            acc -= c2 * x;
            x = acc;
        }
        if (WHICH == 6) {                                            // mbarrier try_wait on a completed phase + dependent branch
            unsigned ok;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(bar)), "r"(1u) : "memory");
            if (!ok) x += 1.0;
        }
        if (WHICH == 7) {                                            // st.shared + ld.shared of the same word (round trip)
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(x) : "memory");
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a) : "memory");
        }
        if (WHICH == 8) {                                            // DDIV chain
            x = y / x;
        }
        if (WHICH == 9) {                                            // integer dependent chain (IMAD)
            a = a * 3 + 1;
        }
        if (WHICH == 11) {                                           // mbarrier.arrive by all 32 lanes (count 32) + DADD
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar) + 8) : "memory");
            x = x - c2;
        }
        if (WHICH == 12) {                                           // mbarrier.arrive by one lane (count 1) + DADD
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar) + 16) : "memory");
            x = x - c2;
        }
        if (WHICH == 13 || WHICH == 14 || WHICH == 15) {             // chain step + record loads + result store (+ barrier traffic)
            const double sJ = __shfl_sync(0xffffffffu, x, (lane + 31) & 31);
            const double sK = __shfl_sync(0xffffffffu, x, (lane + 24) & 31);
            double2 v0, v1, v2;
            const unsigned ra = smem_u32(sm) + lane * 48 + (i & 3) * 1536;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0.x), "=d"(v0.y) : "r"(ra));
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v1.x), "=d"(v1.y) : "r"(ra + 16));
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v2.x), "=d"(v2.y) : "r"(ra + 32));
            const double vJ = (lane & 7) ? sJ : v2.x, vK = (lane >> 3) ? sK : v2.y;
            double acc = v0.x;
            acc -= v0.y * vK;
            acc -= v1.x * vJ;
            acc -= v1.y * x;
            x = acc * 1e-30 + 1.0;
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(smem_u32(sm) + 6400 + lane * 8), "d"(x) : "memory");
            if (WHICH == 14) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar) + 8) : "memory");
            if (WHICH == 15) {
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar) + 16) : "memory");
            }
        }
        if (WHICH >= 16 && WHICH <= 19) {
            // the chain warp's step as written in pencil.cuh: barrier tests (completed phases), record + neighbour loads
            // one step ahead, shuffles, arithmetic, result store, warp sync, two single-lane arrives
            unsigned ok1 = 1, ok2 = 1;
            if (WHICH != 18) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok1) : "r"(smem_u32(bar)), "r"(1u) : "memory");
                if (!ok1) x += 1.0;
            }
            double2 v0, v1;
            const unsigned ra = smem_u32(sm) + lane * 48 + (i & 3) * 1536;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0.x), "=d"(v0.y) : "r"(ra));
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v1.x), "=d"(v1.y) : "r"(ra + 16));
            if (WHICH != 18) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok2) : "r"(smem_u32(bar)), "r"(1u) : "memory");
                if (!ok2) x += 1.0;
            }
            double e0, e1;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(e0) : "r"(smem_u32(sm) + 7000 + (lane & 7) * 8));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(e1) : "r"(smem_u32(sm) + 7104 + (lane >> 3) * 8));
            const double sJ = __shfl_sync(0xffffffffu, x, (lane + 31) & 31);
            const double sK = __shfl_sync(0xffffffffu, x, (lane + 24) & 31);
            const double vJ = (lane & 7) ? sJ : e0, vK = (lane >> 3) ? sK : e1;
            double acc = v0.x;
            acc -= v0.y * vK;
            acc -= v1.x * vJ;
            acc -= v1.y * x;
            x = acc * 1e-30 + 1.0;
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(smem_u32(sm) + 6400 + lane * 8), "d"(x) : "memory");
            if (WHICH != 19) {
                __syncwarp();
                if (lane == 0) {
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar) + 16) : "memory");
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar) + 24) : "memory");
                }
            }
            if (WHICH == 17) {   // + another warp hammering shared memory is started by the host (block of 160 threads)
            }
        }
        if (WHICH == 10) {                                           // vote + dependent branch
            if (__any_sync(0xffffffffu, x == 12345.0)) x += 1.0;
            x = x - c2;
        }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    out[lane] = x + a;
}

template <int W>
void run(const char* name, double* out, long long* cyc) {
    const int n = 20000;
    k<W><<<1, 64>>>(out, cyc, n, 1.25);
    cudaDeviceSynchronize();
    k<W><<<1, 64>>>(out, cyc, n, 1.25);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-55s %7.1f cycles/iteration\n", name, double(h) / n);
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1024);
    cudaMalloc(&cyc, 8);
    run<0>("DADD dependent", out, cyc);
    run<1>("DMUL dependent", out, cyc);
    run<2>("DMUL + DADD dependent", out, cyc);
    run<3>("64-bit shuffle dependent", out, cyc);
    run<4>("LDS.64 dependent", out, cyc);
    run<5>("chain step (2 shuffles, selects, 3 mul, 3 sub)", out, cyc);
    run<6>("mbarrier.try_wait (complete) + branch", out, cyc);
    run<7>("STS + LDS same word", out, cyc);
    run<8>("DDIV dependent", out, cyc);
    run<9>("IMAD dependent", out, cyc);
    run<10>("vote.any + branch + DADD", out, cyc);
    run<11>("mbarrier.arrive x32 lanes + DADD", out, cyc);
    run<12>("mbarrier.arrive x1 lane + DADD", out, cyc);
    run<13>("chain step + 3 LDS.128 + STS", out, cyc);
    run<14>("chain step + 3 LDS.128 + STS + arrive x32", out, cyc);
    run<15>("chain step + 3 LDS.128 + STS + syncwarp + arrive x1", out, cyc);
    run<16>("pencil chain step as written (2 tests, 2 LDS.128, 2 LDS.64, 2 arrives)", out, cyc);
    run<18>("  ... without the two barrier tests", out, cyc);
    run<19>("  ... without syncwarp + arrives", out, cyc);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("clock rate attribute %d kHz\n", clk);
    return 0;
}
