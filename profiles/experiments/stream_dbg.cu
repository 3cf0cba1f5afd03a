// Harness for k_stream_sweep (the shipped kernel, compiled with -DB200LS_STREAM_PROF): forward sweep on an N^3 block,
// result checked bit for bit against the sequential recurrence, cycle attribution per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -DB200LS_STREAM_PROF [-DB200LS_STREAM_R=16]
//        stream_dbg.cu ../../openfoam-dev_b200/csrc/mesh.cpp -o stream_dbg ; ./stream_dbg 128
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../../openfoam-dev_b200/csrc/mesh.hpp"
#include "../../openfoam-dev_b200/csrc/kernels.cuh"
using namespace b200ls;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 128;
    const int reps = argc > 2 ? atoi(argv[2]) : 10;
    std::vector<int32_t> lo, up;
    for (int k = 0; k < N; k++) for (int j = 0; j < N; j++) for (int i = 0; i < N; i++) {
        const int c = i + N * (j + N * k);
        if (i < N - 1) { lo.push_back(c); up.push_back(c + 1); }
        if (j < N - 1) { lo.push_back(c); up.push_back(c + N); }
        if (k < N - 1) { lo.push_back(c); up.push_back(c + N * N); }
    }
    LevelHost L;
    setenv("B200LS_STREAM", "1", 1);
    setenv("B200LS_STREAM_MIN_CELLS", "0", 1);
    buildLevel(L, N * N * N, int(lo.size()), lo.data(), up.data(), {});
    if (!L.fwdStream.valid) { printf("no plan\n"); return 1; }
    const int n = L.nCells, nF = L.nFaces;
    std::vector<double> rD(n), in(n), Lval(nF), ref(n);
    srand(1);
    for (int p = 0; p < n; p++) { rD[p] = 0.15 + 0.05 * (rand() / double(RAND_MAX)); in[p] = rand() / double(RAND_MAX) - 0.5; }
    for (int e = 0; e < nF; e++) Lval[e] = -(0.5 + rand() / double(RAND_MAX));
    // sequential reference in positions (wavefront-major => ascending position is a valid order)
    for (int p = 0; p < n; p++) {
        double acc = rD[p] * in[p];
        for (int e = L.Lptr[p]; e < L.Lptr[p + 1]; e++) acc -= (rD[p] * Lval[e]) * ref[L.Lcol[e]];
        ref[p] = acc;
    }
    double *dRD, *dIn, *dVal, *dOut, *dPack; int4* dRec; int* dEb; int* dPS; int* dErr; unsigned long long* dProf;
    const StreamPlan& pl = L.fwdStream;
    CK(cudaMalloc(&dRD, n * 8 + 16)); CK(cudaMalloc(&dIn, n * 8 + 16)); CK(cudaMalloc(&dVal, size_t(nF) * 8 + 16)); CK(cudaMalloc(&dOut, n * 8 + 16));
    CK(cudaMalloc(&dRec, pl.rec.size() * 16)); CK(cudaMalloc(&dEb, pl.rec.size() * 4)); CK(cudaMalloc(&dPack, pl.rec.size() * 32 + 16)); CK(cudaMalloc(&dPS, pl.partStart.size() * 4));
    CK(cudaMalloc(&dErr, 4)); CK(cudaMemset(dErr, 0, 4));
    CK(cudaMemcpy(dRD, rD.data(), n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dIn, in.data(), n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dVal, Lval.data(), size_t(nF) * 8, cudaMemcpyHostToDevice));
    {
        std::vector<int4> rec(pl.rec.size()); std::vector<int> eb(pl.rec.size());
        for (size_t i = 0; i < pl.rec.size(); i++) { rec[i] = make_int4(pl.rec[i].pos, pl.rec[i].ext0, pl.rec[i].ext1, int(pl.meta[i])); eb[i] = pl.rec[i].ebase; }
        CK(cudaMemcpy(dRec, rec.data(), rec.size() * 16, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dEb, eb.data(), eb.size() * 4, cudaMemcpyHostToDevice));
    }
    k_stream_pack<<<1184, 256>>>(dPack, dRec, dEb, dRD, dVal, pl.rec.size(), 0);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(dPS, pl.partStart.data(), pl.partStart.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem = sizeof(StreamSmem);
    auto kern = k_stream_sweep<false>;
    CK(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int occ = 0, sms = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 64, smem));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int blocks = std::max(1, std::min(occ * sms, pl.nParts));
    const int nWarps = blocks;
    CK(cudaMalloc(&dProf, size_t(nWarps) * 9 * 8)); CK(cudaMemset(dProf, 0, size_t(nWarps) * 9 * 8));
    StreamArgs a{};
    a.partStart = dPS; a.nParts = pl.nParts; a.rec = dRec; a.pack = dPack; a.in = dIn; a.out = dOut;
    a.clear = nullptr; a.err = dErr;
#ifdef B200LS_STREAM_PROF
    a.prof = dProf;
#endif
    std::vector<unsigned long long> sent(n + 2, kSentinelBits);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f, sum = 0;
    for (int r = 0; r < reps; r++) {
        CK(cudaMemcpy(dOut, sent.data(), n * 8 + 16, cudaMemcpyHostToDevice));
        void* args[] = {&a};
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((const void*)kern, dim3(blocks), dim3(64), args, smem, 0));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms); if (r) sum += ms;
    }
    std::vector<double> out(n); CK(cudaMemcpy(out.data(), dOut, n * 8, cudaMemcpyDeviceToHost));
    int err; CK(cudaMemcpy(&err, dErr, 4, cudaMemcpyDeviceToHost));
    size_t bad = 0; for (int p = 0; p < n; p++) if (out[p] != ref[p]) bad++;
    std::vector<unsigned long long> prof(size_t(nWarps) * 9); CK(cudaMemcpy(prof.data(), dProf, prof.size() * 8, cudaMemcpyDeviceToHost));
    double tot = 0, wait = 0, poll = 0, polls = 0, steps = 0, maxTot = 0; int used = 0;
    for (int w = 0; w < nWarps; w++) if (prof[w * 9 + 4]) { used++; tot += prof[w*9]; wait += prof[w*9+1]; poll += prof[w*9+2]; polls += prof[w*9+3]; steps += prof[w*9+4]; maxTot = std::max(maxTot, double(prof[w*9])); }
    printf("N=%d R=%d E=%d parts=%d blocks=%d occ=%d smem=%zu  best %.3f ms avg %.3f ms  mismatches %zu err %d\n", N, kStreamR, kStreamE,
           pl.nParts, blocks, occ, smem, best, reps > 1 ? sum / (reps - 1) : best, bad, err);
    printf("per compute warp (avg over %d): total %.0f cyc (max %.0f), steps %.0f, stage wait %.0f cyc (%.0f/step), poll %.0f cyc in %.0f polls (%.0f/poll), other %.0f/step\n",
           used, tot / used, maxTot, steps / used, wait / used, wait / steps, poll / used, polls / used, polls ? poll / polls : 0.0,
           (tot - wait - poll) / steps);
    for (int w = 0; w < std::min(nWarps, 12); w++) printf("  CTA %d: total %llu wait %llu poll %llu polls %llu steps %llu -> %.0f cyc/step excl. poll | per step: load+poll+sync %.0f, shfl+math %.0f, store+arrive %.0f, prefetch(+wait) %.0f\n", w, prof[w*9], prof[w*9+1], prof[w*9+2], prof[w*9+3], prof[w*9+4], double(prof[w*9]-prof[w*9+2])/std::max(1ull,prof[w*9+4]), double(prof[w*9+5])/prof[w*9+4], double(prof[w*9+6])/prof[w*9+4], double(prof[w*9+7])/prof[w*9+4], double(prof[w*9+8])/prof[w*9+4]);
    return bad || err;
}
