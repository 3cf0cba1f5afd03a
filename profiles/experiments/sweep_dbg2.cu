// debug harness 2: forward sweep variants on an N^3 grid (cold/warm L2, polling strategy, prefetch, clear)
#include <cstdio>
#include <vector>
#include <algorithm>
#include "../../openfoam-dev_b200/csrc/kernels.cuh"
#include "../../openfoam-dev_b200/csrc/mesh.hpp"
using namespace b200ls;
__device__ __forceinline__ bool isS(double v){return __double_as_longlong(v)==(long long)kSentinelBits;}
template <bool DESC>
__device__ __forceinline__ double gather_all(double acc, double scale, int j0, int j1, const int* __restrict__ col,
                                              const double* __restrict__ val, const double* y, int* err) {
    for (int base = 0; base < j1 - j0; base += 4) {
        const int n = min(4, j1 - j0 - base);
        int c[4]; double v[4], w[4];
#pragma unroll
        for (int k = 0; k < 4; k++) if (k < n) { const int j = DESC ? (j1 - 1 - base - k) : (j0 + base + k); c[k] = col[j]; v[k] = val[j]; }
#pragma unroll
        for (int k = 0; k < 4; k++) if (k < n) w[k] = ld_l2(y + c[k]);
        unsigned spins=0;
        while (true) {
            bool pend=false;
#pragma unroll
            for (int k = 0; k < 4; k++) if (k < n && isS(w[k])) pend=true;
            if (!pend) break;
            if (++spins > (1u<<22)) { *err=1; break; }
#pragma unroll
            for (int k = 0; k < 4; k++) if (k < n && isS(w[k])) w[k] = ld_l2(y + c[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) if (k < n) acc -= (scale * v[k]) * w[k];
    }
    return acc;
}
template<int POLL, int PREF, int CLEAR, int SYNC>
__global__ void __launch_bounds__(256) k_dbg(SweepArgs a){
    const double sent = sentinel();
    SWEEP_TASK_LOOP(a) {
        const int2 next = SWEEP_NEXT_TASK(a);
        if (lane < task.y) {
            const int p = task.x + lane;
            if (PREF) asm volatile("prefetch.global.L2 [%0];"::"l"(a.out+p));
            const double rd = a.rD[p];
            const int j0 = a.ptr[p], j1 = a.ptr[p + 1];
            double acc = rd * a.in[p];
            if (POLL==0) acc = gather_deps<false, true>(acc, rd, j0, j1, a.col, a.val, a.out, a.err);
            else acc = gather_all<false>(acc, rd, j0, j1, a.col, a.val, a.out, a.err);
            st_l2(a.out + p, acc);
            if (CLEAR) a.clear[p] = sent;
        }
        if (SYNC) __syncwarp();
        task = next;
    }
}
__global__ void k_flush(double* f,size_t n){ for(size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x;i<n;i+=(size_t)gridDim.x*blockDim.x) f[i]=f[i]*1.0000001+1.0;}
int N,bpsm; SweepArgs a; double* outp; int n; double* flushbuf; size_t nflush=64u<<20; // 512MB
template<int POLL,int PREF,int CLEAR,int SYNC> void run(const char* name,int cold,int refill){
  int blocks=std::min(148*bpsm,(a.nTasks+7)/8); float best=1e9, sum=0; int reps=4;
  for(int rep=0;rep<reps;rep++){
    if(refill||rep==0) k_fill_sentinel<<<1024,256>>>(outp,n);
    if(cold) k_flush<<<2048,256>>>(flushbuf,nflush);
    if(!refill && rep>0){ /* out still has values: refill needed anyway */ k_fill_sentinel<<<1024,256>>>(outp,n); if(cold) k_flush<<<2048,256>>>(flushbuf,nflush);} 
    cudaEvent_t e0,e1;cudaEventCreate(&e0);cudaEventCreate(&e1);cudaEventRecord(e0);
    void* args[]={&a}; cudaLaunchCooperativeKernel((void*)k_dbg<POLL,PREF,CLEAR,SYNC>,dim3(blocks),dim3(256),args,0,0);
    cudaEventRecord(e1);cudaEventSynchronize(e1);float ms;cudaEventElapsedTime(&ms,e0,e1); if(rep>0){best=std::min(best,ms);sum+=ms;}
  }
  printf("%-28s cold %d blocks %4d : best %.3f ms avg %.3f ms  (%s)\n",name,cold,blocks,best,sum/(reps-1),cudaGetErrorString(cudaGetLastError()));
}
int main(int argc,char**argv){
  N=argc>1?atoi(argv[1]):128; bpsm=argc>2?atoi(argv[2]):4;
  std::vector<int32_t> lo,up; 
  for(int k=0;k<N;k++)for(int j=0;j<N;j++)for(int i=0;i<N;i++){int c=i+N*(j+N*k); if(i<N-1){lo.push_back(c);up.push_back(c+1);} if(j<N-1){lo.push_back(c);up.push_back(c+N);} if(k<N-1){lo.push_back(c);up.push_back(c+N*N);}}
  LevelHost L; buildLevel(L,N*N*N,(int)lo.size(),lo.data(),up.data(),{});
  n=L.nCells; int nF=L.nFaces; int nT=L.fwdTasks.size();
  int *Lptr,*Lcol; double *Lval,*rD,*in,*clr; int2* tasks; int* err;
  cudaMalloc(&Lptr,(n+1)*4);cudaMalloc(&Lcol,nF*4);cudaMalloc(&Lval,nF*8);cudaMalloc(&rD,n*8);cudaMalloc(&in,n*8);cudaMalloc(&outp,n*8);cudaMalloc(&clr,n*8);cudaMalloc(&tasks,nT*8);cudaMalloc(&err,4);cudaMalloc(&flushbuf,nflush*8);cudaMemset(flushbuf,0,nflush*8);
  cudaMemcpy(Lptr,L.Lptr.data(),(n+1)*4,cudaMemcpyHostToDevice);cudaMemcpy(Lcol,L.Lcol.data(),nF*4,cudaMemcpyHostToDevice);cudaMemcpy(tasks,L.fwdTasks.data(),nT*8,cudaMemcpyHostToDevice);
  std::vector<double> v(nF,-0.1),d(n,0.5),b(n,1.0); cudaMemcpy(Lval,v.data(),nF*8,cudaMemcpyHostToDevice);cudaMemcpy(rD,d.data(),n*8,cudaMemcpyHostToDevice);cudaMemcpy(in,b.data(),n*8,cudaMemcpyHostToDevice);cudaMemset(err,0,4);
  a=SweepArgs{}; a.tasks=tasks;a.nTasks=nT;a.ptr=Lptr;a.col=Lcol;a.val=Lval;a.rD=rD;a.in=in;a.out=outp;a.err=err;a.clear=clr;
  printf("N %d levels %zu tasks %d\n",N,L.fwdOffsets.size()-1,nT);
  for(int cold=0;cold<2;cold++){
    run<0,0,0,0>("seqpoll",cold,1); run<0,0,0,1>("seqpoll+sync",cold,1); run<1,0,0,0>("allpoll",cold,1); run<1,0,0,1>("allpoll+sync",cold,1); run<1,1,1,1>("allpoll+pref+clear+sync",cold,1); run<1,0,1,1>("allpoll+clear+sync",cold,1);
  }
  int h; cudaMemcpy(&h,err,4,cudaMemcpyDeviceToHost); printf("err flag %d\n",h);
  return 0;}
