// harness 3: what sets the per-wavefront latency? variants of the polling loop on an N^3 grid forward sweep
#include <cstdio>
#include <vector>
#include <algorithm>
#include "../../openfoam-dev_b200/csrc/kernels.cuh"
#include "../../openfoam-dev_b200/csrc/mesh.hpp"
using namespace b200ls;
__device__ __forceinline__ double ld_cv(const double* p){double v; asm volatile("ld.volatile.global.f64 %0, [%1];":"=d"(v):"l"(p):"memory"); return v;}
__device__ __forceinline__ double ld_acq(const double* p){double v; asm volatile("ld.acquire.gpu.global.f64 %0, [%1];":"=d"(v):"l"(p):"memory"); return v;}
// MODE 0: baseline all-poll. 1: only first dep honoured (others read plain, wrong values but same traffic) 2: double interleaved polling 3: volatile loads
template<int MODE>
__device__ __forceinline__ double gather(double acc, double scale, int j0, int j1, const int* __restrict__ col, const double* __restrict__ val, const double* y, int* err){
    const int n = min(4, j1 - j0); int c[4]; double v[4], w[4], w2[4];
#pragma unroll
    for (int k=0;k<4;k++) if(k<n){c[k]=col[j0+k]; v[k]=val[j0+k];}
#pragma unroll
    for (int k=0;k<4;k++) if(k<n) w[k]= (MODE==3)? ld_cv(y+c[k]) : ld_l2(y+c[k]);
    if (MODE==2) {
        __nanosleep(60);
#pragma unroll
        for (int k=0;k<4;k++) if(k<n) w2[k]=ld_l2(y+c[k]);
    }
    unsigned spins=0;
    const int nchk = (MODE==1)? min(n,1) : n;
    while(true){ bool pend=false;
#pragma unroll
        for(int k=0;k<4;k++) if(k<nchk&&is_sentinel(w[k])) pend=true;
        if(!pend) break; if(++spins>kMaxSpins){*err=1;break;}
        if (MODE==2) {
#pragma unroll
            for(int k=0;k<4;k++) if(k<nchk&&is_sentinel(w[k])) { w[k]=w2[k]; w2[k]=ld_l2(y+c[k]); }
        } else {
#pragma unroll
            for(int k=0;k<4;k++) if(k<nchk&&is_sentinel(w[k])) w[k]= (MODE==3)? ld_cv(y+c[k]) : ld_l2(y+c[k]);
        }
    }
#pragma unroll
    for(int k=0;k<4;k++) if(k<n) acc -= (scale*v[k])*(is_sentinel(w[k])?1.0:w[k]);
    return acc;
}
template<int MODE>
__global__ void __launch_bounds__(256) k_dbg(SweepArgs a){
    SWEEP_TASK_LOOP(a) {
        const int2 next = SWEEP_NEXT_TASK(a);
        if (lane < task.y) {
            const int p = task.x + lane;
            const double rd = a.rD[p];
            const int j0 = a.ptr[p], j1 = a.ptr[p + 1];
            double acc = rd * a.in[p];
            acc = gather<MODE>(acc, rd, j0, j1, a.col, a.val, a.out, a.err);
            st_l2(a.out + p, acc);
        }
        __syncwarp();
        task = next;
    }
}
int N,bpsm; SweepArgs a; double* outp; int n;
template<int MODE> void run(const char* name){
  int blocks=std::min(148*bpsm,(a.nTasks+7)/8); float best=1e9;
  for(int rep=0;rep<5;rep++){ k_fill_sentinel<<<1024,256>>>(outp,n);
    cudaEvent_t e0,e1;cudaEventCreate(&e0);cudaEventCreate(&e1);cudaEventRecord(e0);
    void* args[]={&a}; cudaLaunchCooperativeKernel((void*)k_dbg<MODE>,dim3(blocks),dim3(256),args,0,0);
    cudaEventRecord(e1);cudaEventSynchronize(e1);float ms;cudaEventElapsedTime(&ms,e0,e1); if(rep>0) best=std::min(best,ms);}
  printf("%-28s blocks %4d : best %.3f ms (%s)\n",name,blocks,best,cudaGetErrorString(cudaGetLastError()));
}
int main(int argc,char**argv){
  N=argc>1?atoi(argv[1]):128; bpsm=argc>2?atoi(argv[2]):4; int NZ = argc>3?atoi(argv[3]):N;
  std::vector<int32_t> lo,up; 
  for(int k=0;k<NZ;k++)for(int j=0;j<N;j++)for(int i=0;i<N;i++){int c=i+N*(j+N*k); if(i<N-1){lo.push_back(c);up.push_back(c+1);} if(j<N-1){lo.push_back(c);up.push_back(c+N);} if(k<NZ-1){lo.push_back(c);up.push_back(c+N*N);}}
  LevelHost L; buildLevel(L,N*N*NZ,(int)lo.size(),lo.data(),up.data(),{});
  n=L.nCells; int nF=L.nFaces; int nT=L.fwdTasks.size();
  int *Lptr,*Lcol; double *Lval,*rD,*in; int2* tasks; int* err;
  cudaMalloc(&Lptr,(n+1)*4);cudaMalloc(&Lcol,nF*4);cudaMalloc(&Lval,nF*8);cudaMalloc(&rD,n*8);cudaMalloc(&in,n*8);cudaMalloc(&outp,n*8);cudaMalloc(&tasks,nT*8);cudaMalloc(&err,4);
  cudaMemcpy(Lptr,L.Lptr.data(),(n+1)*4,cudaMemcpyHostToDevice);cudaMemcpy(Lcol,L.Lcol.data(),nF*4,cudaMemcpyHostToDevice);cudaMemcpy(tasks,L.fwdTasks.data(),nT*8,cudaMemcpyHostToDevice);
  std::vector<double> v(nF,-0.1),d(n,0.5),b(n,1.0); cudaMemcpy(Lval,v.data(),nF*8,cudaMemcpyHostToDevice);cudaMemcpy(rD,d.data(),n*8,cudaMemcpyHostToDevice);cudaMemcpy(in,b.data(),n*8,cudaMemcpyHostToDevice);cudaMemset(err,0,4);
  a=SweepArgs{}; a.tasks=tasks;a.nTasks=nT;a.ptr=Lptr;a.col=Lcol;a.val=Lval;a.rD=rD;a.in=in;a.out=outp;a.err=err;
  printf("N %d x %d x %d levels %zu tasks %d => ", N,N,NZ,L.fwdOffsets.size()-1,nT);
  printf("\n"); run<0>("baseline"); run<1>("first-dep-only"); run<2>("double-poll"); run<3>("volatile");
  float ms0; { int blocks=std::min(148*bpsm,(a.nTasks+7)/8); (void)blocks; }
  return 0;}
