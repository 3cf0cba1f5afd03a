// debug harness: forward sweep on an N^3 grid with per-task completion timestamps
#include <cstdio>
#include <vector>
#include <algorithm>
#include "../openfoam-dev_b200/csrc/kernels.cuh"
#include "../openfoam-dev_b200/csrc/mesh.hpp"
using namespace b200ls;
__device__ __forceinline__ unsigned long long gtime(){unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;":"=l"(t)); return t;}
__global__ void __launch_bounds__(256) k_dbg(SweepArgs a, unsigned long long* ts, unsigned long long* ts0, int* sm){
    SWEEP_TASK_LOOP(a) {
        const int2 next = SWEEP_NEXT_TASK(a);
        if (lane==0) { ts0[t]=gtime(); unsigned s; asm("mov.u32 %0, %smid;":"=r"(s)); sm[t]=s; }
        if (lane < task.y) {
            const int p = task.x + lane;
            const double rd = a.rD[p];
            const int j0 = a.ptr[p], j1 = a.ptr[p + 1];
            double acc = rd * a.in[p];
            acc = gather_deps<false, true>(acc, rd, j0, j1, a.col, a.val, a.out, a.err);
            st_l2(a.out + p, acc);
        }
        __syncwarp();
        if (lane==0) ts[t]=gtime();
        task = next;
    }
}
int main(int argc,char**argv){
  int N=argc>1?atoi(argv[1]):128; int bpsm=argc>2?atoi(argv[2]):4;
  std::vector<int32_t> lo,up; 
  for(int k=0;k<N;k++)for(int j=0;j<N;j++)for(int i=0;i<N;i++){int c=i+N*(j+N*k); if(i<N-1){lo.push_back(c);up.push_back(c+1);} if(j<N-1){lo.push_back(c);up.push_back(c+N);} if(k<N-1){lo.push_back(c);up.push_back(c+N*N);}}
  LevelHost L; buildLevel(L,N*N*N,(int)lo.size(),lo.data(),up.data(),{});
  int n=L.nCells,nF=L.nFaces; int nT=L.fwdTasks.size();
  int *Lptr,*Lcol; double *Lval,*rD,*in,*out; int2* tasks; int* err; unsigned long long *ts,*ts0; int* sm;
  cudaMalloc(&Lptr,(n+1)*4);cudaMalloc(&Lcol,nF*4);cudaMalloc(&Lval,nF*8);cudaMalloc(&rD,n*8);cudaMalloc(&in,n*8);cudaMalloc(&out,n*8);cudaMalloc(&tasks,nT*8);cudaMalloc(&err,4);cudaMalloc(&ts,nT*8);cudaMalloc(&ts0,nT*8);cudaMalloc(&sm,nT*4);
  cudaMemcpy(Lptr,L.Lptr.data(),(n+1)*4,cudaMemcpyHostToDevice);cudaMemcpy(Lcol,L.Lcol.data(),nF*4,cudaMemcpyHostToDevice);cudaMemcpy(tasks,L.fwdTasks.data(),nT*8,cudaMemcpyHostToDevice);
  std::vector<double> v(nF,-0.1),d(n,0.5),b(n,1.0); cudaMemcpy(Lval,v.data(),nF*8,cudaMemcpyHostToDevice);cudaMemcpy(rD,d.data(),n*8,cudaMemcpyHostToDevice);cudaMemcpy(in,b.data(),n*8,cudaMemcpyHostToDevice);cudaMemset(err,0,4);
  SweepArgs a{}; a.tasks=tasks;a.nTasks=nT;a.ptr=Lptr;a.col=Lcol;a.val=Lval;a.rD=rD;a.in=in;a.out=out;a.err=err;
  int blocks=std::min(148*bpsm,(nT+7)/8);
  for(int rep=0;rep<3;rep++){
    k_fill_sentinel<<<1024,256>>>(out,n);
    cudaEvent_t e0,e1;cudaEventCreate(&e0);cudaEventCreate(&e1);cudaEventRecord(e0);
    void* args[]={&a,&ts,&ts0,&sm}; cudaLaunchCooperativeKernel((void*)k_dbg,dim3(blocks),dim3(256),args,0,0);
    cudaEventRecord(e1);cudaEventSynchronize(e1);float ms;cudaEventElapsedTime(&ms,e0,e1);printf("N %d blocks %d ms %.3f err %s\n",N,blocks,ms,cudaGetErrorString(cudaGetLastError()));
  }
  std::vector<unsigned long long> hts(nT),hts0(nT); std::vector<int> hsm(nT); cudaMemcpy(hts.data(),ts,nT*8,cudaMemcpyDeviceToHost);cudaMemcpy(hts0.data(),ts0,nT*8,cudaMemcpyDeviceToHost);cudaMemcpy(hsm.data(),sm,nT*4,cudaMemcpyDeviceToHost);
  unsigned long long t0=*std::min_element(hts0.begin(),hts0.end());
  // per level stats
  int ti=0; int nLev=L.fwdOffsets.size()-1; 
  for(int k=0;k<nLev;k++){ int nt=(L.fwdOffsets[k+1]-L.fwdOffsets[k]+31)/32; unsigned long long mn=~0ull,mx=0,smn=~0ull,smx=0; for(int q=0;q<nt;q++,ti++){mn=std::min(mn,hts[ti]);mx=std::max(mx,hts[ti]);smn=std::min(smn,hts0[ti]);smx=std::max(smx,hts0[ti]);}
    if(k%20==0||k>nLev-3) printf("level %4d tasks %4d start[min %.2f max %.2f] done[min %.2f max %.2f] us\n",k,nt,(smn-t0)/1e3,(smx-t0)/1e3,(mn-t0)/1e3,(mx-t0)/1e3);}
  // a few tasks of the middle level
  return 0;}
