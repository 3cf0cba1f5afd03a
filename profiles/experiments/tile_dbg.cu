// feasibility harness: warp-tile forward sweep on an N^3 grid (tiles a x b x c = 32 rows, internal levels in-warp)
#include <cstdio>
#include <vector>
#include <algorithm>
#include <numeric>
#include "../../openfoam-dev_b200/csrc/kernels.cuh"
using namespace b200ls;
struct Args { const int2* tasks; int nTasks; const int* ptr; const int* col; const double* val; const double* rD; const double* in; double* out; const unsigned char* lev; int* err; };
template<int JIT>
__global__ void __launch_bounds__(256) k_tile(Args a){
    __shared__ double sy[8][32];
    const int wpb = blockDim.x >> 5, nW = gridDim.x * wpb, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int t = blockIdx.x * wpb + wib; t < a.nTasks; t += nW) {
        const int2 task = a.tasks[t];
        const int start = task.x, count = task.y & 0xff, nLev = task.y >> 8;
        const bool act = lane < count;
        int p = start + lane; double rd=0, acc=0; int n=0; int c[4]; double v[4], w[4]; int lev = -1; bool ext[4];
        if (act) {
            rd = a.rD[p]; acc = rd * a.in[p]; const int j0 = a.ptr[p]; n = a.ptr[p+1]-j0; lev = a.lev[p];
#pragma unroll
            for (int k=0;k<4;k++) if (k<n) { c[k]=a.col[j0+k]; v[k]=a.val[j0+k]; ext[k] = (unsigned)(c[k]-start) >= (unsigned)count; }
#pragma unroll
            for (int k=0;k<4;k++) if (k<n && ext[k]) w[k]=ld_l2(a.out+c[k]);
            if (!JIT) { // wait for all external deps up-front
                unsigned spins=0; while(true){ bool pend=false;
#pragma unroll
                    for(int k=0;k<4;k++) if(k<n&&ext[k]&&is_sentinel(w[k])) pend=true;
                    if(!pend) break; if(++spins>kMaxSpins){*a.err=1;break;}
#pragma unroll
                    for(int k=0;k<4;k++) if(k<n&&ext[k]&&is_sentinel(w[k])) w[k]=ld_l2(a.out+c[k]); }
            }
        }
        __syncwarp();
        for (int q=0;q<nLev;q++){
            if (act && lev==q){
                if (JIT) { unsigned spins=0; while(true){ bool pend=false;
#pragma unroll
                    for(int k=0;k<4;k++) if(k<n&&ext[k]&&is_sentinel(w[k])) pend=true;
                    if(!pend) break; if(++spins>kMaxSpins){*a.err=1;break;}
#pragma unroll
                    for(int k=0;k<4;k++) if(k<n&&ext[k]&&is_sentinel(w[k])) w[k]=ld_l2(a.out+c[k]); } }
#pragma unroll
                for (int k=0;k<4;k++) if (k<n) { const double yk = ext[k] ? w[k] : sy[wib][c[k]-start]; acc -= (rd*v[k])*yk; }
                sy[wib][lane]=acc; st_l2(a.out+p, acc);
            }
            __syncwarp();
        }
    }
}
int main(int argc,char**argv){
  int N=argc>1?atoi(argv[1]):128; int ta=argc>2?atoi(argv[2]):2, tb=argc>3?atoi(argv[3]):4, tc=argc>4?atoi(argv[4]):4; int bpsm=argc>5?atoi(argv[5]):4;
  int n=N*N*N; int TX=N/ta,TY=N/tb,TZ=N/tc; int nT=TX*TY*TZ; int cnt=ta*tb*tc;
  // tile order: by tile level then index
  std::vector<int> order(nT); std::iota(order.begin(),order.end(),0);
  auto tl=[&](int t){int x=t%TX,y=(t/TX)%TY,z=t/(TX*TY);return x+y+z;};
  std::stable_sort(order.begin(),order.end(),[&](int A,int B){return tl(A)<tl(B);});
  std::vector<int> ipos(n), perm(n); std::vector<int2> tasks(nT); std::vector<unsigned char> lev(n);
  int p=0; for(int o=0;o<nT;o++){int t=order[o];int x=t%TX,y=(t/TX)%TY,z=t/(TX*TY); tasks[o]=make_int2(p, cnt | ((ta+tb+tc-2)<<8));
    for(int k=0;k<tc;k++)for(int j=0;j<tb;j++)for(int i=0;i<ta;i++){int c=(x*ta+i)+N*((y*tb+j)+N*(z*tc+k)); perm[p]=c; ipos[c]=p; lev[p]=i+j+k; p++;}}
  std::vector<int> Lptr(n+1,0),Lcol; Lcol.reserve(3*n);
  for(int q=0;q<n;q++){int c=perm[q];int i=c%N,j=(c/N)%N,k=c/(N*N); if(k>0)Lcol.push_back(ipos[c-N*N]); if(j>0)Lcol.push_back(ipos[c-N]); if(i>0)Lcol.push_back(ipos[c-1]); Lptr[q+1]=Lcol.size();}
  int nF=Lcol.size();
  int *dLptr,*dLcol; double *Lval,*rD,*in,*out; int2* dtasks; int* err; unsigned char* dlev;
  cudaMalloc(&dLptr,(n+1)*4);cudaMalloc(&dLcol,nF*4);cudaMalloc(&Lval,nF*8);cudaMalloc(&rD,n*8);cudaMalloc(&in,n*8);cudaMalloc(&out,n*8);cudaMalloc(&dtasks,nT*8);cudaMalloc(&err,4);cudaMalloc(&dlev,n);
  cudaMemcpy(dLptr,Lptr.data(),(n+1)*4,cudaMemcpyHostToDevice);cudaMemcpy(dLcol,Lcol.data(),nF*4,cudaMemcpyHostToDevice);cudaMemcpy(dtasks,tasks.data(),nT*8,cudaMemcpyHostToDevice);cudaMemcpy(dlev,lev.data(),n,cudaMemcpyHostToDevice);
  std::vector<double> v(nF,-0.1),d(n,0.5),b(n,1.0); cudaMemcpy(Lval,v.data(),nF*8,cudaMemcpyHostToDevice);cudaMemcpy(rD,d.data(),n*8,cudaMemcpyHostToDevice);cudaMemcpy(in,b.data(),n*8,cudaMemcpyHostToDevice);cudaMemset(err,0,4);
  Args a{dtasks,nT,dLptr,dLcol,Lval,rD,in,out,dlev,err};
  int blocks=std::min(148*bpsm,(nT+7)/8);
  for(int jit=0;jit<2;jit++){ float best=1e9; for(int rep=0;rep<5;rep++){ k_fill_sentinel<<<1024,256>>>(out,n); cudaEvent_t e0,e1;cudaEventCreate(&e0);cudaEventCreate(&e1);cudaEventRecord(e0); void* args[]={&a};
      if(jit) cudaLaunchCooperativeKernel((void*)k_tile<1>,dim3(blocks),dim3(256),args,0,0); else cudaLaunchCooperativeKernel((void*)k_tile<0>,dim3(blocks),dim3(256),args,0,0);
      cudaEventRecord(e1);cudaEventSynchronize(e1);float ms;cudaEventElapsedTime(&ms,e0,e1); if(rep>0) best=std::min(best,ms);} 
    printf("N %d tile %dx%dx%d blocks %d jit %d: best %.3f ms (%s)\n",N,ta,tb,tc,blocks,jit,best,cudaGetErrorString(cudaGetLastError())); }
  // verify against host reference
  std::vector<double> ho(n), ref(n); cudaMemcpy(ho.data(),out,n*8,cudaMemcpyDeviceToHost);
  for(int c=0;c<n;c++){int i=c%N,j=(c/N)%N,k=c/(N*N); double acc=0.5*1.0; if(k>0) acc-=(0.5*-0.1)*ref[c-N*N]; if(j>0) acc-=(0.5*-0.1)*ref[c-N]; if(i>0) acc-=(0.5*-0.1)*ref[c-1]; ref[c]=acc;}
  double md=0; for(int q=0;q<n;q++) md=std::max(md,fabs(ho[q]-ref[perm[q]])); int h; cudaMemcpy(&h,err,4,cudaMemcpyDeviceToHost); printf("max diff %.3e err %d\n",md,h);
  return 0;}
