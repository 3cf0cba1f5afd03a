// prototype: CTA-owns-a-contiguous-range forward sweep with smem-resident values, cp.async prefetch of row records
// (stage A) and of external dependencies (stage B).  N^3 grid, ranges = R consecutive cells (natural order).
#include <cstdio>
#include <vector>
#include <algorithm>
#include <cmath>
#include <cuda_runtime.h>
static constexpr unsigned long long SENT = 0x7FF4B2005E471AE1ull;
__device__ __forceinline__ double ld_l2(const double* p){double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];":"=d"(v):"l"(p):"memory"); return v;}
__device__ __forceinline__ void st_l2(double* p,double v){asm volatile("st.relaxed.gpu.global.f64 [%0], %1;"::"l"(p),"d"(v):"memory");}
__device__ __forceinline__ bool isS(double v){return __double_as_longlong(v)==(long long)SENT;}
__device__ __forceinline__ void cp16(void* s,const void* g){unsigned a=(unsigned)__cvta_generic_to_shared(s); asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"::"r"(a),"l"(g):"memory");}
__device__ __forceinline__ void cp8(void* s,const void* g){unsigned a=(unsigned)__cvta_generic_to_shared(s); asm volatile("cp.async.ca.shared.global [%0], [%1], 8;"::"r"(a),"l"(g):"memory");}
__device__ __forceinline__ void cp8cg(void* s,const void* g){unsigned a=(unsigned)__cvta_generic_to_shared(s); asm volatile("cp.async.ca.shared.global [%0], [%1], 8;"::"r"(a),"l"(g):"memory");}
__device__ __forceinline__ void cpcommit(){asm volatile("cp.async.commit_group;":::"memory");}
template<int N> __device__ __forceinline__ void cpwait(){asm volatile("cp.async.wait_group %0;"::"n"(N):"memory");}
#ifndef DAV
#define DAV 8
#endif
#ifndef DBV
#define DBV 4
#endif
constexpr int T=128, DA=DAV, DB=DBV;
struct Args { int nRanges; const int* rangeStart; const int* levStart; const int* levOff; // levOff[levStart[r]+l] = first row (relative) of level l in range r
  const int4* cols4; const double* vals4; const double* rD; const double* in; double* out; int* err; int Rmax; long long* prof; };
__global__ void __launch_bounds__(T) k_range(Args a){
  extern __shared__ __align__(16) unsigned char sm[];
  double* sv=(double*)sm;                         // Rmax values
  int4* rcol=(int4*)(sv+a.Rmax);                  // [DA][T]
  double* rval=(double*)(rcol+DA*T);              // [DA][T][4]
  double* rrd=rval+DA*T*4;                        // [DA][T]
  double* rin=rrd+DA*T;                           // [DA][T]
  double* rext=rin+DA*T;                          // [DB][T][4]
  const int t=threadIdx.x;
  for(int r=blockIdx.x;r<a.nRanges;r+=gridDim.x){
    const int s=a.rangeStart[r]; const int* off=a.levOff+a.levStart[r]; const int nL=a.levStart[r+1]-a.levStart[r]-1;
    __syncthreads();
    auto issueA=[&](int lev){ if(lev<nL){ int p=s+off[lev]+t; if(off[lev]+t<off[lev+1]){ int sl=(lev%DA)*T+t; cp16(&rcol[sl],&a.cols4[p]); cp16(&rval[sl*4],&a.vals4[(size_t)p*4]); cp16(&rval[sl*4+2],&a.vals4[(size_t)p*4+2]); cp8(&rrd[sl],&a.rD[p]); cp8(&rin[sl],&a.in[p]); } } };
    auto issueB=[&](int lev){ if(lev<nL){ if(off[lev]+t<off[lev+1]){ int sl=(lev%DA)*T+t; int4 c=rcol[sl]; int sb=((lev%DB)*T+t)*4; int cc[4]={c.x,c.y,c.z,c.w};
#pragma unroll
          for(int k=0;k<4;k++) if(cc[k]>=0 && cc[k]<s) cp8cg(&rext[sb+k],&a.out[cc[k]]); } } };
    // prologue
    for(int l=0;l<DA;l++){ issueA(l); cpcommit(); }
    // stage B for levels 0..DB-1 needs A(0..DB-1): wait for groups
    cpwait<DA-DB>();   // oldest DB groups complete
    for(int l=0;l<DB;l++){ issueB(l); cpcommit(); }
    long long tA=0,tB=0,tC=0,tD=0;
    for(int lev=0;lev<nL;lev++){
      long long c0=clock64();
      cpwait<DB-1>();
      long long c1=clock64();  // everything except the most recent DB-1 groups is complete: A(lev), A(lev+DB), B(lev)
      const bool act = off[lev]+t<off[lev+1];
      if(act){ const int p=s+off[lev]+t; const int sl=(lev%DA)*T+t; const int sb=((lev%DB)*T+t)*4;
        const int4 c=rcol[sl]; const double rd=rrd[sl]; double acc=rd*rin[sl]; const int cc[4]={c.x,c.y,c.z,c.w};
#pragma unroll
        for(int k=0;k<4;k++){ if(cc[k]>=0){ double y; if(cc[k]>=s) y=sv[cc[k]-s]; else { y=rext[sb+k]; unsigned spins=0; while(isS(y)){ y=ld_l2(a.out+cc[k]); if(++spins>(1u<<22)){*a.err=1;break;} } }
            acc -= (rd*rval[sl*4+k])*y; } }
        sv[p-s]=acc; st_l2(a.out+p,acc); }
      long long c2=clock64();
      __syncthreads();
      long long c3=clock64();
      issueA(lev+DA); issueB(lev+DB); cpcommit();
      long long c4=clock64(); tA+=c1-c0; tB+=c2-c1; tC+=c3-c2; tD+=c4-c3;
    }
    if(t==0 && a.prof){ a.prof[0]=tA; a.prof[1]=tB; a.prof[2]=tC; a.prof[3]=tD; a.prof[4]=nL; }
    cpwait<0>();
  }
}
__global__ void fill(double* y,size_t n){ for(size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x;i<n;i+=(size_t)gridDim.x*blockDim.x) y[i]=__longlong_as_double((long long)SENT);} 
int main(int argc,char**argv){
  int N=argc>1?atoi(argv[1]):128; int R=argc>2?atoi(argv[2]):8192; int NZ=argc>3?atoi(argv[3]):N; int n=N*N*NZ; int nR=(n+R-1)/R;
  // positions: range-major, within range by internal level then cell
  std::vector<int> perm(n), ipos(n), rangeStart(nR+1), levStart(nR+1), levOff; 
  for(int r=0;r<nR;r++){ int c0=r*R, c1=std::min(n,c0+R); rangeStart[r]=c0; std::vector<int> lv(c1-c0,0); int maxl=0;
    for(int c=c0;c<c1;c++){ int i=c%N,j=(c/N)%N,k=c/(N*N); int l=0; if(i>0&&c-1>=c0) l=std::max(l,lv[c-1-c0]+1); if(j>0&&c-N>=c0) l=std::max(l,lv[c-N-c0]+1); if(k>0&&c-N*N>=c0) l=std::max(l,lv[c-N*N-c0]+1); lv[c-c0]=l; maxl=std::max(maxl,l);} 
    std::vector<int> cnt(maxl+2,0); for(int x:lv) cnt[x+1]++; for(int l=0;l<=maxl;l++) cnt[l+1]+=cnt[l];
    levStart[r]=levOff.size(); for(int l=0;l<=maxl+1;l++) levOff.push_back(cnt[l]);
    std::vector<int> cur(cnt.begin(),cnt.end()-1); for(int c=c0;c<c1;c++){ int q=c0+cur[lv[c-c0]]++; perm[q]=c; ipos[c]=q; } }
  rangeStart[nR]=n; levStart[nR]=levOff.size();
  int maxW=0; for(int r=0;r<nR;r++) for(int l=levStart[r];l<levStart[r+1]-1;l++) maxW=std::max(maxW,levOff[l+1]-levOff[l]);
  printf("N %d R %d ranges %d levels/range %d max level width %d\n",N,R,nR,levStart[1]-levStart[0]-1,maxW);
  std::vector<int4> cols4(n); std::vector<double> vals4((size_t)n*4,0.0), rD(n), in(n);
  for(int q=0;q<n;q++){ int c=perm[q]; int i=c%N,j=(c/N)%N,k=c/(N*N); int cc[4]={-1,-1,-1,-1}; int m=0; if(k>0){cc[m]=ipos[c-N*N]; vals4[(size_t)q*4+m]=-0.1-0.001*(c%7); m++;} if(j>0){cc[m]=ipos[c-N]; vals4[(size_t)q*4+m]=-0.11-0.001*(c%5); m++;} if(i>0){cc[m]=ipos[c-1]; vals4[(size_t)q*4+m]=-0.12-0.001*(c%3); m++;} cols4[q]=make_int4(cc[0],cc[1],cc[2],cc[3]); rD[q]=0.5+0.01*(c%5); in[q]=1.0+0.1*(c%3);} 
  int *dRS,*dLS,*dLO; int4* dC; double *dV,*drD,*din,*dout; int* err;
  cudaMalloc(&dRS,(nR+1)*4);cudaMalloc(&dLS,(nR+1)*4);cudaMalloc(&dLO,levOff.size()*4);cudaMalloc(&dC,(size_t)n*16);cudaMalloc(&dV,(size_t)n*32);cudaMalloc(&drD,n*8);cudaMalloc(&din,n*8);cudaMalloc(&dout,n*8);cudaMalloc(&err,4);cudaMemset(err,0,4);
  cudaMemcpy(dRS,rangeStart.data(),(nR+1)*4,cudaMemcpyHostToDevice);cudaMemcpy(dLS,levStart.data(),(nR+1)*4,cudaMemcpyHostToDevice);cudaMemcpy(dLO,levOff.data(),levOff.size()*4,cudaMemcpyHostToDevice);cudaMemcpy(dC,cols4.data(),(size_t)n*16,cudaMemcpyHostToDevice);cudaMemcpy(dV,vals4.data(),(size_t)n*32,cudaMemcpyHostToDevice);cudaMemcpy(drD,rD.data(),n*8,cudaMemcpyHostToDevice);cudaMemcpy(din,in.data(),n*8,cudaMemcpyHostToDevice);
  long long* dprof; cudaMalloc(&dprof,64); cudaMemset(dprof,0,64);
  Args a{nR,dRS,dLS,dLO,dC,dV,drD,din,dout,err,R,dprof};
  size_t smem=(size_t)R*8 + (size_t)DA*T*16 + (size_t)DA*T*32 + (size_t)DA*T*8*2 + (size_t)DB*T*32;
  cudaFuncSetAttribute(k_range,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem);
  int occ=0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,k_range,T,smem); int blocks=std::min(nR,occ*148);
  printf("smem %zu bytes occ %d blocks %d\n",smem,occ,blocks);
  float best=1e9; for(int rep=0;rep<5;rep++){ fill<<<1024,256>>>(dout,n); cudaEvent_t e0,e1;cudaEventCreate(&e0);cudaEventCreate(&e1);cudaEventRecord(e0); void* args[]={&a};
    cudaLaunchCooperativeKernel((void*)k_range,dim3(blocks),dim3(T),args,smem,0); cudaEventRecord(e1);cudaEventSynchronize(e1);float ms;cudaEventElapsedTime(&ms,e0,e1); if(rep>0) best=std::min(best,ms);} 
  printf("range sweep: best %.3f ms (%s)\n",best,cudaGetErrorString(cudaGetLastError()));
  std::vector<double> ho(n),ref(n); cudaMemcpy(ho.data(),dout,n*8,cudaMemcpyDeviceToHost);
  // reference in cell order
  std::vector<double> rc(n); for(int c=0;c<n;c++){ int q=ipos[c]; double acc=rD[q]*in[q]; int4 cc=cols4[q]; int ccs[4]={cc.x,cc.y,cc.z,cc.w}; for(int k=0;k<4;k++) if(ccs[k]>=0) acc-=(rD[q]*vals4[(size_t)q*4+k])*rc[perm[ccs[k]]]; rc[c]=acc; }
  long long hp[5]; cudaMemcpy(hp,dprof,40,cudaMemcpyDeviceToHost); printf("cycles/level: wait %.0f compute %.0f barrier %.0f issue %.0f (levels %lld)\n",(double)hp[0]/hp[4],(double)hp[1]/hp[4],(double)hp[2]/hp[4],(double)hp[3]/hp[4],hp[4]);
  int bad=0; for(int q=0;q<n;q++) if(ho[q]!=rc[perm[q]]) bad++; int h; cudaMemcpy(&h,err,4,cudaMemcpyDeviceToHost); printf("mismatches %d err %d\n",bad,h);
  return 0; }
