// microbenchmark: latency of a producer->consumer hop through L2 with the sentinel protocol
#include <cstdio>
#include <cuda_runtime.h>
static constexpr unsigned long long SENT = 0x7FF4B2005E471AE1ull;
__device__ __forceinline__ double ld_l2(const double* p){double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];":"=d"(v):"l"(p):"memory"); return v;}
__device__ __forceinline__ void st_l2(double* p,double v){asm volatile("st.relaxed.gpu.global.f64 [%0], %1;"::"l"(p),"d"(v):"memory");}
// each warp w handles hops w, w+W, ...; hop h waits for y[(h-1)*stride + lane] then writes y[h*stride+lane]
template<int SLEEP>
__global__ void chain(double* y, int nHops, int stride, int lanes){
  int wpb=blockDim.x>>5; int W=gridDim.x*wpb; int lane=threadIdx.x&31;
  for(int h=blockIdx.x*wpb+(threadIdx.x>>5); h<nHops; h+=W){
    if(lane<lanes){
      double v=1.0;
      if(h>0){
        const double* p=y+(size_t)(h-1)*stride+lane;
        v=ld_l2(p); unsigned spins=0;
        while(__double_as_longlong(v)==(long long)SENT){ v=ld_l2(p); if(SLEEP && ++spins>8) __nanosleep(SLEEP);} 
      }
      st_l2(y+(size_t)h*stride+lane, v+1.0);
    }
  }
}
__global__ void fill(double* y,size_t n){ for(size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x;i<n;i+=(size_t)gridDim.x*blockDim.x) y[i]=__longlong_as_double((long long)SENT);} 
template<int SLEEP> float run(double* y,int nHops,int stride,int lanes,int blocks,int threads){
  fill<<<1024,256>>>(y,(size_t)nHops*stride); cudaEvent_t a,b; cudaEventCreate(&a);cudaEventCreate(&b);
  void* args[]={&y,&nHops,&stride,&lanes};
  cudaEventRecord(a); cudaLaunchCooperativeKernel((void*)chain<SLEEP>,dim3(blocks),dim3(threads),args,0,0); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b); cudaError_t e=cudaGetLastError(); if(e) printf("err %s\n",cudaGetErrorString(e)); return ms; }
int main(){ int nHops=4000, stride=32; double* y; cudaMalloc(&y,(size_t)nHops*stride*8);
  int cfg[][2]={{2,32},{148,32},{148,256},{592,256},{1184,256}};
  for(auto& c:cfg){ for(int lanes: {1,32}){
    float t0=run<0>(y,nHops,stride,lanes,c[0],c[1]); t0=run<0>(y,nHops,stride,lanes,c[0],c[1]);
    float t1=run<40>(y,nHops,stride,lanes,c[0],c[1]); t1=run<40>(y,nHops,stride,lanes,c[0],c[1]);
    float t2=run<500>(y,nHops,stride,lanes,c[0],c[1]); t2=run<500>(y,nHops,stride,lanes,c[0],c[1]);
    printf("blocks %4d threads %3d lanes %2d : us/hop nosleep %.3f sleep40 %.3f sleep500 %.3f\n",c[0],c[1],lanes,t0*1e3/nHops,t1*1e3/nHops,t2*1e3/nHops);} }
  return 0; }
