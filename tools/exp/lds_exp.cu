// Does random shared-memory gathering slow the global stream down (shared L1 data path)?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
__device__ __forceinline__ uint32_t pack16(uint4 w){const uint32_t M=0x00820820u;uint32_t p0=(w.x&0x06060606u)*M,p1=(w.y&0x06060606u)*M,p2=(w.z&0x06060606u)*M,p3=(w.w&0x06060606u)*M;return __byte_perm(__byte_perm(p0,p1,0x0073),__byte_perm(p2,p3,0x0073),0x5410);}
__device__ __forceinline__ void pf(const void*p){asm volatile("prefetch.global.L2 [%0];"::"l"(p));}
// NLDS random LDS per tile (conflicting), WORK extra independent-ish ALU/FMA instr pairs
template<int NLDS, int WORK, int BCAST>
__global__ void __launch_bounds__(1024,1) k(const uint4* __restrict__ in, uint32_t n_tiles, uint32_t T, uint32_t nw, unsigned long long* out){
  extern __shared__ uint32_t sm[];
  for(uint32_t i=threadIdx.x;i<nw;i+=blockDim.x) sm[i]=i*2654435761u;
  __syncthreads();
  uint32_t lane=threadIdx.x&31, warp=blockIdx.x*32+(threadIdx.x>>5), nwarps=gridDim.x*32; uint32_t acc=0;
  uint32_t n_spans=(n_tiles+T-1)/T;
  for(uint32_t span=warp; span<n_spans; span+=nwarps){
    uint32_t t=span*T, t1=min(t+T,n_tiles); if(t1+12>n_tiles) continue;
    const uint4* ptr=in+t*32+lane;
    pf((const char*)(in+t*32)+lane*128);
    uint4 wa=__ldcs(ptr), wb=__ldcs(ptr+32);
    for(;t<t1;t+=2){
      pf((const char*)ptr+8*512); pf((const char*)ptr+9*512);
      #pragma unroll
      for(int ph=0;ph<2;++ph){
        uint32_t x=pack16(ph?wb:wa);
        uint32_t y=x;
        #pragma unroll
        for(int i=0;i<NLDS;++i){ uint32_t h=(x+i)*0x9E3779B1u; uint32_t idx=BCAST? (__umulhi(h,nw)&~31u)+lane : __umulhi(h,nw); y+=sm[idx]; }
        #pragma unroll
        for(int i=0;i<WORK;++i){ y = y*0x9E3779B1u + (x>>((i&15)+1)); }
        acc+=y;
        if(ph) wb=__ldcs(ptr+96+32); else wa=__ldcs(ptr+64);
      }
      ptr+=64;
    }
  }
  for(int o=16;o;o>>=1) acc+=__shfl_xor_sync(FULL,acc,o);
  if(lane==0) atomicAdd(out,(unsigned long long)acc);
}
template<class F> float timeit(F f,int it=5){cudaEvent_t a,b;cudaEventCreate(&a);cudaEventCreate(&b);f();f();cudaDeviceSynchronize();float best=1e9;for(int i=0;i<it;++i){cudaEventRecord(a);f();cudaEventRecord(b);cudaEventSynchronize(b);float ms;cudaEventElapsedTime(&ms,a,b);if(ms<best)best=ms;}return best;}
int main(){
  size_t n=(size_t)2<<30; uint4* d; cudaMalloc(&d,n+(1<<20)); cudaMemset(d,0x41,n+(1<<20)); unsigned long long* out; cudaMalloc(&out,8);
  uint32_t n_tiles=n/512; int sm=148; uint32_t nw=50000; size_t smem=nw*4;
  #define R(name,NL,W,B) { cudaFuncSetAttribute(k<NL,W,B>, cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem); float ms=timeit([&]{k<NL,W,B><<<sm,1024,smem>>>(d,n_tiles,64,nw,out);}); cudaError_t e=cudaGetLastError(); printf("%-40s %.3f ms %.0f GB/s %s\n",name,ms,n/ms/1e6,e?cudaGetErrorString(e):""); }
  R("lds=0 work=0",0,0,0); R("lds=0 work=16",0,16,0); R("lds=0 work=24",0,24,0); R("lds=0 work=32",0,32,0);
  R("lds=2 work=16 (random)",2,16,0); R("lds=4 work=16 (random)",4,16,0); R("lds=4 work=8 (random)",4,8,0); R("lds=8 work=8 (random)",8,8,0);
  R("lds=4 work=16 (conflict-free)",4,16,1); R("lds=8 work=8 (conflict-free)",8,8,1);
  return 0;
}
