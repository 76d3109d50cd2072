// Streaming-structure experiments: how fast can 148 persistent CTAs of 32 warps read a
// buffer with 16-byte loads under different assignment / prefetch schemes?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
__device__ __forceinline__ uint32_t pack16(uint4 w){const uint32_t M=0x00820820u;uint32_t p0=(w.x&0x06060606u)*M,p1=(w.y&0x06060606u)*M,p2=(w.z&0x06060606u)*M,p3=(w.w&0x06060606u)*M;return __byte_perm(__byte_perm(p0,p1,0x0073),__byte_perm(p2,p3,0x0073),0x5410);}
__device__ __forceinline__ void pf(const void*p){asm volatile("prefetch.global.L2 [%0];"::"l"(p));}

// mode 0: spans per warp (T tiles), depth-2 register pipeline, L2 prefetch PF tiles ahead
template<int PFT, int WORK>
__global__ void __launch_bounds__(1024,1) k_span(const uint4* __restrict__ in, uint32_t n_tiles, uint32_t T, unsigned long long* out){
  uint32_t lane=threadIdx.x&31, warp=blockIdx.x*32+(threadIdx.x>>5), nwarps=gridDim.x*32; uint32_t acc=0;
  uint32_t n_spans=(n_tiles+T-1)/T;
  for(uint32_t span=warp; span<n_spans; span+=nwarps){
    uint32_t t=span*T, t1=min(t+T,n_tiles); if(t1+PFT+2>n_tiles) continue; // skip tail for simplicity
    const uint4* ptr=in+t*32+lane;
    if(PFT) pf((const char*)(in+t*32)+lane*128);
    uint4 w1=__ldcs(ptr), w2=__ldcs(ptr+32);
    for(;t<t1;++t){
      if(PFT) pf((const char*)ptr+PFT*512);
      uint32_t x=pack16(w1);
      #pragma unroll
      for(int i=0;i<WORK;++i){ x = x*0x9E3779B1u ^ (x>>15); }
      acc+=x;
      w1=w2; w2=__ldcs(ptr+64); ptr+=32;
    }
  }
  for(int o=16;o;o>>=1) acc+=__shfl_xor_sync(FULL,acc,o);
  if(lane==0) atomicAdd(out,(unsigned long long)acc);
}
// mode 1: grid-stride tiles (adjacent warps read adjacent tiles), depth-D register pipeline
template<int D, int WORK>
__global__ void __launch_bounds__(1024,1) k_stride(const uint4* __restrict__ in, uint32_t n_tiles, unsigned long long* out){
  uint32_t lane=threadIdx.x&31, warp=blockIdx.x*32+(threadIdx.x>>5), nwarps=gridDim.x*32; uint32_t acc=0;
  uint4 w[D];
  uint32_t t=warp;
  #pragma unroll
  for(int d=0;d<D;++d){ uint32_t tt=min(t+d*nwarps,n_tiles-1); w[d]=__ldcs(in+tt*32+lane);} 
  for(;t<n_tiles;t+=nwarps){
    uint32_t x=pack16(w[0]);
    #pragma unroll
    for(int i=0;i<WORK;++i){ x = x*0x9E3779B1u ^ (x>>15); }
    acc+=x;
    #pragma unroll
    for(int d=0;d+1<D;++d) w[d]=w[d+1];
    uint32_t tt=min(t+D*nwarps,n_tiles-1); w[D-1]=__ldcs(in+tt*32+lane);
  }
  for(int o=16;o;o>>=1) acc+=__shfl_xor_sync(FULL,acc,o);
  if(lane==0) atomicAdd(out,(unsigned long long)acc);
}
// mode 2: CTA-contiguous: each CTA takes a big contiguous region; within it warps stride by tile (32 warps adjacent)
template<int D, int WORK, int PFT>
__global__ void __launch_bounds__(1024,1) k_cta(const uint4* __restrict__ in, uint32_t n_tiles, uint32_t tiles_per_cta_chunk, unsigned long long* out){
  uint32_t lane=threadIdx.x&31, wid=threadIdx.x>>5; uint32_t acc=0;
  uint32_t n_chunks=(n_tiles+tiles_per_cta_chunk-1)/tiles_per_cta_chunk;
  for(uint32_t ch=blockIdx.x; ch<n_chunks; ch+=gridDim.x){
    uint32_t t0=ch*tiles_per_cta_chunk, t1=min(t0+tiles_per_cta_chunk,n_tiles);
    uint4 w[D]; uint32_t t=t0+wid;
    #pragma unroll
    for(int d=0;d<D;++d){ uint32_t tt=min(t+d*32,n_tiles-1); w[d]=__ldcs(in+tt*32+lane);} 
    for(;t<t1;t+=32){
      if(PFT){ uint32_t tp=min(t+PFT*32,n_tiles-1); pf(in+tp*32+lane);} 
      uint32_t x=pack16(w[0]);
      #pragma unroll
      for(int i=0;i<WORK;++i){ x = x*0x9E3779B1u ^ (x>>15); }
      acc+=x;
      #pragma unroll
      for(int d=0;d+1<D;++d) w[d]=w[d+1];
      uint32_t tt=min(t+D*32,n_tiles-1); w[D-1]=__ldcs(in+tt*32+lane);
    }
  }
  for(int o=16;o;o>>=1) acc+=__shfl_xor_sync(FULL,acc,o);
  if(lane==0) atomicAdd(out,(unsigned long long)acc);
}
template<class F> float timeit(F f,int it=5){cudaEvent_t a,b;cudaEventCreate(&a);cudaEventCreate(&b);f();f();cudaDeviceSynchronize();float best=1e9;for(int i=0;i<it;++i){cudaEventRecord(a);f();cudaEventRecord(b);cudaEventSynchronize(b);float ms;cudaEventElapsedTime(&ms,a,b);if(ms<best)best=ms;}return best;}
int main(){
  size_t n=(size_t)2<<30; uint4* d; cudaMalloc(&d,n+(1<<20)); cudaMemset(d,0x41,n+(1<<20)); unsigned long long* out; cudaMalloc(&out,8);
  uint32_t n_tiles=n/512; int sm=148;
  #define R(name,...) { float ms=timeit([&]{__VA_ARGS__;}); cudaError_t e=cudaGetLastError(); printf("%-44s %.3f ms %.0f GB/s %s\n",name,ms,n/ms/1e6,e?cudaGetErrorString(e):""); }
  R("span T=64 PF=8 work=0", (k_span<8,0><<<sm,1024>>>(d,n_tiles,64,out)));
  R("span T=64 PF=0 work=0", (k_span<0,0><<<sm,1024>>>(d,n_tiles,64,out)));
  R("span T=64 PF=16 work=0", (k_span<16,0><<<sm,1024>>>(d,n_tiles,64,out)));
  R("span T=64 PF=32 work=0", (k_span<32,0><<<sm,1024>>>(d,n_tiles,64,out)));
  R("span T=16 PF=8 work=0", (k_span<8,0><<<sm,1024>>>(d,n_tiles,16,out)));
  R("span T=64 PF=8 work=20", (k_span<8,20><<<sm,1024>>>(d,n_tiles,64,out)));
  R("span T=64 PF=8 work=40", (k_span<8,40><<<sm,1024>>>(d,n_tiles,64,out)));
  R("span T=64 PF=16 work=40", (k_span<16,40><<<sm,1024>>>(d,n_tiles,64,out)));
  R("stride D=2 work=0", (k_stride<2,0><<<sm,1024>>>(d,n_tiles,out)));
  R("stride D=4 work=0", (k_stride<4,0><<<sm,1024>>>(d,n_tiles,out)));
  R("stride D=2 work=40", (k_stride<2,40><<<sm,1024>>>(d,n_tiles,out)));
  R("stride D=4 work=40", (k_stride<4,40><<<sm,1024>>>(d,n_tiles,out)));
  R("cta 2048t D=2 PF=0 work=0", (k_cta<2,0,0><<<sm,1024>>>(d,n_tiles,2048,out)));
  R("cta 2048t D=2 PF=8 work=0", (k_cta<2,0,8><<<sm,1024>>>(d,n_tiles,2048,out)));
  R("cta 2048t D=4 PF=0 work=0", (k_cta<4,0,0><<<sm,1024>>>(d,n_tiles,2048,out)));
  R("cta 2048t D=2 PF=8 work=40", (k_cta<2,40,8><<<sm,1024>>>(d,n_tiles,2048,out)));
  R("cta 2048t D=2 PF=16 work=40", (k_cta<2,40,16><<<sm,1024>>>(d,n_tiles,2048,out)));
  R("cta 2048t D=4 PF=0 work=40", (k_cta<4,40,0><<<sm,1024>>>(d,n_tiles,2048,out)));
  R("cta 256t D=2 PF=8 work=40", (k_cta<2,40,8><<<sm,1024>>>(d,n_tiles,256,out)));
  return 0;
}
