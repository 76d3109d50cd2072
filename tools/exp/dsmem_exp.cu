// How fast are random 4-byte gathers from a filter spread over the shared memory of a
// thread-block cluster (DSMEM), compared with local shared memory?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define FULL 0xffffffffu

template<int CS, int WORK>
__global__ void __launch_bounds__(1024,1) k_gather(uint32_t words_per_cta, uint32_t iters, unsigned long long* out){
  extern __shared__ uint32_t sm[];
  cg::cluster_group cl = cg::this_cluster();
  for(uint32_t i=threadIdx.x;i<words_per_cta;i+=blockDim.x) sm[i]=i*2654435761u+blockIdx.x;
  cl.sync();
  uint32_t base32 = (uint32_t)__cvta_generic_to_shared(sm);
  uint32_t rank = cl.block_rank();
  // shared::cluster address of rank r: mapa
  uint32_t x = threadIdx.x*747796405u + blockIdx.x*2891336453u + 12345u, acc=0;
  const uint32_t total = words_per_cta*CS;
  for(uint32_t it=0; it<iters; ++it){
    x = x*1664525u + 1013904223u;
    uint32_t w = __umulhi(x, total);
    uint32_t r = w / words_per_cta, off = w - r*words_per_cta;
    uint32_t v;
    if (CS==1) { v = sm[off]; }
    else {
      uint32_t addr;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(base32 + off*4), "r"(r));
      asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    }
    uint32_t y=v;
    #pragma unroll
    for(int i=0;i<WORK;++i){ y = y*0x9E3779B1u + (y>>15); }
    acc+=y;
  }
  cl.sync();
  for(int o=16;o;o>>=1) acc+=__shfl_xor_sync(FULL,acc,o);
  if((threadIdx.x&31)==0) atomicAdd(out,(unsigned long long)acc);
  (void)rank;
}
template<int CS,int WORK> float run(uint32_t words, uint32_t iters, unsigned long long* out){
  cudaLaunchConfig_t cfg={}; cfg.gridDim=dim3(148/CS*CS); cfg.blockDim=dim3(1024); cfg.dynamicSmemBytes=words*4;
  cudaLaunchAttribute at[1]; at[0].id=cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x=CS; at[0].val.clusterDim.y=1; at[0].val.clusterDim.z=1;
  cfg.attrs=at; cfg.numAttrs=1;
  cudaFuncSetAttribute(k_gather<CS,WORK>, cudaFuncAttributeMaxDynamicSharedMemorySize, words*4);
  if(CS>8) cudaFuncSetAttribute(k_gather<CS,WORK>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaLaunchKernelEx(&cfg, k_gather<CS,WORK>, words, iters, out); cudaDeviceSynchronize();
  float best=1e9; for(int i=0;i<3;++i){ cudaEventRecord(a); cudaLaunchKernelEx(&cfg, k_gather<CS,WORK>, words, iters, out); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best) best=ms; }
  cudaError_t e=cudaGetLastError();
  double lookups=(double)(148/CS*CS)*1024*iters;
  printf("cluster=%d work=%d: %.3f ms  %.2f G lookups/s  (%.3f lookups/clk/SM @1.9GHz) %s\n",CS,WORK,best,lookups/best/1e6,lookups/best/1e6/148/1.9, e?cudaGetErrorString(e):"");
  return best;
}
int main(){ unsigned long long* out; cudaMalloc(&out,8); uint32_t words=50000, iters=4096;
  run<1,0>(words,iters,out); run<2,0>(words,iters,out); run<4,0>(words,iters,out); run<8,0>(words,iters,out);
  run<1,12>(words,iters,out); run<2,12>(words,iters,out); run<4,12>(words,iters,out); run<8,12>(words,iters,out);
  run<1,24>(words,iters,out); run<2,24>(words,iters,out); run<4,24>(words,iters,out);
  return 0; }
