// microbench_tma.cu — throughput of TMA bulk reduce-add (cp.reduce.async.bulk ... .add.f64) as a scatter
// primitive: every thread stages one 128-byte block (16 doubles) in shared memory and issues ONE bulk
// reduction to global memory, instead of 16 lane-wide REDG.F64.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__device__ __forceinline__ uint32_t hash32(uint32_t x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }
__device__ __forceinline__ uint32_t block_of(uint32_t e, int k, uint32_t NB, int mode){
  if (mode==0) { uint32_t base = (uint32_t)(((uint64_t)e*5)/2); return (base + hash32(e*16u+k)%64u) % NB; }
  return hash32(e*16u+k) % NB;
}
__device__ __forceinline__ void bulk_red_add_f64(double* gdst, const double* ssrc, uint32_t bytes){
  uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" :: "l"(gdst), "r"(s), "r"(bytes) : "memory");
}
// NBUF staging slots per thread, each 128 B.
template<int NBUF>
__global__ void __launch_bounds__(128) tma_red(double* V, uint32_t nEl, uint32_t NB, int mode){
  extern __shared__ __align__(128) double sm[];
  double* mine = sm + (size_t)threadIdx.x*16*NBUF;
  uint32_t t = blockIdx.x*blockDim.x+threadIdx.x, nt = gridDim.x*blockDim.x;
  for (uint32_t e=t; e<nEl; e+=nt){
    for (int k=0;k<16;k++){
      double* slot = mine + (k%NBUF)*16;
      if (k>=NBUF) asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(NBUF-1) : "memory");
      #pragma unroll
      for (int j=0;j<16;j+=2) *reinterpret_cast<double2*>(slot+j) = make_double2(1.0,1.0);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      uint32_t b = block_of(e,k,NB,mode);
      bulk_red_add_f64(V + (size_t)b*16, slot, 128);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}
// warp-cooperative: lane-per-element data but one lane issues 256B..? variant: each thread issues ONE op of 2 KB? not applicable.
template<class F> float timeit(F f, int reps){
  cudaEvent_t a,b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a)); for(int i=0;i<reps;i++) f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms,a,b)); return ms/reps;
}
__global__ void check(const double* V, uint32_t NB, double* out){ double s=0; for (size_t i=threadIdx.x;i<(size_t)NB*16;i+=blockDim.x) s+=V[i]; atomicAdd(out,s);} 
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int nsm=p.multiProcessorCount;
  uint32_t nEl=4000000;
  for (uint32_t NB : {10000000u, 500000u}) {
    double* V; CK(cudaMalloc(&V,(size_t)NB*128)); CK(cudaMemset(V,0,(size_t)NB*128));
    printf("--- NB=%u blocks (%.0f MB)\n", NB, NB*128.0/1e6);
    for (int mode=0;mode<2;mode++){
      for (int bps : {4,8,16}) {
        CK(cudaFuncSetAttribute(tma_red<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128*16*8*2));
        float ms = timeit([&]{ tma_red<2><<<nsm*bps,128,128*16*8*2>>>(V,nEl,NB,mode); },3);
        printf("mode %d tma_red<2> blocks/SM %2d: %.3f ms  %.2f Gel/s  %.1f G blockops/s\n", mode,bps,ms,nEl/ms*1e-6,nEl*16.0/ms*1e-6);
        CK(cudaFuncSetAttribute(tma_red<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128*16*8*4));
        ms = timeit([&]{ tma_red<4><<<nsm*bps,128,128*16*8*4>>>(V,nEl,NB,mode); },3);
        printf("mode %d tma_red<4> blocks/SM %2d: %.3f ms  %.2f Gel/s  %.1f G blockops/s\n", mode,bps,ms,nEl/ms*1e-6,nEl*16.0/ms*1e-6);
      }
    }
    // correctness: total sum must be (#launches)*nEl*256
    cudaFree(V);
  }
  { // correctness check
    uint32_t NB=1000; double* V; CK(cudaMalloc(&V,(size_t)NB*128)); CK(cudaMemset(V,0,(size_t)NB*128));
    CK(cudaFuncSetAttribute(tma_red<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128*16*8*2));
    tma_red<2><<<8,128,128*16*8*2>>>(V,5000,NB,1); CK(cudaDeviceSynchronize());
    double* out; CK(cudaMalloc(&out,8)); CK(cudaMemset(out,0,8)); check<<<1,256>>>(V,NB,out); double h; CK(cudaMemcpy(&h,out,8,cudaMemcpyDeviceToHost));
    printf("check: sum %.1f expected %.1f\n", h, 5000.0*256);
  }
  return 0;
}
