// microbench.cu — design-time measurements on B200 (not part of the product library):
//   1. FP64 FMA peak (independent DFMA chains)
//   2. REDG.F64 throughput for the scatter patterns the assembly kernel could use
//   3. plain vectorised read-modify-write for the same patterns (coloured scatter)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__global__ void fma_peak(double* out, int iters, double a, double b) {
  double x0=threadIdx.x, x1=x0+1, x2=x0+2, x3=x0+3, x4=x0+4, x5=x0+5, x6=x0+6, x7=x0+7;
  for (int i=0;i<iters;i++){
    x0=fma(x0,a,b); x1=fma(x1,a,b); x2=fma(x2,a,b); x3=fma(x3,a,b);
    x4=fma(x4,a,b); x5=fma(x5,a,b); x6=fma(x6,a,b); x7=fma(x7,a,b);
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=x0+x1+x2+x3+x4+x5+x6+x7;
}

__device__ __forceinline__ uint32_t hash32(uint32_t x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

// mode 0: sequential-local blocks (element e -> blocks around 2.5*e, emulating mesh locality)
// mode 1: random blocks
__device__ __forceinline__ uint32_t block_of(uint32_t e, int k, uint32_t NB, int mode){
  if (mode==0) { uint32_t base = (uint32_t)(((uint64_t)e*5)/2); return (base + hash32(e*16u+k)%64u) % NB; }
  return hash32(e*16u+k) % NB;
}

// Pattern A: half-warp per element; for each of the 16 blocks the 16 lanes add 16 contiguous doubles (one 128B line).
__global__ void red_coalesced(double* V, uint32_t nEl, uint32_t NB, int mode){
  uint32_t hw = (blockIdx.x*blockDim.x+threadIdx.x)>>4, l = threadIdx.x&15;
  uint32_t nhw = (gridDim.x*blockDim.x)>>4;
  for (uint32_t e=hw; e<nEl; e+=nhw){
    #pragma unroll 4
    for (int k=0;k<16;k++){
      uint32_t b = block_of(e,k,NB,mode);
      atomicAdd(V + (size_t)b*16 + l, 1.0);
    }
  }
}
// Pattern B: thread per element; each thread adds 16 consecutive doubles per block (lanes hit different lines).
__global__ void red_thread(double* V, uint32_t nEl, uint32_t NB, int mode){
  uint32_t t = blockIdx.x*blockDim.x+threadIdx.x, nt = gridDim.x*blockDim.x;
  for (uint32_t e=t; e<nEl; e+=nt){
    for (int k=0;k<16;k++){
      uint32_t b = block_of(e,k,NB,mode);
      #pragma unroll
      for (int j=0;j<16;j++) atomicAdd(V + (size_t)b*16 + j, 1.0);
    }
  }
}
// Pattern C: 8 lanes per block, each lane does a 16B read-modify-write (non-atomic; coloured scatter).
__global__ void rmw_vec(double* V, uint32_t nEl, uint32_t NB, int mode){
  uint32_t g8 = (blockIdx.x*blockDim.x+threadIdx.x)>>3, l = threadIdx.x&7;
  uint32_t ng = (gridDim.x*blockDim.x)>>3;
  for (uint32_t e=g8; e<nEl; e+=ng){
    #pragma unroll 4
    for (int k=0;k<16;k++){
      uint32_t b = block_of(e,k,NB,mode);
      double2* p = reinterpret_cast<double2*>(V + (size_t)b*16) + l;
      double2 v = *p; v.x += 1.0; v.y += 1.0; *p = v;
    }
  }
}
// Pattern D: plain streaming store of blocks (owner-computes output, write-once).
__global__ void store_vec(double* V, uint32_t NB){
  size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x, n=(size_t)gridDim.x*blockDim.x;
  double2* p = reinterpret_cast<double2*>(V);
  for (size_t k=i; k<(size_t)NB*8; k+=n) p[k] = make_double2(1.0,2.0);
}
// Pattern E: 4 sector-RED: 4 lanes per block? (each lane 4 consecutive doubles = 1 sector) -> 4 RED instr per lane
__global__ void red_sector(double* V, uint32_t nEl, uint32_t NB, int mode){
  uint32_t q = (blockIdx.x*blockDim.x+threadIdx.x)>>2, l = threadIdx.x&3;
  uint32_t nq = (gridDim.x*blockDim.x)>>2;
  for (uint32_t e=q; e<nEl; e+=nq){
    for (int k=0;k<16;k++){
      uint32_t b = block_of(e,k,NB,mode);
      double* p = V + (size_t)b*16 + l*4;
      atomicAdd(p,1.0); atomicAdd(p+1,1.0); atomicAdd(p+2,1.0); atomicAdd(p+3,1.0);
    }
  }
}

template<class F> float timeit(F f, int reps){
  cudaEvent_t a,b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a)); for(int i=0;i<reps;i++) f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms,a,b)); return ms/reps;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  printf("device %s SMs %d\n", p.name, p.multiProcessorCount);
  int nsm = p.multiProcessorCount;
  { double* out; CK(cudaMalloc(&out, sizeof(double)*nsm*8*256));
    int iters=20000;
    for (int bpsm : {2,4,8}) {
      float ms = timeit([&]{ fma_peak<<<nsm*bpsm,256>>>(out,iters,1.0000001,1e-9); },3);
      double fl = 2.0*8*iters*(double)nsm*bpsm*256;
      printf("fp64 fma peak: blocks/SM %d  %.2f TFLOP/s (%.3f ms)\n", bpsm, fl/ms*1e-9, ms);
    }
    cudaFree(out);
  }
  uint32_t nEl = 4000000;
  for (uint32_t NB : {10000000u, 500000u}) {   // 1.28 GB (>> L2) and 64 MB (L2 resident)
    double* V; CK(cudaMalloc(&V,(size_t)NB*128)); CK(cudaMemset(V,0,(size_t)NB*128));
    printf("--- NB=%u blocks (%.0f MB), nEl=%u (x16 blocks x16 doubles)\n", NB, NB*128.0/1e6, nEl);
    for (int mode=0; mode<2; mode++){
      float ms;
      ms = timeit([&]{ red_coalesced<<<nsm*8,256>>>(V,nEl,NB,mode); },3);
      printf("mode %d red_coalesced : %.3f ms  %.2f Gel/s  %.1f G red/s\n", mode, ms, nEl/ms*1e-6, nEl*256.0/ms*1e-6);
      ms = timeit([&]{ red_thread<<<nsm*8,256>>>(V,nEl,NB,mode); },3);
      printf("mode %d red_thread    : %.3f ms  %.2f Gel/s  %.1f G red/s\n", mode, ms, nEl/ms*1e-6, nEl*256.0/ms*1e-6);
      ms = timeit([&]{ red_sector<<<nsm*8,256>>>(V,nEl,NB,mode); },3);
      printf("mode %d red_sector    : %.3f ms  %.2f Gel/s  %.1f G red/s\n", mode, ms, nEl/ms*1e-6, nEl*256.0/ms*1e-6);
      ms = timeit([&]{ rmw_vec<<<nsm*8,256>>>(V,nEl,NB,mode); },3);
      printf("mode %d rmw_vec       : %.3f ms  %.2f Gel/s  %.1f GB/s rd+wr\n", mode, ms, nEl/ms*1e-6, nEl*4096.0/ms*1e-6);
    }
    float ms = timeit([&]{ store_vec<<<nsm*8,256>>>(V,NB); },3);
    printf("store_vec: %.3f ms %.1f GB/s\n", ms, NB*128.0/ms*1e-6);
    ms = timeit([&]{ CK(cudaMemsetAsync(V,0,(size_t)NB*128)); },3);
    printf("memset: %.3f ms %.1f GB/s\n", ms, NB*128.0/ms*1e-6);
    cudaFree(V);
  }
  return 0;
}
