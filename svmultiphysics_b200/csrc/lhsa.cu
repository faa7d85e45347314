// lhsa.cu — sparse-matrix structure from element connectivity, on the device.
//
// Replaces lhsa_ns::lhsa (Code/Source/solver/lhsa.cpp:126-381), whose per-pair insertion sort
// (add_col, lhsa.cpp:13-54) is serial and O(nnz * row length).  Here every (row,col) pair of every
// element becomes one 64-bit key row<<32|col; a radix sort + unique gives the column lists already
// ascending per row, exactly the order the reference builds.  The sort/unique primitives come from
// Thrust/CUB (setup code, run once per mesh, not part of the per-Newton hot path).
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/sort.h>
#include <thrust/unique.h>
#include "svb200_internal.h"

namespace svb {

__global__ void lhsa_pairs_kernel(int eNoN, long long nEl, const int* __restrict__ IEN, unsigned long long* __restrict__ keys)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = nEl * eNoN * eNoN;
  if (t >= total) return;
  const long long e = t / (eNoN * eNoN);
  const int ab = (int)(t % (eNoN * eNoN));
  const unsigned long long r = (unsigned)IEN[e * eNoN + ab / eNoN];
  const unsigned long long c = (unsigned)IEN[e * eNoN + ab % eNoN];
  keys[t] = (r << 32) | c;
}

__global__ void lhsa_split_kernel(long long nnz, int nNo, const unsigned long long* __restrict__ keys, int* __restrict__ rowPtr,
                                  int* __restrict__ colPtr)
{
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const int r = (int)(keys[k] >> 32);
  colPtr[k] = (int)(keys[k] & 0xffffffffu);
  const int rprev = (k == 0) ? -1 : (int)(keys[k - 1] >> 32);
  for (int q = rprev + 1; q <= r; q++) rowPtr[q] = (int)k;      // rows without entries get an empty range
  if (k == nnz - 1)
    for (int q = r + 1; q <= nNo; q++) rowPtr[q] = (int)nnz;
}

}  // namespace svb

using namespace svb;

extern "C" {

int svb200_lhsa_begin(svb200_ctx* ctx, int32_t nNo)
{
  if (!ctx) { set_error("svb200: null context"); return SVB200_ERR_INVALID; }
  SVB_CUDA(cudaSetDevice(ctx->device));
  SVB_REQUIRE(nNo >= 0, "svb200_lhsa_begin: negative node count");
  ctx->lhsa_nNo = nNo;
  ctx->lhsa_n = 0;
  ctx->lhsa_rowPtr.clear();
  ctx->lhsa_colPtr.clear();
  return SVB200_OK;
}

int svb200_lhsa_add_mesh(svb200_ctx* ctx, int32_t eNoN, int32_t nEl, const int32_t* IEN)
{
  if (!ctx) { set_error("svb200: null context"); return SVB200_ERR_INVALID; }
  SVB_CUDA(cudaSetDevice(ctx->device));
  SVB_REQUIRE(eNoN >= 1 && nEl >= 0 && (IEN || nEl == 0), "svb200_lhsa_add_mesh: bad arguments");
  const size_t add = (size_t)nEl * eNoN * eNoN;
  if (add == 0) return SVB200_OK;
  for (size_t k = 0; k < (size_t)nEl * eNoN; k++)
    SVB_REQUIRE(IEN[k] >= 0 && IEN[k] < ctx->lhsa_nNo, "svb200_lhsa_add_mesh: IEN entry out of range");
  if (ctx->lhsa_n + add > ctx->lhsa_cap) {
    unsigned long long* nk = nullptr;
    const size_t cap = (ctx->lhsa_n + add);
    SVB_CUDA(cudaMalloc(&nk, sizeof(unsigned long long) * cap));
    if (ctx->lhsa_n) SVB_CUDA(cudaMemcpyAsync(nk, ctx->d_lhsa_keys, sizeof(unsigned long long) * ctx->lhsa_n, cudaMemcpyDeviceToDevice, ctx->stream));
    SVB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_lhsa_keys);
    ctx->d_lhsa_keys = nk;
    ctx->lhsa_cap = cap;
  }
  int* d_ien = nullptr;
  SVB_CUDA(cudaMalloc(&d_ien, sizeof(int) * (size_t)nEl * eNoN));
  SVB_CUDA(cudaMemcpyAsync(d_ien, IEN, sizeof(int) * (size_t)nEl * eNoN, cudaMemcpyHostToDevice, ctx->stream));
  lhsa_pairs_kernel<<<(unsigned)((add + 255) / 256), 256, 0, ctx->stream>>>(eNoN, nEl, d_ien, ctx->d_lhsa_keys + ctx->lhsa_n);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_ien);
  ctx->lhsa_n += add;
  return SVB200_OK;
}

int svb200_lhsa_finish(svb200_ctx* ctx, int32_t* nnz_out)
{
  if (!ctx) { set_error("svb200: null context"); return SVB200_ERR_INVALID; }
  SVB_CUDA(cudaSetDevice(ctx->device));
  SVB_REQUIRE(nnz_out, "svb200_lhsa_finish: null output");
  const int nNo = ctx->lhsa_nNo;
  long long nnz = 0;
  ctx->lhsa_rowPtr.assign((size_t)nNo + 1, 0);
  if (ctx->lhsa_n > 0) {
    thrust::device_ptr<unsigned long long> k(ctx->d_lhsa_keys);
    try {
      thrust::sort(thrust::cuda::par.on(ctx->stream), k, k + ctx->lhsa_n);
      nnz = thrust::unique(thrust::cuda::par.on(ctx->stream), k, k + ctx->lhsa_n) - k;
    } catch (const std::exception& ex) {
      set_error(std::string("svb200_lhsa_finish: ") + ex.what());
      return SVB200_ERR_CUDA;
    }
    ctx->launches += 2;
    SVB_REQUIRE(nnz < (1ll << 31), "svb200_lhsa_finish: nnz exceeds int32");
    int* d_row = nullptr; int* d_col = nullptr;
    SVB_CUDA(cudaMalloc(&d_row, sizeof(int) * ((size_t)nNo + 1)));
    SVB_CUDA(cudaMalloc(&d_col, sizeof(int) * (size_t)nnz));
    lhsa_split_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, ctx->stream>>>(nnz, nNo, ctx->d_lhsa_keys, d_row, d_col);
    ctx->launches++;
    SVB_CUDA(cudaGetLastError());
    ctx->lhsa_colPtr.resize((size_t)nnz);
    SVB_CUDA(cudaMemcpyAsync(ctx->lhsa_rowPtr.data(), d_row, sizeof(int) * ((size_t)nNo + 1), cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(cudaMemcpyAsync(ctx->lhsa_colPtr.data(), d_col, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_row); cudaFree(d_col);
  }
  cudaFree(ctx->d_lhsa_keys);
  ctx->d_lhsa_keys = nullptr;
  ctx->lhsa_cap = ctx->lhsa_n = 0;
  *nnz_out = (int32_t)nnz;
  return SVB200_OK;
}

int svb200_lhsa_get(svb200_ctx* ctx, int32_t* rowPtr, int32_t* colPtr)
{
  if (!ctx) { set_error("svb200: null context"); return SVB200_ERR_INVALID; }
  SVB_REQUIRE(rowPtr && (colPtr || ctx->lhsa_colPtr.empty()), "svb200_lhsa_get: null output");
  SVB_REQUIRE(!ctx->lhsa_rowPtr.empty(), "svb200_lhsa_get: call svb200_lhsa_finish first");
  std::copy(ctx->lhsa_rowPtr.begin(), ctx->lhsa_rowPtr.end(), rowPtr);
  std::copy(ctx->lhsa_colPtr.begin(), ctx->lhsa_colPtr.end(), colPtr);
  return SVB200_OK;
}

}  // extern "C"
