// graph_kernels.cu — one-time structure kernels: element->CSR slot map, diagonal pointers, node
// permutations between the caller's node order and the FSILS order kept on the device.
//
// The slot map replaces the per-pair binary search of lhsa_ns::do_assem
// (Code/Source/solver/lhsa.cpp:96-107) by a lookup; diagPtr is lhs.diagPtr of
// fsils_lhs_create (Code/Source/linear_solver/lhs.cpp:233-241).
#include "svb200_internal.h"

namespace svb {

// Columns of a row are NOT sorted in the internal numbering when a node permutation is active, so the
// search is linear over the (short: ~15-27 entries) row.
__global__ void slot_map_kernel(int eNoN, int nEl, const int* __restrict__ IEN, const int* __restrict__ rowPtr,
                                const int* __restrict__ colPtr, int* __restrict__ slot, int* __restrict__ err)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)nEl * eNoN * eNoN;
  if (t >= total) return;
  const int e = (int)(t / (eNoN * eNoN));
  const int ab = (int)(t % (eNoN * eNoN));
  const int a = ab / eNoN, b = ab % eNoN;
  const int row = IEN[(size_t)e * eNoN + a];
  const int col = IEN[(size_t)e * eNoN + b];
  int s = -1;
  for (int k = rowPtr[row]; k < rowPtr[row + 1]; k++) {
    if (colPtr[k] == col) { s = k; break; }
  }
  if (s < 0) atomicExch(err, 1);
  slot[t] = s;
}

__global__ void find_diag_kernel(int nNo, const int* __restrict__ rowPtr, const int* __restrict__ colPtr,
                                 int* __restrict__ diagPtr, int* __restrict__ err)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nNo) return;
  int d = -1;
  for (int k = rowPtr[r]; k < rowPtr[r + 1]; k++) {
    if (colPtr[k] == r) { d = k; break; }
  }
  if (d < 0) atomicExch(err, 1);
  diagPtr[r] = d;
}

// inverse == false: dst(:,map[a]) = src(:,a)   (caller order -> device order)
// inverse == true : dst(:,a) = src(:,map[a])   (device order -> caller order)
__global__ void permute_cols_kernel(int rows, int n, const int* __restrict__ map, const double* __restrict__ src,
                                    double* __restrict__ dst, int inverse)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * n) return;
  const int a = (int)(t / rows), i = (int)(t % rows);
  const int m = map[a];
  if (inverse)
    dst[t] = src[(size_t)m * rows + i];
  else
    dst[(size_t)m * rows + i] = src[t];
}

static int check_flag(svb200_ctx* ctx, int* d_err, const char* what)
{
  int h = 0;
  SVB_CUDA(cudaMemcpyAsync(&h, d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  SVB_CUDA(cudaFree(d_err));
  if (h) {
    set_error(std::string("svb200: ") + what);
    return SVB200_ERR_INVALID;
  }
  return SVB200_OK;
}

int launch_build_slot_map(svb200_ctx* ctx, Mesh& m)
{
  const long long total = (long long)m.nEl * m.eNoN * m.eNoN;
  if (m.d_slot) { cudaFree(m.d_slot); m.d_slot = nullptr; }
  if (total == 0) return SVB200_OK;
  SVB_CUDA(cudaMalloc(&m.d_slot, sizeof(int) * total));
  int* d_err = nullptr;
  SVB_CUDA(cudaMalloc(&d_err, sizeof(int)));
  SVB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream));
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  slot_map_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(m.eNoN, m.nEl, m.d_IEN, ctx->d_rowPtr, ctx->d_colPtr,
                                                               m.d_slot, d_err);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return check_flag(ctx, d_err, "mesh connectivity references a node pair that is not in the CSR graph");
}

int launch_find_diag(svb200_ctx* ctx)
{
  int* d_err = nullptr;
  SVB_CUDA(cudaMalloc(&d_err, sizeof(int)));
  SVB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream));
  const int threads = 256;
  find_diag_kernel<<<(ctx->nNo + threads - 1) / threads, threads, 0, ctx->stream>>>(ctx->nNo, ctx->d_rowPtr,
                                                                                   ctx->d_colPtr, ctx->d_diagPtr, d_err);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return check_flag(ctx, d_err, "CSR graph has a row without a diagonal entry");
}

int launch_permute_cols(svb200_ctx* ctx, int rows, int n, const int* d_map, const double* src, double* dst, bool inverse)
{
  const long long total = (long long)rows * n;
  if (total == 0) return SVB200_OK;
  const int threads = 256;
  permute_cols_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, ctx->stream>>>(rows, n, d_map, src, dst,
                                                                                              inverse ? 1 : 0);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

}  // namespace svb
