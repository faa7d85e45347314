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

// CSR blocks between the caller's row order and the internal (FSILS) row order: caller row a is internal row map[a], the
// columns of a row keep the caller's order.  One warp per caller row of the chunk [a0,a1); `staged` holds the chunk's caller
// entries starting at rowPtr_in[a0].
__global__ void __launch_bounds__(256)
permute_row_blocks_kernel(int a0, int a1, int d2, const int* __restrict__ rowPtr_in, const int* __restrict__ rowPtr,
                          const int* __restrict__ map, double* __restrict__ internal, double* __restrict__ staged, int to_caller)
{
  const int a = a0 + (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (a >= a1) return;
  const long long len = (long long)(rowPtr_in[a + 1] - rowPtr_in[a]) * d2;
  double* h = staged + (long long)(rowPtr_in[a] - rowPtr_in[a0]) * d2;
  double* d = internal + (long long)rowPtr[map ? map[a] : a] * d2;
  if (to_caller) for (long long k = lane; k < len; k += 32) h[k] = d[k];
  else for (long long k = lane; k < len; k += 32) d[k] = h[k];
}

// dst[off[k] ...) = the d2-blocks of internal row rows[k] (one warp per listed row); rowPtr == null: a nodal (d2, nNo) array,
// one block per row.
__global__ void __launch_bounds__(256)
gather_row_blocks_kernel(int n, int d2, const int* __restrict__ rows, const int* __restrict__ rowPtr,
                         const long long* __restrict__ off, const double* __restrict__ src, double* __restrict__ dst)
{
  const int k = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= n) return;
  const int r = rows[k];
  const long long len = rowPtr ? (long long)(rowPtr[r + 1] - rowPtr[r]) * d2 : d2;
  const double* s = src + (long long)(rowPtr ? rowPtr[r] : r) * d2;
  double* d = dst + off[k] * d2;
  for (long long q = lane; q < len; q += 32) d[q] = s[q];
}

int launch_permute_row_blocks(svb200_ctx* ctx, int a0, int a1, int d2, const int* d_rowPtr_in, double* internal, double* staged,
                              bool to_caller)
{
  if (a1 <= a0) return SVB200_OK;
  const long long threads = (long long)(a1 - a0) * 32;
  permute_row_blocks_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(
      a0, a1, d2, d_rowPtr_in, ctx->d_rowPtr, ctx->has_map ? ctx->d_map : nullptr, internal, staged, to_caller ? 1 : 0);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

int launch_gather_row_blocks(svb200_ctx* ctx, int n, int d2, const int* d_rows, bool csr, const long long* d_off, const double* src,
                             double* dst)
{
  if (n <= 0) return SVB200_OK;
  const long long threads = (long long)n * 32;
  gather_row_blocks_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(n, d2, d_rows, csr ? ctx->d_rowPtr : nullptr, d_off, src,
                                                                                       dst);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
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

int launch_permute_cols(svb200_ctx* ctx, int rows, int n, const int* d_map, const double* src, double* dst, bool inverse,
                        cudaStream_t stream)
{
  const long long total = (long long)rows * n;
  if (total == 0) return SVB200_OK;
  const int threads = 256;
  permute_cols_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, stream ? stream : ctx->stream>>>(rows, n, d_map, src, dst,
                                                                                                                inverse ? 1 : 0);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

}  // namespace svb
