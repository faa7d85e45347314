// fsils_kernels.cu — device building blocks of the FSILS Krylov solvers.
//
//   bsr_spmv           spar_mul::fsils_spar_mul_vv   Code/Source/linear_solver/spar_mul.cpp:164-231
//   multi_dot          dot::fsils_nc_dot_v           linear_solver/dot.cpp:107-146  (i+2 dots in ONE pass)
//   cgs_update         the Gram-Schmidt sweep + scale of gmres_v, linear_solver/gmres.cpp:541-549
//   lincomb            X += sum_j y_j u_j            linear_solver/gmres.cpp:590-592
//   axpby & friends    omp_la::omp_sum_v / omp_mul_v linear_solver/omp_la.cpp:21-122
//   precond_*          precond::precond_diag         linear_solver/precond.cpp:95-242
//
// All of them are HBM-bandwidth bound; vectors are (dof,nNo) node-major, the matrix is block-CSR
// with dof*dof contiguous doubles per block.  Reductions are two-stage and run in a fixed order, so
// results are bitwise reproducible from run to run.
#include <algorithm>
#include <cstdlib>
#include "svb200_internal.h"
#include "fsils_kernels.h"

namespace svb {

// ----------------------------------------------------------------------------------------------
// SpMV, dof = 4: 8 lanes per row, each lane owns two adjacent entries (one double2 = 16 B) of every
// 4x4 block of that row, so that a lane group reads one whole 128-byte block per load instruction.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bsr_spmv4_kernel(int row0, int nNo, const int* __restrict__ rowPtr, const int* __restrict__ colPtr,
                 const double* __restrict__ Val, const double* __restrict__ U, double* __restrict__ KU)
{
  // rows [row0, nNo)
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = row0 + (t >> 3);
  const int l = t & 7;
  const int i = l >> 1, j0 = (l & 1) << 1;
  double acc = 0.0;
  if (row < nNo) {
    const int k0 = rowPtr[row], k1 = rowPtr[row + 1];
    const double2* V2 = reinterpret_cast<const double2*>(Val) + l;
    int k = k0;
    for (; k + 4 <= k1; k += 4) {
      const int c0 = __ldg(colPtr + k), c1 = __ldg(colPtr + k + 1), c2 = __ldg(colPtr + k + 2), c3 = __ldg(colPtr + k + 3);
      const double2 v0 = __ldcs(V2 + 8 * (size_t)k), v1 = __ldcs(V2 + 8 * (size_t)(k + 1));
      const double2 v2 = __ldcs(V2 + 8 * (size_t)(k + 2)), v3 = __ldcs(V2 + 8 * (size_t)(k + 3));
      const double2 u0 = *reinterpret_cast<const double2*>(U + 4 * (size_t)c0 + j0);
      const double2 u1 = *reinterpret_cast<const double2*>(U + 4 * (size_t)c1 + j0);
      const double2 u2 = *reinterpret_cast<const double2*>(U + 4 * (size_t)c2 + j0);
      const double2 u3 = *reinterpret_cast<const double2*>(U + 4 * (size_t)c3 + j0);
      acc += v0.x * u0.x + v0.y * u0.y;
      acc += v1.x * u1.x + v1.y * u1.y;
      acc += v2.x * u2.x + v2.y * u2.y;
      acc += v3.x * u3.x + v3.y * u3.y;
    }
    for (; k < k1; k++) {
      const int c0 = __ldg(colPtr + k);
      const double2 v0 = __ldcs(V2 + 8 * (size_t)k);
      const double2 u0 = *reinterpret_cast<const double2*>(U + 4 * (size_t)c0 + j0);
      acc += v0.x * u0.x + v0.y * u0.y;
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if (row < nNo && (l & 1) == 0) KU[4 * (size_t)row + i] = acc;
}

// Generic dof: one thread per (row, i).
template <int DOF>
__global__ void __launch_bounds__(256)
bsr_spmv_kernel(int nNo, const int* __restrict__ rowPtr, const int* __restrict__ colPtr,
                const double* __restrict__ Val, const double* __restrict__ U, double* __restrict__ KU)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nNo * DOF) return;
  const int row = (int)(t / DOF), i = (int)(t % DOF);
  double acc = 0.0;
  for (int k = rowPtr[row]; k < rowPtr[row + 1]; k++) {
    const int c = colPtr[k];
    const double* v = Val + (size_t)k * DOF * DOF + i * DOF;
    const double* u = U + (size_t)c * DOF;
#pragma unroll
    for (int j = 0; j < DOF; j++) acc += v[j] * u[j];
  }
  KU[t] = acc;
}

// dof = 3 and dof = 1 go through the lane-group kernels of spmv_lanegroup.cu (3x3 and 1x1 blocks).  A warp-per-row form of the
// dof = 3 product (consecutive lanes on consecutive doubles, block row picked per double) was measured at 2.36 TB/s against 5.15 TB/s
// of the thread-per-(row, i) kernel above (profiles/r1n_spmv3_ab.txt) and removed.
int launch_spmv(svb200_ctx* ctx, int dof, const double* Val, const double* U, double* KU)
{
  const int nNo = ctx->nNo;
  if (nNo == 0) return SVB200_OK;
  if (dof == 4) {
    const long long threads = (long long)nNo * 8;
    bsr_spmv4_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(0, nNo, ctx->d_rowPtr, ctx->d_colPtr, Val, U, KU);
  } else {
    const long long threads = (long long)nNo * dof;
    const unsigned blocks = (unsigned)((threads + 255) / 256);
    if (dof == 3) return spmv_rc(ctx, 3, 3, Val, U, KU);
    if (dof == 1) return spmv_rc(ctx, 1, 1, Val, U, KU);
    if (dof != 2) {
      set_error("svb200: SpMV supports dof 1..4");
      return SVB200_ERR_UNSUPPORTED;
    }
    bsr_spmv_kernel<2><<<blocks, 256, 0, ctx->stream>>>(nNo, ctx->d_rowPtr, ctx->d_colPtr, Val, U, KU);
  }
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// dof = 4, rows [row0, row0 + nrows) only (the interior rows of the overlapped multi-GPU SpMV, comm.cu).
int launch_spmv4_rows(svb200_ctx* ctx, int row0, int nrows, const double* Val, const double* U, double* KU)
{
  if (nrows <= 0) return SVB200_OK;
  const long long threads = (long long)nrows * 8;
  bsr_spmv4_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(row0, row0 + nrows, ctx->d_rowPtr, ctx->d_colPtr, Val, U, KU);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// ----------------------------------------------------------------------------------------------
// multi_dot: out[j] = sum_{k<n} u_j[k] * v[k], j = 0..nvec-1, u_j = ubase + j*stride.
// A CTA keeps a tile of v in registers and streams the matching tile of every u_j past it, so v and
// each u_j are read from HBM exactly once.  Stage 1 leaves per-CTA partials in `part`
// (gridDim.x, nvec); stage 2 (one CTA) adds them in CTA order.
// ----------------------------------------------------------------------------------------------
constexpr int DOT_THREADS = 256;
constexpr int DOT_PER_THREAD = 4;   // doubles of v per thread per tile (2 x double2)
constexpr int DOT_TILE = DOT_THREADS * DOT_PER_THREAD;

__global__ void __launch_bounds__(DOT_THREADS)
multi_dot_stage1(long long n, int nvec, const double* __restrict__ ubase, long long stride,
                 const double* __restrict__ v, double* __restrict__ part)
{
  extern __shared__ double sm[];   // [DOT_THREADS/32][nvec]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* mine = sm + (size_t)warp * nvec;
  for (int j = lane; j < nvec; j += 32) mine[j] = 0.0;
  __syncwarp();
  const long long ntiles = (n + DOT_TILE - 1) / DOT_TILE;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long k0 = tile * DOT_TILE + (long long)threadIdx.x * 2;
    const long long k1 = k0 + DOT_TILE / 2;
    double2 va = make_double2(0.0, 0.0), vb = make_double2(0.0, 0.0);
    const bool fa = k0 + 1 < n, fb = k1 + 1 < n;
    if (fa) va = *reinterpret_cast<const double2*>(v + k0);
    else if (k0 < n) va.x = v[k0];
    if (fb) vb = *reinterpret_cast<const double2*>(v + k1);
    else if (k1 < n) vb.x = v[k1];
    for (int j = 0; j < nvec; j++) {
      const double* u = ubase + (long long)j * stride;
      double2 ua = make_double2(0.0, 0.0), ub = make_double2(0.0, 0.0);
      if (fa) ua = *reinterpret_cast<const double2*>(u + k0);
      else if (k0 < n) ua.x = u[k0];
      if (fb) ub = *reinterpret_cast<const double2*>(u + k1);
      else if (k1 < n) ub.x = u[k1];
      double s = va.x * ua.x + va.y * ua.y + vb.x * ub.x + vb.y * ub.y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) mine[j] += s;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nvec; j += DOT_THREADS) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < DOT_THREADS / 32; w++) s += sm[(size_t)w * nvec + j];
    part[(size_t)blockIdx.x * nvec + j] = s;
  }
}

__global__ void multi_dot_stage2(int nblocks, int nvec, const double* __restrict__ part, double* __restrict__ out)
{
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;     // one warp per output
  const int lane = threadIdx.x & 31;
  if (j >= nvec) return;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += part[(size_t)b * nvec + j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[j] = s;
}

int multi_dot(svb200_ctx* ctx, long long n, int nvec, const double* ubase, long long stride, const double* v, double* d_out,
              bool reduce_ranks)
{
  if (nvec <= 0) return SVB200_OK;
  const long long ntiles = (n + DOT_TILE - 1) / DOT_TILE;
  int blocks = (int)std::min<long long>(std::max<long long>(ntiles, 1), 148 * 4);
  const size_t need = (size_t)blocks * nvec;
  if (need > ctx->red_cap) {
    if (ctx->d_red) cudaFree(ctx->d_red);
    ctx->red_cap = need * 2;
    SVB_CUDA(cudaMalloc(&ctx->d_red, sizeof(double) * ctx->red_cap));
  }
  const size_t smem = sizeof(double) * (DOT_THREADS / 32) * nvec;
  multi_dot_stage1<<<blocks, DOT_THREADS, smem, ctx->stream>>>(n, nvec, ubase, stride, v, ctx->d_red);
  ctx->launches++;
  if (reduce_ranks && ctx->nranks > 1) {
    // the cross-rank sum as the tail of the second stage (one launch, comm.cu); same local summation order
    bool handled = false;
    int rc = dot_stage2_allreduce(ctx, blocks, nvec, ctx->d_red, d_out, &handled);
    if (rc) return rc;
    if (handled) return SVB200_OK;
  }
  multi_dot_stage2<<<(nvec * 32 + 127) / 128, 128, 0, ctx->stream>>>(blocks, nvec, ctx->d_red, d_out);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  if (reduce_ranks && ctx->nranks > 1) return allreduce_sum(ctx, d_out, nvec);
  return SVB200_OK;
}

// ----------------------------------------------------------------------------------------------
// cgs_update: classical Gram-Schmidt sweep of gmres_v fused with the normalisation:
//   v <- (v - sum_{j<=i} h_j u_j) / hn,   hn = sqrt|h_{i+1} - sum_j h_j^2|
// h (i+2 doubles) is read from device memory (it may just have been all-reduced); every thread
// recomputes hn with the reference's sequential order, thread 0 of block 0 stores it to h[i+1].
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cgs_update_kernel(long long n, int nprev, const double* __restrict__ ubase, long long stride, double* __restrict__ v,
                  double* __restrict__ h, double* __restrict__ hn_out)
{
  extern __shared__ double hs[];   // nprev + 1
  for (int j = threadIdx.x; j <= nprev; j += blockDim.x) hs[j] = h[j];
  __syncthreads();
  double hn = hs[nprev];
  for (int j = 0; j < nprev; j++) hn = __dsub_rn(hn, __dmul_rn(hs[j], hs[j]));   // no FMA: host repeats this bit for bit
  hn = sqrt(fabs(hn));
  const double inv = 1.0 / hn;
  if (blockIdx.x == 0 && threadIdx.x == 0) *hn_out = hn;
  const long long k0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (k0 >= n) return;
  if (k0 + 1 < n) {
    double2 x = *reinterpret_cast<double2*>(v + k0);
    for (int j = 0; j < nprev; j++) {
      const double2 u = *reinterpret_cast<const double2*>(ubase + (long long)j * stride + k0);
      x.x -= hs[j] * u.x;
      x.y -= hs[j] * u.y;
    }
    x.x *= inv;
    x.y *= inv;
    *reinterpret_cast<double2*>(v + k0) = x;
  } else {
    double x = v[k0];
    for (int j = 0; j < nprev; j++) x -= hs[j] * ubase[(long long)j * stride + k0];
    v[k0] = x * inv;
  }
}

int cgs_update(svb200_ctx* ctx, long long n, int nprev, const double* ubase, long long stride, double* v, double* d_h,
               double* d_hn)
{
  const long long threads = (n + 1) / 2;
  cgs_update_kernel<<<(unsigned)((threads + 255) / 256), 256, sizeof(double) * (nprev + 1), ctx->stream>>>(
      n, nprev, ubase, stride, v, d_h, d_hn);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// X += sum_{j<ny} y_j u_j with y passed by value (<= 256 coefficients).
__global__ void __launch_bounds__(256)
lincomb_kernel(long long n, int ny, const __grid_constant__ Coefs y, const double* __restrict__ ubase, long long stride,
               double* __restrict__ X)
{
  const long long k0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (k0 >= n) return;
  if (k0 + 1 < n) {
    double2 x = *reinterpret_cast<double2*>(X + k0);
    for (int j = 0; j < ny; j++) {
      const double2 u = *reinterpret_cast<const double2*>(ubase + (long long)j * stride + k0);
      x.x += y.c[j] * u.x;
      x.y += y.c[j] * u.y;
    }
    *reinterpret_cast<double2*>(X + k0) = x;
  } else {
    double x = X[k0];
    for (int j = 0; j < ny; j++) x += y.c[j] * ubase[(long long)j * stride + k0];
    X[k0] = x;
  }
}

int lincomb(svb200_ctx* ctx, long long n, int ny, const Coefs& y, const double* ubase, long long stride, double* X)
{
  if (ny <= 0) return SVB200_OK;
  const long long threads = (n + 1) / 2;
  lincomb_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(n, ny, y, ubase, stride, X);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// z = a*x + b*y (any of the pointers may alias); x or y may be null when its factor is zero.
__global__ void __launch_bounds__(256)
axpby_kernel(long long n, double a, const double* x, double b, const double* y, double* z)
{
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double r = 0.0;
  if (x) r = a * x[k];
  if (y) r += b * y[k];
  z[k] = r;
}

int axpby(svb200_ctx* ctx, long long n, double a, const double* x, double b, const double* y, double* z)
{
  if (n == 0) return SVB200_OK;
  axpby_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, a, x, b, y, z);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// z = x * y element-wise.
__global__ void __launch_bounds__(256) hadamard_kernel(long long n, const double* x, const double* y, double* z)
{
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) z[k] = x[k] * y[k];
}

int hadamard(svb200_ctx* ctx, long long n, const double* x, const double* y, double* z)
{
  if (n == 0) return SVB200_OK;
  hadamard_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, x, y, z);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// ----------------------------------------------------------------------------------------------
// precond_diag pieces.
// ----------------------------------------------------------------------------------------------
__global__ void diag_extract_kernel(int nNo, int dof, const int* __restrict__ diagPtr, const double* __restrict__ Val,
                                    double* __restrict__ W)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nNo * dof) return;
  const int r = (int)(t / dof), i = (int)(t % dof);
  W[t] = Val[(size_t)diagPtr[r] * dof * dof + i * dof + i];
}

__global__ void w_invsqrt_kernel(long long n, double* __restrict__ W)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double w = W[t];
  if (w == 0.0) w = 1.0;
  W[t] = 1.0 / sqrt(fabs(w));
}

// W(i,glob(a)) *= val(i,a), i < nd  (Dirichlet faces, precond.cpp:177-198)
__global__ void face_scale_w_kernel(int fnNo, int fdof, int nd, int dof, const int* __restrict__ glob,
                                    const double* __restrict__ val, double* __restrict__ W)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= fnNo * nd) return;
  const int a = t / nd, i = t % nd;
  W[(size_t)glob[a] * dof + i] *= val[(size_t)a * fdof + i];
}

// valM(i,a) = val(i,a) * W(i,glob(a))  (coupled faces, precond.cpp:218-241)
__global__ void face_valm_kernel(int fnNo, int fdof, int nd, int dof, const int* __restrict__ glob,
                                 const double* __restrict__ val, const double* __restrict__ W, double* __restrict__ valM)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= fnNo * fdof) return;
  const int a = t / fdof, i = t % fdof;
  valM[t] = (i < nd) ? val[t] * W[(size_t)glob[a] * dof + i] : 0.0;
}

// Val(dof*i+j, k) = (Val * Wr(i,row)) * Wc(j,col(k)) — pre_mul then pos_mul in ONE pass (same rounding
// sequence as the reference's two passes, precond.cpp:534-611 / 19-94); one warp per row, coalesced over Val.
__global__ void __launch_bounds__(256)
scale_matrix_kernel(int nNo, int dof, const int* __restrict__ rowPtr, const int* __restrict__ colPtr,
                    const double* __restrict__ Wr, const double* __restrict__ Wc, double* __restrict__ Val)
{
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= nNo) return;
  const int d2 = dof * dof;
  const long long e0 = (long long)rowPtr[row] * d2, e1 = (long long)rowPtr[row + 1] * d2;
  for (long long e = e0 + lane; e < e1; e += 32) {
    const int k = (int)(e / d2), r = (int)(e % d2);
    const int i = r / dof, j = r % dof;
    const double wi = Wr[(size_t)row * dof + i];
    const double wj = Wc[(size_t)colPtr[k] * dof + j];
    Val[e] = (Val[e] * wi) * wj;
  }
}

// dof = 4: 8 lanes per row, lane l owns the double2 (i = l/2, j0 = 2(l&1)) of every block, 4 blocks in flight;
// a lane group moves whole 128-byte blocks, Wr(i,row) stays in a register.
__global__ void __launch_bounds__(256)
scale_matrix4_kernel(int nNo, const int* __restrict__ rowPtr, const int* __restrict__ colPtr,
                     const double* __restrict__ Wr, const double* __restrict__ Wc, double* __restrict__ Val)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = t >> 3;
  if (row >= nNo) return;
  const int l = t & 7, i = l >> 1, j0 = (l & 1) << 1;
  const double wi = Wr[4 * (size_t)row + i];
  const int k0 = rowPtr[row], k1 = rowPtr[row + 1];
  double2* V2 = reinterpret_cast<double2*>(Val) + l;
  int k = k0;
  for (; k + 4 <= k1; k += 4) {
    double2 v[4], w[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int c = __ldg(colPtr + k + q);
      v[q] = __ldcs(V2 + 8 * (size_t)(k + q));
      w[q] = *reinterpret_cast<const double2*>(Wc + 4 * (size_t)c + j0);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      v[q].x = (v[q].x * wi) * w[q].x;
      v[q].y = (v[q].y * wi) * w[q].y;
      __stcs(V2 + 8 * (size_t)(k + q), v[q]);
    }
  }
  for (; k < k1; k++) {
    const int c = __ldg(colPtr + k);
    double2 v = __ldcs(V2 + 8 * (size_t)k);
    const double2 w = *reinterpret_cast<const double2*>(Wc + 4 * (size_t)c + j0);
    v.x = (v.x * wi) * w.x;
    v.y = (v.y * wi) * w.y;
    __stcs(V2 + 8 * (size_t)k, v);
  }
}

// ----------------------------------------------------------------------------------------------
// precond_rcs pieces (row-and-column max-norm equilibration, precond.cpp:251-523).
// ----------------------------------------------------------------------------------------------
__global__ void fill_kernel(long long n, double* __restrict__ W, double v)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) W[t] = v;
}

// The renormalisation of the summed Dirichlet mask (precond.cpp:289-291): Wr-0.5 -> sign -> {0,1}.
__global__ void rcs_renorm_kernel(long long n, double* __restrict__ W)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double w = W[t] - 0.5;
  w = w / fabs(w);
  W[t] = (w + fabs(w)) * 0.5;
}

// Val(ii,diag) = Wr(i)*(Val(ii,diag)-1)+1 (precond.cpp:305-349).
__global__ void rcs_diag_one_kernel(int nNo, int dof, const int* __restrict__ diagPtr, const double* __restrict__ Wr,
                                    double* __restrict__ Val)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nNo * dof) return;
  const int r = (int)(t / dof), i = (int)(t % dof);
  double* v = Val + (size_t)diagPtr[r] * dof * dof + i * dof + i;
  *v = Wr[t] * (*v - 1.0) + 1.0;
}

// Max norms along rows (warp reduction) and columns (atomic max on the bit pattern of a non-negative double, which
// is order preserving; max is exact, so the result does not depend on the order).  Wr, Wc zeroed by the caller.
__global__ void __launch_bounds__(256)
rcs_rowcol_max_kernel(int nNo, int dof, const int* __restrict__ rowPtr, const int* __restrict__ colPtr,
                      const double* __restrict__ Val, double* __restrict__ Wr, double* __restrict__ Wc)
{
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= nNo) return;
  const int d2 = dof * dof;
  const long long e0 = (long long)rowPtr[row] * d2, e1 = (long long)rowPtr[row + 1] * d2;
  double rmax[4] = {0.0, 0.0, 0.0, 0.0};    // dof <= 4
  for (long long e = e0 + lane; e < e1; e += 32) {
    const int k = (int)(e / d2), r = (int)(e % d2);
    const int i = r / dof, j = r % dof;
    const double v = fabs(Val[e]);
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (q == i) rmax[q] = fmax(rmax[q], v);
    double* wc = Wc + (size_t)colPtr[k] * dof + j;
    if (v > *wc) atomicMax(reinterpret_cast<unsigned long long*>(wc), (unsigned long long)__double_as_longlong(v));
  }
#pragma unroll
  for (int q = 0; q < 4; q++) {
    for (int o = 16; o > 0; o >>= 1) rmax[q] = fmax(rmax[q], __shfl_xor_sync(0xffffffffu, rmax[q], o));
    if (lane == 0 && q < dof) Wr[(size_t)row * dof + q] = rmax[q];
  }
}

// out[which] = max |1 - W| (bit-pattern atomic max; out zeroed by the caller).
__global__ void __launch_bounds__(256)
rcs_dev_from_one_kernel(long long n, const double* __restrict__ W, double* __restrict__ out)
{
  __shared__ double red[256];
  double m = 0.0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    m = fmax(m, fabs(1.0 - W[t]));
  red[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(red[0]));
}

// W = 1/sqrt(W); Wacc *= W (precond.cpp:496-505).
__global__ void rcs_invsqrt_acc_kernel(long long n, double* __restrict__ W, double* __restrict__ Wacc)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double w = 1.0 / sqrt(W[t]);
  W[t] = w;
  Wacc[t] = Wacc[t] * w;
}

#define SVB_LAUNCH_1D(kernel, n, ...)                                                        \
  do {                                                                                       \
    if ((n) > 0) {                                                                           \
      kernel<<<(unsigned)(((n) + 255) / 256), 256, 0, ctx->stream>>>(__VA_ARGS__);           \
      ctx->launches++;                                                                       \
      SVB_CUDA(cudaGetLastError());                                                          \
    }                                                                                        \
  } while (0)

int fill(svb200_ctx* ctx, long long n, double* W, double v)
{
  SVB_LAUNCH_1D(fill_kernel, n, n, W, v);
  return SVB200_OK;
}

int rcs_renorm(svb200_ctx* ctx, long long n, double* W)
{
  SVB_LAUNCH_1D(rcs_renorm_kernel, n, n, W);
  return SVB200_OK;
}

int rcs_diag_one(svb200_ctx* ctx, int dof, const double* Wr, double* Val)
{
  const long long n = (long long)ctx->nNo * dof;
  SVB_LAUNCH_1D(rcs_diag_one_kernel, n, ctx->nNo, dof, ctx->d_diagPtr, Wr, Val);
  return SVB200_OK;
}

int rcs_rowcol_max(svb200_ctx* ctx, int dof, const double* Val, double* Wr, double* Wc)
{
  const long long threads = (long long)ctx->nNo * 32;
  SVB_LAUNCH_1D(rcs_rowcol_max_kernel, threads, ctx->nNo, dof, ctx->d_rowPtr, ctx->d_colPtr, Val, Wr, Wc);
  return SVB200_OK;
}

// d_out[0] = max|1-Wr|, d_out[1] = max|1-Wc|
int rcs_dev_from_one(svb200_ctx* ctx, long long n, const double* Wr, const double* Wc, double* d_out)
{
  SVB_CUDA(cudaMemsetAsync(d_out, 0, 2 * sizeof(double), ctx->stream));
  if (n == 0) return SVB200_OK;
  const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, 148 * 8);
  rcs_dev_from_one_kernel<<<blocks, 256, 0, ctx->stream>>>(n, Wr, d_out);
  rcs_dev_from_one_kernel<<<blocks, 256, 0, ctx->stream>>>(n, Wc, d_out + 1);
  ctx->launches += 2;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

int rcs_invsqrt_acc(svb200_ctx* ctx, long long n, double* W, double* Wacc)
{
  SVB_LAUNCH_1D(rcs_invsqrt_acc_kernel, n, n, W, Wacc);
  return SVB200_OK;
}

int precond_extract_diag(svb200_ctx* ctx, int dof, const double* Val, double* W)
{
  const long long n = (long long)ctx->nNo * dof;
  if (n == 0) return SVB200_OK;
  diag_extract_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->nNo, dof, ctx->d_diagPtr, Val, W);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

int precond_invsqrt(svb200_ctx* ctx, int dof, double* W)
{
  const long long n = (long long)ctx->nNo * dof;
  if (n == 0) return SVB200_OK;
  w_invsqrt_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, W);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

int precond_face_scale(svb200_ctx* ctx, const Face& f, int dof, double* W)
{
  const int nd = f.dof < dof ? f.dof : dof;
  const int n = f.nNo * nd;
  if (n == 0) return SVB200_OK;
  face_scale_w_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(f.nNo, f.dof, nd, dof, f.d_glob, f.d_val, W);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

int precond_face_valm(svb200_ctx* ctx, const Face& f, int dof, const double* W)
{
  const int nd = f.dof < dof ? f.dof : dof;
  const int n = f.nNo * f.dof;
  if (n == 0 && !(f.has_cap && f.cap_n > 0)) return SVB200_OK;
  if (n > 0) face_valm_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(f.nNo, f.dof, nd, dof, f.d_glob, f.d_val, W, f.d_valM);
  ctx->launches++;
  if (f.has_cap && f.cap_n > 0) {     // cap_valM = cap_val * W (precond.cpp:229-237)
    const int nc = f.cap_n * f.dof;
    face_valm_kernel<<<(nc + 255) / 256, 256, 0, ctx->stream>>>(f.cap_n, f.dof, nd, dof, f.d_cap_glob, f.d_cap_val, W, f.d_cap_valM);
    ctx->launches++;
  }
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

int precond_scale_matrix(svb200_ctx* ctx, int dof, const double* Wr, const double* Wc, double* Val)
{
  if (ctx->nNo == 0) return SVB200_OK;
  if (dof == 4) {
    const long long threads = (long long)ctx->nNo * 8;
    scale_matrix4_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(ctx->nNo, ctx->d_rowPtr, ctx->d_colPtr, Wr, Wc, Val);
  } else {
    const long long threads = (long long)ctx->nNo * 32;
    scale_matrix_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(ctx->nNo, dof, ctx->d_rowPtr, ctx->d_colPtr, Wr, Wc, Val);
  }
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// ----------------------------------------------------------------------------------------------
// NS (Schur complement) solver pieces, linear_solver/ns_solver.cpp.
// ----------------------------------------------------------------------------------------------
// The rectangular-block products (3x3, 3x1, 1x3, 1x1: fsils_spar_mul_vv/sv/vs/ss) live in spmv_lanegroup.cu.

// ---- device-resident Schur-complement CG (cgrad::schur, linear_solver/cgrad.cpp:23-133) ----------------------------
// Scalars of the iteration live in a small device array `cg` so that no kernel waits for the host:
//   cg[0] errO (|r|^2 entering the iteration)   cg[1] err (|r|^2 after the update)   cg[2] <p, S p>   cg[3] eps
//   cg[4] done flag   cg[5] iterations executed   cg[6] errO of the last executed iteration
// The kernels that change X, R, P or the scalars return at once when `done` is set, so iterations the host enqueued
// ahead of the device's stopping test are no-ops.
// SP = L p - D (G p): one pass over the row, two accumulators so that both sums keep the reference's order
// (spar_mul_ss on L, spar_mul_vs on D, then SP = -DGP + SP; cgrad.cpp:77-84).
template <int NSD>
__global__ void __launch_bounds__(256)
schur_sp_kernel(int nNo, const int* __restrict__ rowPtr, const int* __restrict__ colPtr, const double* __restrict__ L,
                const double* __restrict__ D, const double* __restrict__ P, const double* __restrict__ GP, double* __restrict__ SP)
{
  // 4 threads per row, one 8-byte stream each: q = 0 walks L (times P), q = 1..NSD walk component q-1 of the 1 x NSD
  // blocks of D (times GP) — the interleaving of bsr_spmv_rc_kernel<3,1>, which reaches 4.7 TB/s where a thread reading a
  // whole 24-byte D block per step reaches 2.7 (profiles/r1j_ns_launches.csv).  The NSD partial sums of D are added
  // component by component instead of block by block (round-off level difference to cgrad.cpp:77-84).
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int row = (int)(t >> 2), q = (int)(t & 3);
  double acc = 0.0;
  if (row < nNo && q <= NSD) {
    const int k1 = rowPtr[row + 1];
    if (q == 0) {
      for (int k = rowPtr[row]; k < k1; k++) acc += L[k] * P[colPtr[k]];
    } else {
      const double* d = D + (q - 1);
      const double* g = GP + (q - 1);
      for (int k = rowPtr[row]; k < k1; k++) acc += d[(size_t)k * NSD] * g[(size_t)colPtr[k] * NSD];
    }
  }
  const double a1 = __shfl_down_sync(0xffffffffu, acc, 1), a2 = __shfl_down_sync(0xffffffffu, acc, 2),
               a3 = __shfl_down_sync(0xffffffffu, acc, 3);
  if (row < nNo && q == 0) {
    const double accD = (NSD == 3) ? (a1 + a2) + a3 : a1 + a2;
    SP[row] = -1.0 * accD + acc;
  }
}

// X = alpha P + X, R = -alpha SP + R with alpha = errO / <p, S p> taken from the device scalars.
__global__ void __launch_bounds__(256)
cg_xr_kernel(long long n, const double* __restrict__ cg, const double* __restrict__ P, const double* __restrict__ SP,
             double* __restrict__ X, double* __restrict__ R)
{
  if (cg[4] != 0.0) return;
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double alpha = cg[0] / cg[2];
  X[k] = alpha * P[k] + X[k];
  R[k] = -alpha * SP[k] + R[k];
}

// P = (errO/err) R + P, then P = (err/errO) P  (the two omp_sum_s / omp_mul_s calls of cgrad.cpp:95-96).
__global__ void __launch_bounds__(256)
cg_p_kernel(long long n, const double* __restrict__ cg, const double* __restrict__ R, double* __restrict__ P)
{
  if (cg[4] != 0.0) return;
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double errO = cg[0];
  double err = sqrt(cg[1]);      // err = norm(R); err = err*err as in cgrad.cpp:92-93 (cg[1] holds the raw dot product)
  err = err * err;
  double p = (errO / err) * R[k] + P[k];
  P[k] = (err / errO) * p;
}

// End of an iteration: count it, shift err -> errO, raise `done` when the NEXT iteration's test `err < eps` would fire.
__global__ void cg_advance_kernel(double* cg)
{
  if (cg[4] != 0.0) return;
  double err = sqrt(cg[1]);
  err = err * err;
  cg[6] = cg[0];
  cg[0] = err;
  cg[5] += 1.0;
  if (err < cg[3]) cg[4] = 1.0;
}

int schur_sp(svb200_ctx* ctx, int nsd, const double* L, const double* D, const double* P, const double* GP, double* SP)
{
  const int nNo = ctx->nNo;
  if (nNo == 0) return SVB200_OK;
  const unsigned blocks = (unsigned)(((long long)nNo * 4 + 255) / 256);
  if (nsd == 3) schur_sp_kernel<3><<<blocks, 256, 0, ctx->stream>>>(nNo, ctx->d_rowPtr, ctx->d_colPtr, L, D, P, GP, SP);
  else schur_sp_kernel<2><<<blocks, 256, 0, ctx->stream>>>(nNo, ctx->d_rowPtr, ctx->d_colPtr, L, D, P, GP, SP);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

int cg_step_kernels(svb200_ctx* ctx, int which, long long n, double* cg, const double* P, const double* SP, double* X, double* R,
                    double* Pw)
{
  if (n == 0) return SVB200_OK;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (which == 0) cg_xr_kernel<<<blocks, 256, 0, ctx->stream>>>(n, cg, P, SP, X, R);
  else if (which == 1) cg_p_kernel<<<blocks, 256, 0, ctx->stream>>>(n, cg, R, Pw);
  else cg_advance_kernel<<<1, 1, 0, ctx->stream>>>(cg);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// tslot[j] = slot l of row col(j) whose column is the row of j (the transposed entry), or -1.
__global__ void transpose_slot_kernel(int nNo, const int* __restrict__ rowPtr, const int* __restrict__ colPtr,
                                      int* __restrict__ tslot)
{
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3);
  const int l8 = threadIdx.x & 7;
  if (row >= nNo) return;
  for (int j = rowPtr[row] + l8; j < rowPtr[row + 1]; j += 8) {
    const int k = colPtr[j];
    int t = -1;
    for (int l = rowPtr[k]; l < rowPtr[k + 1]; l++)
      if (colPtr[l] == row) { t = l; break; }
    tslot[j] = t;
  }
}

int build_transpose_slots(svb200_ctx* ctx, int* d_tslot)
{
  if (ctx->nNo == 0) return SVB200_OK;
  const long long threads = (long long)ctx->nNo * 8;
  transpose_slot_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(ctx->nNo, ctx->d_rowPtr, ctx->d_colPtr, d_tslot);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// ns_solver::depart (ns_solver.cpp:64-132): split the (nsd+1)^2 blocks of Val into mK, mG, mD, mL and
// Gt(:, tslot(j)) = -mG(:, j).  Gt must be zeroed by the caller (entries without a transposed slot).
__global__ void __launch_bounds__(256)
depart_kernel(long long nnz, int nsd, const double* __restrict__ Val, const int* __restrict__ tslot, double* __restrict__ mK,
              double* __restrict__ mG, double* __restrict__ mD, double* __restrict__ mL, double* __restrict__ Gt,
              double* __restrict__ DL)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  const int dof = nsd + 1;
  const double* v = Val + (size_t)j * dof * dof;
  const int t = tslot[j];
  for (int a = 0; a < nsd; a++) {
    for (int b = 0; b < nsd; b++) mK[(size_t)j * nsd * nsd + a * nsd + b] = v[a * dof + b];
    const double g = v[a * dof + nsd];
    mG[(size_t)j * nsd + a] = g;
    if (t >= 0) {
      Gt[(size_t)t * nsd + a] = -g;
      if (DL) DL[(size_t)t * 4 + a] = -g;      // interleaved { Gt(0..2), L } of schur_sp4_kernel (nsd = 3)
    }
    mD[(size_t)j * nsd + a] = v[nsd * dof + a];
  }
  mL[j] = v[nsd * dof + nsd];
  if (DL) DL[(size_t)j * 4 + 3] = v[nsd * dof + nsd];
}

int ns_depart(svb200_ctx* ctx, int nsd, const double* Val, const int* d_tslot, double* mK, double* mG, double* mD, double* mL,
              double* Gt, double* DL)
{
  const long long nnz = ctx->nnz;
  if (nnz == 0) return SVB200_OK;
  SVB_CUDA(cudaMemsetAsync(Gt, 0, sizeof(double) * nnz * nsd, ctx->stream));
  if (DL) SVB_CUDA(cudaMemsetAsync(DL, 0, sizeof(double) * nnz * 4, ctx->stream));
  depart_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, ctx->stream>>>(nnz, nsd, Val, d_tslot, mK, mG, mD, mL, Gt, DL);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// Split Ri(dof,nNo) into Rm(nsd,nNo), Rc(nNo) and merge back.
__global__ void split_kernel(long long nNo, int dof, const double* __restrict__ Ri, double* __restrict__ Rm, double* __restrict__ Rc)
{
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nNo) return;
  for (int i = 0; i < dof - 1; i++) Rm[a * (dof - 1) + i] = Ri[a * dof + i];
  Rc[a] = Ri[a * dof + dof - 1];
}
__global__ void merge_kernel(long long nNo, int dof, const double* __restrict__ Rm, const double* __restrict__ Rc, double* __restrict__ Ri)
{
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nNo) return;
  for (int i = 0; i < dof - 1; i++) Ri[a * dof + i] = Rm[a * (dof - 1) + i];
  Ri[a * dof + dof - 1] = Rc[a];
}
int ns_split(svb200_ctx* ctx, int dof, const double* Ri, double* Rm, double* Rc)
{
  if (ctx->nNo == 0) return SVB200_OK;
  split_kernel<<<(ctx->nNo + 255) / 256, 256, 0, ctx->stream>>>(ctx->nNo, dof, Ri, Rm, Rc);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}
int ns_merge(svb200_ctx* ctx, int dof, const double* Rm, const double* Rc, double* Ri)
{
  if (ctx->nNo == 0) return SVB200_OK;
  merge_kernel<<<(ctx->nNo + 255) / 256, 256, 0, ctx->stream>>>(ctx->nNo, dof, Rm, Rc, Ri);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// ----------------------------------------------------------------------------------------------
// FP64 FMA peak (independent DFMA chains), for the roofline denominator of the assembly kernel.
// ----------------------------------------------------------------------------------------------
__global__ void fma_peak_kernel(double* out, int iters, double a, double b)
{
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int fp64_peak(svb200_ctx* ctx, double* tflops)
{
  int nsm = 0;
  SVB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device));
  const int blocks = nsm * 8, threads = 256, iters = 20000;
  double* out = nullptr;
  SVB_CUDA(cudaMalloc(&out, sizeof(double) * blocks * threads));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    SVB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    fma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(out, iters, 1.0000001, 1e-9);
    ctx->launches++;
    SVB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    SVB_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    SVB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (rep > 0 && ms < best) best = ms;
  }
  SVB_CUDA(cudaFree(out));
  *tflops = 2.0 * 8 * iters * (double)blocks * threads / (best * 1e-3) * 1e-12;
  return SVB200_OK;
}

}  // namespace svb
