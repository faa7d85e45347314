// spmv_lanegroup.cu — block-CSR SpMV for every block shape other than 4x4 (which has its own kernel in fsils_kernels.cu).
//
//   fsils_spar_mul_vv   (3x3: struct / lElas / mesh equations, the momentum block mK of the NS solver)
//   fsils_spar_mul_sv   (3x1: G of the NS solver)      fsils_spar_mul_vs   (1x3: D / Gt)      fsils_spar_mul_ss   (1x1: L, heat)
//   Code/Source/linear_solver/spar_mul.cpp:19-231
//
// All of them are HBM bound: a row is ONE contiguous run of len*R*C doubles of the matrix plus len column ids, read exactly once.
// The kernel below generalises the lane mapping of bsr_spmv4_kernel: a GROUP of LPR = G*R*JL lanes owns a row,
//   g  in [0,G)   which of the G consecutive blocks the group reads per step,
//   i  in [0,R)   the block row,
//   jl in [0,JL)  which C/JL-wide part of the block row,
// so that per step the group reads G*R*C consecutive doubles (coalesced), and UN steps are issued before the first use
// (UN*G blocks of every row in flight).  The partial sums of a row are combined with shuffles in a fixed order, so the result
// is bitwise reproducible from run to run; it differs from the reference's strictly sequential sum over k by round-off only
// (tests: 1e-13).  What bounds these kernels on B200 is the L1 request pipeline, not DRAM (ncu: l1tex 70-92 % busy, DRAM 55-75 %, 90 % of the
// stalls long_scoreboard at 90 % occupancy): every warp-wide 8-byte load of a thread-per-(row, i) mapping touches ~11 different lines
// and the column gather of U costs as many requests as the matrix itself.  Measured consequences: ld.global.cs on the matrix is 20-25 %
// SLOWER than plain ld.global (a group re-reads the sectors it straddles in its next step; L1 must keep them), a serial tail loop costs
// more than predicated extra steps on 15-block rows, and more lanes per row only pay while the group still reads >= 96 contiguous bytes.
// `variant` picks (JL, G, UN, LD); the defaults are the ones measured fastest on B200 (profiles/r2q_spmv_variants.txt, r2q_spmv_ncu_table.txt),
// SVB200_SPMV_VARIANT_<R><C>=v overrides them, variant 0 is the round-1 thread-per-(row, i) kernel.
#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include "svb200_internal.h"
#include "fsils_kernels.h"

namespace svb {

template <int R, int C>
__global__ void __launch_bounds__(256)
bsr_spmv_rc_kernel(int nNo, const int* __restrict__ rowPtr, const int* __restrict__ colPtr, const double* __restrict__ K,
                   const double* __restrict__ U, double* __restrict__ KU)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nNo * R) return;
  const int row = (int)(t / R), i = (int)(t % R);
  double acc = 0.0;
  for (int k = rowPtr[row]; k < rowPtr[row + 1]; k++) {
    const int c = colPtr[k];
    const double* v = K + (size_t)k * R * C + i * C;
    const double* u = U + (size_t)c * C;
#pragma unroll
    for (int j = 0; j < C; j++) acc += v[j] * u[j];
  }
  KU[t] = acc;
}

// LD: how the matrix is loaded — 0: ld.global (allocates in L1: the sectors a group straddles are reused by its next step),
// 1: ld.global.cs (streaming / evict first).
template <int LD>
__device__ __forceinline__ double ld_mat(const double* p) { return LD == 1 ? __ldcs(p) : *p; }
template <int LD>
__device__ __forceinline__ double2 ld_mat2(const double2* p) { return LD == 1 ? __ldcs(p) : *p; }

template <int R, int C, int JL, int G, int UN, int LD>
__global__ void __launch_bounds__(256)
bsr_spmv_lg_kernel(int nNo, const int* __restrict__ rowPtr, const int* __restrict__ colPtr, const double* __restrict__ K,
                   const double* __restrict__ U, double* __restrict__ KU)
{
  constexpr int CW = C / JL;      // doubles of a block row one lane reads
  constexpr int LPB = R * JL;     // lanes per block
  constexpr int LPR = G * LPB;    // lanes per row
  constexpr int RPW = 32 / LPR;   // rows per warp (lanes >= RPW*LPR idle)
  static_assert(C % JL == 0 && LPR <= 32, "bad lane mapping");
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int rg = lane / LPR, lg = lane - rg * LPR;
  const int g = lg / LPB, e = lg - g * LPB, i = e / JL, jl = e - i * JL;
  const long long row = warp * RPW + rg;
  const bool live = rg < RPW && row < nNo;
  double acc = 0.0;
  if (live) {
    const int k0 = __ldg(rowPtr + row), k1 = __ldg(rowPtr + row + 1);
    const double* v = K + i * C + jl * CW;
    const double* u = U + jl * CW;
    // UN steps in flight; the steps past the end of the row are predicated (they re-read the row's first block and add 0),
    // so a short row costs one batch of loads instead of a serial tail
    for (int k = k0 + g; k < k1; k += UN * G) {
      int kk[UN], c[UN];
      bool ok[UN];
      double vv[UN][CW], uu[UN][CW];
#pragma unroll
      for (int q = 0; q < UN; q++) {
        ok[q] = k + q * G < k1;
        kk[q] = ok[q] ? k + q * G : k0;
        c[q] = __ldg(colPtr + kk[q]);
      }
#pragma unroll
      for (int q = 0; q < UN; q++)
#pragma unroll
        for (int j = 0; j < CW; j++) vv[q][j] = ld_mat<LD>(v + (size_t)kk[q] * (R * C) + j);
#pragma unroll
      for (int q = 0; q < UN; q++)
#pragma unroll
        for (int j = 0; j < CW; j++) uu[q][j] = __ldg(u + (size_t)c[q] * C + j);
#pragma unroll
      for (int q = 0; q < UN; q++)
#pragma unroll
        for (int j = 0; j < CW; j++) acc += ok[q] ? vv[q][j] * uu[q][j] : 0.0;
    }
  }
  if (G * JL > 1) {
    double s = 0.0;
#pragma unroll
    for (int gg = 0; gg < G; gg++)
#pragma unroll
      for (int jj = 0; jj < JL; jj++) s += __shfl_sync(0xffffffffu, acc, (rg * LPR + gg * LPB + i * JL + jj) & 31);
    acc = s;
  }
  if (live && g == 0 && jl == 0) KU[(size_t)row * R + i] = acc;
}

template <int R, int C, int JL, int G, int UN, int LD>
static void launch_lg(svb200_ctx* ctx, const double* K, const double* U, double* KU)
{
  constexpr int RPW = 32 / (G * R * JL);
  const long long warps = ((long long)ctx->nNo + RPW - 1) / RPW;
  const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
  bsr_spmv_lg_kernel<R, C, JL, G, UN, LD><<<blocks, 256, 0, ctx->stream>>>(ctx->nNo, ctx->d_rowPtr, ctx->d_colPtr, K, U, KU);
}

template <int R, int C>
static void launch_rc_old(svb200_ctx* ctx, const double* K, const double* U, double* KU)
{
  const unsigned blocks = (unsigned)(((long long)ctx->nNo * R + 255) / 256);
  bsr_spmv_rc_kernel<R, C><<<blocks, 256, 0, ctx->stream>>>(ctx->nNo, ctx->d_rowPtr, ctx->d_colPtr, K, U, KU);
}

// Number of variants per shape (for the A/B tool) and the dispatch table.
int spmv_rc_num_variants(int R, int C)
{
  if (R == 3 && C == 3) return 8;
  if (R == 3 && C == 1) return 7;
  if (R == 1 && C == 3) return 7;
  if (R == 1 && C == 1) return 7;
  return 1;
}

int spmv_rc_variant(svb200_ctx* ctx, int R, int C, int variant, const double* K, const double* U, double* KU)
{
  if (ctx->nNo == 0) return SVB200_OK;
  bool ok = true;
#define V(v, ...) case v: __VA_ARGS__(ctx, K, U, KU); break;
  if (R == 3 && C == 3) {
    switch (variant) {
      V(0, launch_rc_old<3, 3>)
      V(1, launch_lg<3, 3, 1, 1, 4, 1>)
      V(2, launch_lg<3, 3, 1, 1, 4, 0>)
      V(3, launch_lg<3, 3, 3, 1, 4, 1>)
      V(4, launch_lg<3, 3, 3, 1, 4, 0>)
      V(5, launch_lg<3, 3, 1, 2, 4, 0>)
      V(6, launch_lg<3, 3, 3, 1, 8, 0>)
      V(7, launch_lg<3, 3, 1, 1, 8, 0>)
      default: ok = false;
    }
  } else if (R == 3 && C == 1) {
    switch (variant) {
      V(0, launch_rc_old<3, 1>)
      V(1, launch_lg<3, 1, 1, 1, 4, 1>)
      V(2, launch_lg<3, 1, 1, 1, 4, 0>)
      V(3, launch_lg<3, 1, 1, 1, 8, 0>)
      V(4, launch_lg<3, 1, 1, 2, 4, 0>)
      V(5, launch_lg<3, 1, 1, 2, 4, 1>)
      V(6, launch_lg<3, 1, 1, 4, 2, 0>)
      default: ok = false;
    }
  } else if (R == 1 && C == 3) {
    switch (variant) {
      V(0, launch_rc_old<1, 3>)
      V(1, launch_lg<1, 3, 3, 1, 4, 0>)
      V(2, launch_lg<1, 3, 1, 4, 2, 0>)
      V(3, launch_lg<1, 3, 1, 4, 4, 0>)
      V(4, launch_lg<1, 3, 1, 4, 2, 1>)
      V(5, launch_lg<1, 3, 1, 2, 4, 0>)
      V(6, launch_lg<1, 3, 1, 8, 2, 0>)
      default: ok = false;
    }
  } else if (R == 1 && C == 1) {
    switch (variant) {
      V(0, launch_rc_old<1, 1>)
      V(1, launch_lg<1, 1, 1, 4, 2, 0>)
      V(2, launch_lg<1, 1, 1, 4, 4, 0>)
      V(3, launch_lg<1, 1, 1, 4, 2, 1>)
      V(4, launch_lg<1, 1, 1, 8, 2, 0>)
      V(5, launch_lg<1, 1, 1, 2, 4, 0>)
      V(6, launch_lg<1, 1, 1, 4, 4, 1>)
      default: ok = false;
    }
  } else if (variant == 0 && R == 2 && C == 2) { launch_rc_old<2, 2>(ctx, K, U, KU);
  } else if (variant == 0 && R == 2 && C == 1) { launch_rc_old<2, 1>(ctx, K, U, KU);
  } else if (variant == 0 && R == 1 && C == 2) { launch_rc_old<1, 2>(ctx, K, U, KU);
  } else ok = false;
#undef V
  if (!ok) {
    set_error("svb200: unsupported block shape / variant in spmv_rc");
    return SVB200_ERR_UNSUPPORTED;
  }
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// Default variant per shape: measured on B200 (profiles/r2q_spmv_variants.txt, r2q_spmv_ncu_table.txt); SVB200_SPMV_VARIANT_33 / _31 / _13 / _11 override.
static int default_variant(int R, int C)
{
  static int v33 = -1, v31 = -1, v13 = -1, v11 = -1;
  if (v33 < 0) {
    auto rd = [](const char* name, int dflt) { const char* s = getenv(name); return s ? atoi(s) : dflt; };
    v33 = rd("SVB200_SPMV_VARIANT_33", 5);
    v31 = rd("SVB200_SPMV_VARIANT_31", 2);
    v13 = rd("SVB200_SPMV_VARIANT_13", 1);
    v11 = rd("SVB200_SPMV_VARIANT_11", 2);
  }
  if (R == 3 && C == 3) return v33;
  if (R == 3 && C == 1) return v31;
  if (R == 1 && C == 3) return v13;
  if (R == 1 && C == 1) return v11;
  return 0;
}

int spmv_rc(svb200_ctx* ctx, int R, int C, const double* K, const double* U, double* KU)
{
  return spmv_rc_variant(ctx, R, C, default_variant(R, C), K, U, KU);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Schur-complement operator of the NS solver, nsd = 3 (cgrad::schur, linear_solver/cgrad.cpp:77-84):
//   SP = L p - Gt (G p)    with  DL(:,k) = { Gt(0..2,k), L(k) }  interleaved by ns_depart: 32 bytes = one sector per block.
// Two lanes per block (one double2 each): lane h = 0 multiplies (Gt0, Gt1) with (GP0, GP1), lane h = 1 (Gt2, L) with (GP2, p);
// G blocks per row group and step, UN steps in flight.  The D and the L sums stay separate until the end (SP = -DGP + LP as in the
// reference).  DOT: the kernel also leaves the per-CTA partial sums of <p, SP> over the owned rows in `part` (single partition: SP
// needs no shared-node sum), rows are handed out grid-stride so that the number of partials is the grid size.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int SCHUR_THREADS = 256;

template <int G, int UN, int LD, bool DOT>
__global__ void __launch_bounds__(SCHUR_THREADS)
schur_sp4_kernel(int nNo, int mynNo, const int* __restrict__ rowPtr, const int* __restrict__ colPtr, const double* __restrict__ DL,
                 const double* __restrict__ P, const double* __restrict__ GP, double* __restrict__ SP, double* __restrict__ part)
{
  constexpr int LPR = 2 * G, RPW = 32 / LPR;
  __shared__ double wsum[SCHUR_THREADS / 32];
  const int lane = threadIdx.x & 31;
  const int rg = lane / LPR, lg = lane & (LPR - 1), g = lg >> 1, h = lg & 1;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long ngroups = ((long long)nNo + RPW - 1) / RPW;
  double dot = 0.0;
  for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < ngroups; w += nwarps) {
    const long long row = w * RPW + rg;
    double accA = 0.0, accB = 0.0;
    if (row < nNo) {
      const int k0 = __ldg(rowPtr + row), k1 = __ldg(rowPtr + row + 1);
      const double2* M = reinterpret_cast<const double2*>(DL) + h;
      for (int k = k0 + g; k < k1; k += UN * G) {
        int kk[UN], c[UN];
        bool ok[UN];
        double2 m[UN];
        double xa[UN], xb[UN];
#pragma unroll
        for (int q = 0; q < UN; q++) {
          ok[q] = k + q * G < k1;
          kk[q] = ok[q] ? k + q * G : k0;
          c[q] = __ldg(colPtr + kk[q]);
        }
#pragma unroll
        for (int q = 0; q < UN; q++) m[q] = ld_mat2<LD>(M + 2 * (size_t)kk[q]);
#pragma unroll
        for (int q = 0; q < UN; q++) {
          xa[q] = __ldg(GP + 3 * (size_t)c[q] + 2 * h);
          xb[q] = __ldg(h ? P + c[q] : GP + 3 * (size_t)c[q] + 1);
        }
#pragma unroll
        for (int q = 0; q < UN; q++) {
          accA += ok[q] ? m[q].x * xa[q] : 0.0;
          accB += ok[q] ? m[q].y * xb[q] : 0.0;
        }
      }
    }
    double dpart = h ? accA : accA + accB;
    double lpart = h ? accB : 0.0;
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
      dpart += __shfl_xor_sync(0xffffffffu, dpart, o);
      lpart += __shfl_xor_sync(0xffffffffu, lpart, o);
    }
    if (row < nNo && lg == 0) {
      const double sp = -1.0 * dpart + lpart;
      SP[row] = sp;
      if (DOT && row < mynNo) dot += __ldg(P + row) * sp;
    }
  }
  if (DOT) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) wsum[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < SCHUR_THREADS / 32; w++) s += wsum[w];
      part[blockIdx.x] = s;
    }
  }
}

int schur_sp4_num_variants() { return 6; }

// part == nullptr: no fused dot.  Returns the number of partials written (grid size) through nparts.
int schur_sp4(svb200_ctx* ctx, int variant, const double* DL, const double* P, const double* GP, double* SP, double* part, int* nparts)
{
  const int nNo = ctx->nNo;
  if (nNo == 0) { if (nparts) *nparts = 0; return SVB200_OK; }
  static int dflt = -1;
  if (dflt < 0) { const char* s = getenv("SVB200_SCHUR_VARIANT"); dflt = s ? atoi(s) : 2; }
  if (variant < 0) variant = dflt;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
#define SV(v, G_, UN_, LD_)                                                                                                        \
  case v: {                                                                                                                     \
    constexpr int RPW = 32 / (2 * G_);                                                                                          \
    const long long ngroups = ((long long)nNo + RPW - 1) / RPW;                                                                 \
    const long long want = (ngroups * 32 + SCHUR_THREADS - 1) / SCHUR_THREADS;                                                  \
    const int blocks = (int)std::min<long long>(want, (long long)sms * 8);                                                      \
    if (part) schur_sp4_kernel<G_, UN_, LD_, true><<<blocks, SCHUR_THREADS, 0, ctx->stream>>>(nNo, ctx->mynNo, ctx->d_rowPtr, ctx->d_colPtr, DL, P, GP, SP, part); \
    else schur_sp4_kernel<G_, UN_, LD_, false><<<blocks, SCHUR_THREADS, 0, ctx->stream>>>(nNo, ctx->mynNo, ctx->d_rowPtr, ctx->d_colPtr, DL, P, GP, SP, nullptr); \
    if (nparts) *nparts = blocks;                                                                                               \
  } break;
  switch (variant) {
    SV(0, 1, 4, 0) SV(1, 2, 2, 0) SV(2, 2, 4, 0) SV(3, 1, 4, 1) SV(4, 4, 2, 0) SV(5, 1, 8, 0)
    default:
      set_error("svb200: unknown schur_sp4 variant");
      return SVB200_ERR_INVALID;
  }
#undef SV
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// ---- the vector part of a Schur CG iteration on a single partition: two kernels + two single-CTA reductions ------------------
// cg scalars as in fsils_kernels.cu: cg[0] errO, cg[1] <r,r>, cg[2] <p,Sp>, cg[3] eps, cg[4] done, cg[5] iterations, cg[6] errO of the
// last executed iteration.
constexpr int CGF_THREADS = 256;

// out = sum of part[0..n) in a fixed order (thread t adds part[t], part[t + 256], ...; then a fixed tree).  ADVANCE: `out` is <r,r> of
// the finished iteration: count it, shift err -> errO, raise `done` when the next iteration's test err < eps would fire
// (cgrad.cpp:70-76, 92-93).
template <bool ADVANCE>
__global__ void __launch_bounds__(CGF_THREADS) cg_reduce_kernel(int n, const double* __restrict__ part, double* __restrict__ cg, int slot)
{
  __shared__ double sm[CGF_THREADS];
  if (cg[4] != 0.0) return;
  double s = 0.0;
  for (int k = threadIdx.x; k < n; k += CGF_THREADS) s += part[k];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = CGF_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double v = sm[0];
    cg[slot] = v;
    if (ADVANCE) {
      double err = sqrt(v);
      err = err * err;
      cg[6] = cg[0];
      cg[0] = err;
      cg[5] += 1.0;
      if (err < cg[3]) cg[4] = 1.0;
    }
  }
}

// X = alpha P + X, R = -alpha SP + R (alpha = errO / <p,Sp>) and the per-CTA partial sums of <r,r> over the owned rows.
__global__ void __launch_bounds__(CGF_THREADS)
cg_xr_dot_kernel(long long n, long long nOwned, const double* __restrict__ cg, const double* __restrict__ P, const double* __restrict__ SP,
                 double* __restrict__ X, double* __restrict__ R, double* __restrict__ part)
{
  __shared__ double wsum[CGF_THREADS / 32];
  if (cg[4] != 0.0) return;
  const double alpha = cg[0] / cg[2];
  double s = 0.0;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    X[k] = alpha * P[k] + X[k];
    const double r = -alpha * SP[k] + R[k];
    R[k] = r;
    if (k < nOwned) s += r * r;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < CGF_THREADS / 32; w++) t += wsum[w];
    part[blockIdx.x] = t;
  }
}

// P = (errO/err) R + P, then P = (err/errO) P (cgrad.cpp:95-96), with errO = cg[6], err = cg[0] after the advance.
__global__ void __launch_bounds__(CGF_THREADS)
cg_p_adv_kernel(long long n, const double* __restrict__ cg, const double* __restrict__ R, double* __restrict__ P)
{
  if (cg[4] != 0.0) return;
  const double errO = cg[6], err = cg[0];
  const double a = errO / err, b = err / errO;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    const double p = a * R[k] + P[k];
    P[k] = b * p;
  }
}

// One fused vector update of the Schur CG (after SP and the partials of <p,Sp> are there): 4 launches.
int schur_cg_fused_tail(svb200_ctx* ctx, double* cg, int npart_psp, double* part_psp, const double* SP, double* P, double* X, double* R,
                        double* part_rr)
{
  const long long n = ctx->nNo;
  if (n == 0) return SVB200_OK;
  const int blocks = (int)std::min<long long>((n + CGF_THREADS - 1) / CGF_THREADS, 148 * 4);
  cg_reduce_kernel<false><<<1, CGF_THREADS, 0, ctx->stream>>>(npart_psp, part_psp, cg, 2);
  cg_xr_dot_kernel<<<blocks, CGF_THREADS, 0, ctx->stream>>>(n, ctx->mynNo, cg, P, SP, X, R, part_rr);
  cg_reduce_kernel<true><<<1, CGF_THREADS, 0, ctx->stream>>>(blocks, part_rr, cg, 1);
  cg_p_adv_kernel<<<blocks, CGF_THREADS, 0, ctx->stream>>>(n, cg, R, P);
  ctx->launches += 4;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

}  // namespace svb
