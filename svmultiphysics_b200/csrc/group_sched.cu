// group_sched.cu — one-time "group schedule" of a mesh: the pre-reduction plan of the grouped scatter.
//
// The reference scatters every element matrix entry by entry (lhsa_ns::do_assem,
// Code/Source/solver/lhsa.cpp:70-114).  On the device the FP64 atomic-add rate of the L2
// (≈ 550 G adds/s, profiles/r1_microbench_fp64_red.txt) caps a one-RED-per-entry scatter at
// ≈ 2.1 G tet4/s.  Neighbouring elements hit the same CSR blocks (6.3 contributions per block on a
// tet mesh), so a CTA that owns GROUP consecutive elements first sums, on chip, everything its
// elements add to one block and only then touches global memory — once per DISTINCT block.
//
// For every group g (elements [g*GROUP, (g+1)*GROUP)) this file builds, on the device:
//   uent[uptr[g] .. uptr[g]+nuniq[g])  : {target, start | count << 16}; target = CSR slot (tangent
//                                        schedule) or node id (residual schedule), ordered by
//                                        count descending so that the lanes of a warp see equal trip counts;
//   contrib[g*GROUP*PER_EL + start ..]  : 16-bit ids  e_local*PER_EL + pair  of the contributions to that
//                                        target, ascending (=> a fixed summation order).
// Built by one CTA per group: bitonic sort of (target, id) keys in shared memory, run twice
// (count, host prefix sum over groups, fill).
#include "svb200_internal.h"

namespace svb {

constexpr int SCHED_THREADS = 256;

template <int N>
__device__ __forceinline__ void bitonic_sort(unsigned long long* s)
{
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < N; i += SCHED_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = s[i], y = s[ixj];
          const bool asc = (i & k) == 0;
          if ((x > y) == asc) { s[i] = y; s[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
}

// PER_EL targets per element (16 slots or 4 nodes); IDXBITS = log2(GROUP*PER_EL).
template <int PER_EL, int IDXBITS>
__global__ void __launch_bounds__(SCHED_THREADS)
build_sched_kernel(const int* __restrict__ src, int nEl, int fill, int* __restrict__ nuniq,
                   const int* __restrict__ uptr, int2* __restrict__ uent, unsigned short* __restrict__ contrib)
{
  constexpr int N = ASM_GROUP * PER_EL;
  static_assert(N == (1 << IDXBITS), "group size");
  __shared__ unsigned long long keys[N];
  __shared__ unsigned long long keys2[N];
  __shared__ int scan[SCHED_THREADS];
  __shared__ int total;
  const int g = blockIdx.x;
  const unsigned long long BAD = ~0ull;

  for (int i = threadIdx.x; i < N; i += SCHED_THREADS) {
    const long long e = (long long)g * ASM_GROUP + i / PER_EL;
    unsigned long long key = BAD;
    if (e < nEl) {
      const int t = src[e * PER_EL + i % PER_EL];
      if (t >= 0) key = ((unsigned long long)(unsigned)t << IDXBITS) | (unsigned)i;
    }
    keys[i] = key;
  }
  __syncthreads();
  bitonic_sort<N>(keys);

  // heads of runs of equal targets; each thread owns N/SCHED_THREADS consecutive positions
  constexpr int PER_T = N / SCHED_THREADS;
  const int p0 = threadIdx.x * PER_T;
  int nh = 0;
  for (int i = p0; i < p0 + PER_T; i++) {
    const bool valid = keys[i] != BAD;
    const bool head = valid && (i == 0 || (keys[i] >> IDXBITS) != (keys[i - 1] >> IDXBITS));
    nh += head;
  }
  scan[threadIdx.x] = nh;
  __syncthreads();
  for (int d = 1; d < SCHED_THREADS; d <<= 1) {   // inclusive Hillis-Steele scan
    const int v = threadIdx.x >= d ? scan[threadIdx.x - d] : 0;
    __syncthreads();
    scan[threadIdx.x] += v;
    __syncthreads();
  }
  if (threadIdx.x == SCHED_THREADS - 1) total = scan[threadIdx.x];
  for (int i = threadIdx.x; i < N; i += SCHED_THREADS) keys2[i] = BAD;
  __syncthreads();
  const int nU = total;
  if (!fill) {
    if (threadIdx.x == 0) nuniq[g] = nU;
    return;
  }
  // keys2[u] = start position of unique u (temporarily), then (count-desc, start) sort keys
  int u = scan[threadIdx.x] - nh;
  for (int i = p0; i < p0 + PER_T; i++) {
    const bool valid = keys[i] != BAD;
    const bool head = valid && (i == 0 || (keys[i] >> IDXBITS) != (keys[i - 1] >> IDXBITS));
    if (head) keys2[u++] = (unsigned long long)i;
  }
  __syncthreads();
  // number of valid keys = position of the first BAD (all BAD keys sort last)
  int nValidLocal = 0;
  for (int i = p0; i < p0 + PER_T; i++) nValidLocal += keys[i] != BAD;
  scan[threadIdx.x] = nValidLocal;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int t = 0; t < SCHED_THREADS; t++) s += scan[t];
    total = s;
  }
  __syncthreads();
  const int nValid = total;
  unsigned long long mine[(N + SCHED_THREADS - 1) / SCHED_THREADS];
  int cnt_i = 0;
  for (int k = threadIdx.x; k < nU; k += SCHED_THREADS) {
    const int start = (int)keys2[k];
    const int end = (k + 1 < nU) ? (int)keys2[k + 1] : nValid;
    const int count = end - start;
    mine[cnt_i++] = ((unsigned long long)(0xFFFF - count) << 16) | (unsigned)start;
  }
  __syncthreads();
  cnt_i = 0;
  for (int k = threadIdx.x; k < nU; k += SCHED_THREADS) keys2[k] = mine[cnt_i++];
  __syncthreads();
  bitonic_sort<N>(keys2);
  const int base = uptr[g];
  for (int k = threadIdx.x; k < nU; k += SCHED_THREADS) {
    const unsigned long long q = keys2[k];
    const int start = (int)(q & 0xFFFF);
    const int count = 0xFFFF - (int)(q >> 16);
    const int target = (int)(keys[start] >> IDXBITS);
    uent[base + k] = make_int2(target, start | (count << 16));
  }
  for (int i = threadIdx.x; i < N; i += SCHED_THREADS)
    contrib[(size_t)g * N + i] = keys[i] == BAD ? (unsigned short)0xFFFF : (unsigned short)(keys[i] & (N - 1));
}

template <int PER_EL, int IDXBITS>
static int build_one(svb200_ctx* ctx, const int* d_src, int nEl, GroupSched& S)
{
  const int nGrp = (nEl + ASM_GROUP - 1) / ASM_GROUP;
  S.nGrp = nGrp;
  if (nGrp == 0) return SVB200_OK;
  SVB_CUDA(cudaMalloc(&S.d_nuniq, sizeof(int) * nGrp));
  SVB_CUDA(cudaMalloc(&S.d_uptr, sizeof(int) * (nGrp + 1)));
  build_sched_kernel<PER_EL, IDXBITS><<<nGrp, SCHED_THREADS, 0, ctx->stream>>>(d_src, nEl, 0, S.d_nuniq, nullptr, nullptr,
                                                                              nullptr);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  std::vector<int> nu(nGrp), ptr(nGrp + 1, 0);
  SVB_CUDA(cudaMemcpyAsync(nu.data(), S.d_nuniq, sizeof(int) * nGrp, cudaMemcpyDeviceToHost, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  long long tot = 0;
  for (int g = 0; g < nGrp; g++) {
    ptr[g] = (int)tot;
    tot += nu[g];
    if (tot > 0x7fffffffLL) {
      set_error("svb200: group schedule exceeds 2^31 entries");
      return SVB200_ERR_UNSUPPORTED;
    }
  }
  ptr[nGrp] = (int)tot;
  S.total = tot;
  SVB_CUDA(cudaMemcpyAsync(S.d_uptr, ptr.data(), sizeof(int) * (nGrp + 1), cudaMemcpyHostToDevice, ctx->stream));
  SVB_CUDA(cudaMalloc(&S.d_uent, sizeof(int2) * std::max<long long>(tot, 1)));
  SVB_CUDA(cudaMalloc(&S.d_contrib, sizeof(unsigned short) * (size_t)nGrp * ASM_GROUP * PER_EL));
  build_sched_kernel<PER_EL, IDXBITS><<<nGrp, SCHED_THREADS, 0, ctx->stream>>>(d_src, nEl, 1, S.d_nuniq, S.d_uptr,
                                                                              S.d_uent, S.d_contrib);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SVB200_OK;
}

void free_group_sched(GroupSched& S)
{
  cudaFree(S.d_nuniq); cudaFree(S.d_uptr); cudaFree(S.d_uent); cudaFree(S.d_contrib);
  S = GroupSched();
}

// Tangent schedule from the element->slot map (16 per tet4), residual schedule from IEN (4 per tet4).
int build_group_schedules(svb200_ctx* ctx, Mesh& m)
{
  free_group_sched(m.schedK);
  free_group_sched(m.schedR);
  if (m.eNoN != 4 || m.nEl == 0) return SVB200_OK;
  int rc = build_one<16, 11>(ctx, m.d_slot, m.nEl, m.schedK);
  if (rc) return rc;
  return build_one<4, 9>(ctx, m.d_IEN, m.nEl, m.schedR);
}

}  // namespace svb
