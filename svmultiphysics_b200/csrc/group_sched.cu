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
//                                        plan) or node id (residual plan), ordered by count
//                                        descending so that the lanes of a warp see equal trip counts;
//   upartner[...]                       : tangent plan only: the transposed slot of an edge, -1 for a diagonal block;
//   contrib[g*GROUP*PER_EL + start ..]  : 16-bit ids (e_local, local nodes) of the contributions to that
//                                        target, ascending (=> a fixed summation order).
// Built by one CTA per group: bitonic sort of (target, id) keys in shared memory, run twice
// (count, host prefix sum over groups, fill).
#include <algorithm>
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

// MODE 0: residual plan — 4 targets per element (its nodes), id = e_local << 2 | a.
// MODE 1: tangent plan — 10 targets per element: the 4 diagonal blocks (a,a) and the 6 edges {a,b}.  An edge
//         is keyed by the CSR slot (lo,hi) with lo = the local node with the smaller global id, so that every
//         element sharing the edge produces the same key; the thread that owns it emits BOTH blocks (lo,hi) and
//         (hi,lo) from one set of loads (partner = slot (hi,lo), -1 for a diagonal).  id = e_local << 4 | lo << 2 | hi.
template <int MODE>
struct SchedTraits;
template <>
struct SchedTraits<0> { static constexpr int PER_EL = 4, NSORT = 512; };
template <>
struct SchedTraits<1> { static constexpr int PER_EL = 10, NSORT = 2048; };
constexpr int IDBITS = 11;

template <int MODE>
__global__ void __launch_bounds__(SCHED_THREADS)
build_sched_kernel(const int* __restrict__ IEN, const int* __restrict__ slot, int nEl, int fill, int* __restrict__ nuniq,
                   const int* __restrict__ uptr, int2* __restrict__ uent, int* __restrict__ upartner,
                   unsigned short* __restrict__ contrib)
{
  constexpr int PER_EL = SchedTraits<MODE>::PER_EL;
  constexpr int N = SchedTraits<MODE>::NSORT;
  constexpr int NC = ASM_GROUP * PER_EL;
  static_assert(NC <= N, "sort size");
  __shared__ unsigned long long keys[N];
  __shared__ unsigned long long keys2[N];
  __shared__ int scan[SCHED_THREADS];
  __shared__ int total;
  const int g = blockIdx.x;
  const unsigned long long BAD = ~0ull;

  for (int i = threadIdx.x; i < N; i += SCHED_THREADS) {
    unsigned long long key = BAD;
    const int el = i / PER_EL, q = i % PER_EL;
    const long long e = (long long)g * ASM_GROUP + el;
    if (i < NC && e < nEl) {
      int t, id;
      if (MODE == 0) {
        t = IEN[e * 4 + q];
        id = el << 2 | q;
      } else {
        int a, b;
        if (q < 4) { a = q; b = q; }
        else if (q < 7) { a = 0; b = q - 3; }
        else if (q < 9) { a = 1; b = q - 5; }
        else { a = 2; b = 3; }
        if (IEN[e * 4 + a] > IEN[e * 4 + b]) { const int tmp = a; a = b; b = tmp; }
        t = slot[e * 16 + a * 4 + b];
        id = el << 4 | a << 2 | b;
      }
      if (t >= 0) key = ((unsigned long long)(unsigned)t << IDBITS) | (unsigned)id;
    }
    keys[i] = key;
  }
  __syncthreads();
  bitonic_sort<N>(keys);

  // heads of runs of equal targets; each thread owns N/SCHED_THREADS consecutive positions
  constexpr int PER_T = N / SCHED_THREADS;
  const int p0 = threadIdx.x * PER_T;
  int nh = 0, nv = 0;
  for (int i = p0; i < p0 + PER_T; i++) {
    const bool valid = keys[i] != BAD;
    const bool head = valid && (i == 0 || (keys[i] >> IDBITS) != (keys[i - 1] >> IDBITS));
    nh += head;
    nv += valid;
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
    const bool head = valid && (i == 0 || (keys[i] >> IDBITS) != (keys[i - 1] >> IDBITS));
    if (head) keys2[u++] = (unsigned long long)i;
  }
  __syncthreads();
  scan[threadIdx.x] = nv;   // number of valid keys = position of the first BAD (BAD keys sort last)
  __syncthreads();
  if (threadIdx.x == 0) {
    int sum = 0;
    for (int t = 0; t < SCHED_THREADS; t++) sum += scan[t];
    total = sum;
  }
  __syncthreads();
  const int nValid = total;
  unsigned long long mine[(N + SCHED_THREADS - 1) / SCHED_THREADS];
  int cnt_i = 0;
  for (int k = threadIdx.x; k < nU; k += SCHED_THREADS) {
    const int start = (int)keys2[k];
    const int end = (k + 1 < nU) ? (int)keys2[k + 1] : nValid;
    int count = end - start;   // <= ASM_GROUP
    if (MODE == 1) {           // diagonal blocks first: a warp then runs one kind of target (one code path)
      const int id = (int)(keys[start] & ((1 << IDBITS) - 1));
      if (((id >> 2) & 3) == (id & 3)) count += 256;
    }
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
    const int count = (0xFFFF - (int)(q >> 16)) & 255;
    const int target = (int)(keys[start] >> IDBITS);
    uent[base + k] = make_int2(target, start | (count << 16));
    if (MODE == 1) {
      const int id = (int)(keys[start] & ((1 << IDBITS) - 1));
      const int el = id >> 4, lo = (id >> 2) & 3, hi = id & 3;
      upartner[base + k] = (lo != hi) ? slot[((long long)g * ASM_GROUP + el) * 16 + hi * 4 + lo] : -1;
    }
  }
  for (int i = threadIdx.x; i < NC; i += SCHED_THREADS)
    contrib[(size_t)g * NC + i] = keys[i] == BAD ? (unsigned short)0xFFFF : (unsigned short)(keys[i] & ((1 << IDBITS) - 1));
}

template <int MODE>
static int build_one(svb200_ctx* ctx, const Mesh& m, GroupSched& S)
{
  constexpr int NC = ASM_GROUP * SchedTraits<MODE>::PER_EL;
  const int nEl = m.nEl;
  const int nGrp = (nEl + ASM_GROUP - 1) / ASM_GROUP;
  S.nGrp = nGrp;
  if (nGrp == 0) return SVB200_OK;
  SVB_CUDA(cudaMalloc(&S.d_nuniq, sizeof(int) * nGrp));
  SVB_CUDA(cudaMalloc(&S.d_uptr, sizeof(int) * (nGrp + 1)));
  build_sched_kernel<MODE><<<nGrp, SCHED_THREADS, 0, ctx->stream>>>(m.d_IEN, m.d_slot, nEl, 0, S.d_nuniq, nullptr, nullptr,
                                                                   nullptr, nullptr);
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
  if (MODE == 1) SVB_CUDA(cudaMalloc(&S.d_upartner, sizeof(int) * std::max<long long>(tot, 1)));
  SVB_CUDA(cudaMalloc(&S.d_contrib, sizeof(unsigned short) * (size_t)nGrp * NC));
  build_sched_kernel<MODE><<<nGrp, SCHED_THREADS, 0, ctx->stream>>>(m.d_IEN, m.d_slot, nEl, 1, S.d_nuniq, S.d_uptr, S.d_uent,
                                                                   S.d_upartner, S.d_contrib);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SVB200_OK;
}

// Highest CSR slot each 128-element group adds to (one CTA per group, max over its elements' 16 slots).
__global__ void __launch_bounds__(ASM_GROUP) group_max_slot_kernel(const int* __restrict__ slot, int nEl, int* __restrict__ gmax)
{
  __shared__ int red[ASM_GROUP];
  const long long e = (long long)blockIdx.x * ASM_GROUP + threadIdx.x;
  int mx = -1;
  if (e < nEl)
    for (int k = 0; k < 16; k++) mx = max(mx, slot[e * 16 + k]);
  red[threadIdx.x] = mx;
  __syncthreads();
  for (int o = ASM_GROUP / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = max(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) gmax[blockIdx.x] = red[0];
}

// grp_need[g] = 1 + the highest slot any of the groups 0..g adds to: Val[0, grp_need[g]) must be zero before group g runs.
// On a mesh whose numbering has locality this grows with g, so the zeroing of Val can run AHEAD of the element kernel on a
// second stream instead of in front of it (run_assemble, svb200_api.cu).
int build_group_slot_need(svb200_ctx* ctx, Mesh& m)
{
  m.grp_need.clear();
  if (m.eNoN != 4 || m.nEl == 0 || !m.d_slot) return SVB200_OK;
  const int nGrp = (m.nEl + ASM_GROUP - 1) / ASM_GROUP;
  int* d_gmax = nullptr;
  SVB_CUDA(cudaMalloc(&d_gmax, sizeof(int) * nGrp));
  group_max_slot_kernel<<<nGrp, ASM_GROUP, 0, ctx->stream>>>(m.d_slot, m.nEl, d_gmax);
  ctx->launches++;
  std::vector<int> gmax(nGrp);
  cudaError_t e = cudaMemcpyAsync(gmax.data(), d_gmax, sizeof(int) * nGrp, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_gmax);
  if (e != cudaSuccess) return cuda_fail(e, "build_group_slot_need", __FILE__, __LINE__);
  m.grp_need.resize(nGrp);
  long long run = 0;
  for (int g = 0; g < nGrp; g++) {
    run = std::max(run, (long long)gmax[g] + 1);
    m.grp_need[g] = run;
  }
  return SVB200_OK;
}

void free_group_sched(GroupSched& S)
{
  cudaFree(S.d_nuniq); cudaFree(S.d_uptr); cudaFree(S.d_uent); cudaFree(S.d_upartner); cudaFree(S.d_contrib);
  S = GroupSched();
}

// Tangent plan (diagonal blocks + edges, 10 per tet4) and residual plan (nodes, 4 per tet4).
int build_group_schedules(svb200_ctx* ctx, Mesh& m)
{
  free_group_sched(m.schedK);
  free_group_sched(m.schedR);
  if (m.eNoN != 4 || m.nEl == 0) return SVB200_OK;
  int rc = build_one<1>(ctx, m, m.schedK);
  if (rc) return rc;
  return build_one<0>(ctx, m, m.schedR);
}

}  // namespace svb
