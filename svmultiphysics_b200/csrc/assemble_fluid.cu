// assemble_fluid.cu — fused element loop + scatter for fluid TET4 (K1 of SURVEY.md §7).
//
// Replaces, per Newton iteration, fluid::construct_fluid (Code/Source/solver/fluid.cpp:480-762),
// nn::gnn (solver/nn.cpp:862-899), fluid_3d_m / fluid_3d_c (fluid.cpp:1768-2237 / 1443-1760) and the
// lhsa_ns::do_assem scatter (solver/lhsa.cpp:70-114).
//
// Mapping: one thread per element computes the ~60 element moments of fluid_elem.cuh in registers;
// the 16 tangent blocks are then emitted one (a,b) pair at a time, transposed through a per-warp
// shared-memory tile so that a half-warp adds the 16 contiguous doubles of ONE 128-byte CSR block
// with one fully coalesced REDG.F64 (measured on B200: 554 G red/s coalesced vs 174 G red/s when
// every lane walks its own block, profiles/r1_microbench_fp64_red.txt).  The CSR slot of every
// (a,b) pair comes from a precomputed element->slot map, so the binary search of do_assem is gone.
// Scatter modes: ATOMIC (red.global.add.f64, all elements in one launch) and COLORED (one launch per
// colour, plain read-modify-write, bitwise reproducible).
#include <cstdlib>
#include "fluid_elem.cuh"

namespace svb {

constexpr int ASM_THREADS = 128;
constexpr int TILE_LD = 17;   // 16 doubles + 1 pad: conflict-free 64-bit column reads

template <bool ATOMIC>
__device__ __forceinline__ void add_f64(double* p, double v)
{
  if (ATOMIC) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
  } else {
    *p += v;
  }
}

__device__ __forceinline__ int pick_domain(const FluidArgs& P, int e)
{
  int iD = 0;
  for (int d = 0; d < P.nDmn; d++) {
    iD = d;
    if (P.dmn[d].Id == -1) break;
    if (P.eId != nullptr && ((P.eId[e] >> P.dmn[d].Id) & 1)) break;
  }
  return iD;
}

template <bool ATOMIC, int MINB>
__global__ void __launch_bounds__(ASM_THREADS, MINB)
assemble_fluid_tet4_kernel(const __grid_constant__ FluidArgs P)
{
  __shared__ double tile[ASM_THREADS / 32][32 * TILE_LD];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  double* T = tile[warp];

  const int idx = P.e0 + blockIdx.x * ASM_THREADS + threadIdx.x;
  bool active = idx < P.e1;
  int e = 0;
  if (active) e = P.perm ? P.perm[idx] : idx;

  Tet4Elem E;
  int node[4] = {0, 0, 0, 0};
  if (active) {
    const int iD = pick_domain(P, e);
    if (!P.dmn[iD].isFluid || (P.emask != nullptr && P.emask[e] != P.emask_val)) {
      active = false;
    } else {
      const int4 n4 = *reinterpret_cast<const int4*>(P.IEN + 4 * (size_t)e);
      node[0] = n4.x; node[1] = n4.y; node[2] = n4.z; node[3] = n4.w;
      double xl[4][3], yl[4][4], uc[4][3], ab[4][3];
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const size_t n = (size_t)node[a];
        const double* xp = P.x + 3 * n;
        const double* bp = P.Bf + 3 * n;
        const double* ap = P.Ag + (size_t)P.tDof * n;
        const double* yp = P.Yg + (size_t)P.tDof * n;
#pragma unroll
        for (int i = 0; i < 3; i++) {
          xl[a][i] = __ldg(xp + i) + (P.ale ? __ldg(P.Dg + (size_t)P.tDof * n + 4 + i) : 0.0);
          ab[a][i] = P.bfZero ? __ldg(ap + i) : __ldg(ap + i) - __ldg(bp + i);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) yl[a][i] = __ldg(yp + i);
#pragma unroll
        for (int i = 0; i < 3; i++) uc[a][i] = yl[a][i] - (P.mvMsh ? __ldg(yp + 4 + i) : 0.0);
      }
      tet4_element(P, P.dmn[iD], xl, yl, uc, ab, E);
    }
  }

  // ---- residual: R(:,node_a) += lR(:,a); 16 doubles per element through the tile -------------------
  const int half = lane >> 4, j = lane & 15;
  if (active) {
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int i = 0; i < 4; i++) T[lane * TILE_LD + 4 * a + i] = E.lR[a][i];
  }
  __syncwarp();
#pragma unroll 4
  for (int r = 0; r < 16; r++) {
    const int src = 2 * r + half;
    const int n0 = __shfl_sync(0xffffffffu, node[0], src);
    const int n1 = __shfl_sync(0xffffffffu, node[1], src);
    const int n2 = __shfl_sync(0xffffffffu, node[2], src);
    const int n3 = __shfl_sync(0xffffffffu, node[3], src);
    const bool act = __shfl_sync(0xffffffffu, (int)active, src);
    const int a = j >> 2;
    const int n = a == 0 ? n0 : (a == 1 ? n1 : (a == 2 ? n2 : n3));
    if (act) add_f64<ATOMIC>(P.R + 4 * (size_t)n + (j & 3), T[src * TILE_LD + j]);
  }
  __syncwarp();

  // ---- tangent: one (a,b) block per step ------------------------------------------------------------
  const int* slotp = P.slot + 16 * (size_t)e;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    int4 s4 = make_int4(-1, -1, -1, -1);
    if (active) s4 = *reinterpret_cast<const int4*>(slotp + 4 * a);
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int myslot = b == 0 ? s4.x : (b == 1 ? s4.y : (b == 2 ? s4.z : s4.w));
      if (active) {
        double K[16];
        tet4_block(E, a, b, K);
#pragma unroll
        for (int i = 0; i < 16; i++) T[lane * TILE_LD + i] = K[i];
      }
      __syncwarp();
#pragma unroll 4
      for (int r = 0; r < 16; r++) {
        const int src = 2 * r + half;
        const int s = __shfl_sync(0xffffffffu, myslot, src);
        if (s >= 0) add_f64<ATOMIC>(P.Val + 16 * (size_t)s + j, T[src * TILE_LD + j]);
      }
      __syncwarp();
    }
  }
}

// =====================================================================================================
// Grouped scatter (default ATOMIC path).  One CTA owns ASM_GROUP consecutive elements:
//   phase 1  one thread per element: gather, gnn, Gauss loop -> the element's ~60 moments go to shared memory;
//   phase 2  one thread per DISTINCT diagonal block / mesh edge the group touches (plan: group_sched.cu): it
//            re-emits the 4x4 blocks of all the group's contributions from the moments (an edge yields both
//            blocks (a,b) and (b,a) from one set of loads; their velocity parts are transposes), sums
//            them in registers in a fixed order and adds the block through a per-warp transposition tile with ONE
//            coalesced 128-byte RED per half-warp — 5.4 block reductions per element instead of 16.  (A TMA
//            bulk reduce per block, cp.reduce.async.bulk...add.f64, was measured slower here: UBLKRED is a
//            uniform-datapath instruction and serialises over the lanes of a warp.)
//   (between the two) the same for the residual rows (distinct nodes of the group, 4 REDs each).
// Phase 1 uses tet4_element_staged (fluid_elem.cuh): 168 registers, so three CTAs are resident per SM.
// Only blocks shared between groups still meet in the L2 atomic units, so the result is bitwise reproducible
// up to the order of those few cross-group additions.
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* sdst, const void* gsrc)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}

__device__ __forceinline__ void cp_async4(void* sdst, const void* gsrc)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}

constexpr int ENT_CACHE = 512;              // plan entries of the group kept in shared memory (the rest: global)
constexpr int ASM_WARPS = ASM_GROUP / 32;

// Warp-wide scatter of one 4x4 block per lane (slot < 0: none): the blocks are transposed through the warp's
// tile so that a half-warp adds the 16 contiguous doubles of one block with ONE coalesced RED; 4 in flight.
__device__ __forceinline__ void red_blocks(double* __restrict__ Val, double* T, const int lane, const int myslot,
                                           const double (&K)[16])
{
  const int half = lane >> 4, j = lane & 15;
  if (myslot >= 0) {
#pragma unroll
    for (int i = 0; i < 16; i++) T[lane * TILE_LD + i] = K[i];
  }
  __syncwarp();
  double* base = Val + j;
#pragma unroll
  for (int r0 = 0; r0 < 16; r0 += 4) {
    int sl[4];
    double v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int src = 2 * (r0 + q) + half;
      sl[q] = __shfl_sync(0xffffffffu, myslot, src);
      v[q] = T[src * TILE_LD + j];
    }
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (sl[q] >= 0) add_f64<true>(base + 16 * (size_t)sl[q], v[q]);
  }
  __syncwarp();
}

template <bool NN>
__global__ void __launch_bounds__(ASM_GROUP, NN ? 2 : 3)
assemble_fluid_tet4_grouped_kernel(const __grid_constant__ FluidArgs P)
{
  constexpr int RS = NN ? REC_NN : REC_NEWT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* tiles = reinterpret_cast<double*>(smem_raw);                             // [warps][32*TILE_LD]
  double* rec = tiles + ASM_WARPS * 32 * TILE_LD;                                  // [ASM_GROUP][RS]
  int2* entc = reinterpret_cast<int2*>(rec + ASM_GROUP * RS);                      // [ENT_CACHE]
  int* entp = reinterpret_cast<int*>(entc + ENT_CACHE);                            // [ENT_CACHE]
  unsigned short* ctr = reinterpret_cast<unsigned short*>(entp + ENT_CACHE);       // [ASM_GROUP*10]
  unsigned short* ctrR = ctr + ASM_GROUP * 10;                                     // [ASM_GROUP*4]
  unsigned char* act = reinterpret_cast<unsigned char*>(ctrR + ASM_GROUP * 4);     // [ASM_GROUP]

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int g = P.gperm ? __ldg(P.gperm + P.g0 + blockIdx.x) : P.g0 + (int)blockIdx.x;
  double* T = tiles + (tid >> 5) * 32 * TILE_LD;
  // the group's plan goes to shared memory with cp.async (lands while phase 1 computes)
  {
    const unsigned short* srcK = P.kContrib + (size_t)g * ASM_GROUP * 10;   // 2560 bytes = 160 x 16
    cp_async16(ctr + 8 * tid, srcK + 8 * tid);
    if (tid < ASM_GROUP * 10 * 2 / 16 - ASM_GROUP) cp_async16(ctr + 8 * (tid + ASM_GROUP), srcK + 8 * (tid + ASM_GROUP));
    if (tid < ASM_GROUP * 4 * 2 / 16) cp_async16(ctrR + 8 * tid, P.rContrib + (size_t)g * ASM_GROUP * 4 + 8 * tid);
  }
  const int ub = __ldg(P.kU_ptr + g);
  const int G = __ldg(P.kU_ptr + g + 1) - ub;
  for (int k = tid; k < min(G, ENT_CACHE); k += ASM_GROUP) {
    cp_async8(entc + k, P.kU_ent + ub + k);
    cp_async4(entp + k, P.kU_partner + ub + k);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  // ---- phase 1: element record -> shared memory, element residual -> this lane's tile row ------------------
  const int e = g * ASM_GROUP + tid;
  bool active = e < P.e1;
  if (active) {
    const int iD = pick_domain(P, e);
    if (!P.dmn[iD].isFluid || (P.emask != nullptr && P.emask[e] != P.emask_val)) {
      active = false;
    } else {
      const int4 n4 = *reinterpret_cast<const int4*>(P.IEN + 4 * (size_t)e);
      const int node[4] = {n4.x, n4.y, n4.z, n4.w};
      double xl[4][3], yl[4][4], uc[4][3], ab[4][3];
      if (P.tDof == 4 && P.bfZero == 2 && !P.ale) {
        // the common fluid case (tDof = 4: a node's state is one aligned 32-byte sector; no body-force array): 128-bit loads,
        // 5 instead of 13 load instructions per node
#pragma unroll
        for (int a = 0; a < 4; a++) {
          const size_t n = (size_t)node[a];
          const double2 a01 = __ldg(reinterpret_cast<const double2*>(P.Ag + 4 * n));
          const double a2 = __ldg(P.Ag + 4 * n + 2);
          const double2 y01 = __ldg(reinterpret_cast<const double2*>(P.Yg + 4 * n));
          const double2 y23 = __ldg(reinterpret_cast<const double2*>(P.Yg + 4 * n) + 1);
          const double* xp = P.x + 3 * n;
          xl[a][0] = __ldg(xp); xl[a][1] = __ldg(xp + 1); xl[a][2] = __ldg(xp + 2);
          ab[a][0] = a01.x; ab[a][1] = a01.y; ab[a][2] = a2;
          yl[a][0] = y01.x; yl[a][1] = y01.y; yl[a][2] = y23.x; yl[a][3] = y23.y;
          uc[a][0] = y01.x; uc[a][1] = y01.y; uc[a][2] = y23.x;
        }
      } else {
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const size_t n = (size_t)node[a];
        const double* xp = P.x + 3 * n;
        const double* bp = P.Bf + 3 * n;
        const double* ap = P.Ag + (size_t)P.tDof * n;
        const double* yp = P.Yg + (size_t)P.tDof * n;
#pragma unroll
        for (int i = 0; i < 3; i++) {
          xl[a][i] = __ldg(xp + i) + (P.ale ? __ldg(P.Dg + (size_t)P.tDof * n + 4 + i) : 0.0);
          ab[a][i] = P.bfZero ? __ldg(ap + i) : __ldg(ap + i) - __ldg(bp + i);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) yl[a][i] = __ldg(yp + i);
#pragma unroll
        for (int i = 0; i < 3; i++) uc[a][i] = yl[a][i] - (P.mvMsh ? __ldg(yp + 4 + i) : 0.0);
      }
      }
      tet4_element_staged(P, P.dmn[iD], xl, yl, uc, ab, NN, rec + tid * RS, T + lane * TILE_LD);
    }
  }
  act[tid] = active ? 1 : 0;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  const bool allActive = __syncthreads_and(active);

  // plan entries beyond the shared-memory cache (targets 512.. of the ~690 of a group, processed in the last passes of phase 3):
  // requested now, so that their L2 latency hides behind phase 2 instead of stalling the start of those passes
  constexpr int PF = 2;
  int2 pfe[PF];
  int pfp[PF];
#pragma unroll
  for (int q = 0; q < PF; q++) {
    const int k = ENT_CACHE + q * ASM_GROUP + tid;
    pfe[q] = make_int2(0, 0);
    pfp[q] = -1;
    if (k < G) { pfe[q] = __ldg(P.kU_ent + ub + k); pfp[q] = __ldg(P.kU_partner + ub + k); }
  }

  // ---- phase 2: residual rows of the group's distinct nodes (element residuals sit in the tiles) ------------
  {
    const int rb = __ldg(P.rU_ptr + g);
    const int GR = __ldg(P.rU_ptr + g + 1) - rb;
    for (int k = tid; k < GR; k += ASM_GROUP) {
      const int2 ent = __ldg(P.rU_ent + rb + k);
      const int start = ent.y & 0xFFFF, end = start + (ent.y >> 16);
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      bool any = false;
      for (int c = start; c < end; c++) {
        const int id = ctrR[c];
        const int el = id >> 2, a = id & 3;
        if (!act[el]) continue;
        any = true;
        const double* r = tiles + el * TILE_LD + 4 * a;   // tile rows of consecutive warps are contiguous
        s0 += r[0]; s1 += r[1]; s2 += r[2]; s3 += r[3];
      }
      if (any) {
        double* dst = P.R + 4 * (size_t)ent.x;
        add_f64<true>(dst, s0); add_f64<true>(dst + 1, s1); add_f64<true>(dst + 2, s2); add_f64<true>(dst + 3, s3);
      }
    }
  }
  __syncthreads();

  // ---- phase 3: one thread per distinct diagonal block / edge (both blocks of the edge) ----------------------
  // The contributions of the group's elements are accumulated in the bilinear form of fluid_elem.cuh (EdgeAcc / DiagAcc)
  // and the 4x4 blocks are assembled once per target (SVB200_ASM_BLOCKWISE=1 at build time restores the per-contribution
  // block form).
  {
    for (int k0 = (tid & ~31); k0 < G; k0 += ASM_GROUP) {   // warp-uniform trip count
      const int k = k0 + lane;
      int slot1 = -1, slot2 = -1;
      double K1[16];
      bool isEdge = false;
      EdgeAcc EA;
      if (k < G) {
        int2 ent;
        int partner;
        if (k < ENT_CACHE) { ent = entc[k]; partner = entp[k]; }
        else if (k < ENT_CACHE + PF * ASM_GROUP) {
          // k = ENT_CACHE + q ASM_GROUP + tid: this thread's own prefetched entry (ENT_CACHE is a multiple of ASM_GROUP)
          const int q = (k - ENT_CACHE) / ASM_GROUP;
          ent = q == 0 ? pfe[0] : pfe[1];
          partner = q == 0 ? pfp[0] : pfp[1];
        } else { ent = __ldg(P.kU_ent + ub + k); partner = __ldg(P.kU_partner + ub + k); }
        const int start = ent.y & 0xFFFF, end = start + (ent.y >> 16);
        if (partner >= 0) {
          isEdge = true;
          edge_acc_zero(EA, NN);
          for (int c = start; c < end; c++) {
            const int id = ctr[c];
            const int el = id >> 4;
            if (!allActive && !act[el]) continue;
            slot1 = ent.x;
            slot2 = partner;
            tet4_edge_rec_acc(rec + el * RS, NN, (id >> 2) & 3, id & 3, EA);
          }
          edge_acc_block(EA, NN, 0, K1);
        } else {
          DiagAcc DA;
          diag_acc_zero(DA, NN);
          for (int c = start; c < end; c++) {
            const int id = ctr[c];
            const int el = id >> 4;
            if (!allActive && !act[el]) continue;
            slot1 = ent.x;
            tet4_diag_rec_acc(rec + el * RS, NN, id & 3, DA);
          }
          diag_acc_block(DA, NN, K1);
        }
      }
      red_blocks(P.Val, T, lane, slot1, K1);
      if (__any_sync(0xffffffffu, slot2 >= 0)) {
        if (isEdge) edge_acc_block(EA, NN, 1, K1);
        red_blocks(P.Val, T, lane, slot2, K1);
      }
    }
  }
}

// nn::gnn of a linear tetrahedron is constant over the element: Jac = det[x0-x3, x1-x3, x2-x3].  construct_fluid /
// construct_fsi throw when utils::is_zero(Jac) (fluid.cpp:637-639, fsi.cpp:185); the assembly kernels are register-tuned,
// so the test runs here, once per set of coordinates (every assembly when the geometry moves, ALE).
__global__ void __launch_bounds__(256) tet4_jacobian_check_kernel(const __grid_constant__ FluidArgs P)
{
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.e1) return;
  if (!P.dmn[pick_domain(P, e)].isFluid) return;
  const int4 n4 = *reinterpret_cast<const int4*>(P.IEN + 4 * (size_t)e);
  const int node[4] = {n4.x, n4.y, n4.z, n4.w};
  double xl[4][3];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int i = 0; i < 3; i++)
      xl[a][i] = __ldg(P.x + 3 * (size_t)node[a] + i) + (P.ale ? __ldg(P.Dg + (size_t)P.tDof * node[a] + 4 + i) : 0.0);
  double xXi[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int k = 0; k < 3; k++) xXi[i][k] += xl[a][i] * P.Nxi[0][a][k];
  const double Jac = xXi[0][0] * xXi[1][1] * xXi[2][2] + xXi[0][1] * xXi[1][2] * xXi[2][0] + xXi[0][2] * xXi[1][0] * xXi[2][1] -
                     xXi[0][0] * xXi[1][2] * xXi[2][1] - xXi[0][1] * xXi[1][0] * xXi[2][2] - xXi[0][2] * xXi[1][1] * xXi[2][0];
  if (is_zero(Jac)) atomicMax(P.err, e + 1);
}

int launch_tet4_jacobian_check(svb200_ctx* ctx, const Mesh& m, const FluidArgs& args)
{
  if (m.eNoN != 4 || m.nEl == 0) return SVB200_OK;
  FluidArgs A = args;
  A.e0 = 0; A.e1 = m.nEl; A.perm = nullptr;
  tet4_jacobian_check_kernel<<<(m.nEl + 255) / 256, 256, 0, ctx->stream>>>(A);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

template <bool NN>
static int launch_grouped(svb200_ctx* ctx, const FluidArgs& args)
{
  constexpr int RS = NN ? REC_NN : REC_NEWT;
  constexpr size_t smem = sizeof(double) * ASM_GROUP * (TILE_LD + RS) + sizeof(int2) * ENT_CACHE + sizeof(int) * ENT_CACHE + 2 * ASM_GROUP * 14 + ASM_GROUP;
  static bool configured = false;
  if (!configured) {
    SVB_CUDA(cudaFuncSetAttribute(assemble_fluid_tet4_grouped_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    configured = true;
  }
  // gperm: the groups gperm[g0 .. g0 + nGrpLaunch) (one group colour); else nGrpLaunch > 0: the group range [g0, g0 + nGrpLaunch)
  // (chunked launches behind the overlapped zeroing of Val); else all groups
  const int nGrp = (args.gperm || args.nGrpLaunch > 0) ? args.nGrpLaunch : (args.e1 + ASM_GROUP - 1) / ASM_GROUP;
  if (nGrp <= 0) return SVB200_OK;
  assemble_fluid_tet4_grouped_kernel<NN><<<nGrp, ASM_GROUP, smem, ctx->stream>>>(args);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

int launch_assemble_fluid(svb200_ctx* ctx, const Mesh& m, const FluidArgs& args)
{
  if (m.eNoN != 4) {
    set_error("svb200: fluid assembly is implemented for TET4 meshes (eNoN=4)");
    return SVB200_ERR_UNSUPPORTED;
  }
  const int n = args.e1 - args.e0;
  if (n <= 0) return SVB200_OK;
  const int blocks = (n + ASM_THREADS - 1) / ASM_THREADS;
  static const int variant = getenv("SVB200_ASM_MINB") ? atoi(getenv("SVB200_ASM_MINB")) : 2;   // tuning knob
  static const bool legacy = getenv("SVB200_ASM_LEGACY") != nullptr;   // A/B knob: per-entry RED scatter
  if ((args.atomic || args.gperm) && !legacy && args.kU_ptr && args.e0 == 0) {
    bool nn = false;
    for (int d = 0; d < args.nDmn; d++) nn |= (args.dmn[d].viscType != SVB200_VISC_CONST);
    return nn ? launch_grouped<true>(ctx, args) : launch_grouped<false>(ctx, args);
  }
  if (args.atomic) {
    if (variant == 4) assemble_fluid_tet4_kernel<true, 4><<<blocks, ASM_THREADS, 0, ctx->stream>>>(args);
    else if (variant == 3) assemble_fluid_tet4_kernel<true, 3><<<blocks, ASM_THREADS, 0, ctx->stream>>>(args);
    else assemble_fluid_tet4_kernel<true, 2><<<blocks, ASM_THREADS, 0, ctx->stream>>>(args);
  } else {
    assemble_fluid_tet4_kernel<false, 2><<<blocks, ASM_THREADS, 0, ctx->stream>>>(args);
  }
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

}  // namespace svb
