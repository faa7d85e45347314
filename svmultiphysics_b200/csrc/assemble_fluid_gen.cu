// assemble_fluid_gen.cu — fused element loop + scatter of the VMS fluid for elements that are not linear
// tetrahedra (HEX8; TET4 also runs through it as a cross-check of the specialised kernel of assemble_fluid.cu).
//
// Replaces fluid::construct_fluid (Code/Source/solver/fluid.cpp:480-762) with nn::gnn + nn::gn_nxx PER GAUSS POINT
// (solver/nn.cpp:862-899, 1172-1283), fluid_3d_m / fluid_3d_c (fluid.cpp:1768-2237 / 1443-1760) and the do_assem
// scatter (solver/lhsa.cpp:70-114).  The algebra is fluid_gen.cuh.
//
// Mapping (same idea as the solid kernel): ENON lanes per element, 32/ENON elements per warp, nG == ENON.
//   phase A  lane g evaluates Gauss point g once: Jacobian, physical first and second derivatives, the interpolated
//            state, viscosity, tau_M/C/B, the fine-scale velocity — and leaves a FluidGP (55 doubles) plus one
//            FluidNode (11 doubles) per element node in shared memory.  The lane of the LAST Gauss point publishes its
//            second derivatives first, because the reference's continuity loop uses them at every Gauss point.
//   phase B  lane a = element node a = one block row of the element matrix: residual row, then FG_NB column nodes at a
//            time: per Gauss point the (g, a) part of the tangent is folded into a FluidRow once (fluid_gen_row) and the
//            FG_NB 4x4 blocks accumulate in registers (~80 FMAs and 11 shared loads per block and Gauss point), then each
//            is scattered as 16 contiguous doubles = one CSR block.
#include <cstdlib>
#include <vector>
#include <cub/cub.cuh>
#include "fluid_gen.cuh"

namespace svb {

struct FluidGenArgs {
  const int* IEN;
  const int* eId;
  const int* slot;
  const int* perm;
  const double* x;
  const double* Ag;
  const double* Yg;
  const double* Bf;
  const double* Dg;
  const double* tab;    // per Gauss point: w | N[ENON] | Nxi[ENON][3] | Nxi2[ENON][6]
  int* err;             // 1 + index of an element with a zero Jacobian (construct_fluid throws, fluid.cpp:645-647)
  double* R;
  double* Val;
  int e0, e1;
  int tDof, mvMsh, nDmn, atomic, ale;
  int lShpF;            // mshType::lShpF (nn_elem_props.h): TET4 and WDG — nn::gnn / gn_nxx are evaluated at Gauss point 0 only
                        // (fluid.cpp:641, 707), also for the wedge, whose gradients are NOT constant: reproduced as is
  double dt, af, am, gam;
  FluidDmn dmn[MAX_DMN];
  const double* uris;   // URIS valves (svb200_set_uris) or null
  int nUris;
  svb200_uris urisP[SVB200_MAX_URIS];
};

template <bool ATOMIC>
__device__ __forceinline__ void fg_add(double* p, double v)
{
  if (ATOMIC) asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
  else *p += v;
}

constexpr int FG_THREADS = 64;
constexpr int FG_NB = 4;      // column nodes per pass of phase B
__host__ __device__ constexpr int fg_tab_ld(int enon) { return 1 + enon * 10; }
// stride between the FluidNode tables of consecutive Gauss points: odd, so that the phase-A stores of the lanes of an
// element (lane = Gauss point) fall into different banks
__host__ __device__ constexpr int fg_nd_ld(int enon) { return enon * FLUID_NODEC_DOUBLES + 1; }
// per element: nodal inputs (x 3, al 3, yl 4, bfl 3, ym 3 = 16) | NxxL 6 | FluidGP per Gauss point | FluidNodeC tables;
// padded to 8 mod 16 doubles (the two elements of a half-warp then use disjoint banks)
__host__ __device__ constexpr int fg_per_el(int enon, int ng)
{
  const int n = enon * 16 + enon * 6 + ng * FLUID_GP_DOUBLES + ng * fg_nd_ld(enon);
  return n + ((8 - (n % 16)) + 16) % 16;
}
__host__ __device__ constexpr bool fg_tab_in_smem(int enon, int ng) { return ng * fg_tab_ld(enon) <= 2048; }
// lanes per element: one per element node in phase B and one per Gauss point in phase A
__host__ __device__ constexpr int fg_lpe(int enon, int ng) { return enon > ng ? enon : ng; }
// CTA size: two warps, one for the big quadratic elements (HEX20 / HEX27 need ~75 KB of shared memory per element)
__host__ __device__ constexpr int fg_threads(int enon, int ng) { return fg_per_el(enon, ng) * (32 / fg_lpe(enon, ng)) > 6000 ? 32 : FG_THREADS; }

// NG = number of Gauss points: NG == ENON for TET4 / HEX8 / WDG, 15 for TET10, 27 for HEX20 / HEX27 (nn_elem_props.h).
template <int ENON, int NG, bool ATOMIC>
__global__ void __launch_bounds__(fg_threads(ENON, NG))
assemble_fluid_gen_kernel(const __grid_constant__ FluidGenArgs P)
{
  constexpr int LPE = fg_lpe(ENON, NG);
  constexpr int EPW = 32 / LPE;
  constexpr int PER_EL = fg_per_el(ENON, NG);
  constexpr int TLD = fg_tab_ld(ENON);
  constexpr int THREADS = fg_threads(ENON, NG);
  // reference-element tables: staged in shared memory when small (TET4 .. TET10), read in place (L2-resident, every CTA reads the
  // same 58 KB) for the 27-point rules, where they would cost a resident CTA
  constexpr bool TAB_SMEM = fg_tab_in_smem(ENON, NG);
  extern __shared__ double sm[];
  const double* stab = TAB_SMEM ? sm : P.tab;
  if (TAB_SMEM) {
    // all loads of a thread in flight at once (a plain copy loop waits for each L2 round trip before the next one: ten dependent
    // round trips per thread at 6 warps per SM were 13 % of the HEX8 kernel's samples, profiles/r2al_ncu_fluid_hex8.txt)
    constexpr int NT = (NG * TLD + THREADS - 1) / THREADS;
    double tmp[NT];
#pragma unroll
    for (int k = 0; k < NT; k++) {
      const int t = threadIdx.x + k * THREADS;
      tmp[k] = (t < NG * TLD) ? __ldg(P.tab + t) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < NT; k++) {
      const int t = threadIdx.x + k * THREADS;
      if (t < NG * TLD) sm[t] = tmp[k];
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a = lane % LPE, el = lane / LPE;
  double* se = sm + (TAB_SMEM ? NG * TLD : 0) + (size_t)(warp * EPW + el) * PER_EL;
  double(*sx)[3] = reinterpret_cast<double(*)[3]>(se);
  double(*sal)[3] = reinterpret_cast<double(*)[3]>(se + 3 * ENON);
  double(*syl)[4] = reinterpret_cast<double(*)[4]>(se + 6 * ENON);
  double(*sbf)[3] = reinterpret_cast<double(*)[3]>(se + 10 * ENON);
  double(*sym)[3] = reinterpret_cast<double(*)[3]>(se + 13 * ENON);
  double(*sNxxL)[6] = reinterpret_cast<double(*)[6]>(se + 16 * ENON);
  FluidGP* sgp = reinterpret_cast<FluidGP*>(se + 22 * ENON);
  constexpr int NLD = fg_nd_ld(ENON);
  double* sndd = se + 22 * ENON + NG * FLUID_GP_DOUBLES;        // FluidNode tables, one per Gauss point, stride NLD
  auto snd = [&](int g) { return reinterpret_cast<FluidNodeC*>(sndd + g * NLD); };

  const long long idx = (long long)P.e0 + ((long long)blockIdx.x * (THREADS / 32) + warp) * EPW + el;
  bool active = (lane < EPW * LPE) && idx < P.e1;
  int e = 0;
  if (active) e = P.perm ? P.perm[idx] : (int)idx;
  int iD = 0;
  if (active) {
    for (int d = 0; d < P.nDmn; d++) {
      iD = d;
      if (P.dmn[d].Id == -1) break;
      if (P.eId != nullptr && ((P.eId[e] >> P.dmn[d].Id) & 1)) break;
    }
    if (!P.dmn[iD].isFluid) active = false;
  }
  const FluidDmn& dm = P.dmn[iD];
  int node = 0;
  if (active && a < ENON) {
    node = P.IEN[(size_t)e * ENON + a];
    const size_t n = (size_t)node;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      sx[a][i] = __ldg(P.x + 3 * n + i) + (P.ale ? __ldg(P.Dg + (size_t)P.tDof * n + 4 + i) : 0.0);
      sal[a][i] = __ldg(P.Ag + (size_t)P.tDof * n + i);
      sbf[a][i] = __ldg(P.Bf + 3 * n + i);
      sym[a][i] = P.mvMsh ? __ldg(P.Yg + (size_t)P.tDof * n + 4 + i) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) syl[a][i] = __ldg(P.Yg + (size_t)P.tDof * n + i);
  }
  __syncthreads();      // tables + nodal inputs

  // ---- phase A ------------------------------------------------------------------------------------------------
  {
    const int g = a;
    const double* tg = stab + g * TLD;
    const double* tgd = P.lShpF ? stab : tg;       // derivative tables: Gauss point 0 for the lShpF elements
    const double(*Nxi)[3] = reinterpret_cast<const double(*)[3]>(tgd + 1 + ENON);
    const double(*Nxi2)[6] = reinterpret_cast<const double(*)[6]>(tgd + 1 + 4 * ENON);
    double Nx[ENON][3], Nxx[ENON][6], xiX[3][3], ks[3][3];
    double Jac = 1.0;
    const bool gp = active && g < NG;
    if (gp) {
      Jac = gnn3_full<ENON>(Nxi, sx, Nx, xiX, ks);
      if (is_zero(Jac)) atomicMax(P.err, e + 1);
      gn_nxx3<ENON>(Nxi2, sx, xiX, Nx, Nxx);
      if (g == NG - 1) {
#pragma unroll
        for (int b = 0; b < ENON; b++)
#pragma unroll
          for (int v = 0; v < 6; v++) sNxxL[b][v] = Nxx[b][v];
      }
    }
    __syncwarp();
    // URIS valves: resistance factor and valve-velocity term at this Gauss point (fluid.cpp:622-672)
    double uF = 0.0, uV[3] = {0.0, 0.0, 0.0};
    if (gp && P.uris != nullptr) {
      int nodes[ENON];
#pragma unroll
      for (int b = 0; b < ENON; b++) nodes[b] = __ldg(P.IEN + (size_t)e * ENON + b);
      uris_factor<ENON>(P.uris, P.nUris, P.urisP, tg + 1, nodes, uF, uV);
    }
    if (gp)
      fluid_gen_gauss_point<ENON>(dm, P.dt, P.af, P.am, P.gam, tg[0] * Jac, ks, tg + 1, Nx, Nxx, sNxxL, sal, syl, sbf,
                                  P.mvMsh ? sym : nullptr, sgp[g], snd(g), uF, uV);
  }
  __syncwarp();
  // HEX8: the 4 x 4 blocks leave through a transposition tile (below), which needs every lane of the warp — lanes without an element
  // stay and move other lanes' entries
  constexpr bool TILE = (ENON == 8 && NG == 8);
  if (TILE) {
    if (__ballot_sync(0xffffffffu, active) == 0u) return;
  } else if (!active || a >= ENON) return;
  const int NGa = active ? NG : 0;

  // ---- phase B ------------------------------------------------------------------------------------------------
  double lR[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
  for (int g = 0; g < NGa; g++) fluid_gen_residual(sgp[g], fluid_node_expand(sgp[g], snd(g)[a]), lR);
  if (active)
#pragma unroll
    for (int i = 0; i < 4; i++) fg_add<ATOMIC>(P.R + 4 * (size_t)node + i, lR[i]);
  const int* sl = P.slot + (size_t)e * ENON * ENON;
  constexpr int NB = (ENON % FG_NB == 0) ? FG_NB : (ENON % 2 == 0 ? 2 : 1);
#pragma unroll 1
  for (int b0 = 0; b0 < ENON; b0 += NB) {
    // CSR slots of the NB blocks, requested before the Gauss loop so that the load latency hides behind the FMAs
    int slots[NB];
#pragma unroll
    for (int bb = 0; bb < NB; bb++) slots[bb] = active ? __ldg(sl + a * ENON + b0 + bb) : -1;
    double K[NB][16];
#pragma unroll
    for (int bb = 0; bb < NB; bb++)
#pragma unroll
      for (int i = 0; i < 16; i++) K[bb][i] = 0.0;
#pragma unroll 1
    for (int g = 0; g < NGa; g++) {
      const FluidNodeC* nd = snd(g);
      FluidRow row;
      fluid_gen_row(sgp[g], fluid_node_expand(sgp[g], nd[a]), row);
#pragma unroll
      for (int bb = 0; bb < NB; bb++) fluid_gen_block_row(row, fluid_node_expand(sgp[g], nd[b0 + bb]), K[bb]);
    }
    if (TILE) {
      // one block = 16 contiguous doubles = one 128-byte line: lane (el, a) parks its block in the dead nodal-input area of its element
      // (row stride 17), then each half-warp adds one block per step with consecutive lanes on consecutive doubles: 4 L2 sectors per
      // block instead of 16
      double* trow = se + a * 17;
      const double* wbase = sm + (TAB_SMEM ? NG * TLD : 0) + (size_t)(warp * EPW) * PER_EL;
#pragma unroll
      for (int bb = 0; bb < NB; bb++) {
        __syncwarp();
        if (active)
#pragma unroll
          for (int i = 0; i < 16; i++) trow[i] = K[bb][i];
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 16; it++) {
          const int src = 2 * it + (lane >> 4), i = lane & 15;
          const int s_ = __shfl_sync(0xffffffffu, slots[bb], src);
          if (s_ >= 0) fg_add<ATOMIC>(P.Val + 16 * (size_t)s_ + i, wbase[(size_t)(src >> 3) * PER_EL + (src & 7) * 17 + i]);
        }
      }
    } else {
#pragma unroll
      for (int bb = 0; bb < NB; bb++) {
        double* v = P.Val + 16 * (size_t)slots[bb];
#pragma unroll
        for (int i = 0; i < 16; i++) fg_add<ATOMIC>(v + i, K[bb][i]);
      }
    }
  }
}

template <int ENON, int NG>
static int launch_gen(svb200_ctx* ctx, const FluidGenArgs& A)
{
  constexpr int THREADS = fg_threads(ENON, NG);
  constexpr int EPB = (THREADS / 32) * (32 / fg_lpe(ENON, NG));
  const long long n = (long long)A.e1 - A.e0;
  if (n <= 0) return SVB200_OK;
  const unsigned blocks = (unsigned)((n + EPB - 1) / EPB);
  constexpr size_t smem = sizeof(double) * ((fg_tab_in_smem(ENON, NG) ? (size_t)NG * fg_tab_ld(ENON) : 0) + (size_t)EPB * fg_per_el(ENON, NG));
  static_assert(smem <= 227 * 1024, "element does not fit in shared memory");
  static bool configured = false;
  if (!configured) {
    SVB_CUDA(cudaFuncSetAttribute(assemble_fluid_gen_kernel<ENON, NG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SVB_CUDA(cudaFuncSetAttribute(assemble_fluid_gen_kernel<ENON, NG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  if (A.atomic) assemble_fluid_gen_kernel<ENON, NG, true><<<blocks, THREADS, smem, ctx->stream>>>(A);
  else assemble_fluid_gen_kernel<ENON, NG, false><<<blocks, THREADS, smem, ctx->stream>>>(A);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// Device copy of the reference-element tables in the layout the kernel stages into shared memory.
int upload_fluid_gen_tables(svb200_ctx* ctx, Mesh& m)
{
  const int E = m.eNoN, G = m.nG, LD = fg_tab_ld(E);
  std::vector<double> t((size_t)G * LD, 0.0);
  for (int g = 0; g < G; g++) {
    double* p = t.data() + (size_t)g * LD;
    p[0] = m.w[g];
    for (int a = 0; a < E; a++) {
      p[1 + a] = m.N[(size_t)g * E + a];
      for (int k = 0; k < 3; k++) p[1 + E + 3 * a + k] = m.Nx[((size_t)g * E + a) * 3 + k];
      for (int k = 0; k < 6; k++) p[1 + 4 * E + 6 * a + k] = m.Nxx.empty() ? 0.0 : m.Nxx[((size_t)g * E + a) * 6 + k];
    }
  }
  if (m.d_gtab) cudaFree(m.d_gtab);
  m.d_gtab = nullptr;
  SVB_CUDA(cudaMalloc(&m.d_gtab, sizeof(double) * t.size()));
  SVB_CUDA(cudaMemcpyAsync(m.d_gtab, t.data(), sizeof(double) * t.size(), cudaMemcpyHostToDevice, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SVB200_OK;
}

// URIS split launch.  An element whose nodes all lie outside every valve's (and scaffold's) thickness has a zero valve factor at
// every Gauss point (dist = sum_a N_a |sdf_a| >= min_a |sdf_a| >= deps for the non-negative shape functions of a linear element), so
// the closed-form TET4 kernel is exact for it; the band around the valves goes through the per-Gauss-point kernel.
__global__ void uris_element_mask_kernel(int nEl, int eNoN, const int* __restrict__ IEN, const double* __restrict__ nodal, int nUris,
                                         const __grid_constant__ FluidGenArgs P, unsigned char* __restrict__ mask)
{
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nEl) return;
  bool near = false;
  for (int a = 0; a < eNoN; a++) {
    const double* r = nodal + (size_t)IEN[(size_t)e * eNoN + a] * nUris * 5;
    for (int v = 0; v < nUris; v++) {
      near |= r[5 * v] < P.urisP[v].sdf_deps;
      if (P.urisP[v].scaffold) near |= r[5 * v + 1] < P.urisP[v].scaffold_deps;
    }
  }
  mask[e] = near ? 1 : 0;
}

int build_uris_element_mask(svb200_ctx* ctx, const Mesh& m)
{
  if (m.uris_version == ctx->uris_version && m.d_uris_mask) return SVB200_OK;
  if (!m.d_uris_mask) {
    SVB_CUDA(cudaMalloc(&m.d_uris_mask, std::max(m.nEl, 1)));
    SVB_CUDA(cudaMalloc(&m.d_uris_list, sizeof(int) * std::max(m.nEl, 1)));
  }
  FluidGenArgs A;
  memset(&A, 0, sizeof(A));
  for (int v = 0; v < ctx->nUris; v++) A.urisP[v] = ctx->urisP[v];
  uris_element_mask_kernel<<<(m.nEl + 255) / 256, 256, 0, ctx->stream>>>(m.nEl, m.eNoN, m.d_IEN, ctx->d_uris, ctx->nUris, A, m.d_uris_mask);
  // compact list in element order (deterministic): cub select over a counting iterator
  int* d_n = nullptr;
  SVB_CUDA(cudaMalloc(&d_n, sizeof(int)));
  size_t tmp = 0;
  cub::CountingInputIterator<int> it(0);
  SVB_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp, it, m.d_uris_mask, m.d_uris_list, d_n, m.nEl, ctx->stream));
  void* d_tmp = nullptr;
  SVB_CUDA(cudaMalloc(&d_tmp, std::max<size_t>(tmp, 1)));
  cudaError_t ce = cub::DeviceSelect::Flagged(d_tmp, tmp, it, m.d_uris_mask, m.d_uris_list, d_n, m.nEl, ctx->stream);
  int n = 0;
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(&n, d_n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_tmp); cudaFree(d_n);
  SVB_CUDA(ce);
  ctx->launches += 2;
  m.n_uris_el = n;
  m.uris_version = ctx->uris_version;
  return SVB200_OK;
}

int run_assemble_fluid_gen(svb200_ctx* ctx, const Mesh& m, const FluidArgs& F)
{
  // element types of nn_elem_props.h with their quadrature rules: TET4 (4), HEX8 (8), WDG (6), TET10 (15), HEX20 / HEX27 (27)
  const int key = m.eNoN * 100 + m.nG;
  SVB_REQUIRE(key == 404 || key == 808 || key == 606 || key == 1015 || key == 2027 || key == 2727,
              "svb200_assemble: the general fluid kernel covers TET4, HEX8, WDG, TET10, HEX20 and HEX27 meshes with the reference's quadrature rules");
  SVB_REQUIRE(m.d_gtab, "svb200_assemble: element tables missing");
  if (m.eNoN != 4 && m.Nxx.empty()) {
    set_error("svb200_assemble: a fluid mesh of non-linear elements needs the second-derivative table (svb200_set_mesh_nxx)");
    return SVB200_ERR_INVALID;
  }
  FluidGenArgs A;
  memset(&A, 0, sizeof(A));
  A.IEN = F.IEN; A.eId = F.eId; A.slot = F.slot; A.perm = nullptr;
  A.x = F.x; A.Ag = F.Ag; A.Yg = F.Yg; A.Bf = F.Bf; A.Dg = F.Dg; A.tab = m.d_gtab; A.R = F.R; A.Val = F.Val; A.err = F.err;
  A.e0 = 0; A.e1 = m.nEl;
  if (F.emask != nullptr) {            // URIS split launch: only the band around the valves (compact list, atomic mode)
    A.perm = m.d_uris_list; A.e1 = m.n_uris_el;
  }
  A.tDof = F.tDof; A.mvMsh = F.mvMsh; A.nDmn = F.nDmn; A.atomic = F.atomic; A.ale = F.ale;
  A.lShpF = (m.eNoN == 4 || m.eNoN == 6) ? 1 : 0;
  A.dt = F.dt; A.af = F.af; A.am = F.am; A.gam = F.gam;
  for (int d = 0; d < MAX_DMN; d++) A.dmn[d] = F.dmn[d];
  A.uris = F.uris; A.nUris = F.nUris;
  for (int v = 0; v < F.nUris; v++) A.urisP[v] = F.urisP[v];
  auto launch = [&](const FluidGenArgs& B) {
    switch (key) {
      case 404: return launch_gen<4, 4>(ctx, B);
      case 808: return launch_gen<8, 8>(ctx, B);
      case 606: return launch_gen<6, 6>(ctx, B);
      case 1015: return launch_gen<10, 15>(ctx, B);
      case 2027: return launch_gen<20, 27>(ctx, B);
      default: return launch_gen<27, 27>(ctx, B);
    }
  };
  if (A.atomic) return launch(A);
  A.perm = m.d_color_perm;
  for (size_t c = 0; c + 1 < m.color_off.size(); c++) {
    A.e0 = m.color_off[c];
    A.e1 = m.color_off[c + 1];
    int rc = launch(A);
    if (rc) return rc;
  }
  return SVB200_OK;
}

}  // namespace svb
