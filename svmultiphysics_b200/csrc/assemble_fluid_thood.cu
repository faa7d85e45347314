// assemble_fluid_thood.cu — Navier-Stokes element loop + scatter on Taylor-Hood meshes (mshType::nFs = 2: P2-P1 tetrahedra TET10 / TET4,
// Q2-Q1 hexahedra HEX27 / HEX8), the vmsStab = false branch of fluid::construct_fluid (Code/Source/solver/fluid.cpp:494-500, 596-748)
// with the function spaces of fs::get_thood_fs (solver/fs.cpp:73-178), and fs::thood_val_rc (fs.cpp:394-466).  Algebra: fluid_thood.cuh.
//
// The fluid elements of fsi::construct_fsi on such a mesh (fsi.cpp:84-88, 170-216, 300-316) are the same loops on the moved geometry
// x + Dg(4..6) with K_darcy = 0 (ale = 1).
//
// Mapping (that of assemble_fluid_gen.cu): LPE = max(eNoN, nG1) lanes per element.
//   phase A1  lane g < nG1: Gauss point g of the VELOCITY rule — gnn + gn_nxx of the velocity space, gnn of the pressure space (at Gauss
//             point 0 only when that space is linear, fluid.cpp:620-626), thood_gauss_point_m -> FluidGP + Nq, Nqx + node records;
//   phase A2  lane g < nG2: Gauss point g of the PRESSURE rule — gnn of the velocity space there, div u, and the weight with the Jacobian
//             the reference ends up with: the pressure space's when its gnn ran last (g = 0, or every g for HEX8), else the velocity
//             space's at this point (fluid.cpp:700-720);
//   phase B   lane a < eNoN: row a — momentum rows over the nG1 records, the continuity row (pressure nodes a < eNoNq only) over the nG2
//             records, two column nodes per pass, 16 scalar adds per block.
// Correct first: the tables are read in place (L2-resident), the scatter is entry by entry.
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "fluid_thood.cuh"

namespace svb {

struct ThoodArgs {
  const int* IEN;
  const int* eId;
  const int* slot;
  const int* perm;
  const double* x;
  const double* Ag;
  const double* Yg;
  const double* Bf;
  const double* Dg;       // displacement state: dofs 4..6 = mesh displacement (FSI / ALE geometry, fsi.cpp:140-146)
  const double* tab;      // velocity space, velocity rule: w | N | Nxi | Nxi2 per Gauss point (assemble_fluid_gen.cu layout)
  const double* thtab;    // nG1 x [Nq1 | Nqxi1], then nG2 x [w2 | Nw2 | Nwxi2 | Nq2 | Nqxi2]
  int* err;
  double* R;
  double* Val;
  int e0, e1;
  int tDof, mvMsh, nDmn, atomic, lShpFq, ale;
  double dt, af, am, gam;
  FluidDmn dmn[MAX_DMN];
  const double* uris;     // URIS valves (svb200_set_uris) or null: the momentum loop sees the factor at the velocity rule's points
  int nUris;
  svb200_uris urisP[SVB200_MAX_URIS];
};

template <bool ATOMIC>
__device__ __forceinline__ void th_add(double* p, double v)
{
  if (ATOMIC) asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
  else *p += v;
}

__host__ __device__ constexpr int th_lpe(int enon, int ng) { return enon > ng ? enon : ng; }
__host__ __device__ constexpr int th_rec1(int enon, int enonq) { return FLUID_GP_DOUBLES + 4 * enonq + enon * FLUID_NODEC_DOUBLES + 1; }
__host__ __device__ constexpr int th_rec2(int enon) { return 2 + 3 * enon + 1; }
__host__ __device__ constexpr int th_per_el(int enon, int enonq, int ng1, int ng2)
{
  return 16 * enon + ng1 * th_rec1(enon, enonq) + ng2 * th_rec2(enon);
}
constexpr int TH_THREADS = 64;

template <int ENON, int ENONQ, int NG1, int NG2, bool ATOMIC>
__global__ void __launch_bounds__(TH_THREADS)
assemble_fluid_thood_kernel(const __grid_constant__ ThoodArgs P)
{
  constexpr int LPE = th_lpe(ENON, NG1);
  constexpr int EPW = 32 / LPE;
  constexpr int PER_EL = th_per_el(ENON, ENONQ, NG1, NG2);
  constexpr int TLD = 1 + 10 * ENON;               // velocity table stride
  constexpr int TQ1 = 4 * ENONQ;                   // thtab stride, loop 1
  constexpr int T2 = 1 + 4 * ENON + 4 * ENONQ;     // thtab stride, loop 2
  constexpr int REC1 = th_rec1(ENON, ENONQ), REC2 = th_rec2(ENON);
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a = lane % LPE, el = lane / LPE;
  double* se = sm + (size_t)(warp * EPW + (el < EPW ? el : 0)) * PER_EL;
  double(*sx)[3] = reinterpret_cast<double(*)[3]>(se);
  double(*sal)[3] = reinterpret_cast<double(*)[3]>(se + 3 * ENON);
  double(*syl)[4] = reinterpret_cast<double(*)[4]>(se + 6 * ENON);
  double(*sbf)[3] = reinterpret_cast<double(*)[3]>(se + 10 * ENON);
  double(*sym)[3] = reinterpret_cast<double(*)[3]>(se + 13 * ENON);
  double* srec1 = se + 16 * ENON;
  double* srec2 = srec1 + NG1 * REC1;

  const long long idx = (long long)P.e0 + ((long long)blockIdx.x * (TH_THREADS / 32) + warp) * EPW + el;
  bool active = (el < EPW) && idx < P.e1;
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
  __syncwarp();

  // ---- phase A1: velocity rule ---------------------------------------------------------------------------------
  if (active && a < NG1) {
    const int g = a;
    const double* tg = P.tab + (size_t)g * TLD;
    const double(*Nxi)[3] = reinterpret_cast<const double(*)[3]>(tg + 1 + ENON);
    const double(*Nxi2)[6] = reinterpret_cast<const double(*)[6]>(tg + 1 + 4 * ENON);
    double Nx[ENON][3], Nxx[ENON][6], xiX[3][3], ks[3][3];
    const double Jac = gnn3_full<ENON>(Nxi, sx, Nx, xiX, ks);
    if (is_zero(Jac)) atomicMax(P.err, e + 1);
    gn_nxx3<ENON>(Nxi2, sx, xiX, Nx, Nxx);
    // pressure space at this point: shape functions from the table, gradients by gnn on the first ENONQ nodes
    const double* tq = P.thtab + (size_t)g * TQ1;
    const double* tqd = P.thtab + (size_t)(P.lShpFq ? 0 : g) * TQ1;
    double Nqx[ENONQ][3], xiXq[3][3], ksq[3][3];
    const double Jq = gnn3_full<ENONQ>(reinterpret_cast<const double(*)[3]>(tqd + ENONQ), sx, Nqx, xiXq, ksq);
    if (is_zero(Jq)) atomicMax(P.err, e + 1);
    double* r = srec1 + (size_t)g * REC1;
    FluidGP* q = reinterpret_cast<FluidGP*>(r);
    double* rq = r + FLUID_GP_DOUBLES;
#pragma unroll
    for (int b = 0; b < ENONQ; b++) {
      rq[b] = tq[b];
#pragma unroll
      for (int i = 0; i < 3; i++) rq[ENONQ + 3 * b + i] = Nqx[b][i];
    }
    // URIS valves (fluid.cpp:622-672): with vmsFlag false the continuity loop has no URIS term left (tauM = 0, fluid.cpp:1711-1715)
    double uF = 0.0, uV[3] = {0.0, 0.0, 0.0};
    if (P.uris != nullptr) {
      int nodes[ENON];
#pragma unroll
      for (int b = 0; b < ENON; b++) nodes[b] = __ldg(P.IEN + (size_t)e * ENON + b);
      uris_factor<ENON>(P.uris, P.nUris, P.urisP, tg + 1, nodes, uF, uV);
    }
    thood_gauss_point_m<ENON, ENONQ>(dm, P.dt, P.af, P.am, P.gam, tg[0] * Jac, ks, tg + 1, Nx, Nxx, tq, Nqx, sal, syl, sbf,
                                     P.mvMsh ? sym : nullptr, *q, reinterpret_cast<FluidNodeC*>(rq + 4 * ENONQ), uF, uV);
  }
  // ---- phase A2: pressure rule ----------------------------------------------------------------------------------
  if (active && a < NG2) {
    const int g = a;
    const double* t2 = P.thtab + (size_t)NG1 * TQ1 + (size_t)g * T2;
    const double* t2d = P.thtab + (size_t)NG1 * TQ1 + (size_t)(P.lShpFq ? 0 : g) * T2;
    double Nx[ENON][3], xiX[3][3], ks[3][3];
    const double Jw = gnn3_full<ENON>(reinterpret_cast<const double(*)[3]>(t2 + 1 + ENON), sx, Nx, xiX, ks);
    if (is_zero(Jw)) atomicMax(P.err, e + 1);
    double Nqx[ENONQ][3];
    const double Jq = gnn3_full<ENONQ>(reinterpret_cast<const double(*)[3]>(t2d + 1 + 4 * ENON + ENONQ), sx, Nqx, xiX, ks);
    // fluid.cpp:700-720: Jac is what the LAST gnn left — the pressure space's at g = 0 (and at every g when it is not linear)
    const double Jac = (g == 0 || !P.lShpFq) ? Jq : Jw;
    double divU = 0.0;
#pragma unroll
    for (int b = 0; b < ENON; b++) divU += Nx[b][0] * syl[b][0] + Nx[b][1] * syl[b][1] + Nx[b][2] * syl[b][2];
    double* r = srec2 + (size_t)g * REC2;
    r[0] = t2[0] * Jac;
    r[1] = divU;
#pragma unroll
    for (int b = 0; b < ENON; b++)
#pragma unroll
      for (int i = 0; i < 3; i++) r[2 + 3 * b + i] = Nx[b][i];
  }
  __syncwarp();
  if (!active || a >= ENON) return;

  // ---- phase B -------------------------------------------------------------------------------------------------
  const double T1 = P.af * P.gam * P.dt;
  double lR[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
  for (int g = 0; g < NG1; g++) {
    const double* r = srec1 + (size_t)g * REC1;
    const FluidGP& q = *reinterpret_cast<const FluidGP*>(r);
    const FluidNodeC* nd = reinterpret_cast<const FluidNodeC*>(r + FLUID_GP_DOUBLES + 4 * ENONQ);
    thood_residual_m(q, fluid_node_expand(q, nd[a]), lR);
  }
  if (a < ENONQ) {
#pragma unroll 1
    for (int g = 0; g < NG2; g++) {
      const double* r = srec2 + (size_t)g * REC2;
      const double Nqa = __ldg(P.thtab + (size_t)NG1 * TQ1 + (size_t)g * T2 + 1 + 4 * ENON + a);
      lR[3] += r[0] * (Nqa * r[1]);                 // fluid.cpp:1726-1729 with up = 0
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) th_add<ATOMIC>(P.R + 4 * (size_t)node + i, lR[i]);
  const int* sl = P.slot + (size_t)e * ENON * ENON;
  constexpr int NB = (ENON % 2 == 0) ? 2 : 1;
#pragma unroll 1
  for (int b0 = 0; b0 < ENON; b0 += NB) {
    int slots[NB];
#pragma unroll
    for (int bb = 0; bb < NB; bb++) slots[bb] = __ldg(sl + a * ENON + b0 + bb);
    double K[NB][16];
#pragma unroll
    for (int bb = 0; bb < NB; bb++)
#pragma unroll
      for (int i = 0; i < 16; i++) K[bb][i] = 0.0;
#pragma unroll 1
    for (int g = 0; g < NG1; g++) {
      const double* r = srec1 + (size_t)g * REC1;
      const FluidGP& q = *reinterpret_cast<const FluidGP*>(r);
      const double* rq = r + FLUID_GP_DOUBLES;
      const FluidNodeC* nd = reinterpret_cast<const FluidNodeC*>(rq + 4 * ENONQ);
      FluidRow row;
      thood_row(q, fluid_node_expand(q, nd[a]), row);
#pragma unroll
      for (int bb = 0; bb < NB; bb++) {
        const int b = b0 + bb;
        const bool pb = b < ENONQ;
        thood_block_m(row, fluid_node_expand(q, nd[b]), pb ? rq + b : nullptr, pb ? rq + ENONQ + 3 * b : nullptr, K[bb]);
      }
    }
    if (a < ENONQ) {
#pragma unroll 1
      for (int g = 0; g < NG2; g++) {
        const double* r = srec2 + (size_t)g * REC2;
        const double c = r[0] * T1 * __ldg(P.thtab + (size_t)NG1 * TQ1 + (size_t)g * T2 + 1 + 4 * ENON + a);   // wl Nq_a
#pragma unroll
        for (int bb = 0; bb < NB; bb++)
#pragma unroll
          for (int j = 0; j < 3; j++) K[bb][12 + j] += c * r[2 + 3 * (b0 + bb) + j];      // fluid.cpp:1733-1749 with tauM = 0
      }
    }
#pragma unroll
    for (int bb = 0; bb < NB; bb++) {
      double* v = P.Val + 16 * (size_t)slots[bb];
#pragma unroll
      for (int i = 0; i < 16; i++) th_add<ATOMIC>(v + i, K[bb][i]);
    }
  }
}

template <int ENON, int ENONQ, int NG1, int NG2>
static int launch_thood(svb200_ctx* ctx, const ThoodArgs& A)
{
  constexpr int EPB = (TH_THREADS / 32) * (32 / th_lpe(ENON, NG1));
  const long long n = (long long)A.e1 - A.e0;
  if (n <= 0) return SVB200_OK;
  const unsigned blocks = (unsigned)((n + EPB - 1) / EPB);
  constexpr size_t smem = sizeof(double) * (size_t)EPB * th_per_el(ENON, ENONQ, NG1, NG2);
  static_assert(smem <= 227 * 1024, "element does not fit in shared memory");
  static bool configured = false;
  if (!configured) {
    SVB_CUDA(cudaFuncSetAttribute(assemble_fluid_thood_kernel<ENON, ENONQ, NG1, NG2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SVB_CUDA(cudaFuncSetAttribute(assemble_fluid_thood_kernel<ENON, ENONQ, NG1, NG2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  if (A.atomic) assemble_fluid_thood_kernel<ENON, ENONQ, NG1, NG2, true><<<blocks, TH_THREADS, smem, ctx->stream>>>(A);
  else assemble_fluid_thood_kernel<ENON, ENONQ, NG1, NG2, false><<<blocks, TH_THREADS, smem, ctx->stream>>>(A);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// Device copy of the Taylor-Hood tables (svb200_set_mesh_thood), Fortran-ordered inputs as fs::get_thood_fs holds them.
int upload_thood_tables(svb200_ctx* ctx, Mesh& m, int eNoNq, int nG2, const double* Nq1, const double* Nqxi1, const double* w2,
                        const double* Nw2, const double* Nwxi2, const double* Nq2, const double* Nqxi2)
{
  const int E = m.eNoN, Q = eNoNq, G1 = m.nG, G2 = nG2;
  const int TQ1 = 4 * Q, T2 = 1 + 4 * E + 4 * Q;
  std::vector<double> t((size_t)G1 * TQ1 + (size_t)G2 * T2, 0.0);
  for (int g = 0; g < G1; g++) {
    double* p = t.data() + (size_t)g * TQ1;
    for (int a = 0; a < Q; a++) {
      p[a] = Nq1[(size_t)g * Q + a];
      for (int k = 0; k < 3; k++) p[Q + 3 * a + k] = Nqxi1[((size_t)g * Q + a) * 3 + k];
    }
  }
  for (int g = 0; g < G2; g++) {
    double* p = t.data() + (size_t)G1 * TQ1 + (size_t)g * T2;
    p[0] = w2[g];
    for (int a = 0; a < E; a++) {
      p[1 + a] = Nw2[(size_t)g * E + a];
      for (int k = 0; k < 3; k++) p[1 + E + 3 * a + k] = Nwxi2[((size_t)g * E + a) * 3 + k];
    }
    for (int a = 0; a < Q; a++) {
      p[1 + 4 * E + a] = Nq2[(size_t)g * Q + a];
      for (int k = 0; k < 3; k++) p[1 + 4 * E + Q + 3 * a + k] = Nqxi2[((size_t)g * Q + a) * 3 + k];
    }
  }
  if (m.d_thtab) cudaFree(m.d_thtab);
  m.d_thtab = nullptr;
  SVB_CUDA(cudaMalloc(&m.d_thtab, sizeof(double) * t.size()));
  SVB_CUDA(cudaMemcpyAsync(m.d_thtab, t.data(), sizeof(double) * t.size(), cudaMemcpyHostToDevice, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SVB200_OK;
}

int run_assemble_fluid_thood(svb200_ctx* ctx, const Mesh& m, const FluidArgs& F)
{
  const int key = (m.eNoN * 100 + m.th_eNoNq) * 10000 + m.nG * 100 + m.th_nG2;
  SVB_REQUIRE(key == 10041504 || key == 27082708 || key == 20082708,
              "svb200_assemble: Taylor-Hood fluid elements: TET10 / TET4 (15 + 4 Gauss points), HEX27 / HEX8 and HEX20 / HEX8 (27 + 8)");
  SVB_REQUIRE(m.d_gtab && m.d_thtab && !m.Nxx.empty(), "svb200_assemble: Taylor-Hood mesh needs svb200_set_mesh_nxx and svb200_set_mesh_thood");
  ThoodArgs A;
  memset(&A, 0, sizeof(A));
  A.IEN = F.IEN; A.eId = F.eId; A.slot = F.slot; A.perm = nullptr;
  A.x = F.x; A.Ag = F.Ag; A.Yg = F.Yg; A.Bf = F.Bf; A.Dg = F.Dg; A.ale = F.ale; A.tab = m.d_gtab; A.thtab = m.d_thtab; A.R = F.R; A.Val = F.Val; A.err = F.err;
  A.e0 = 0; A.e1 = m.nEl;
  A.tDof = F.tDof; A.mvMsh = F.mvMsh; A.nDmn = F.nDmn; A.atomic = F.atomic; A.lShpFq = m.th_lShpFq;
  A.dt = F.dt; A.af = F.af; A.am = F.am; A.gam = F.gam;
  for (int d = 0; d < MAX_DMN; d++) A.dmn[d] = F.dmn[d];
  A.uris = F.uris; A.nUris = F.nUris;
  for (int v = 0; v < F.nUris; v++) A.urisP[v] = F.urisP[v];
  auto launch = [&](const ThoodArgs& B) {
    switch (key) {
      case 10041504: return launch_thood<10, 4, 15, 4>(ctx, B);
      case 27082708: return launch_thood<27, 8, 27, 8>(ctx, B);
      default: return launch_thood<20, 8, 27, 8>(ctx, B);
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

// fs::thood_val_rc (fs.cpp:394-466): the pressure dof of a node that is only ever an edge / face / centre node of Taylor-Hood elements
// carries no equation — R(3, a) = 0 and its row of the pressure-pressure entries becomes the identity.
__global__ void thood_mark_kernel(int eNoN, int eNoNq, int nEl, const int* __restrict__ IEN, int* __restrict__ flag)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nh = eNoN - eNoNq;
  if (t >= (long long)nEl * nh) return;
  const int e = (int)(t / nh), a = eNoNq + (int)(t % nh);
  flag[IEN[(size_t)e * eNoN + a]] = 1;
}

__global__ void thood_val_rc_kernel(int nNo, const int* __restrict__ flag, const int* __restrict__ rowPtr, const int* __restrict__ colPtr,
                                    double* __restrict__ R, double* __restrict__ Val)
{
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nNo || !flag[a]) return;
  R[4 * (size_t)a + 3] = 0.0;
  for (int k = rowPtr[a]; k < rowPtr[a + 1]; k++) Val[16 * (size_t)k + 15] = (colPtr[k] == a) ? 1.0 : 0.0;
}

int run_thood_val_rc(svb200_ctx* ctx)
{
  bool any = false;
  for (auto& m : ctx->mesh) any |= (m.set && m.th_eNoNq > 0);
  if (!any || ctx->nNo == 0) return SVB200_OK;
  SVB_REQUIRE(ctx->dof == 4 && ctx->d_R && ctx->d_Val, "svb200_thood_val_rc: a dof = 4 system must be allocated");
  int* flag = nullptr;
  SVB_CUDA(cudaMalloc(&flag, sizeof(int) * (size_t)ctx->nNo));
  SVB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int) * (size_t)ctx->nNo, ctx->stream));
  for (auto& m : ctx->mesh) {
    if (!m.set || m.th_eNoNq <= 0 || m.nEl == 0) continue;
    const long long n = (long long)m.nEl * (m.eNoN - m.th_eNoNq);
    thood_mark_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(m.eNoN, m.th_eNoNq, m.nEl, m.d_IEN, flag);
    ctx->launches++;
  }
  thood_val_rc_kernel<<<(ctx->nNo + 255) / 256, 256, 0, ctx->stream>>>(ctx->nNo, flag, ctx->d_rowPtr, ctx->d_colPtr, ctx->d_R, ctx->d_Val);
  ctx->launches++;
  cudaError_t ce = cudaStreamSynchronize(ctx->stream);
  cudaFree(flag);
  SVB_CUDA(ce);
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

}  // namespace svb
