// assemble_heat.cu — element loop + scatter of the scalar heat equations.
//
// Replaces heats::construct_heats / heats_3d (Code/Source/solver/heats.cpp:30-119, 186-233) and
// heatf::construct_heatf / heatf_3d (Code/Source/solver/heatf.cpp:52-139, 238-331) with the do_assem scatter
// (Code/Source/solver/lhsa.cpp:70-114) for TET4 and HEX8 meshes; dof = 1, so the CSR "blocks" are single doubles.
//
// Mapping: ENON lanes per element, lane a owns row a of the element matrix (ENON doubles in registers) and one
// residual entry; the per-Gauss-point scalars (grad T, u, tauM, the discontinuity-capturing conductivity) are
// evaluated by every lane of the element (a few hundred flops; the kernel is bound by the gather and the scatter:
// 20 B/node of state in, ENON^2 + ENON adds out).
#include <cstdlib>
#include "svb200_internal.h"
#include "heat_elem.cuh"

namespace svb {

struct HeatArgs {
  const int* IEN;
  const int* eId;
  const int* slot;
  const int* perm;
  const double* x;
  const double* Ag;
  const double* Yg;
  double* R;
  double* Val;
  int* err;
  int e0, e1;
  int tDof, s, nDmn, nG, mvMsh, pad;
  double dt, af, am, gam;
  double w[MAX_NG];
  double N[MAX_NG][MAX_ENON];
  double Nxi[MAX_NG][MAX_ENON][3];
  HeatDmn dmn[MAX_DMN];
  const double* tab;     // per-mesh device copy of the element tables (w | N | Nxi | Nxi2 per Gauss point) for TET10 / HEX20 / HEX27
};

template <bool ATOMIC>
__device__ __forceinline__ void heat_add(double* p, double v)
{
  if (ATOMIC) asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
  else *p += v;
}

template <int ENON, bool ATOMIC, bool FLUID>
__global__ void __launch_bounds__(128)
assemble_heat_kernel(const __grid_constant__ HeatArgs P)
{
  constexpr int EPW = 32 / ENON;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a = lane % ENON, el = lane / ENON;
  constexpr bool TAB = ENON > MAX_ENON;
  constexpr int TLD = 1 + 10 * ENON;
  if (el >= EPW) return;             // 32 is not a multiple of ENON (WDG, TET10, HEX20, HEX27): the last lanes of the warp idle
  const long long idx = (long long)P.e0 + ((long long)blockIdx.x * 4 + warp) * EPW + el;
  if (idx >= P.e1) return;
  const int e = P.perm ? P.perm[idx] : (int)idx;
  int iD = 0;
  for (int d = 0; d < P.nDmn; d++) {
    iD = d;
    if (P.dmn[d].Id == -1) break;
    if (P.eId != nullptr && ((P.eId[e] >> P.dmn[d].Id) & 1)) break;
  }
  const HeatDmn& dm = P.dmn[iD];
  if (!dm.active) return;
  int node[ENON];
  double xl[ENON][3], Tl[ENON], Tdl[ENON], ul[ENON][3];
#pragma unroll
  for (int b = 0; b < ENON; b++) {
    node[b] = P.IEN[(size_t)e * ENON + b];
    const size_t n = (size_t)node[b];
    const double* y = P.Yg + (size_t)P.tDof * n;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      xl[b][i] = __ldg(P.x + 3 * n + i);
      // heatf_3d: convective velocity from state dofs 0..2, minus the mesh velocity of dofs 4..6 when mvMsh (heatf.cpp:276-291)
      ul[b][i] = FLUID ? (__ldg(y + i) - (P.mvMsh ? __ldg(y + 4 + i) : 0.0)) : 0.0;
    }
    Tl[b] = __ldg(y + P.s);
    Tdl[b] = __ldg(P.Ag + (size_t)P.tDof * n + P.s);
  }
  const double T1 = P.af * P.gam * P.dt;
  double lK[ENON], lR = 0.0;
#pragma unroll
  for (int b = 0; b < ENON; b++) lK[b] = 0.0;
  double Nx[ENON][3], ks[3][3], Jac = 1.0;
#pragma unroll 1
  for (int g = 0; g < P.nG; g++) {
    const double* tg = TAB ? P.tab + (size_t)g * TLD : nullptr;
    const double* Ng = TAB ? tg + 1 : P.N[g];
    if (g == 0 || (ENON != 4 && ENON != 6)) {        // TET4 / WDG: lShpF, one gnn per element (heats.cpp:89, heatf.cpp:112)
      Jac = gnn3_metric<ENON>(TAB ? reinterpret_cast<const double(*)[3]>(tg + 1 + ENON) : P.Nxi[g], xl, Nx, ks);
      if (fabs(Jac) < 10.0 * 2.220446049250313e-16 * 2.220446049250313e-16) { if (a == 0) atomicMax(P.err, e + 1); return; }
    }
    const double w = (TAB ? tg[0] : P.w[g]) * Jac;
    HeatGP q;
    heat_gauss_point<ENON, FLUID>(dm, P.dt, P.af, P.am, P.gam, Ng, Nx, ks, Tl, Tdl, ul, q);
    double Na = 0.0, Nxa[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int b = 0; b < ENON; b++)
      if (b == a) { Na = Ng[b]; Nxa[0] = Nx[b][0]; Nxa[1] = Nx[b][1]; Nxa[2] = Nx[b][2]; }
    heat_row<ENON>(q, w, w * T1, Na, Nxa, Ng, Nx, lR, lK);
  }
  int na = 0;
#pragma unroll
  for (int b = 0; b < ENON; b++)
    if (b == a) na = node[b];
  heat_add<ATOMIC>(P.R + na, lR);
  const int* sl = P.slot + (size_t)e * ENON * ENON + a * ENON;
#pragma unroll
  for (int b = 0; b < ENON; b++) heat_add<ATOMIC>(P.Val + sl[b], lK[b]);
}

// HEX8: one lane per Gauss point in phase A.  In the kernel above every lane of an element evaluates gnn and the Gauss-point scalars
// of ALL Gauss points (8 x redundant for a hexahedron: 22 of the 30 kflop per element); here lane g evaluates Gauss point g once and
// leaves its gradients Nx_g(8,3), the HeatGP record and the weight in shared memory (37 doubles, odd stride: conflict-free stores),
// then lane a accumulates row a over the 8 records (broadcast reads: the lanes of an element read the same words).
constexpr int HEAT_GP_LD = 37;
template <bool ATOMIC, bool FLUID>
__global__ void __launch_bounds__(128)
assemble_heat_hex8_kernel(const __grid_constant__ HeatArgs P)
{
  constexpr int ENON = 8, EPW = 4;
  __shared__ double sgp[4][EPW][ENON][HEAT_GP_LD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a = lane % ENON, el = lane / ENON;
  const long long idx = (long long)P.e0 + ((long long)blockIdx.x * 4 + warp) * EPW + el;
  bool active = idx < P.e1;
  int e = 0, iD = 0;
  if (active) {
    e = P.perm ? P.perm[idx] : (int)idx;
    for (int d = 0; d < P.nDmn; d++) {
      iD = d;
      if (P.dmn[d].Id == -1) break;
      if (P.eId != nullptr && ((P.eId[e] >> P.dmn[d].Id) & 1)) break;
    }
    active = P.dmn[iD].active != 0;
  }
  const HeatDmn& dm = P.dmn[iD];
  int na = 0;
  if (active) {
    // ---- phase A: lane g = a evaluates Gauss point g ----
    const int g = a;
    double xl[ENON][3], Tl[ENON], Tdl[ENON], ul[ENON][3];
#pragma unroll
    for (int b = 0; b < ENON; b++) {
      const int nb = P.IEN[(size_t)e * ENON + b];
      if (b == a) na = nb;
      const size_t n = (size_t)nb;
      const double* y = P.Yg + (size_t)P.tDof * n;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        xl[b][i] = __ldg(P.x + 3 * n + i);
        ul[b][i] = FLUID ? (__ldg(y + i) - (P.mvMsh ? __ldg(y + 4 + i) : 0.0)) : 0.0;
      }
      Tl[b] = __ldg(y + P.s);
      Tdl[b] = __ldg(P.Ag + (size_t)P.tDof * n + P.s);
    }
    double Nx[ENON][3], ks[3][3];
    const double Jac = gnn3_metric<ENON>(P.Nxi[g], xl, Nx, ks);
    if (fabs(Jac) < 10.0 * 2.220446049250313e-16 * 2.220446049250313e-16) atomicMax(P.err, e + 1);
    HeatGP q;
    heat_gauss_point<ENON, FLUID>(dm, P.dt, P.af, P.am, P.gam, P.N[g], Nx, ks, Tl, Tdl, ul, q);
    double* r = sgp[warp][el][g];
#pragma unroll
    for (int b = 0; b < ENON; b++)
#pragma unroll
      for (int i = 0; i < 3; i++) r[3 * b + i] = Nx[b][i];
    r[24] = q.c0; r[25] = q.nu; r[26] = q.Tp; r[27] = q.tauM; r[28] = q.amd;
    r[29] = q.Tx[0]; r[30] = q.Tx[1]; r[31] = q.Tx[2]; r[32] = q.u[0]; r[33] = q.u[1]; r[34] = q.u[2];
    r[35] = P.w[g] * Jac;
  }
  __syncwarp();
  if (!active) return;
  // ---- phase B: lane a = row a ----
  const double T1 = P.af * P.gam * P.dt;
  double lK[ENON], lR = 0.0;
#pragma unroll
  for (int b = 0; b < ENON; b++) lK[b] = 0.0;
#pragma unroll 1
  for (int g = 0; g < ENON; g++) {
    const double* r = sgp[warp][el][g];
    HeatGP q;
    q.c0 = r[24]; q.nu = r[25]; q.Tp = r[26]; q.tauM = r[27]; q.amd = r[28];
    q.Tx[0] = r[29]; q.Tx[1] = r[30]; q.Tx[2] = r[31]; q.u[0] = r[32]; q.u[1] = r[33]; q.u[2] = r[34];
    const double w = r[35];
    const double Nxa[3] = {r[3 * a], r[3 * a + 1], r[3 * a + 2]};
    heat_row<ENON>(q, w, w * T1, P.N[g][a], Nxa, P.N[g], reinterpret_cast<const double(*)[3]>(r), lR, lK);
  }
  heat_add<ATOMIC>(P.R + na, lR);
  const int* sl = P.slot + (size_t)e * ENON * ENON + a * ENON;
#pragma unroll
  for (int b = 0; b < ENON; b++) heat_add<ATOMIC>(P.Val + sl[b], lK[b]);
}

template <int ENON, bool FLUID>
static int launch_heat(svb200_ctx* ctx, const HeatArgs& A, bool atomic)
{
  constexpr int EPB = 4 * (32 / ENON);
  const long long n = (long long)A.e1 - A.e0;
  if (n <= 0) return SVB200_OK;
  const unsigned blocks = (unsigned)((n + EPB - 1) / EPB);
  static const bool lane_rows = getenv("SVB200_HEAT_HEX8_LEGACY") != nullptr;      // A/B: the lane-per-row kernel for HEX8
  if (ENON == 8 && !lane_rows) {
    if (atomic) assemble_heat_hex8_kernel<true, FLUID><<<blocks, 128, 0, ctx->stream>>>(A);
    else assemble_heat_hex8_kernel<false, FLUID><<<blocks, 128, 0, ctx->stream>>>(A);
  } else if (atomic) assemble_heat_kernel<ENON, true, FLUID><<<blocks, 128, 0, ctx->stream>>>(A);
  else assemble_heat_kernel<ENON, false, FLUID><<<blocks, 128, 0, ctx->stream>>>(A);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

int run_assemble_heat(svb200_ctx* ctx, const Mesh& m, const svb200_eqparams* eq, const svb200_dmnparams* dmn, int nDmn)
{
  const bool fluid = (eq->phys == SVB200_PHYS_HEATF);
  SVB_REQUIRE(nDmn >= 1 && nDmn <= MAX_DMN, "svb200_assemble: between 1 and 8 domains are supported");
  SVB_REQUIRE(eq->dof == 1 && ctx->dof == 1, "svb200_assemble: the heat equations have dof = 1 (call svb200_alloc(1))");
  SVB_REQUIRE(eq->tDof == ctx->tDof && ctx->d_Yg && ctx->d_Ag, "svb200_assemble: state not set or tDof mismatch");
  SVB_REQUIRE(eq->s >= 0 && eq->s < eq->tDof, "svb200_assemble: eq.s out of range");
  SVB_REQUIRE(!fluid || eq->tDof >= 4, "svb200_assemble: heatF reads the fluid velocity from state dofs 0..2 (tDof >= 4)");
  SVB_REQUIRE(!fluid || !eq->mvMsh || eq->tDof >= 7, "svb200_assemble: mvMsh needs the mesh velocity in state dofs 4..6");
  SVB_REQUIRE(ctx->d_x, "svb200_assemble: coordinates not set");
  const int key = m.eNoN * 100 + m.nG;
  SVB_REQUIRE(key == 404 || key == 808 || key == 606 || key == 1015 || key == 2027 || key == 2727,
              "svb200_assemble: the heat equations cover TET4, HEX8, WDG, TET10, HEX20 and HEX27 meshes with the reference's quadrature rules");
  SVB_REQUIRE(m.eNoN <= MAX_ENON || m.d_gtab, "svb200_assemble: element tables missing");
  HeatArgs A;
  memset(&A, 0, sizeof(A));
  A.IEN = m.d_IEN; A.eId = m.d_eId; A.slot = m.d_slot; A.perm = nullptr;
  A.x = ctx->d_x; A.Ag = ctx->d_Ag; A.Yg = ctx->d_Yg; A.R = ctx->d_R; A.Val = ctx->d_Val;
  if (!ctx->d_err) {
    SVB_CUDA(cudaMalloc(&ctx->d_err, sizeof(int)));
    SVB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
  }
  A.err = ctx->d_err;
  A.e0 = 0; A.e1 = m.nEl;
  A.tDof = eq->tDof; A.s = eq->s; A.nDmn = nDmn; A.nG = m.nG; A.mvMsh = eq->mvMsh;
  A.dt = eq->dt; A.af = eq->af; A.am = eq->am; A.gam = eq->gam;
  A.tab = m.d_gtab;
  if (m.nG <= MAX_NG && m.eNoN <= MAX_ENON)
    for (int g = 0; g < m.nG; g++) {
      A.w[g] = m.w[g];
      for (int a = 0; a < m.eNoN; a++) {
        A.N[g][a] = m.N[(size_t)g * m.eNoN + a];
        for (int k = 0; k < 3; k++) A.Nxi[g][a][k] = m.Nx[((size_t)g * m.eNoN + a) * 3 + k];
      }
    }
  bool whole = false;
  for (int d = 0; d < nDmn; d++) {
    HeatDmn& o = A.dmn[d];
    o.rho = dmn[d].rho; o.nu = dmn[d].conductivity; o.s = dmn[d].source_term;
    o.Id = dmn[d].Id;
    o.active = (dmn[d].phys == eq->phys);
    SVB_REQUIRE(o.Id >= -1 && o.Id < 31, "svb200_assemble: domain Id out of range");
    whole |= (o.Id == -1);
  }
  if (!whole && !m.d_eId) { set_error("eId is not allocated"); return SVB200_ERR_INVALID; }
  const bool atomic = (eq->scatter == SVB200_SCATTER_ATOMIC);
  auto launch = [&](const HeatArgs& B) {
    switch (m.eNoN) {
      case 8: return fluid ? launch_heat<8, true>(ctx, B, atomic) : launch_heat<8, false>(ctx, B, atomic);
      case 6: return fluid ? launch_heat<6, true>(ctx, B, atomic) : launch_heat<6, false>(ctx, B, atomic);
      case 10: return fluid ? launch_heat<10, true>(ctx, B, atomic) : launch_heat<10, false>(ctx, B, atomic);
      case 20: return fluid ? launch_heat<20, true>(ctx, B, atomic) : launch_heat<20, false>(ctx, B, atomic);
      case 27: return fluid ? launch_heat<27, true>(ctx, B, atomic) : launch_heat<27, false>(ctx, B, atomic);
      default: return fluid ? launch_heat<4, true>(ctx, B, atomic) : launch_heat<4, false>(ctx, B, atomic);
    }
  };
  int rc = SVB200_OK;
  if (atomic) rc = launch(A);
  else {
    A.perm = m.d_color_perm;
    for (size_t c = 0; c + 1 < m.color_off.size() && rc == SVB200_OK; c++) {
      A.e0 = m.color_off[c];
      A.e1 = m.color_off[c + 1];
      rc = launch(A);
    }
  }
  if (rc) return rc;
  // construct_heats / construct_heatf throw when utils::is_zero(Jac) (heats.cpp:96-98, heatf.cpp:116-118)
  int e = 0;
  SVB_CUDA(cudaMemcpyAsync(&e, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (e != 0) {
    SVB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
    set_error(std::string(fluid ? "[construct_heatf]" : "[construct_heats]") + " Jacobian for element " + std::to_string(e - 1) + " is < 0.");
    return SVB200_ERR_NUMERIC;
  }
  return SVB200_OK;
}

}  // namespace svb
