// assemble_struct.cu — fused element loop + scatter for the displacement-based solid (K2 of SURVEY.md §7).
//
// Replaces struct_ns::construct_dsolid / struct_3d (Code/Source/solver/sv_struct.cpp:184-341, 541-826),
// mat_models::compute_pk2cc (Code/Source/solver/mat_models.cpp:291-817), nn::gnn per Gauss point
// (Code/Source/solver/nn.cpp:862-899) and the do_assem scatter (Code/Source/solver/lhsa.cpp:70-114), for
// TET4 and HEX8 meshes, neo-Hookean / Mooney-Rivlin / Guccione / St.Venant-Kirchhoff with Quad/ST91/M94
// penalties, as a struct equation (dof = 3) or as the solid part of an FSI equation (dof = 4).
//
// Mapping: ENON lanes per element, 32/ENON elements per warp, two phases.  Phase A: lane g evaluates Gauss point g
// (nn::gnn, F, compute_pk2cc) ONCE per element and leaves xiX, w, F, S, Dm, ud in shared memory.  Phase B: lane a
// = element node a; per Gauss point it contracts its Bm_a with Dm and F into H_a(3,3,3) so that a block needs only
// grad N_b of the other node (see phase B); the element matrix is symmetric for a hyperelastic solid
// (K_ab = K_ba^T), so lane a only accumulates the blocks (a, a+k mod ENON), k = 0..ENON/2, in registers (5 x 9
// doubles for HEX8) and scatters each of them twice (as is, and transposed) — 36 of 64 block products instead of 64.
#include <cstdlib>
#include <vector>
#include "struct_elem.cuh"

namespace svb {

struct StructArgs {
  const int* IEN;
  const int* eId;
  const int* slot;
  const int* perm;
  const double* fN;
  const double* x;
  const double* Ag;
  const double* Yg;
  const double* Dg;
  const double* Bf;
  const double* Ya;      // nodal active tensions (3, nNo): Ya_f, Ya_s, Ya_n (cep_mod.cem), or nullptr
  const double* pS0;     // nodal prestress com_mod.pS0 (6, nNo), Voigt 11,22,33,12,23,31, or nullptr
  double* pSn;           // pstEq: accumulators com_mod.pSn (6, nNo) and pSa (nNo) (sv_struct.cpp:327-336), else nullptr
  double* pSa;
  double* R;
  double* Val;
  int e0, e1;
  int tDof, dof, s, nFn, nDmn, atomic, nG;
  int lShpF;             // mshType::lShpF (nn_elem_props.h): gnn at Gauss point 0 only (TET4, WDG)
  const double* tab;     // per-mesh device copy of the element tables (quadratic elements)
  double dt, af, am, gam, beta;
  double w[MAX_NG];
  double N[MAX_NG][MAX_ENON];
  double Nxi[MAX_NG][MAX_ENON][3];
  StructDmn dmn[MAX_DMN];
  CannRow cann[MAX_CANN_ROWS];
};

template <bool ATOMIC>
__device__ __forceinline__ void add64(double* p, double v)
{
  if (ATOMIC) asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
  else *p += v;
}

constexpr int STRUCT_THREADS = 128;

// Shared-memory layout per element (doubles): nodal inputs x,d,q (9 ENON) | Gauss-point data (GP_LD per point) |
// per-node exchange of grad N_a of the current Gauss point (double-buffered).
constexpr int GP_LD = 64;   // xiX 9 | w 1 | F 9 | S 6 (11,22,33,12,23,31) | Dm 36 | ud 3
// nodal inputs 9 ENON | Gauss data GP_LD ENON | grad N exchange, double-buffered, 6 ENON; the stride is padded to
// 8 mod 16 doubles so that the elements of a half-warp read the same offset from different banks
__host__ __device__ constexpr int struct_per_el(int enon, int ng)
{
  const int n = enon * (9 + 6) + ng * GP_LD;
  return n + ((8 - (n % 16)) + 16) % 16;
}

// NG = number of Gauss points: ENON for TET4 / HEX8 / WDG; 15 for TET10, 27 for HEX20 / HEX27 (tables then come from P.tab, the
// per-mesh device copy of w | N | Nxi, because the fixed-size argument arrays stop at 8).
template <int ENON, int NG, bool ATOMIC, bool VISC, bool CANN>
__global__ void __launch_bounds__(STRUCT_THREADS)
assemble_struct_kernel(const __grid_constant__ StructArgs P)
{
  constexpr int LPE = ENON > NG ? ENON : NG; // lanes per element: Gauss points in phase A, element nodes in phase B
  constexpr int EPW = 32 / LPE;             // elements per warp
  constexpr bool ODD = (ENON & 1) != 0;
  constexpr int KMAX = ENON / 2;            // lane a owns blocks (a, a+k), k = 0..KMAX (even ENON: k = KMAX only for a < KMAX;
                                            // odd ENON: k = 1..KMAX for every a covers each unordered pair once)
  constexpr int PER_EL = struct_per_el(ENON, NG);
  constexpr int NTAB = NG * ENON * 4 + NG;  // Nxi[g][a][3], N[g][a], w[g]
  extern __shared__ double sm[];
  double* tab = sm;                          // reference-element tables (lane-dependent Gauss point in phase A)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a = lane % LPE, el = lane / LPE;
  double* se = sm + NTAB + (size_t)(warp * EPW + el) * PER_EL;
  double(*sx)[3] = reinterpret_cast<double(*)[3]>(se);
  double(*sd)[3] = reinterpret_cast<double(*)[3]>(se + 3 * ENON);
  double(*sq)[3] = reinterpret_cast<double(*)[3]>(se + 6 * ENON);
  double* sgp = se + 9 * ENON;
  double(*sNx)[3] = reinterpret_cast<double(*)[3]>(se + 9 * ENON + NG * GP_LD);
  double(*tNxi)[ENON][3] = reinterpret_cast<double(*)[ENON][3]>(tab);
  double(*tN)[ENON] = reinterpret_cast<double(*)[ENON]>(tab + NG * ENON * 3);
  double* tw = tab + NG * ENON * 4;
  for (int t = threadIdx.x; t < NG * ENON; t += STRUCT_THREADS) {
    const int g = t / ENON, b = t % ENON;
    if (NG <= MAX_NG && ENON <= MAX_ENON) {
      tNxi[g][b][0] = P.Nxi[g][b][0]; tNxi[g][b][1] = P.Nxi[g][b][1]; tNxi[g][b][2] = P.Nxi[g][b][2];
      tN[g][b] = P.N[g][b];
      if (b == 0) tw[g] = P.w[g];
    } else {
      // per Gauss point in P.tab: w | N[ENON] | Nxi[ENON][3] | Nxi2[ENON][6]  (fg_tab_ld = 1 + 10 ENON doubles)
      const double* tg = P.tab + (size_t)g * (1 + 10 * ENON);
      tNxi[g][b][0] = __ldg(tg + 1 + ENON + 3 * b); tNxi[g][b][1] = __ldg(tg + 2 + ENON + 3 * b); tNxi[g][b][2] = __ldg(tg + 3 + ENON + 3 * b);
      tN[g][b] = __ldg(tg + 1 + b);
      if (b == 0) tw[g] = __ldg(tg);
    }
  }

  const long long idx = (long long)P.e0 + ((long long)blockIdx.x * (STRUCT_THREADS / 32) + warp) * EPW + el;
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
    if (!P.dmn[iD].isStruct) active = false;
  }
  const StructDmn& dm = P.dmn[iD];
  const int DOF = P.dof;
  const bool rowLane = active && a < ENON;      // phase B / nodal loads: one lane per element node
  int node = 0;
  if (rowLane) {
    node = P.IEN[(size_t)e * ENON + a];
    const size_t n = (size_t)node;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      sx[a][i] = __ldg(P.x + 3 * n + i);
      sd[a][i] = __ldg(P.Dg + (size_t)P.tDof * n + P.s + i);
      sq[a][i] = dm.rho * (__ldg(P.Ag + (size_t)P.tDof * n + P.s + i) - __ldg(P.Bf + 3 * n + i)) +
                 dm.dmp * __ldg(P.Yg + (size_t)P.tDof * n + P.s + i);
    }
  }
  __syncthreads();

  // ---- phase A: lane a of the element evaluates Gauss point g = a: nn::gnn, F, compute_pk2cc -----
  // (mshType::lShpF elements — TET4, WDG — take the derivatives of Gauss point 0 everywhere, sv_struct.cpp:297-303)
  if (active && a < NG) {
    const int g = a;
    const int gd = P.lShpF ? 0 : g;
    double fN[2][3] = {{0, 0, 0}, {0, 0, 0}};
    if (P.fN != nullptr)
      for (int k = 0; k < P.nFn && k < 2; k++)
#pragma unroll
        for (int i = 0; i < 3; i++) fN[k][i] = __ldg(P.fN + (size_t)3 * P.nFn * e + 3 * k + i);
    double xXi[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int b = 0; b < ENON; b++)
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int k = 0; k < 3; k++) xXi[i][k] += sx[b][i] * tNxi[gd][b][k];
    const double Jac = xXi[0][0] * xXi[1][1] * xXi[2][2] + xXi[0][1] * xXi[1][2] * xXi[2][0] + xXi[0][2] * xXi[1][0] * xXi[2][1] -
                       xXi[0][0] * xXi[1][2] * xXi[2][1] - xXi[0][1] * xXi[1][0] * xXi[2][2] - xXi[0][2] * xXi[1][1] * xXi[2][0];
    const double iJ = 1.0 / Jac;
    double xiX[3][3];
    xiX[0][0] = (xXi[1][1] * xXi[2][2] - xXi[1][2] * xXi[2][1]) * iJ;
    xiX[0][1] = (xXi[2][1] * xXi[0][2] - xXi[2][2] * xXi[0][1]) * iJ;
    xiX[0][2] = (xXi[0][1] * xXi[1][2] - xXi[0][2] * xXi[1][1]) * iJ;
    xiX[1][0] = (xXi[1][2] * xXi[2][0] - xXi[1][0] * xXi[2][2]) * iJ;
    xiX[1][1] = (xXi[2][2] * xXi[0][0] - xXi[2][0] * xXi[0][2]) * iJ;
    xiX[1][2] = (xXi[0][2] * xXi[1][0] - xXi[0][0] * xXi[1][2]) * iJ;
    xiX[2][0] = (xXi[1][0] * xXi[2][1] - xXi[1][1] * xXi[2][0]) * iJ;
    xiX[2][1] = (xXi[2][0] * xXi[0][1] - xXi[2][1] * xXi[0][0]) * iJ;
    xiX[2][2] = (xXi[0][0] * xXi[1][1] - xXi[0][1] * xXi[1][0]) * iJ;
    double F[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double ud[3] = {-dm.rho * dm.f[0], -dm.rho * dm.f[1], -dm.rho * dm.f[2]};
#pragma unroll
    for (int b = 0; b < ENON; b++) {
      const double Nb = tN[g][b];
      double Nxb[3];
#pragma unroll
      for (int j = 0; j < 3; j++) Nxb[j] = tNxi[gd][b][0] * xiX[0][j] + tNxi[gd][b][1] * xiX[1][j] + tNxi[gd][b][2] * xiX[2][j];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        ud[i] += Nb * sq[b][i];
#pragma unroll
        for (int j = 0; j < 3; j++) F[i][j] += Nxb[j] * sd[b][i];
      }
    }
    // active tensions at the Gauss point: ya_g = sum_b N_b Ya(:, node_b)  (sv_struct.cpp:640-642)
    double ya[3] = {0.0, 0.0, 0.0};
    const bool act = CANN && (P.Ya != nullptr) && dm.active;
    if (act) {
#pragma unroll
      for (int b = 0; b < ENON; b++) {
        const size_t nb = (size_t)P.IEN[(size_t)e * ENON + b];
        const double Nb = tN[g][b];
#pragma unroll
        for (int i = 0; i < 3; i++) ya[i] += Nb * __ldg(P.Ya + 3 * nb + i);
      }
    }
    double S[3][3], Dm[6][6];
    pk2cc_voigt<CANN>(dm, F, fN, act ? ya : nullptr, P.cann, P.nFn, S, Dm);
    if (VISC && dm.viscType != SVB200_SOLID_VISC_NONE) {
      // S += Svis (sv_struct.cpp:666-669); the viscous tangent is added by assemble_struct_visc_kernel
      double vx[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, Svis[3][3];
#pragma unroll
      for (int b = 0; b < ENON; b++) {
        const size_t nb = (size_t)P.IEN[(size_t)e * ENON + b];
        double Nxb[3];
#pragma unroll
        for (int j = 0; j < 3; j++) Nxb[j] = tNxi[gd][b][0] * xiX[0][j] + tNxi[gd][b][1] * xiX[1][j] + tNxi[gd][b][2] * xiX[2][j];
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const double y = __ldg(P.Yg + (size_t)P.tDof * nb + P.s + i);
#pragma unroll
          for (int j = 0; j < 3; j++) vx[i][j] += Nxb[j] * y;
        }
      }
      ViscGP vgp;
      visc_gauss_point(dm.viscType, dm.visc_mu, 0.0, 0.0, F, vx, Svis, vgp);
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) S[i][j] += Svis[i][j];
    }
    // Prestress (sv_struct.cpp:671-680, 327-336): pSl = S before the nodal prestress S0 = sum_b N_b pS0(:, node_b) is added;
    // a prestress equation accumulates w N_a pSl and w N_a per node (the corrector divides them, Integrator.cpp:912-924).
    if (P.pSn != nullptr) {
      const double wg = tw[g] * Jac;
      const double pSl[6] = {S[0][0], S[1][1], S[2][2], S[0][1], S[1][2], S[2][0]};
#pragma unroll
      for (int b = 0; b < ENON; b++) {
        const size_t nb = (size_t)P.IEN[(size_t)e * ENON + b];
        const double wN = wg * tN[g][b];
        add64<true>(P.pSa + nb, wN);
#pragma unroll
        for (int i = 0; i < 6; i++) add64<true>(P.pSn + 6 * nb + i, wN * pSl[i]);
      }
    }
    if (P.pS0 != nullptr) {
      double S0[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int b = 0; b < ENON; b++) {
        const size_t nb = (size_t)P.IEN[(size_t)e * ENON + b];
        const double Nb = tN[g][b];
#pragma unroll
        for (int i = 0; i < 6; i++) S0[i] += Nb * __ldg(P.pS0 + 6 * nb + i);
      }
      S[0][0] += S0[0]; S[1][1] += S0[1]; S[2][2] += S0[2];
      S[0][1] += S0[3]; S[1][0] += S0[3]; S[1][2] += S0[4]; S[2][1] += S0[4]; S[2][0] += S0[5]; S[0][2] += S0[5];
    }
    double* q = sgp + g * GP_LD;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) { q[3 * i + j] = xiX[i][j]; q[10 + 3 * i + j] = F[i][j]; }
    q[9] = tw[g] * Jac;
    q[19] = S[0][0]; q[20] = S[1][1]; q[21] = S[2][2]; q[22] = S[0][1]; q[23] = S[1][2]; q[24] = S[2][0];
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int c = 0; c < 6; c++) q[25 + 6 * r + c] = Dm[r][c];
    q[61] = ud[0]; q[62] = ud[1]; q[63] = ud[2];
  }
  __syncwarp();

  // ---- phase B: lane a = element node a.  Material tangent in contracted form: with G_a = Bm_a^T Dm (3x6) arranged as
  // three symmetric 3x3 matrices Gs_i, Bm_b(r,j) being linear in grad N_b gives
  //     Bm_a(:,i) . Dm Bm_b(:,j) = sum_L H_a(i,j,L) dN_b/dx_L ,   H_a(i,j,L) = sum_M Gs_i(L,M) F(j,M),
  // so only grad N_b (3 doubles per node and Gauss point) is exchanged between lanes and a block costs 27 + 8 FMAs
  // (sv_struct.cpp:736-825 evaluates the same sums as Bm_a^T (Dm Bm_b): 54 FMAs and an 18-double DBm_b per block). -----
  const double afu = P.af * P.beta * P.dt * P.dt;
  const double amd = P.am * dm.rho + P.af * P.gam * P.dt * dm.dmp;
  double acc[KMAX + 1][3][3];
  double lR[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k <= KMAX; k++)
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) acc[k][i][j] = 0.0;

  // CSR slots of this lane's blocks (and of their transposes), requested before the Gauss loop: the scatter at the end
  // otherwise waits a full global-memory latency per block with nothing to overlap it (15 % of the kernel under ncu)
  int slotA[KMAX + 1], slotT[KMAX + 1];
#pragma unroll
  for (int k = 0; k <= KMAX; k++) { slotA[k] = 0; slotT[k] = 0; }
  if (rowLane) {
    const int* sl = P.slot + (size_t)e * ENON * ENON;
#pragma unroll
    for (int k = 0; k <= KMAX; k++) {
      const int b = (a + k) % ENON;
      slotA[k] = __ldg(sl + a * ENON + b);
      slotT[k] = __ldg(sl + b * ENON + a);
    }
  }
#pragma unroll 1
  for (int g = 0; g < NG; g++) {
    double H[3][3][3], SNx[3];
    double wamdNa = 0.0;
    double(*pNx)[3] = sNx + (g & 1) * ENON;       // double-buffered exchange: one __syncwarp per Gauss point
    const int gd = P.lShpF ? 0 : g;
    if (rowLane) {
      const double* q = sgp + g * GP_LD;
      const double w = q[9];
      const double wafu = w * afu;
      double F[3][3], Nxa[3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        Nxa[i] = tNxi[gd][a][0] * q[i] + tNxi[gd][a][1] * q[3 + i] + tNxi[gd][a][2] * q[6 + i];
        pNx[a][i] = Nxa[i];
#pragma unroll
        for (int j = 0; j < 3; j++) F[i][j] = q[10 + 3 * i + j];
      }
      {
        double Bm[6][3], G[3][6];
        make_Bm(Nxa, F, Bm);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int c = 0; c < 6; c++) G[i][c] = 0.0;
        // G(i,c) = sum_r Bm_a(r,i) Dm(r,c), Dm read row by row from shared memory (broadcast within the element)
#pragma unroll
        for (int r = 0; r < 6; r++) {
          double d[6];
#pragma unroll
          for (int c = 0; c < 6; c++) d[c] = q[25 + 6 * r + c];
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int c = 0; c < 6; c++) G[i][c] += Bm[r][i] * d[c];
        }
        // Voigt order 11,22,33,12,23,31 -> symmetric matrix Gs_i(L,M); H(i,j,L) = w afu sum_M Gs_i(L,M) F(j,M)
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const double g00 = wafu * G[i][0], g11 = wafu * G[i][1], g22 = wafu * G[i][2];
          const double g01 = wafu * G[i][3], g12 = wafu * G[i][4], g20 = wafu * G[i][5];
#pragma unroll
          for (int j = 0; j < 3; j++) {
            H[i][j][0] = g00 * F[j][0] + g01 * F[j][1] + g20 * F[j][2];
            H[i][j][1] = g01 * F[j][0] + g11 * F[j][1] + g12 * F[j][2];
            H[i][j][2] = g20 * F[j][0] + g12 * F[j][1] + g22 * F[j][2];
          }
        }
      }
      const double S00 = q[19], S11 = q[20], S22 = q[21], S01 = q[22], S12 = q[23], S20 = q[24];
      SNx[0] = Nxa[0] * S00 + Nxa[1] * S01 + Nxa[2] * S20;
      SNx[1] = Nxa[0] * S01 + Nxa[1] * S11 + Nxa[2] * S12;
      SNx[2] = Nxa[0] * S20 + Nxa[1] * S12 + Nxa[2] * S22;
      const double Na = tN[g][a];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        // P = F S ; lR(i,a) += w (N_a ud_i + sum_j Nx(j,a) P(i,j))
        const double PNx = F[i][0] * SNx[0] + F[i][1] * SNx[1] + F[i][2] * SNx[2];
        lR[i] += w * (Na * q[61 + i] + PNx);
      }
      wamdNa = w * amd * Na;
#pragma unroll
      for (int i = 0; i < 3; i++) SNx[i] *= wafu;
    }
    __syncwarp();
    if (rowLane) {
#pragma unroll
      for (int k = 0; k <= KMAX; k++) {
        if (!ODD && k == KMAX && a >= KMAX) continue;
        const int b = (a + k) % ENON;
        const double n0 = pNx[b][0], n1 = pNx[b][1], n2 = pNx[b][2];
        // delta_ij w (amd Na Nb + afu gradNa.S.gradNb)
        const double T1 = wamdNa * tN[g][b] + (SNx[0] * n0 + SNx[1] * n1 + SNx[2] * n2);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++)
            acc[k][i][j] += (H[i][j][0] * n0 + H[i][j][1] * n1 + H[i][j][2] * n2) + (i == j ? T1 : 0.0);
      }
    }
  }

  // ---- scatter ------------------------------------------------------------------------------------------
  if (rowLane) {
#pragma unroll
    for (int i = 0; i < 3; i++) add64<ATOMIC>(P.R + (size_t)DOF * node + i, lR[i]);
  }
  // Blocks go out through a per-warp transposition tile (the warp's own Gauss-point area, dead by now): every lane deposits
  // a finished 3x3 block, then the warp adds the 32 blocks with consecutive lanes on consecutive doubles, first as they
  // are, then transposed into the mirrored slots.  Lane-strided REDs (each lane walking its own block) reach a third of
  // the coalesced rate at best (profiles/r1_microbench_fp64_red.txt: 174 vs 554 G adds/s, 42 vs 285 when Val streams
  // from DRAM) and were what bounded this kernel.
  double* tile = sm + NTAB + (size_t)warp * EPW * PER_EL;
  int* tsl = reinterpret_cast<int*>(tile + 32 * 9);
  const int DD = DOF * DOF;
#pragma unroll
  for (int k = 0; k <= KMAX; k++) {
    const bool mine = rowLane && !(!ODD && k == KMAX && a >= KMAX);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) tile[lane * 9 + 3 * i + j] = acc[k][i][j];
    tsl[lane] = mine ? slotA[k] : -1;
    tsl[32 + lane] = (mine && k > 0) ? slotT[k] : -1;
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 9; it++) {
      const int p = it * 32 + lane, src = p / 9, idx = p - 9 * src, i = idx / 3, j = idx - 3 * i;
      const double v = tile[p];
      const int s0 = tsl[src];
      if (s0 >= 0) add64<ATOMIC>(P.Val + (size_t)DD * s0 + DOF * i + j, v);
    }
    if (k > 0) {
#pragma unroll
      for (int it = 0; it < 9; it++) {
        // walk the TRANSPOSED block in its own memory order so that the adds stay on consecutive doubles
        const int p = it * 32 + lane, src = p / 9, idx = p - 9 * src, i = idx / 3, j = idx - 3 * i;
        const int s1 = tsl[32 + src];
        if (s1 >= 0) add64<ATOMIC>(P.Val + (size_t)DD * s1 + DOF * i + j, tile[src * 9 + 3 * j + i]);
      }
    }
  }
}

// ---- linear tetrahedra: one thread per element ----------------------------------------------------------------
// For TET4 the shape-function gradients are constant (lShpF, sv_struct.cpp:297-303), hence F, S, Dm, P and every Bm_a
// are the same at the four Gauss points: struct_3d evaluates compute_pk2cc four times with identical arguments.  Only
// the inertia / mass terms see N_a(g).  With W = Jac sum_g w_g, m_a = Jac sum_g w_g N_a(g), M_ab = Jac sum_g w_g N_a(g) N_b(g):
//     lR(i,a)   = -rho f_i m_a + sum_b M_ab q_b(i) + W (F S grad N_a)_i
//     lK(ij,ab) = delta_ij (amd M_ab + afu W grad N_a . S grad N_b) + afu W Bm_a(:,i) . Dm Bm_b(:,j)
// (the Gauss sum of sv_struct.cpp:697-825 taken in closed form; differences to it are rounding only).  One thread does
// one element: one compute_pk2cc, 4 DBm_b, the 10 blocks a <= b (K_ba = K_ab^T), each block leaving at once through the
// per-warp transposition tile (coalesced adds).  Without solid viscosity only (that tangent is not symmetric and keeps
// the general kernel).  FSI pipe C5: 4.4 M solid tets 5.85 ms with the general kernel.
constexpr int STET_THREADS = 128;

// Whole-warp scatter of one 3x3 block per lane: every lane deposits its block in the warp's tile, then the 32 blocks are
// added with consecutive lanes on consecutive doubles — into slot `s_ab` as they are and, when `mirrored` (uniform over the
// warp), transposed into `s_ba`.  Lanes with nothing to add pass slot -1.  Must be called by all 32 lanes.
template <bool ATOMIC>
__device__ __forceinline__ void warp_scatter_block_pair(double* tile, int* tsl, int lane, const double K[3][3], int s_ab, int s_ba,
                                                        bool mirrored, double* Val, int DD, int DOF)
{
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) tile[lane * 9 + 3 * i + j] = K[i][j];
  tsl[lane] = s_ab;
  tsl[32 + lane] = s_ba;
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 9; it++) {
    const int p = it * 32 + lane, src = p / 9, idx9 = p - 9 * src, i = idx9 / 3, j = idx9 - 3 * i;
    const int s0 = tsl[src];
    if (s0 >= 0) add64<ATOMIC>(Val + (size_t)DD * s0 + DOF * i + j, tile[p]);
  }
  if (mirrored) {
#pragma unroll
    for (int it = 0; it < 9; it++) {
      const int p = it * 32 + lane, src = p / 9, idx9 = p - 9 * src, i = idx9 / 3, j = idx9 - 3 * i;
      const int s1 = tsl[32 + src];
      if (s1 >= 0) add64<ATOMIC>(Val + (size_t)DD * s1 + DOF * i + j, tile[src * 9 + 3 * j + i]);
    }
  }
}


template <bool ATOMIC, bool CANN>
__global__ void __launch_bounds__(STET_THREADS)
assemble_struct_tet4_kernel(const __grid_constant__ StructArgs P)
{
  __shared__ double s_tile[STET_THREADS / 32][32 * 9];
  __shared__ int s_slot[STET_THREADS / 32][64];
  __shared__ double s_Dm[36][STET_THREADS];      // the element's Dm, one column per thread: 36 registers less in the block loop
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* tile = s_tile[warp];
  int* tsl = s_slot[warp];
  const long long idx = (long long)P.e0 + (long long)blockIdx.x * STET_THREADS + threadIdx.x;
  bool active = idx < P.e1;
  int e = 0;
  if (active) e = P.perm ? P.perm[idx] : (int)idx;
  int iD = 0;
  if (active) {
    for (int d = 0; d < P.nDmn; d++) {
      iD = d;
      if (P.dmn[d].Id == -1) break;
      if (P.eId != nullptr && ((P.eId[e] >> P.dmn[d].Id) & 1)) break;
    }
    if (!P.dmn[iD].isStruct) active = false;
  }
  const StructDmn& dm = P.dmn[iD];
  const int DOF = P.dof, DD = DOF * DOF;
  const double afu = P.af * P.beta * P.dt * P.dt;
  const double amd = P.am * dm.rho + P.af * P.gam * P.dt * dm.dmp;

  int node[4] = {0, 0, 0, 0};
  int sl[16];
#pragma unroll
  for (int k = 0; k < 16; k++) sl[k] = -1;
  double Nx[4][3], F[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, S[3][3], M2[4][4];
  double W = 0.0, Jac = 0.0;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    Nx[a][0] = Nx[a][1] = Nx[a][2] = 0.0;
#pragma unroll
    for (int b = 0; b < 4; b++) M2[a][b] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) S[i][j] = 0.0;

  if (active) {
    const int4 nn = __ldg(reinterpret_cast<const int4*>(P.IEN) + e);
    node[0] = nn.x; node[1] = nn.y; node[2] = nn.z; node[3] = nn.w;
    const int4* sp = reinterpret_cast<const int4*>(P.slot) + (size_t)e * 4;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int4 v = __ldg(sp + k);
      sl[4 * k] = v.x; sl[4 * k + 1] = v.y; sl[4 * k + 2] = v.z; sl[4 * k + 3] = v.w;
    }
    double xl[4][3], dl[4][3], ql[4][3];
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const size_t n = (size_t)node[a];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        xl[a][i] = __ldg(P.x + 3 * n + i);
        dl[a][i] = __ldg(P.Dg + (size_t)P.tDof * n + P.s + i);
        ql[a][i] = dm.rho * (__ldg(P.Ag + (size_t)P.tDof * n + P.s + i) - __ldg(P.Bf + 3 * n + i)) +
                   dm.dmp * __ldg(P.Yg + (size_t)P.tDof * n + P.s + i);
      }
    }
    Jac = gnn3<4>(P.Nxi[0], xl, Nx);
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) F[i][j] += Nx[a][j] * dl[a][i];
    double fN[2][3] = {{0, 0, 0}, {0, 0, 0}};
    if (P.fN != nullptr)
      for (int k = 0; k < P.nFn && k < 2; k++)
#pragma unroll
        for (int i = 0; i < 3; i++) fN[k][i] = __ldg(P.fN + (size_t)3 * P.nFn * e + 3 * k + i);
    // active tensions (full twin only): S and Dm are affine in (Tfa, Tsa, Tna), so the Gauss sum of struct_3d equals one
    // evaluation at the weighted mean sum_g w_g ya_g / W = sum_a m1_a Ya_a / W (the weights' Jacobian cancels)
    double ya[3] = {0.0, 0.0, 0.0};
    const bool act = CANN && (P.Ya != nullptr) && dm.active;
    if (act) {
      double wsum = 0.0;
#pragma unroll
      for (int g = 0; g < 4; g++) {
        wsum += P.w[g];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int i = 0; i < 3; i++) ya[i] += P.w[g] * P.N[g][a] * __ldg(P.Ya + 3 * (size_t)node[a] + i);
      }
#pragma unroll
      for (int i = 0; i < 3; i++) ya[i] /= wsum;
    }
    double Dm[6][6];
    pk2cc_voigt<CANN>(dm, F, fN, act ? ya : nullptr, P.cann, P.nFn, S, Dm);
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int c = 0; c < 6; c++) s_Dm[6 * r + c][threadIdx.x] = Dm[r][c];
    // moments of the quadrature rule, residual (struct_elem.cuh: the same routines the CPU suite checks against the reference);
    // after compute_pk2cc, so that the 21 moments are not live across it
    Tet4Mom q;
    tet4_moments(P.w, &P.N[0][0], MAX_ENON, Jac, q);
    W = q.W;
    double Pk[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Pk[i][j] = F[i][0] * S[0][j] + F[i][1] * S[1][j] + F[i][2] * S[2][j];
#pragma unroll
    for (int a = 0; a < 4; a++) {
      double r[3];
      struct_tet4_residual(dm, q, a, Nx[a], Pk, ql, r);
#pragma unroll
      for (int i = 0; i < 3; i++) add64<ATOMIC>(P.R + (size_t)DOF * node[a] + i, r[i]);
    }
    // from here on only amd M_ab is needed
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) M2[a][b] = amd * q.M2[a][b];
  }

  // tangent blocks a <= b; every lane of the warp takes part in the deposit / add steps (lanes without an element carry
  // slot -1)
  const double wafu = W * afu;
#pragma unroll
  for (int b = 0; b < 4; b++) {
    double DBmb[6][3], SNb[3];
    {
      double Bmb[6][3];
      make_Bm(Nx[b], F, Bmb);
#pragma unroll
      for (int r = 0; r < 6; r++) {
        double d[6];
#pragma unroll
        for (int c = 0; c < 6; c++) d[c] = active ? s_Dm[6 * r + c][threadIdx.x] : 0.0;
#pragma unroll
        for (int j = 0; j < 3; j++)
          DBmb[r][j] = d[0] * Bmb[0][j] + d[1] * Bmb[1][j] + d[2] * Bmb[2][j] + d[3] * Bmb[3][j] + d[4] * Bmb[4][j] + d[5] * Bmb[5][j];
      }
    }
#pragma unroll
    for (int i = 0; i < 3; i++) SNb[i] = S[i][0] * Nx[b][0] + S[i][1] * Nx[b][1] + S[i][2] * Nx[b][2];
#pragma unroll
    for (int a = 0; a <= b; a++) {
      double Bma[6][3], K[3][3];
      make_Bm(Nx[a], F, Bma);
      struct_tet4_block(wafu, M2[a][b], Nx[a], SNb, Bma, DBmb, K);
      warp_scatter_block_pair<ATOMIC>(tile, tsl, lane, K, sl[4 * a + b], (a != b) ? sl[4 * b + a] : -1, a != b, P.Val, DD, DOF);
    }
  }
}

// ---- viscous tangent of the solid (mat_models.cpp:1583-1762, sv_struct.cpp:759-823) ---------------------
// Launched after assemble_struct_kernel<.., VISC = true> for the domains with a solid viscosity model; adds
// w (afu Kvis_u + afv Kvis_v) for ALL ENON x ENON node pairs (this part of the element matrix is not symmetric).
// Same lane mapping: in phase A lane g evaluates Gauss point g (nn::gnn, F, vx, visc_gauss_point) and leaves the
// three vectors V1..V3 of every element node plus M and w c in shared memory; in phase B lane a accumulates block
// (a, b) over the Gauss points for one b at a time and scatters it.
constexpr int VISC_THREADS = 128;
__host__ __device__ constexpr int visc_per_el(int enon)
{
  const int n = enon * (enon * 9 + 10);
  return n + ((8 - (n % 16)) + 16) % 16;
}

template <int ENON, bool ATOMIC>
__global__ void __launch_bounds__(VISC_THREADS)
assemble_struct_visc_kernel(const __grid_constant__ StructArgs P)
{
  constexpr int EPW = 32 / ENON;
  constexpr int PER_EL = visc_per_el(ENON);
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a = lane % ENON, el = lane / ENON;
  double* se = sm + (size_t)(warp * EPW + el) * PER_EL;
  double* sV = se;                           // [g][b][9]
  double* sM = se + ENON * ENON * 9;         // [g][10]: M (9), w c
  const long long idx = (long long)P.e0 + ((long long)blockIdx.x * (VISC_THREADS / 32) + warp) * EPW + el;
  bool active = idx < P.e1;
  int e = 0;
  if (active) e = P.perm ? P.perm[idx] : (int)idx;
  int iD = 0;
  if (active) {
    for (int d = 0; d < P.nDmn; d++) {
      iD = d;
      if (P.dmn[d].Id == -1) break;
      if (P.eId != nullptr && ((P.eId[e] >> P.dmn[d].Id) & 1)) break;
    }
    if (!P.dmn[iD].isStruct || P.dmn[iD].viscType == SVB200_SOLID_VISC_NONE) active = false;
  }
  const StructDmn& dm = P.dmn[iD];
  const int DOF = P.dof;
  const double afu = P.af * P.beta * P.dt * P.dt, afv = P.af * P.gam * P.dt;
  if (active) {
    const int g = a;
    double xl[ENON][3], Nx[ENON][3];
    double F[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, vx[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int b = 0; b < ENON; b++) {
      const size_t n = (size_t)P.IEN[(size_t)e * ENON + b];
#pragma unroll
      for (int i = 0; i < 3; i++) xl[b][i] = __ldg(P.x + 3 * n + i);
    }
    const double Jac = gnn3<ENON>(P.Nxi[g], xl, Nx);
#pragma unroll
    for (int b = 0; b < ENON; b++) {
      const size_t n = (size_t)P.IEN[(size_t)e * ENON + b];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double d = __ldg(P.Dg + (size_t)P.tDof * n + P.s + i), y = __ldg(P.Yg + (size_t)P.tDof * n + P.s + i);
#pragma unroll
        for (int j = 0; j < 3; j++) { F[i][j] += Nx[b][j] * d; vx[i][j] += Nx[b][j] * y; }
      }
    }
    ViscGP vgp;
    double Svis[3][3];
    visc_gauss_point(dm.viscType, dm.visc_mu, afu, afv, F, vx, Svis, vgp);
#pragma unroll
    for (int b = 0; b < ENON; b++) {
      double V[9];
      visc_node(vgp, Nx[b], V);
#pragma unroll
      for (int k = 0; k < 9; k++) sV[(g * ENON + b) * 9 + k] = V[k];
    }
#pragma unroll
    for (int k = 0; k < 9; k++) sM[g * 10 + k] = vgp.M[k / 3][k % 3];
    sM[g * 10 + 9] = P.w[g] * Jac * vgp.c;
  }
  __syncwarp();
  if (!active) return;
  const int node = P.IEN[(size_t)e * ENON + a];
  (void)node;
  const int* sl = P.slot + (size_t)e * ENON * ENON;
#pragma unroll 1
  for (int b = 0; b < ENON; b++) {
    double K[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll 1
    for (int g = 0; g < ENON; g++) {
      double Va[9], Vb[9], M[9];
#pragma unroll
      for (int k = 0; k < 9; k++) {
        Va[k] = sV[(g * ENON + a) * 9 + k];
        Vb[k] = sV[(g * ENON + b) * 9 + k];
        M[k] = sM[g * 10 + k];
      }
      visc_block(dm.viscType, sM[g * 10 + 9], afu, afv, M, Va, Vb, K);
    }
    double* v = P.Val + (size_t)DOF * DOF * sl[a * ENON + b];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) add64<ATOMIC>(v + DOF * i + j, K[i][j]);
  }
}

// ---- mesh-motion equation: linear elasticity on the configuration at t_n --------------------------------
// mesh::construct_mesh (Code/Source/solver/mesh.cpp:22-135) + l_elas::l_elas_3d (Code/Source/solver/l_elas.cpp:249-365):
// geometry x + Do(is..), displacement Dg(is..) - Do(is..), and the Gauss weight WITHOUT the Jacobian
// (w = lM.w(g), mesh.cpp:122: Jacobian-based stiffening).  ENON lanes per element, lane a owns row a.
// LELAS = true: the linear-elasticity EQUATION, l_elas::construct_l_elas (l_elas.cpp:36-145): reference geometry, the
// displacement itself, w = lM.w(g) * Jac and the nodal body force Bf in the inertia term.
template <int ENON, bool ATOMIC, bool LELAS>
__global__ void __launch_bounds__(128)
assemble_mesh_kernel(const __grid_constant__ StructArgs P, const double* __restrict__ Do)
{
  constexpr int EPW = 32 / ENON;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a = lane % ENON, el = lane / ENON;
  // tables: the fixed-size argument arrays for TET4 / HEX8 / WDG, the per-mesh device copy (w | N | Nxi | Nxi2 per Gauss point,
  // assemble_fluid_gen.cu) for TET10 / HEX20 / HEX27
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
  const StructDmn& dm = P.dmn[iD];
  if (!dm.isStruct) return;          // isStruct marks the domains this launch handles (mesh domains here)
  const int DOF = P.dof, is = P.s;
  int node[ENON];
  double xl[ENON][3], dl[ENON][3];
#pragma unroll
  for (int b = 0; b < ENON; b++) {
    node[b] = P.IEN[(size_t)e * ENON + b];
    const size_t n = (size_t)node[b];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const double dol = LELAS ? 0.0 : __ldg(Do + (size_t)P.tDof * n + is + i);
      xl[b][i] = __ldg(P.x + 3 * n + i) + dol;
      dl[b][i] = __ldg(P.Dg + (size_t)P.tDof * n + is + i) - dol;
    }
  }
  // elasticity_modulus and poisson_ratio travel in C10 / C01 for this kernel
  const double elM = dm.C10, nu = dm.C01, rho = dm.rho;
  const double lambda = elM * nu / (1.0 + nu) / (1.0 - 2.0 * nu);
  const double mu = elM * 0.5 / (1.0 + nu);
  const double lDm = lambda / mu;
  const double T1c = P.af * P.beta * P.dt * P.dt;
  const double amd = P.am / T1c * rho;
  double K[ENON][3][3], lR[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int b = 0; b < ENON; b++)
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) K[b][i][j] = 0.0;
  double Nx[ENON][3];
  double Jac = 1.0;
#pragma unroll 1
  for (int g = 0; g < P.nG; g++) {
    const double* tg = TAB ? P.tab + (size_t)g * TLD : nullptr;
    const double* Ng = TAB ? tg + 1 : P.N[g];
    const double wg = TAB ? tg[0] : P.w[g];
    // TET4 / WDG: lShpF, one gnn per element (l_elas.cpp:112, mesh.cpp:114) — for the wedge too, whose gradients are not constant
    if (g == 0 || (ENON != 4 && ENON != 6))
      Jac = gnn3<ENON>(TAB ? reinterpret_cast<const double(*)[3]>(tg + 1 + ENON) : P.Nxi[g], xl, Nx);
    const double w = LELAS ? wg * Jac : wg;
    const double wl = w * T1c * mu;
    double ud[3] = {-dm.f[0], -dm.f[1], -dm.f[2]}, ed[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int b = 0; b < ENON; b++) {
      const double Nb = Ng[b];
      const size_t n = (size_t)node[b];
#pragma unroll
      for (int i = 0; i < 3; i++)
        ud[i] += Nb * (__ldg(P.Ag + (size_t)P.tDof * n + is + i) - (LELAS ? __ldg(P.Bf + 3 * n + i) : 0.0));
      ed[0] += Nx[b][0] * dl[b][0];
      ed[1] += Nx[b][1] * dl[b][1];
      ed[2] += Nx[b][2] * dl[b][2];
      ed[3] += Nx[b][1] * dl[b][0] + Nx[b][0] * dl[b][1];
      ed[4] += Nx[b][2] * dl[b][1] + Nx[b][1] * dl[b][2];
      ed[5] += Nx[b][0] * dl[b][2] + Nx[b][2] * dl[b][0];
    }
    const double divD = lambda * (ed[0] + ed[1] + ed[2]);
    double S0 = divD + 2.0 * mu * ed[0], S1 = divD + 2.0 * mu * ed[1], S2 = divD + 2.0 * mu * ed[2];
    double S3 = mu * ed[3], S4 = mu * ed[4], S5 = mu * ed[5];
    const double Na = Ng[a];
    if (LELAS) {
      // prestress of the linear-elasticity equation (l_elas.cpp:321-338, 130-140): lane a accumulates its own node
      if (P.pSn != nullptr) {
        const double wN = w * Na;
        add64<true>(P.pSa + node[a], wN);
        add64<true>(P.pSn + 6 * (size_t)node[a] + 0, wN * S0); add64<true>(P.pSn + 6 * (size_t)node[a] + 1, wN * S1);
        add64<true>(P.pSn + 6 * (size_t)node[a] + 2, wN * S2); add64<true>(P.pSn + 6 * (size_t)node[a] + 3, wN * S3);
        add64<true>(P.pSn + 6 * (size_t)node[a] + 4, wN * S4); add64<true>(P.pSn + 6 * (size_t)node[a] + 5, wN * S5);
      }
      if (P.pS0 != nullptr) {
        double p0[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int b = 0; b < ENON; b++)
#pragma unroll
          for (int i = 0; i < 6; i++) p0[i] += Ng[b] * __ldg(P.pS0 + 6 * (size_t)node[b] + i);
        S0 += p0[0]; S1 += p0[1]; S2 += p0[2]; S3 += p0[3]; S4 += p0[4]; S5 += p0[5];
      }
    }
    lR[0] += w * (rho * Na * ud[0] + Nx[a][0] * S0 + Nx[a][1] * S3 + Nx[a][2] * S5);
    lR[1] += w * (rho * Na * ud[1] + Nx[a][0] * S3 + Nx[a][1] * S1 + Nx[a][2] * S4);
    lR[2] += w * (rho * Na * ud[2] + Nx[a][0] * S5 + Nx[a][1] * S4 + Nx[a][2] * S2);
#pragma unroll
    for (int b = 0; b < ENON; b++) {
      const double NxdNx = Nx[a][0] * Nx[b][0] + Nx[a][1] * Nx[b][1] + Nx[a][2] * Nx[b][2];
      const double T1 = amd * Na * Ng[b] / mu + NxdNx;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
          K[b][i][j] += wl * ((i == j ? T1 + (1.0 + lDm) * Nx[a][i] * Nx[b][i] : lDm * Nx[a][i] * Nx[b][j] + Nx[a][j] * Nx[b][i]));
    }
  }
#pragma unroll
  for (int i = 0; i < 3; i++) add64<ATOMIC>(P.R + (size_t)DOF * node[a] + i, lR[i]);
  const int* sl = P.slot + (size_t)e * ENON * ENON + a * ENON;
#pragma unroll
  for (int b = 0; b < ENON; b++) {
    double* v = P.Val + (size_t)DOF * DOF * sl[b];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) add64<ATOMIC>(v + DOF * i + j, K[b][i][j]);
  }
}

// HEX8 variant with one lane per Gauss point in phase A (the kernel above lets every lane of an element evaluate gnn and the stress
// at ALL 8 Gauss points): lane g leaves Nx_g(8,3), the stress S(6), the interpolated prestress p0(6), the inertia vector ud(3) and the
// weight in shared memory (41 doubles, odd stride), lane a then accumulates row a over the 8 records.  Same expressions, same order
// of the Gauss-point sums, so the results are those of the lane-per-row kernel.
constexpr int MESH_GP_LD = 41;
template <bool ATOMIC, bool LELAS>
__global__ void __launch_bounds__(128)
assemble_mesh_hex8_kernel(const __grid_constant__ StructArgs P, const double* __restrict__ Do)
{
  constexpr int ENON = 8, EPW = 4;
  __shared__ double sgp[4][EPW][ENON][MESH_GP_LD];
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
    active = P.dmn[iD].isStruct != 0;
  }
  const StructDmn& dm = P.dmn[iD];
  const int DOF = P.dof, is = P.s;
  const double elM = dm.C10, nu = dm.C01, rho = dm.rho;
  const double lambda = elM * nu / (1.0 + nu) / (1.0 - 2.0 * nu);
  const double mu = elM * 0.5 / (1.0 + nu);
  const double lDm = lambda / mu;
  const double T1c = P.af * P.beta * P.dt * P.dt;
  const double amd = P.am / T1c * rho;
  int na = 0;
  if (active) {
    const int g = a;
    int node[ENON];
    double xl[ENON][3], dl[ENON][3];
#pragma unroll
    for (int b = 0; b < ENON; b++) {
      node[b] = P.IEN[(size_t)e * ENON + b];
      if (b == a) na = node[b];
      const size_t n = (size_t)node[b];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double dol = LELAS ? 0.0 : __ldg(Do + (size_t)P.tDof * n + is + i);
        xl[b][i] = __ldg(P.x + 3 * n + i) + dol;
        dl[b][i] = __ldg(P.Dg + (size_t)P.tDof * n + is + i) - dol;
      }
    }
    double Nx[ENON][3];
    const double Jac = gnn3<ENON>(P.Nxi[g], xl, Nx);
    const double w = LELAS ? P.w[g] * Jac : P.w[g];
    double ud[3] = {-dm.f[0], -dm.f[1], -dm.f[2]}, ed[6] = {0, 0, 0, 0, 0, 0}, p0[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int b = 0; b < ENON; b++) {
      const double Nb = P.N[g][b];
      const size_t n = (size_t)node[b];
#pragma unroll
      for (int i = 0; i < 3; i++)
        ud[i] += Nb * (__ldg(P.Ag + (size_t)P.tDof * n + is + i) - (LELAS ? __ldg(P.Bf + 3 * n + i) : 0.0));
      ed[0] += Nx[b][0] * dl[b][0];
      ed[1] += Nx[b][1] * dl[b][1];
      ed[2] += Nx[b][2] * dl[b][2];
      ed[3] += Nx[b][1] * dl[b][0] + Nx[b][0] * dl[b][1];
      ed[4] += Nx[b][2] * dl[b][1] + Nx[b][1] * dl[b][2];
      ed[5] += Nx[b][0] * dl[b][2] + Nx[b][2] * dl[b][0];
      if (LELAS && P.pS0 != nullptr)
#pragma unroll
        for (int i = 0; i < 6; i++) p0[i] += Nb * __ldg(P.pS0 + 6 * n + i);
    }
    const double divD = lambda * (ed[0] + ed[1] + ed[2]);
    double* r = sgp[warp][el][g];
#pragma unroll
    for (int b = 0; b < ENON; b++)
#pragma unroll
      for (int i = 0; i < 3; i++) r[3 * b + i] = Nx[b][i];
    r[24] = divD + 2.0 * mu * ed[0]; r[25] = divD + 2.0 * mu * ed[1]; r[26] = divD + 2.0 * mu * ed[2];
    r[27] = mu * ed[3]; r[28] = mu * ed[4]; r[29] = mu * ed[5];
#pragma unroll
    for (int i = 0; i < 6; i++) r[30 + i] = p0[i];
    r[36] = ud[0]; r[37] = ud[1]; r[38] = ud[2];
    r[39] = w;
  }
  __syncwarp();
  const unsigned amask = __ballot_sync(0xffffffffu, active);     // lanes with an element (whole elements: 8 lanes each)
  if (amask == 0u) return;
  // lanes without an element stay for the transposed scatter of the dof = 3 path (they carry no rows but move other lanes' entries)
  if (!active && P.dof != 3) return;
  double K[ENON][3][3], lR[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int b = 0; b < ENON; b++)
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) K[b][i][j] = 0.0;
#pragma unroll 1
  for (int g = 0; g < (active ? ENON : 0); g++) {
    const double* r = sgp[warp][el][g];
    const double w = r[39], wl = w * T1c * mu;
    const double Na = P.N[g][a];
    const double Nxa[3] = {r[3 * a], r[3 * a + 1], r[3 * a + 2]};
    double S0 = r[24], S1 = r[25], S2 = r[26], S3 = r[27], S4 = r[28], S5 = r[29];
    if (LELAS) {
      if (P.pSn != nullptr) {           // l_elas.cpp:321-338, 130-140: the accumulators see the stress WITHOUT the prestress
        const double wN = w * Na;
        add64<true>(P.pSa + na, wN);
        add64<true>(P.pSn + 6 * (size_t)na + 0, wN * S0); add64<true>(P.pSn + 6 * (size_t)na + 1, wN * S1);
        add64<true>(P.pSn + 6 * (size_t)na + 2, wN * S2); add64<true>(P.pSn + 6 * (size_t)na + 3, wN * S3);
        add64<true>(P.pSn + 6 * (size_t)na + 4, wN * S4); add64<true>(P.pSn + 6 * (size_t)na + 5, wN * S5);
      }
      if (P.pS0 != nullptr) { S0 += r[30]; S1 += r[31]; S2 += r[32]; S3 += r[33]; S4 += r[34]; S5 += r[35]; }
    }
    lR[0] += w * (rho * Na * r[36] + Nxa[0] * S0 + Nxa[1] * S3 + Nxa[2] * S5);
    lR[1] += w * (rho * Na * r[37] + Nxa[0] * S3 + Nxa[1] * S1 + Nxa[2] * S4);
    lR[2] += w * (rho * Na * r[38] + Nxa[0] * S5 + Nxa[1] * S4 + Nxa[2] * S2);
#pragma unroll
    for (int b = 0; b < ENON; b++) {
      const double Nxb[3] = {r[3 * b], r[3 * b + 1], r[3 * b + 2]};
      const double NxdNx = Nxa[0] * Nxb[0] + Nxa[1] * Nxb[1] + Nxa[2] * Nxb[2];
      const double T1 = amd * Na * P.N[g][b] / mu + NxdNx;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
          K[b][i][j] += wl * ((i == j ? T1 + (1.0 + lDm) * Nxa[i] * Nxb[i] : lDm * Nxa[i] * Nxb[j] + Nxa[j] * Nxb[i]));
    }
  }
  if (active)
#pragma unroll
    for (int i = 0; i < 3; i++) add64<ATOMIC>(P.R + (size_t)DOF * na + i, lR[i]);
  const int* sl = P.slot + (size_t)e * ENON * ENON + a * ENON;
  if (DOF == 3) {
    // dof = 3: a block is 9 contiguous doubles.  The 32 blocks (a, b) of the warp for one b go through a 32 x 9 tile (aliased onto the
    // Gauss-point records, which are dead now) and out as 9 warp-wide adds over contiguous runs: 3 L2 sectors per block instead of 9.
    double* tile = &sgp[warp][0][0][0];
    __syncwarp();                                        // every lane of the warp is done with the records
#pragma unroll
    for (int b = 0; b < ENON; b++) {
      const int myslot = active ? __ldg(sl + b) : -1;
      if (active)
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) tile[lane * 9 + 3 * i + j] = K[b][i][j];
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 9; it++) {
        const int p = it * 32 + lane, src = p / 9, i = p - 9 * src;
        const int s_ = __shfl_sync(0xffffffffu, myslot, src);
        if (s_ >= 0) add64<ATOMIC>(P.Val + (size_t)9 * s_ + i, tile[src * 9 + i]);
      }
      __syncwarp();
    }
    return;
  }
#pragma unroll
  for (int b = 0; b < ENON; b++) {
    double* v = P.Val + (size_t)DOF * DOF * sl[b];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) add64<ATOMIC>(v + DOF * i + j, K[b][i][j]);
  }
}

// ---- mesh-motion / linear-elasticity equation on linear tetrahedra: one thread per element ------------------------
// l_elas_3d with constant gradients (l_elas.cpp:249-365): the strain and stress are constant over the element, only the
// inertia term sees N_a(g).  With the weights w = w_g (mesh equation: Jacobian-free, mesh.cpp:122) or w_g Jac (lElas),
// W = sum w, m_a = sum w N_a, M_ab = sum w N_a N_b:
//     lR(i,a)   = rho (-f_i m_a + sum_b M_ab q_b(i)) + W (S grad N_a)_i
//     lK(ij,ab) = T1c [ delta_ij (amd M_ab + mu W grad N_a . grad N_b) + mu W (lDm Nx_a(i) Nx_b(j) + Nx_a(j) Nx_b(i)) ]
// (for i = j the last bracket is (1 + lDm) Nx_a(i) Nx_b(i), the same expression).  K_ba = K_ab^T: 10 blocks, scattered like
// the solid's.
template <bool ATOMIC, bool LELAS>
__global__ void __launch_bounds__(STET_THREADS)
assemble_mesh_tet4_kernel(const __grid_constant__ StructArgs P, const double* __restrict__ Do)
{
  __shared__ double s_tile[STET_THREADS / 32][32 * 9];
  __shared__ int s_slot[STET_THREADS / 32][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* tile = s_tile[warp];
  int* tsl = s_slot[warp];
  const long long idx = (long long)P.e0 + (long long)blockIdx.x * STET_THREADS + threadIdx.x;
  bool active = idx < P.e1;
  int e = 0;
  if (active) e = P.perm ? P.perm[idx] : (int)idx;
  int iD = 0;
  if (active) {
    for (int d = 0; d < P.nDmn; d++) {
      iD = d;
      if (P.dmn[d].Id == -1) break;
      if (P.eId != nullptr && ((P.eId[e] >> P.dmn[d].Id) & 1)) break;
    }
    if (!P.dmn[iD].isStruct) active = false;
  }
  const StructDmn& dm = P.dmn[iD];
  const int DOF = P.dof, DD = DOF * DOF, is = P.s;
  const double elM = dm.C10, nu = dm.C01, rho = dm.rho;        // elasticity_modulus / poisson_ratio travel in C10 / C01
  const double lambda = elM * nu / (1.0 + nu) / (1.0 - 2.0 * nu);
  const double mu = elM * 0.5 / (1.0 + nu);
  const double lDm = lambda / mu;
  const double T1c = P.af * P.beta * P.dt * P.dt;
  const double amd = P.am / T1c * rho;
  int sl[16];
#pragma unroll
  for (int k = 0; k < 16; k++) sl[k] = -1;
  double Nx[4][3], M2[4][4], W = 0.0;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    Nx[a][0] = Nx[a][1] = Nx[a][2] = 0.0;
#pragma unroll
    for (int b = 0; b < 4; b++) M2[a][b] = 0.0;
  }
  if (active) {
    int node[4];
    const int4 nn = __ldg(reinterpret_cast<const int4*>(P.IEN) + e);
    node[0] = nn.x; node[1] = nn.y; node[2] = nn.z; node[3] = nn.w;
    const int4* sp = reinterpret_cast<const int4*>(P.slot) + (size_t)e * 4;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int4 v = __ldg(sp + k);
      sl[4 * k] = v.x; sl[4 * k + 1] = v.y; sl[4 * k + 2] = v.z; sl[4 * k + 3] = v.w;
    }
    double xl[4][3], dl[4][3], ql[4][3];
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const size_t n = (size_t)node[a];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double dol = LELAS ? 0.0 : __ldg(Do + (size_t)P.tDof * n + is + i);
        xl[a][i] = __ldg(P.x + 3 * n + i) + dol;
        dl[a][i] = __ldg(P.Dg + (size_t)P.tDof * n + is + i) - dol;
        ql[a][i] = __ldg(P.Ag + (size_t)P.tDof * n + is + i) - (LELAS ? __ldg(P.Bf + 3 * n + i) : 0.0);
      }
    }
    const double Jac = gnn3<4>(P.Nxi[0], xl, Nx);
    Tet4Mom q;
    tet4_moments(P.w, &P.N[0][0], MAX_ENON, LELAS ? Jac : 1.0, q);
    W = q.W;
    double S[6];
    lelas_tet4_stress(lambda, mu, Nx, dl, S);
#pragma unroll
    for (int a = 0; a < 4; a++) {
      double r[3];
      lelas_tet4_residual(rho, dm.f, q, a, Nx[a], S, ql, r);
#pragma unroll
      for (int i = 0; i < 3; i++) add64<ATOMIC>(P.R + (size_t)DOF * node[a] + i, r[i]);
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) M2[a][b] = q.M2[a][b];
  }
  const double c0 = T1c * amd, c1 = T1c * mu * W;
#pragma unroll
  for (int b = 0; b < 4; b++)
#pragma unroll
    for (int a = 0; a <= b; a++) {
      double K[3][3];
      lelas_tet4_block(c0 * M2[a][b], c1, lDm, Nx[a], Nx[b], K);
      warp_scatter_block_pair<ATOMIC>(tile, tsl, lane, K, sl[4 * a + b], (a != b) ? sl[4 * b + a] : -1, a != b, P.Val, DD, DOF);
    }
}

// Active stress flag and CANN table of one solid domain (shared by the struct and the ustruct argument blocks).
int fill_solid_extras(svb200_ctx* ctx, const svb200_dmnparams& p, StructDmn& o, CannRow* table, int* used)
{
  o.active = (p.active_stress != 0 && ctx->d_Ya != nullptr) ? 1 : 0;
  if (p.active_stress != 0 && ctx->d_Ya == nullptr) {
    set_error("svb200_assemble: the domain has an active-stress model but svb200_set_active_tension was not called");
    return SVB200_ERR_INVALID;
  }
  const bool dirs = (o.isoType == SVB200_ISO_GUCCIONE || o.isoType == SVB200_ISO_HO || o.isoType == SVB200_ISO_HO_MA);
  if (o.active && !dirs && ctx->ya_sn_positive) {
    // mat_models.cpp:334-340
    set_error("Directional distribution of active stress (eta_s > 0 or eta_n > 0) is only supported for Guccione, "
              "Holzapfel-Ogden (HO), and Holzapfel-Ogden Modified Anisotropy (HO-ma) models.");
    return SVB200_ERR_INVALID;
  }
  o.cann_off = 0;
  o.cann_rows = 0;
  if (o.isoType == SVB200_ISO_CANN) {
    const int n = p.cann_rows;
    if (n < 1 || n > SVB200_CANN_MAX_ROWS || *used + n > MAX_CANN_ROWS) {
      set_error("svb200_assemble: CANN parameter table must have 1..16 rows (32 over all domains)");
      return SVB200_ERR_INVALID;
    }
    o.cann_off = *used;
    o.cann_rows = n;
    for (int r = 0; r < n; r++) {
      CannRow& q = table[*used + r];
      q.inv = p.cann_inv[r]; q.a0 = p.cann_act[r][0]; q.a1 = p.cann_act[r][1]; q.a2 = p.cann_act[r][2];
      q.w0 = p.cann_w[r][0]; q.w1 = p.cann_w[r][1]; q.w2 = p.cann_w[r][2];
      if (q.inv < 1 || q.inv > 9 || q.a0 < 1 || q.a0 > 3 || q.a1 < 1 || q.a1 > 2 || q.a2 < 1 || q.a2 > 3) {
        set_error("svb200_assemble: CANN row out of range (invariant 1..9, activation functions (1..3, 1..2, 1..3))");
        return SVB200_ERR_INVALID;
      }
    }
    *used += n;
  }
  return SVB200_OK;
}

int fill_struct_args(svb200_ctx* ctx, const Mesh& m, const svb200_eqparams* eq, const svb200_dmnparams* dmn, int nDmn, StructArgs& A)
{
  SVB_REQUIRE(nDmn >= 1 && nDmn <= MAX_DMN, "svb200_assemble: between 1 and 8 domains are supported");
  SVB_REQUIRE(eq->dof == ctx->dof && (eq->dof == 3 || eq->dof == 4), "svb200_assemble: the solid needs dof = 3 (struct) or 4 (FSI)");
  SVB_REQUIRE(eq->tDof == ctx->tDof && ctx->d_Dg && ctx->d_Yg && ctx->d_Ag, "svb200_assemble: state not set or tDof mismatch");
  SVB_REQUIRE(eq->s >= 0 && eq->s + 3 <= eq->tDof, "svb200_assemble: eq.s out of range");
  SVB_REQUIRE(ctx->d_x, "svb200_assemble: coordinates not set");
  memset(&A, 0, sizeof(A));
  A.tab = m.d_gtab;
  A.lShpF = (m.eNoN == 4 || m.eNoN == 6) ? 1 : 0;
  A.IEN = m.d_IEN; A.eId = m.d_eId; A.slot = m.d_slot; A.perm = nullptr; A.fN = m.d_fN;
  A.x = ctx->d_x; A.Ag = ctx->d_Ag; A.Yg = ctx->d_Yg; A.Dg = ctx->d_Dg; A.Bf = ctx->d_Bf;
  A.Ya = ctx->d_Ya;
  A.pS0 = ctx->d_pS0;
  if (eq->reserved & SVB200_EQ_PRESTRESS) {
    // pstEq: the accumulators live until the next svb200_alloc (which zeroes them like Integrator::initiator does)
    const size_t n = std::max<size_t>((size_t)ctx->nNo, 1);
    if (!ctx->d_pSn) {
      SVB_CUDA(cudaMalloc(&ctx->d_pSn, sizeof(double) * 7 * n));
      SVB_CUDA(cudaMemsetAsync(ctx->d_pSn, 0, sizeof(double) * 7 * n, ctx->stream));
    }
    A.pSn = ctx->d_pSn;
    A.pSa = ctx->d_pSn + 6 * n;
  }
  A.R = ctx->d_R; A.Val = ctx->d_Val;
  A.e0 = 0; A.e1 = m.nEl;
  A.tDof = eq->tDof; A.dof = eq->dof; A.s = eq->s; A.nFn = m.nFn; A.nDmn = nDmn; A.nG = m.nG;
  int cann_used = 0;
  A.atomic = (eq->scatter == SVB200_SCATTER_ATOMIC);
  A.dt = eq->dt; A.af = eq->af; A.am = eq->am; A.gam = eq->gam; A.beta = eq->beta;
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
    StructDmn& o = A.dmn[d];
    o.rho = dmn[d].rho;
    for (int k = 0; k < 3; k++) o.f[k] = dmn[d].f[k];
    o.dmp = dmn[d].dmp;
    o.Kpen = dmn[d].Kpen; o.C10 = dmn[d].C10; o.C01 = dmn[d].C01;
    o.bff = dmn[d].bff; o.bss = dmn[d].bss; o.bfs = dmn[d].bfs;
    o.st_a = dmn[d].st_a; o.st_b = dmn[d].st_b; o.aff = dmn[d].aff; o.ass = dmn[d].ass; o.afs = dmn[d].afs;
    o.kap = dmn[d].kap; o.khs = dmn[d].khs;
    o.isoType = dmn[d].isoType; o.volType = dmn[d].volType;
    o.visc_mu = dmn[d].solid_visc_mu;
    o.viscType = SVB200_SOLID_VISC_NONE;
    if (dmn[d].phys == SVB200_PHYS_STRUCT && dmn[d].solid_visc_mu != 0.0)
      o.viscType = (dmn[d].solidViscType == SVB200_SOLID_VISC_POTENTIAL) ? SVB200_SOLID_VISC_POTENTIAL : SVB200_SOLID_VISC_NEWTONIAN;
    o.Id = dmn[d].Id;
    o.isStruct = (dmn[d].phys == SVB200_PHYS_STRUCT);
    SVB_REQUIRE(o.Id >= -1 && o.Id < 31, "svb200_assemble: domain Id out of range");
    if (o.isStruct) {
      SVB_REQUIRE(o.isoType >= SVB200_ISO_NHK && o.isoType <= SVB200_ISO_CANN, "svb200_assemble: constitutive model not implemented");
      int rc = fill_solid_extras(ctx, dmn[d], o, A.cann, &cann_used);
      if (rc) return rc;
      // compute_pk2cc throws for the fibre models without two fibre families (mat_models.cpp:472-474, 514-516, 584-586)
      const bool fibres = (m.nFn == 2 && m.d_fN);
      if (o.isoType == SVB200_ISO_GUCCIONE && !fibres) {
        set_error("[compute_pk2cc] Min fiber directions not defined for Guccione material model.");
        return SVB200_ERR_INVALID;
      }
      if (o.isoType == SVB200_ISO_HGO && !fibres) {
        set_error("[compute_pk2cc] Min fiber directions not defined for HGO material model.");
        return SVB200_ERR_INVALID;
      }
      if ((o.isoType == SVB200_ISO_HO || o.isoType == SVB200_ISO_HO_MA) && !fibres) {
        set_error("[compute_pk2cc] Min fiber directions not defined for Holzapfel material model.");
        return SVB200_ERR_INVALID;
      }
    }
    whole |= (o.Id == -1);
  }
  if (!whole && !m.d_eId) { set_error("eId is not allocated"); return SVB200_ERR_INVALID; }
  return SVB200_OK;
}

template <int ENON, int NG>
static int launch_one(svb200_ctx* ctx, const StructArgs& A)
{
  constexpr int LPE = ENON > NG ? ENON : NG;
  constexpr int EPB = (STRUCT_THREADS / 32) * (32 / LPE);   // elements per CTA
  const long long n = (long long)A.e1 - A.e0;
  if (n <= 0) return SVB200_OK;
  const unsigned blocks = (unsigned)((n + EPB - 1) / EPB);
  constexpr size_t smem = sizeof(double) * ((size_t)NG * ENON * 4 + NG + (size_t)EPB * struct_per_el(ENON, NG));
  static_assert(smem <= 227 * 1024, "element does not fit in shared memory");
  constexpr bool LINEAR = (NG == ENON && (ENON == 4 || ENON == 8));     // TET4 / HEX8: the solid-viscosity kernels exist
  constexpr int EV = LINEAR ? ENON : 8;
  constexpr int EPBV = (VISC_THREADS / 32) * (32 / EV);
  constexpr size_t smemV = sizeof(double) * (size_t)EPBV * visc_per_el(EV);
  // the hot instantiations (TET4 / HEX8 without viscosity) exist with and without the CANN branch (struct_elem.cuh: pk2cc_voigt<CANN>);
  // everything else is built with it
  constexpr bool NC = !LINEAR;                  // "no-CANN twin" collapses onto the CANN build for the quadratic elements
  bool cannM = false;
  for (int d = 0; d < A.nDmn; d++) cannM |= (A.dmn[d].isStruct && (A.dmn[d].isoType == SVB200_ISO_CANN || A.dmn[d].active));
  static bool configured = false;
  if (!configured) {
    SVB_CUDA(cudaFuncSetAttribute(assemble_struct_kernel<ENON, NG, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SVB_CUDA(cudaFuncSetAttribute(assemble_struct_kernel<ENON, NG, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SVB_CUDA(cudaFuncSetAttribute(assemble_struct_kernel<ENON, NG, true, false, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SVB_CUDA(cudaFuncSetAttribute(assemble_struct_kernel<ENON, NG, false, false, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if constexpr (LINEAR) {
      SVB_CUDA(cudaFuncSetAttribute(assemble_struct_kernel<ENON, NG, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      SVB_CUDA(cudaFuncSetAttribute(assemble_struct_kernel<ENON, NG, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      SVB_CUDA(cudaFuncSetAttribute(assemble_struct_visc_kernel<EV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemV));
      SVB_CUDA(cudaFuncSetAttribute(assemble_struct_visc_kernel<EV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemV));
    }
    configured = true;
  }
  if (A.nG != NG) { set_error("svb200: the solid kernel expects the reference's quadrature rule for this element type"); return SVB200_ERR_UNSUPPORTED; }
  bool visc = false;
  for (int d = 0; d < A.nDmn; d++) visc |= (A.dmn[d].isStruct && A.dmn[d].viscType != SVB200_SOLID_VISC_NONE);
  if (!visc) {
    if (cannM) {
      if (A.atomic) assemble_struct_kernel<ENON, NG, true, false, true><<<blocks, STRUCT_THREADS, smem, ctx->stream>>>(A);
      else assemble_struct_kernel<ENON, NG, false, false, true><<<blocks, STRUCT_THREADS, smem, ctx->stream>>>(A);
    } else {
      if (A.atomic) assemble_struct_kernel<ENON, NG, true, false, NC><<<blocks, STRUCT_THREADS, smem, ctx->stream>>>(A);
      else assemble_struct_kernel<ENON, NG, false, false, NC><<<blocks, STRUCT_THREADS, smem, ctx->stream>>>(A);
    }
    ctx->launches++;
  } else {
    if constexpr (LINEAR) {
      const unsigned blocksV = (unsigned)((n + EPBV - 1) / EPBV);
      if (A.atomic) {
        assemble_struct_kernel<ENON, NG, true, true, true><<<blocks, STRUCT_THREADS, smem, ctx->stream>>>(A);
        assemble_struct_visc_kernel<EV, true><<<blocksV, VISC_THREADS, smemV, ctx->stream>>>(A);
      } else {
        assemble_struct_kernel<ENON, NG, false, true, true><<<blocks, STRUCT_THREADS, smem, ctx->stream>>>(A);
        assemble_struct_visc_kernel<EV, false><<<blocksV, VISC_THREADS, smemV, ctx->stream>>>(A);
      }
      ctx->launches += 2;
    } else {
      set_error("svb200_assemble: solid viscosity is implemented for TET4 and HEX8 meshes");
      return SVB200_ERR_UNSUPPORTED;
    }
  }
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

static int launch_tet4(svb200_ctx* ctx, const StructArgs& A)
{
  const long long n = (long long)A.e1 - A.e0;
  if (n <= 0) return SVB200_OK;
  if (A.nG != 4) { set_error("svb200: the TET4 solid kernel expects the 4-point rule"); return SVB200_ERR_UNSUPPORTED; }
  const unsigned blocks = (unsigned)((n + STET_THREADS - 1) / STET_THREADS);
  bool cannM = false;
  for (int d = 0; d < A.nDmn; d++) cannM |= (A.dmn[d].isStruct && (A.dmn[d].isoType == SVB200_ISO_CANN || A.dmn[d].active));
  if (cannM) {
    if (A.atomic) assemble_struct_tet4_kernel<true, true><<<blocks, STET_THREADS, 0, ctx->stream>>>(A);
    else assemble_struct_tet4_kernel<false, true><<<blocks, STET_THREADS, 0, ctx->stream>>>(A);
  } else {
    if (A.atomic) assemble_struct_tet4_kernel<true, false><<<blocks, STET_THREADS, 0, ctx->stream>>>(A);
    else assemble_struct_tet4_kernel<false, false><<<blocks, STET_THREADS, 0, ctx->stream>>>(A);
  }
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

int run_assemble_struct(svb200_ctx* ctx, const Mesh& m, const svb200_eqparams* eq, const svb200_dmnparams* dmn, int nDmn)
{
  StructArgs A;
  int rc = fill_struct_args(ctx, m, eq, dmn, nDmn, A);
  if (rc) return rc;
  // linear tets without solid viscosity: the one-thread-per-element kernel (SVB200_STRUCT_GENERAL=1 keeps the general
  // two-phase kernel, used by the tests to cross-check the two)
  bool visc = false;
  for (int d = 0; d < A.nDmn; d++) visc |= (A.dmn[d].isStruct && A.dmn[d].viscType != SVB200_SOLID_VISC_NONE);
  static const bool force_general = (getenv("SVB200_STRUCT_GENERAL") != nullptr && atoi(getenv("SVB200_STRUCT_GENERAL")) != 0);
  // (the closed-form kernel has no prestress terms: S0 and pSl vary over the Gauss points with N_a)
  const bool tet4 = (m.eNoN == 4 && !visc && !force_general && A.pS0 == nullptr && A.pSn == nullptr);
  auto launch = [&](const StructArgs& B) {
    if (tet4) return launch_tet4(ctx, B);
    switch (m.eNoN * 100 + m.nG) {
      case 404: return launch_one<4, 4>(ctx, B);
      case 808: return launch_one<8, 8>(ctx, B);
      case 606: return launch_one<6, 6>(ctx, B);
      case 1015: return launch_one<10, 15>(ctx, B);
      case 2027: return launch_one<20, 27>(ctx, B);
      case 2727: return launch_one<27, 27>(ctx, B);
    }
    set_error("svb200_assemble: the solid covers TET4, HEX8, WDG, TET10, HEX20 and HEX27 meshes with the reference's quadrature rules");
    return (int)SVB200_ERR_UNSUPPORTED;
  };
  if (A.atomic) return launch(A);
  A.perm = m.d_color_perm;
  for (size_t c = 0; c + 1 < m.color_off.size(); c++) {
    A.e0 = m.color_off[c];
    A.e1 = m.color_off[c + 1];
    rc = launch(A);
    if (rc) return rc;
  }
  return SVB200_OK;
}

template <int ENON>
static int launch_mesh(svb200_ctx* ctx, const StructArgs& A, const double* Do)
{
  constexpr int EPB = 4 * (32 / ENON);
  const long long n = (long long)A.e1 - A.e0;
  if (n <= 0) return SVB200_OK;
  const unsigned blocks = (unsigned)((n + EPB - 1) / EPB);
  static const bool lane_rows = getenv("SVB200_MESH_HEX8_LEGACY") != nullptr;      // A/B: the lane-per-row kernel for HEX8
  if (ENON == 8 && !lane_rows) {
    if (Do == nullptr) {
      if (A.atomic) assemble_mesh_hex8_kernel<true, true><<<blocks, 128, 0, ctx->stream>>>(A, nullptr);
      else assemble_mesh_hex8_kernel<false, true><<<blocks, 128, 0, ctx->stream>>>(A, nullptr);
    } else {
      if (A.atomic) assemble_mesh_hex8_kernel<true, false><<<blocks, 128, 0, ctx->stream>>>(A, Do);
      else assemble_mesh_hex8_kernel<false, false><<<blocks, 128, 0, ctx->stream>>>(A, Do);
    }
  } else if (Do == nullptr) {      // linear-elasticity equation
    if (A.atomic) assemble_mesh_kernel<ENON, true, true><<<blocks, 128, 0, ctx->stream>>>(A, nullptr);
    else assemble_mesh_kernel<ENON, false, true><<<blocks, 128, 0, ctx->stream>>>(A, nullptr);
  } else {
    if (A.atomic) assemble_mesh_kernel<ENON, true, false><<<blocks, 128, 0, ctx->stream>>>(A, Do);
    else assemble_mesh_kernel<ENON, false, false><<<blocks, 128, 0, ctx->stream>>>(A, Do);
  }
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

int run_assemble_mesh(svb200_ctx* ctx, const Mesh& m, const svb200_eqparams* eq, const svb200_dmnparams* dmn, int nDmn)
{
  const bool lelas = (eq->phys == SVB200_PHYS_LELAS);
  SVB_REQUIRE(lelas || ctx->d_Do, "svb200_assemble: the mesh equation needs the old displacement (svb200_set_old_disp)");
  SVB_REQUIRE(eq->dof == 3 && ctx->dof == 3, "svb200_assemble: the mesh / linear-elasticity equation has dof = 3");
  const int key = m.eNoN * 100 + m.nG;
  SVB_REQUIRE(key == 404 || key == 808 || key == 606 || key == 1015 || key == 2027 || key == 2727,
              "svb200_assemble: the mesh / linear-elasticity equation covers TET4, HEX8, WDG, TET10, HEX20 and HEX27 meshes with the reference's quadrature rules");
  SVB_REQUIRE(m.eNoN <= MAX_ENON || m.d_gtab, "svb200_assemble: element tables missing");
  // reuse the solid argument block: mark mesh domains as the ones to assemble, E / nu travel in C10 / C01
  std::vector<svb200_dmnparams> d(dmn, dmn + nDmn);
  for (auto& q : d) {
    const bool isMesh = (q.phys == (lelas ? SVB200_PHYS_LELAS : SVB200_PHYS_MESH));
    q.phys = isMesh ? SVB200_PHYS_STRUCT : SVB200_PHYS_FLUID;
    q.isoType = SVB200_ISO_NHK;
    q.solid_visc_mu = 0.0;
    q.C10 = q.E;
    q.C01 = q.nu;
  }
  StructArgs A;
  int rc = fill_struct_args(ctx, m, eq, d.data(), nDmn, A);
  if (rc) return rc;
  if (!lelas) { A.pS0 = nullptr; A.pSn = nullptr; A.pSa = nullptr; }     // construct_mesh passes a zero pS0l (mesh.cpp)
  const double* Do = lelas ? nullptr : ctx->d_Do;
  static const bool force_general = (getenv("SVB200_STRUCT_GENERAL") != nullptr && atoi(getenv("SVB200_STRUCT_GENERAL")) != 0);
  auto launch_t4 = [&](const StructArgs& B) {
    const long long n = (long long)B.e1 - B.e0;
    if (n <= 0) return (int)SVB200_OK;
    const unsigned blocks = (unsigned)((n + STET_THREADS - 1) / STET_THREADS);
    if (Do == nullptr) {
      if (B.atomic) assemble_mesh_tet4_kernel<true, true><<<blocks, STET_THREADS, 0, ctx->stream>>>(B, nullptr);
      else assemble_mesh_tet4_kernel<false, true><<<blocks, STET_THREADS, 0, ctx->stream>>>(B, nullptr);
    } else {
      if (B.atomic) assemble_mesh_tet4_kernel<true, false><<<blocks, STET_THREADS, 0, ctx->stream>>>(B, Do);
      else assemble_mesh_tet4_kernel<false, false><<<blocks, STET_THREADS, 0, ctx->stream>>>(B, Do);
    }
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) { set_error("assemble_mesh_tet4_kernel launch failed"); return (int)SVB200_ERR_CUDA; }
    return (int)SVB200_OK;
  };
  auto launch = [&](const StructArgs& B) {
    if (m.eNoN == 4 && m.nG == 4 && !force_general && !(lelas && (A.pS0 || A.pSn))) return launch_t4(B);
    switch (m.eNoN) {
      case 8: return launch_mesh<8>(ctx, B, Do);
      case 6: return launch_mesh<6>(ctx, B, Do);
      case 10: return launch_mesh<10>(ctx, B, Do);
      case 20: return launch_mesh<20>(ctx, B, Do);
      case 27: return launch_mesh<27>(ctx, B, Do);
      default: return launch_mesh<4>(ctx, B, Do);
    }
  };
  if (A.atomic) return launch(A);
  A.perm = m.d_color_perm;
  for (size_t c = 0; c + 1 < m.color_off.size(); c++) {
    A.e0 = m.color_off[c];
    A.e1 = m.color_off[c + 1];
    rc = launch(A);
    if (rc) return rc;
  }
  return SVB200_OK;
}

}  // namespace svb
