// ustruct_elem.cuh — Gauss-point algebra of the mixed velocity-pressure solid ("ustruct", Liu & Marsden 2018),
// __host__ __device__ so that the CPU suite checks it against the compiled reference before a GPU is involved.
//
//   ustruct::ustruct_3d_m   Code/Source/solver/ustruct.cpp:1165-1591   momentum rows, blocks 0..11 of lK and 0..8 of lKd
//   ustruct::ustruct_3d_c   Code/Source/solver/ustruct.cpp:629-871     continuity row,  blocks 12..15 of lK and 9..11 of lKd
//   mat_models::g_vol_pen   Code/Source/solver/mat_models.cpp:1511-1560   rho(p), beta(p) and their derivatives
//   mat_models::compute_tau Code/Source/solver/mat_models.cpp:1470-1493   tauM, tauC
//
// With equal-order (VMS) elements both Gauss loops of construct_usolid (ustruct.cpp:304-396) visit the same points
// with the same shape functions (fs[0] == fs[1], Nq == Nw), and without active strain Ja = 1, so everything the two
// reference routines compute before their node loops is shared: one UGP per Gauss point serves all 16 + 12 entries.
// lK is the tangent with respect to (v, p) [4 x 4 per node pair, row-major], lKd the tangent with respect to the
// displacement [4 x 3]; the reference adds (af/am) lKd into the velocity columns of lK (ustruct.cpp:1457, 1463, ...).
#pragma once
#include "struct_elem.cuh"

namespace svb {

struct UstructDmn {
  StructDmn st;              // rho = solid_density, f, stM parameters; st.Kpen is the ustruct bulk modulus of g_vol_pen
  double E, nu, ctM, ctC;    // elasticity_modulus, poisson_ratio, ctau_M, ctau_C (compute_tau)
};

struct UGP {
  double xiX[3][3];          // d xi / d X of the reference map: Nx_a = xiX^T Nxi_a
  double F[3][3], Fi[3][3];
  double S[3][3];            // isochoric (+ viscous) 2nd Piola-Kirchhoff stress
  double Dm[6][6];
  double Pdev[3][3];         // F S
  double VxFi[3][3];         // grad_X v F^-1
  double vd[3], PxFi[3], rM[3];
  double w, J, rho, beta, drho, dbeta, tauM, tauC, rC, rCl, pd;
};

struct UNode {
  double N, Nx[3], NxFi[3], VxNx[3], rMNx;
};

// nn::gnn (nn.cpp:862-899): xiX and the Jacobian of the reference map at one Gauss point.
template <int ENON>
SVB_HD double ustruct_xiX(const double Nxi[][3], const double xl[][3], double xiX[3][3])
{
  double xXi[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
  for (int a = 0; a < ENON; a++)
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int k = 0; k < 3; k++) xXi[i][k] += xl[a][i] * Nxi[a][k];
  const double Jac = xXi[0][0] * xXi[1][1] * xXi[2][2] + xXi[0][1] * xXi[1][2] * xXi[2][0] + xXi[0][2] * xXi[1][0] * xXi[2][1] -
                     xXi[0][0] * xXi[1][2] * xXi[2][1] - xXi[0][1] * xXi[1][0] * xXi[2][2] - xXi[0][2] * xXi[1][1] * xXi[2][0];
  const double iJ = 1.0 / Jac;
  xiX[0][0] = (xXi[1][1] * xXi[2][2] - xXi[1][2] * xXi[2][1]) * iJ;
  xiX[0][1] = (xXi[2][1] * xXi[0][2] - xXi[2][2] * xXi[0][1]) * iJ;
  xiX[0][2] = (xXi[0][1] * xXi[1][2] - xXi[0][2] * xXi[1][1]) * iJ;
  xiX[1][0] = (xXi[1][2] * xXi[2][0] - xXi[1][0] * xXi[2][2]) * iJ;
  xiX[1][1] = (xXi[2][2] * xXi[0][0] - xXi[2][0] * xXi[0][2]) * iJ;
  xiX[1][2] = (xXi[0][2] * xXi[1][0] - xXi[0][0] * xXi[1][2]) * iJ;
  xiX[2][0] = (xXi[1][0] * xXi[2][1] - xXi[1][1] * xXi[2][0]) * iJ;
  xiX[2][1] = (xXi[2][0] * xXi[0][1] - xXi[2][1] * xXi[0][0]) * iJ;
  xiX[2][2] = (xXi[0][0] * xXi[1][1] - xXi[0][1] * xXi[1][0]) * iJ;
  return Jac;
}

SVB_HD void ustruct_grad(const double xiX[3][3], const double Nxi[3], double Nx[3])
{
#pragma unroll
  for (int i = 0; i < 3; i++) Nx[i] = Nxi[0] * xiX[0][i] + Nxi[1] * xiX[1][i] + Nxi[2] * xiX[2][i];
}

// mat_models::g_vol_pen (mat_models.cpp:1511-1560) with Ja = 1: rho(p), beta(p) and their derivatives.
SVB_HD void ustruct_vol_pen(const StructDmn& st, double p, double& rho, double& beta, double& drho, double& dbeta)
{
  const double Kp = st.Kpen;
  rho = st.rho; beta = 0.0; drho = 0.0; dbeta = 0.0;
  if (is_zero(Kp)) return;
  if (st.volType == SVB200_VOL_QUAD) {
    const double r1 = 1.0 / (Kp - p);
    rho = rho * Kp * r1; beta = r1; drho = rho * r1; dbeta = r1 * r1;
  } else if (st.volType == SVB200_VOL_ST91) {
    const double r1 = rho / Kp, r2 = sqrt(p * p + Kp * Kp);
    rho = r1 * (p + r2); beta = 1.0 / r2; drho = rho * beta; dbeta = -beta * p / (p * p + Kp * Kp);
  } else if (st.volType == SVB200_VOL_M94) {
    const double r1 = rho / Kp, r2 = Kp + p;
    rho = r1 * r2; beta = 1.0 / r2; drho = r1; dbeta = -beta * beta;
  }
}

// mat_models::compute_tau (mat_models.cpp:1470-1493).
SVB_HD void ustruct_tau(const UstructDmn& dm, double Je, double detF, double& tauM, double& tauC)
{
  const double he = 0.5 * pow(Je, 1.0 / 3.0);
  const double rho0 = dm.st.rho, mu = 0.5 * dm.E / (1.0 + dm.nu);
  double c;
  if (is_zero(dm.nu - 0.5)) c = sqrt(mu / rho0);
  else c = sqrt((2.0 * mu * dm.nu / (1.0 - 2.0 * dm.nu) + 2.0 * mu) / rho0);
  tauM = dm.ctM * (he / c) * (detF / rho0);
  tauC = dm.ctC * (he * c) * (rho0 / detF);
}

// Everything ustruct_3d_m / ustruct_3d_c evaluate before their node loops.
//   ql[a] = al(i..k,a) - bfl(:,a);  vl, dl: nodal velocity / displacement;  pl, pdl: nodal pressure and its rate.
// Returns 0, 1 for an unsupported constitutive model.
template <int ENON, bool CANN = true>
SVB_HD int ustruct_gauss_point(const UstructDmn& dm, double dt, double af_eq, double am, double gam, double wg, const double N[],
                               const double Nxi[][3], const double xl[][3], const double ql[][3], const double vl[][3],
                               const double dl[][3], const double pl[], const double pdl[], const double fN[2][3], UGP& q,
                               ViscGP* gu = nullptr, ViscGP* gv = nullptr, const double* ya = nullptr, const CannRow* cann = nullptr,
                               int nFn = 2)
{
  const double Je = ustruct_xiX<ENON>(Nxi, xl, q.xiX);
  q.w = wg * Je;
  double vx[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, px[3] = {0, 0, 0};
  double p = 0.0;
  q.pd = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    q.vd[i] = -dm.st.f[i];
#pragma unroll
    for (int j = 0; j < 3; j++) q.F[i][j] = (i == j) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int a = 0; a < ENON; a++) {
    double Nx[3];
    ustruct_grad(q.xiX, Nxi[a], Nx);
    p += N[a] * pl[a];
    q.pd += N[a] * pdl[a];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      q.vd[i] += N[a] * ql[a][i];
      px[i] += Nx[i] * pl[a];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        vx[i][j] += Nx[j] * vl[a][i];
        q.F[i][j] += Nx[j] * dl[a][i];
      }
    }
  }
  const double (*F)[3] = q.F;
  q.J = F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) - F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0]) +
        F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
  const double iJ = 1.0 / q.J;
  q.Fi[0][0] = (F[1][1] * F[2][2] - F[1][2] * F[2][1]) * iJ;
  q.Fi[0][1] = (F[0][2] * F[2][1] - F[0][1] * F[2][2]) * iJ;
  q.Fi[0][2] = (F[0][1] * F[1][2] - F[0][2] * F[1][1]) * iJ;
  q.Fi[1][0] = (F[1][2] * F[2][0] - F[1][0] * F[2][2]) * iJ;
  q.Fi[1][1] = (F[0][0] * F[2][2] - F[0][2] * F[2][0]) * iJ;
  q.Fi[1][2] = (F[0][2] * F[1][0] - F[0][0] * F[1][2]) * iJ;
  q.Fi[2][0] = (F[1][0] * F[2][1] - F[1][1] * F[2][0]) * iJ;
  q.Fi[2][1] = (F[0][1] * F[2][0] - F[0][0] * F[2][1]) * iJ;
  q.Fi[2][2] = (F[0][0] * F[1][1] - F[0][1] * F[1][0]) * iJ;

  // compute_pk2cc with the ustruct flag: isochoric part only (mat_models.cpp:311-312, 395-405)
  StructDmn iso = dm.st;
  iso.Kpen = 0.0;
  if (pk2cc_voigt<CANN>(iso, q.F, fN, ya, cann, nFn, q.S, q.Dm)) return 1;
  // compute_visc_stress_and_tangent (ustruct.cpp:1255-1259, mat_models.cpp:1583-1762): Siso += Svis (:1278); the tangent
  // terms are kept as two ViscGP sets, gu for Kvis_u alone (afu = 1, afv = 0) and gv for Kvis_v alone (0, 1)
  if (gu != nullptr && dm.st.viscType != SVB200_SOLID_VISC_NONE) {
    double Svis[3][3];
    visc_gauss_point(dm.st.viscType, dm.st.visc_mu, 1.0, 0.0, q.F, vx, Svis, *gu);
    visc_gauss_point(dm.st.viscType, dm.st.visc_mu, 0.0, 1.0, q.F, vx, Svis, *gv);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) q.S[i][j] += Svis[i][j];
  }

  // g_vol_pen with Ja = 1
  ustruct_vol_pen(dm.st, p, q.rho, q.beta, q.drho, q.dbeta);
  ustruct_tau(dm, Je, q.J, q.tauM, q.tauC);
  double divV = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      q.Pdev[i][j] = F[i][0] * q.S[0][j] + F[i][1] * q.S[1][j] + F[i][2] * q.S[2][j];
      q.VxFi[i][j] = vx[i][0] * q.Fi[0][j] + vx[i][1] * q.Fi[1][j] + vx[i][2] * q.Fi[2][j];
    }
    q.PxFi[i] = px[0] * q.Fi[0][i] + px[1] * q.Fi[1][i] + px[2] * q.Fi[2][i];
    divV += q.VxFi[i][i];
  }
  q.rC = q.beta * q.pd + divV;
  q.rCl = -p + q.tauC * q.rC;
#pragma unroll
  for (int i = 0; i < 3; i++) q.rM[i] = q.rho * q.vd[i] + q.PxFi[i];
  (void)dt; (void)af_eq; (void)am; (void)gam;
  return 0;
}

SVB_HD void ustruct_node(const UGP& q, double Na, const double Nxia[3], UNode& n)
{
  n.N = Na;
  ustruct_grad(q.xiX, Nxia, n.Nx);
#pragma unroll
  for (int i = 0; i < 3; i++) n.NxFi[i] = n.Nx[0] * q.Fi[0][i] + n.Nx[1] * q.Fi[1][i] + n.Nx[2] * q.Fi[2][i];
#pragma unroll
  for (int j = 0; j < 3; j++) n.VxNx[j] = q.VxFi[0][j] * n.NxFi[0] + q.VxFi[1][j] * n.NxFi[1] + q.VxFi[2][j] * n.NxFi[2];
  n.rMNx = q.rM[0] * n.NxFi[0] + q.rM[1] * n.NxFi[1] + q.rM[2] * n.NxFi[2];
}

// lR(0..3, a) of one Gauss point (ustruct.cpp:1311-1330 and 800-804).
SVB_HD void ustruct_resid(const UGP& q, const UNode& a, double lR[4])
{
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double T1 = q.J * q.rho * q.vd[i] * a.N;
    const double T2 = q.Pdev[i][0] * a.Nx[0] + q.Pdev[i][1] * a.Nx[1] + q.Pdev[i][2] * a.Nx[2];
    const double T3 = q.J * q.rCl * a.NxFi[i];
    lR[i] += q.w * (T1 + T2 + T3);
  }
  lR[3] += q.w * q.J * (a.N * q.rC + q.tauM * a.rMNx);
}

// Blocks lK(0..15, a, b) (row-major 4x4) and lKd(0..11, a, b) (row-major 4x3) of one Gauss point.
// af = eq.af * eq.gam * dt, am = eq.am.
SVB_HD void ustruct_block(const UGP& q, double af, double am, const UNode& a, const UNode& b, const double Bma[6][3],
                          const double DBmb[6][3], double K[16], double Kd[12])
{
  const double afm = af / am, w = q.w, J = q.J;
  // NxSNx = grad N_a . S . grad N_b
  double SNb[3];
#pragma unroll
  for (int i = 0; i < 3; i++) SNb[i] = q.S[i][0] * b.Nx[0] + q.S[i][1] * b.Nx[1] + q.S[i][2] * b.Nx[2];
  const double NxSNx = a.Nx[0] * SNb[0] + a.Nx[1] * SNb[1] + a.Nx[2] * SNb[2];
  const double mass = am * J * q.rho * a.N * b.N;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double BtDB = 0.0;
#pragma unroll
      for (int r = 0; r < 6; r++) BtDB += Bma[r][i] * DBmb[r][j];
      const double T1 = J * q.rho * q.vd[i] * a.N * b.NxFi[j];
      const double T2 = -q.tauC * J * a.NxFi[i] * b.VxNx[j];
      const double T3 = (i == j) ? NxSNx : J * q.rCl * (a.NxFi[i] * b.NxFi[j] - a.NxFi[j] * b.NxFi[i]);
      const double Ku = w * af * (T1 + T2 + T3 + BtDB);
      Kd[3 * i + j] += Ku;
      const double Tm = ((i == j) ? mass : 0.0) + af * J * q.tauC * q.rho * a.NxFi[i] * b.NxFi[j];
      K[4 * i + j] += w * Tm + afm * Ku;
    }
  // dR_m / dp (ustruct.cpp:1577-1590)
  {
    const double T0 = am * q.tauC * q.beta + af * (q.tauC * q.dbeta * q.pd - 1.0);
#pragma unroll
    for (int i = 0; i < 3; i++) K[4 * i + 3] += w * J * (T0 * a.NxFi[i] * b.N + af * q.drho * q.vd[i] * a.N * b.N);
  }
  // continuity row (ustruct.cpp:820-869)
  const double NxNx = a.NxFi[0] * b.NxFi[0] + a.NxFi[1] * b.NxFi[1] + a.NxFi[2] * b.NxFi[2];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const double T0 = a.N * (q.rC * b.NxFi[j] - b.VxNx[j]);
    const double T1 = q.tauM * (a.rMNx * b.NxFi[j] - b.rMNx * a.NxFi[j]);
    const double T2 = -q.tauM * NxNx * q.PxFi[j];
    const double Ku = w * af * J * (T0 + T1 + T2);
    Kd[9 + j] += Ku;
    K[12 + j] += w * J * ((am * q.tauM * q.rho) * a.NxFi[j] * b.N + af * a.N * b.NxFi[j]) + afm * Ku;
  }
  {
    const double T0 = (am * q.beta + af * q.dbeta * q.pd) * a.N * b.N;
    const double T1 = a.NxFi[0] * q.vd[0] + a.NxFi[1] * q.vd[1] + a.NxFi[2] * q.vd[2];
    K[15] += w * J * (T0 + af * q.tauM * (NxNx + q.drho * T1 * b.N));
  }
}


// ---- linear tetrahedra in closed form (assemble_ustruct_tet4_kernel) -----------------------------------------------------
// For TET4 the geometry (F, F^-1, J, S, Dm, grad v F^-1, grad p F^-1, tauM, tauC) is constant over the element; the Gauss
// points differ only through N_a(g) and the scalars that depend on p(g), pd(g), vd(g): rho, beta, drho, dbeta, rC, rCl, rM.
// Every term of ustruct_3d_m / ustruct_3d_c is (such a scalar) x (N_a and / or N_b, or neither) x (a constant), so the Gauss sums
// are taken once per element as weighted moments and every block is assembled from them — one compute_pk2cc and four Dm Bm_b
// per element instead of four of each per Gauss point and row lane.
struct UTet4Const {
  double Nx[4][3], NxFi[4][3], VxNx[4][3];
  double F[3][3], S[3][3], Pdev[3][3], PxFi[3];
  double J, tauM, tauC;
};

// Per-Gauss-point scalars (weights included) and the few plain sums; the N-weighted sums a block needs are formed from
// these when the block is assembled (4 products each) — keeping ~100 moments alive instead costs more in spills than the
// recomputation costs in flops.
struct UTet4GP {
  double W, s_rho, s_rCl, s_rM[3];            // sum w, sum w rho, sum w rCl, sum w rM
  double w[4], rho[4], rC[4], T0m[4], T0c[4], drho[4], vd[4][3];   // w_g Je and the scalars of Gauss point g
};
using UTet4Mom = UTet4GP;

// Element constants, Dm and the Gauss-point scalars.  w[g] are the reference weights, N is indexed [g][a] with row stride ldN.
// Returns 0, 1 for an unsupported constitutive model; Je (Jacobian of the reference map) is returned through *Je_out.
template <bool CANN = true>
SVB_HD int ustruct_tet4_setup(const UstructDmn& dm, double af, double am, const double* w, const double* N, int ldN,
                              const double Nxi[][3], const double xl[4][3], const double ql[4][3], const double vl[4][3],
                              const double dl[4][3], const double pl[4], const double pdl[4], const double fN[2][3],
                              UTet4Const& C, UTet4GP& M, double Dm[6][6], double* Je_out, const double* ya = nullptr,
                              const CannRow* cann = nullptr, int nFn = 2)
{
  double xiX[3][3];
  const double Je = ustruct_xiX<4>(Nxi, xl, xiX);
  *Je_out = Je;
  double vx[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, px[3] = {0, 0, 0}, Fi[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C.F[i][j] = (i == j) ? 1.0 : 0.0;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    ustruct_grad(xiX, Nxi[a], C.Nx[a]);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      px[i] += C.Nx[a][i] * pl[a];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        vx[i][j] += C.Nx[a][j] * vl[a][i];
        C.F[i][j] += C.Nx[a][j] * dl[a][i];
      }
    }
  }
  const double (*F)[3] = C.F;
  C.J = F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) - F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0]) +
        F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
  const double iJ = 1.0 / C.J;
  Fi[0][0] = (F[1][1] * F[2][2] - F[1][2] * F[2][1]) * iJ;
  Fi[0][1] = (F[0][2] * F[2][1] - F[0][1] * F[2][2]) * iJ;
  Fi[0][2] = (F[0][1] * F[1][2] - F[0][2] * F[1][1]) * iJ;
  Fi[1][0] = (F[1][2] * F[2][0] - F[1][0] * F[2][2]) * iJ;
  Fi[1][1] = (F[0][0] * F[2][2] - F[0][2] * F[2][0]) * iJ;
  Fi[1][2] = (F[0][2] * F[1][0] - F[0][0] * F[1][2]) * iJ;
  Fi[2][0] = (F[1][0] * F[2][1] - F[1][1] * F[2][0]) * iJ;
  Fi[2][1] = (F[0][1] * F[2][0] - F[0][0] * F[2][1]) * iJ;
  Fi[2][2] = (F[0][0] * F[1][1] - F[0][1] * F[1][0]) * iJ;
  StructDmn iso = dm.st;
  iso.Kpen = 0.0;
  if (pk2cc_voigt<CANN>(iso, C.F, fN, ya, cann, nFn, C.S, Dm)) return 1;
  ustruct_tau(dm, Je, C.J, C.tauM, C.tauC);
  double VxFi[3][3], divV = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      C.Pdev[i][j] = F[i][0] * C.S[0][j] + F[i][1] * C.S[1][j] + F[i][2] * C.S[2][j];
      VxFi[i][j] = vx[i][0] * Fi[0][j] + vx[i][1] * Fi[1][j] + vx[i][2] * Fi[2][j];
    }
    C.PxFi[i] = px[0] * Fi[0][i] + px[1] * Fi[1][i] + px[2] * Fi[2][i];
    divV += VxFi[i][i];
  }
#pragma unroll
  for (int a = 0; a < 4; a++) {
#pragma unroll
    for (int i = 0; i < 3; i++) C.NxFi[a][i] = C.Nx[a][0] * Fi[0][i] + C.Nx[a][1] * Fi[1][i] + C.Nx[a][2] * Fi[2][i];
#pragma unroll
    for (int j = 0; j < 3; j++) C.VxNx[a][j] = VxFi[0][j] * C.NxFi[a][0] + VxFi[1][j] * C.NxFi[a][1] + VxFi[2][j] * C.NxFi[a][2];
  }
  M.W = M.s_rho = M.s_rCl = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++) M.s_rM[i] = 0.0;
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const double wg = w[g] * Je;
    const double* Ng = N + g * ldN;
    double p = 0.0, pd = 0.0, vd[3] = {-dm.st.f[0], -dm.st.f[1], -dm.st.f[2]};
#pragma unroll
    for (int a = 0; a < 4; a++) {
      p += Ng[a] * pl[a];
      pd += Ng[a] * pdl[a];
#pragma unroll
      for (int i = 0; i < 3; i++) vd[i] += Ng[a] * ql[a][i];
    }
    double rho, beta, drho, dbeta;
    ustruct_vol_pen(dm.st, p, rho, beta, drho, dbeta);
    const double rC = beta * pd + divV;
    const double rCl = -p + C.tauC * rC;
    M.w[g] = wg; M.rho[g] = rho; M.rC[g] = rC; M.drho[g] = drho;
    M.T0m[g] = am * C.tauC * beta + af * (C.tauC * dbeta * pd - 1.0);
    M.T0c[g] = am * beta + af * dbeta * pd;
    M.W += wg;
    M.s_rho += wg * rho;
    M.s_rCl += wg * rCl;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      M.vd[g][i] = vd[i];
      M.s_rM[i] += wg * (rho * vd[i] + C.PxFi[i]);
    }
  }
  return 0;
}

// lR(0..3, a) of the element.
SVB_HD void ustruct_tet4_resid(const UTet4Const& C, const UTet4GP& M, const double* N, int ldN, int a, double lR[4])
{
  double n_rC = 0.0, n_rvd[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const double wn = M.w[g] * N[g * ldN + a];
    n_rC += wn * M.rC[g];
#pragma unroll
    for (int i = 0; i < 3; i++) n_rvd[i] += wn * M.rho[g] * M.vd[g][i];
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
    lR[i] = C.J * n_rvd[i] + M.W * (C.Pdev[i][0] * C.Nx[a][0] + C.Pdev[i][1] * C.Nx[a][1] + C.Pdev[i][2] * C.Nx[a][2]) +
            C.J * M.s_rCl * C.NxFi[a][i];
  lR[3] = C.J * (n_rC + C.tauM * (M.s_rM[0] * C.NxFi[a][0] + M.s_rM[1] * C.NxFi[a][1] + M.s_rM[2] * C.NxFi[a][2]));
}

// lK(0..15, a, b) and lKd(0..11, a, b) of the element; af = eq.af eq.gam dt, am = eq.am.
SVB_HD void ustruct_tet4_block(const UTet4Const& C, const UTet4GP& M, const double* N, int ldN, double af, double am, int a, int b,
                               const double Bma[6][3], const double DBmb[6][3], double K[16], double Kd[12])
{
  const double afm = af / am, J = C.J, W = M.W;
  const double* Fa = C.NxFi[a];
  const double* Fb = C.NxFi[b];
  // the N-weighted Gauss sums of this node pair
  double n1a = 0.0, n_rCa = 0.0, n_rhob = 0.0, n_T0mb = 0.0, nn_rho = 0.0, nn_T0c = 0.0;
  double n_rvda[3] = {0.0, 0.0, 0.0}, n_dvdb[3] = {0.0, 0.0, 0.0}, nn_dvd[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const double wa = M.w[g] * N[g * ldN + a], wb = M.w[g] * N[g * ldN + b], wab = wa * N[g * ldN + b];
    n1a += wa;
    n_rCa += wa * M.rC[g];
    n_rhob += wb * M.rho[g];
    n_T0mb += wb * M.T0m[g];
    nn_rho += wab * M.rho[g];
    nn_T0c += wab * M.T0c[g];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      n_rvda[i] += wa * M.rho[g] * M.vd[g][i];
      n_dvdb[i] += wb * M.drho[g] * M.vd[g][i];
      nn_dvd[i] += wab * M.drho[g] * M.vd[g][i];
    }
  }
  double SNb[3];
#pragma unroll
  for (int i = 0; i < 3; i++) SNb[i] = C.S[i][0] * C.Nx[b][0] + C.S[i][1] * C.Nx[b][1] + C.S[i][2] * C.Nx[b][2];
  const double NxSNx = C.Nx[a][0] * SNb[0] + C.Nx[a][1] * SNb[1] + C.Nx[a][2] * SNb[2];
  const double NxNx = Fa[0] * Fb[0] + Fa[1] * Fb[1] + Fa[2] * Fb[2];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double BtDB = 0.0;
#pragma unroll
      for (int r = 0; r < 6; r++) BtDB += Bma[r][i] * DBmb[r][j];
      const double T1 = J * n_rvda[i] * Fb[j];
      const double T2 = W * (-C.tauC * J * Fa[i] * C.VxNx[b][j] + BtDB);
      const double T3 = (i == j) ? W * NxSNx : J * M.s_rCl * (Fa[i] * Fb[j] - Fa[j] * Fb[i]);
      const double Ku = af * (T1 + T2 + T3);
      Kd[3 * i + j] = Ku;
      const double Tm = ((i == j) ? am * J * nn_rho : 0.0) + af * J * C.tauC * M.s_rho * Fa[i] * Fb[j];
      K[4 * i + j] = Tm + afm * Ku;
    }
#pragma unroll
  for (int i = 0; i < 3; i++) K[4 * i + 3] = J * (n_T0mb * Fa[i] + af * nn_dvd[i]);
  const double rMa = M.s_rM[0] * Fa[0] + M.s_rM[1] * Fa[1] + M.s_rM[2] * Fa[2];
  const double rMb = M.s_rM[0] * Fb[0] + M.s_rM[1] * Fb[1] + M.s_rM[2] * Fb[2];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const double T0 = n_rCa * Fb[j] - n1a * C.VxNx[b][j];
    const double T1 = C.tauM * (rMa * Fb[j] - rMb * Fa[j]);
    const double T2 = -C.tauM * NxNx * C.PxFi[j] * W;
    const double Ku = af * J * (T0 + T1 + T2);
    Kd[9 + j] = Ku;
    K[12 + j] = J * (am * C.tauM * n_rhob * Fa[j] + af * n1a * Fb[j]) + afm * Ku;
  }
  K[15] = J * (nn_T0c + af * C.tauM * (W * NxNx + Fa[0] * n_dvdb[0] + Fa[1] * n_dvdb[1] + Fa[2] * n_dvdb[2]));
}

// Viscous part of a block: lKd(0..8) += w af Kvis_u, lK(v,v) += w af Kvis_v + (af/am) w af Kvis_u (ustruct.cpp:1455-1572).
SVB_HD void ustruct_visc_block(int viscType, const UGP& q, double af, double am, const ViscGP& gu, const ViscGP& gv,
                               const double Nxa[3], const double Nxb[3], double K[16], double Kd[12])
{
  double Vau[9], Vbu[9], Vav[9], Vbv[9];
  visc_node(gu, Nxa, Vau); visc_node(gu, Nxb, Vbu);
  visc_node(gv, Nxa, Vav); visc_node(gv, Nxb, Vbv);
  double Tu[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, Tv[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  visc_block(viscType, q.w * af * gu.c, 1.0, 0.0, &gu.M[0][0], Vau, Vbu, Tu);
  visc_block(viscType, q.w * af * gv.c, 0.0, 1.0, &gv.M[0][0], Vav, Vbv, Tv);
  const double afm = af / am;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      Kd[3 * i + j] += Tu[i][j];
      K[4 * i + j] += Tv[i][j] + afm * Tu[i][j];
    }
}

}  // namespace svb
