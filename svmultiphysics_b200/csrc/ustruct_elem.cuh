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

// Everything ustruct_3d_m / ustruct_3d_c evaluate before their node loops.
//   ql[a] = al(i..k,a) - bfl(:,a);  vl, dl: nodal velocity / displacement;  pl, pdl: nodal pressure and its rate.
// Returns 0, 1 for an unsupported constitutive model.
template <int ENON>
SVB_HD int ustruct_gauss_point(const UstructDmn& dm, double dt, double af_eq, double am, double gam, double wg, const double N[],
                               const double Nxi[][3], const double xl[][3], const double ql[][3], const double vl[][3],
                               const double dl[][3], const double pl[], const double pdl[], const double fN[2][3], UGP& q,
                               ViscGP* gu = nullptr, ViscGP* gv = nullptr)
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
  if (pk2cc_voigt(iso, q.F, fN, q.S, q.Dm)) return 1;
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
  const double Kp = dm.st.Kpen;
  q.rho = dm.st.rho; q.beta = 0.0; q.drho = 0.0; q.dbeta = 0.0;
  if (!is_zero(Kp)) {
    if (dm.st.volType == SVB200_VOL_QUAD) {
      const double r1 = 1.0 / (Kp - p);
      q.rho = q.rho * Kp * r1; q.beta = r1; q.drho = q.rho * r1; q.dbeta = r1 * r1;
    } else if (dm.st.volType == SVB200_VOL_ST91) {
      const double r1 = q.rho / Kp, r2 = sqrt(p * p + Kp * Kp);
      q.rho = r1 * (p + r2); q.beta = 1.0 / r2; q.drho = q.rho * q.beta; q.dbeta = -q.beta * p / (p * p + Kp * Kp);
    } else if (dm.st.volType == SVB200_VOL_M94) {
      const double r1 = q.rho / Kp, r2 = Kp + p;
      q.rho = r1 * r2; q.beta = 1.0 / r2; q.drho = r1; q.dbeta = -q.beta * q.beta;
    }
  }
  // compute_tau
  {
    const double he = 0.5 * pow(Je, 1.0 / 3.0);
    const double rho0 = dm.st.rho, mu = 0.5 * dm.E / (1.0 + dm.nu);
    double c;
    if (is_zero(dm.nu - 0.5)) c = sqrt(mu / rho0);
    else c = sqrt((2.0 * mu * dm.nu / (1.0 - 2.0 * dm.nu) + 2.0 * mu) / rho0);
    q.tauM = dm.ctM * (he / c) * (q.J / rho0);
    q.tauC = dm.ctC * (he * c) * (rho0 / q.J);
  }
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
