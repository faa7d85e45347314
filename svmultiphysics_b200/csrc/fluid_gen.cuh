// fluid_gen.cuh — Gauss-point algebra of the VMS Navier-Stokes element for ANY 3-D Lagrange element with equal-order
// velocity/pressure (HEX8 today; TET4 runs through it in the tests as a cross-check of the specialised kernel).
//
// Reference: fluid::construct_fluid (Code/Source/solver/fluid.cpp:480-762) with nn::gnn per Gauss point
// (solver/nn.cpp:862-899), nn::gn_nxx (solver/nn.cpp:1172-1283), fluid_3d_m (fluid.cpp:1768-2237) and
// fluid_3d_c (fluid.cpp:1443-1760).  Unlike fluid_elem.cuh (linear tetrahedra) nothing is constant over the element
// and the second derivatives of the shape functions do not vanish: they enter through the viscous part of the
// strong momentum residual (rS), the viscosity gradient mu_x and the fine-scale velocity tangent updu.
//
// Layout of the work: everything that does not depend on the node pair (a,b) is evaluated ONCE per Gauss point
// (FluidGP, 55 doubles) together with 11 numbers per element node (FluidNode); a tangent block is then a short
// bilinear expression in (FluidNode_a, FluidNode_b) — see fluid_gen_block.
//
// Two reference behaviours that parity needs and that are easy to miss:
//  * gn_nxx solves a 6x6 system K X = B with LAPACK dgesv, K built from the Jacobian dx/dxi.  K is the Voigt form of
//    the transformation of a symmetric second-order tensor, so K^-1 is the same matrix built from dxi/dx, which gnn
//    has already computed: the solve becomes a 6x6 product (agreement with dgesv to round-off).
//  * the continuity loop (second Gauss loop of construct_fluid, fluid.cpp:697-745) recomputes gnn per Gauss point but
//    NOT gn_nxx: fluid_3d_c sees the Nwxx of the LAST Gauss point of the momentum loop at every Gauss point.  The
//    continuity rows therefore use their own d2u2 / mu_x / up / updu (the *_c members below).
#pragma once
#include "fluid_elem.cuh"

namespace svb {

struct FluidGP {
  double w, wl, wr, rho, amd, mu, mu_g, muKd, tauM, tauC, tauB, divU;
  double u[3], up[3], rV[3], rM[3][3];
  double mu_x[3], d2u2[3];
  double up_c[3], mu_x_c[3], d2u2_c[3];
  double es[6];      // strain rate es = ux + ux^T: 00, 11, 22, 01, 12, 02 (for FluidNodeC::expand)
  // URIS valves (uris::eval_uris_ris_factors_quadrature): muKdT = mu Kd + urisFactorTotal is what the tangent and T1 see
  // (fluid.cpp:2126-2129, 2166-2204); rB_j = urisFactorTotal u_j - urisValveVelTermTotal_j enters the residual (:2228-2234).
  // Without valves muKdT = muKd and rB = 0: every sum below then gains an exact + 0.0.
  double muKdT, rB[3];
};
constexpr int FLUID_GP_DOUBLES = sizeof(FluidGP) / sizeof(double);

struct FluidNode {
  double N, Nx[3], esNx[3], uNx, upNx, T1b, T1b_c;
};
constexpr int FLUID_NODE_DOUBLES = sizeof(FluidNode) / sizeof(double);

// What the kernel keeps per (Gauss point, node) in shared memory: the part of a FluidNode that cannot be rebuilt from the Gauss-point
// record (T1b / T1b_c contain the node's second derivatives); esNx, uNx, upNx are 15 FMAs from Nx and q.es, q.u, q.up.
// 6 instead of 11 doubles: one more CTA per SM for HEX8, and 6 instead of 11 shared loads per block in phase B.
struct FluidNodeC {
  double N, Nx[3], T1b, T1b_c;
};
constexpr int FLUID_NODEC_DOUBLES = sizeof(FluidNodeC) / sizeof(double);

SVB_HD void fluid_node_store(FluidNode& d, const FluidNode& s) { d = s; }
SVB_HD void fluid_node_store(FluidNodeC& d, const FluidNode& s)
{
  d.N = s.N; d.Nx[0] = s.Nx[0]; d.Nx[1] = s.Nx[1]; d.Nx[2] = s.Nx[2]; d.T1b = s.T1b; d.T1b_c = s.T1b_c;
}

// nn::gnn, insd = 3: Nx[a][i], xiX[k][i] = d xi_k / d x_i, ks = xiX^T xiX, returns Jac.
template <int ENON>
SVB_HD double gnn3_full(const double Nxi[][3], const double xl[][3], double Nx[][3], double xiX[3][3], double ks[3][3])
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
  xiX[0][0] = (xXi[1][1] * xXi[2][2] - xXi[1][2] * xXi[2][1]) / Jac;
  xiX[0][1] = (xXi[2][1] * xXi[0][2] - xXi[2][2] * xXi[0][1]) / Jac;
  xiX[0][2] = (xXi[0][1] * xXi[1][2] - xXi[0][2] * xXi[1][1]) / Jac;
  xiX[1][0] = (xXi[1][2] * xXi[2][0] - xXi[1][0] * xXi[2][2]) / Jac;
  xiX[1][1] = (xXi[2][2] * xXi[0][0] - xXi[2][0] * xXi[0][2]) / Jac;
  xiX[1][2] = (xXi[0][2] * xXi[1][0] - xXi[0][0] * xXi[1][2]) / Jac;
  xiX[2][0] = (xXi[1][0] * xXi[2][1] - xXi[1][1] * xXi[2][0]) / Jac;
  xiX[2][1] = (xXi[2][0] * xXi[0][1] - xXi[2][1] * xXi[0][0]) / Jac;
  xiX[2][2] = (xXi[0][0] * xXi[1][1] - xXi[0][1] * xXi[1][0]) / Jac;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) ks[i][j] = xiX[0][i] * xiX[0][j] + xiX[1][i] * xiX[1][j] + xiX[2][i] * xiX[2][j];
#pragma unroll
  for (int a = 0; a < ENON; a++)
#pragma unroll
    for (int i = 0; i < 3; i++) Nx[a][i] = Nxi[a][0] * xiX[0][i] + Nxi[a][1] * xiX[1][i] + Nxi[a][2] * xiX[2][i];
  return Jac;
}

// nn::gn_nxx, insd = 3.  Voigt order (00, 11, 22, 01, 12, 02) in both the parametric and the physical frame.
template <int ENON>
SVB_HD void gn_nxx3(const double Nxi2[][6], const double xl[][3], const double xiX[3][3], const double Nx[][3], double Nxx[][6])
{
  const int vi[6] = {0, 1, 2, 0, 1, 0}, vj[6] = {0, 1, 2, 1, 2, 2};
  double xXi2[3][6];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int v = 0; v < 6; v++) xXi2[i][v] = 0.0;
#pragma unroll
  for (int a = 0; a < ENON; a++)
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int v = 0; v < 6; v++) xXi2[i][v] += xl[a][i] * Nxi2[a][v];
  // Kinv(v_phys, v_par): N_,kl = sum_ij xiX(i,k) xiX(j,l) B_ij with B symmetric
  double Kinv[6][6];
#pragma unroll
  for (int p = 0; p < 6; p++) {
    const int k = vi[p], l = vj[p];
#pragma unroll
    for (int q = 0; q < 6; q++) {
      const int i = vi[q], j = vj[q];
      Kinv[p][q] = (i == j) ? xiX[i][k] * xiX[i][l] : xiX[i][k] * xiX[j][l] + xiX[j][k] * xiX[i][l];
    }
  }
#pragma unroll
  for (int a = 0; a < ENON; a++) {
    double B[6];
#pragma unroll
    for (int v = 0; v < 6; v++) B[v] = Nxi2[a][v] - Nx[a][0] * xXi2[0][v] - Nx[a][1] * xXi2[1][v] - Nx[a][2] * xXi2[2][v];
#pragma unroll
    for (int p = 0; p < 6; p++) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < 6; q++) s += Kinv[p][q] * B[q];
      Nxx[a][p] = s;
    }
  }
}

// d2u2 (Laplacian of the velocity) and mu_x / mu_g (gradient of the shear rate, before the mu_g factor) from the
// physical second derivatives Nxx (fluid.cpp:1846-1886, 1944-1973).
template <int ENON>
SVB_HD void second_derivative_terms(const double Nxx[][6], const double yl[][3], const double es[3][3], double d2u2[3], double gx[3])
{
  // H[i][v] = sum_a Nxx[a][v] u_i(a): Hessian of velocity component i in Voigt form
  double H[3][6];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int v = 0; v < 6; v++) H[i][v] = 0.0;
#pragma unroll
  for (int a = 0; a < ENON; a++)
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int v = 0; v < 6; v++) H[i][v] += Nxx[a][v] * yl[a][i];
  const int vv[3][3] = {{0, 3, 5}, {3, 1, 4}, {5, 4, 2}};      // Voigt index of the pair (k,l)
#pragma unroll
  for (int i = 0; i < 3; i++) d2u2[i] = H[i][0] + H[i][1] + H[i][2];
  // gx[k] = 1/2 sum_ij es_x[i][j][k] es[i][j],  es_x[i][j][k] = d/dx_k (u_j,i + u_i,j) = H[j][(i,k)] + H[i][(j,k)]
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) s += (H[j][vv[i][k]] + H[i][vv[j][k]]) * es[i][j];
    gx[k] = 0.5 * s;
  }
}

// uris::eval_uris_ris_factors_quadrature (solver/uris.cpp:1577-1673) at one Gauss point: the total resistance factor and the valve
// velocity term of the URIS valves.  `nodal` holds, per node and valve, |sdf|, |scaffold udf| and the valve velocity (5 doubles,
// node-major: nodal[(node * nUris + v) * 5 + k]); N = shape functions at the Gauss point, nodes = the element's node ids.
template <int ENON>
SVB_HD void uris_factor(const double* nodal, int nUris, const svb200_uris* vp, const double N[], const int nodes[], double& uF, double uV[3])
{
  uF = 0.0; uV[0] = uV[1] = uV[2] = 0.0;
  for (int v = 0; v < nUris; v++) {
    double dist = 0.0, dsc = 0.0, vel[3] = {0.0, 0.0, 0.0};
    for (int a = 0; a < ENON; a++) {
      const double* r = nodal + ((size_t)nodes[a] * nUris + v) * 5;
      dist += N[a] * r[0];
      if (vp[v].scaffold) dsc += N[a] * r[1];
      if (vp[v].include_velocity)
        for (int i = 0; i < 3; i++) vel[i] += N[a] * r[2 + i];
    }
    const double pi = 3.141592653589793238462643383279502884;
    double delta = 0.0, delta_sc = 0.0;
    const double deps = vp[v].sdf_deps;
    if (dist < deps && deps > 0.0) delta = (1 + cos(pi * dist / deps)) / (2 * deps * deps);
    if (vp[v].scaffold) {
      const double sd = vp[v].scaffold_deps;
      if (dsc < sd && sd > 0.0) delta_sc = (1 + cos(pi * dsc / sd)) / (2 * sd * sd);
    }
    uF += vp[v].resistance * (delta + delta_sc);
    if (vp[v].include_velocity)
      for (int i = 0; i < 3; i++) uV[i] += vp[v].resistance * delta * vel[i];
  }
}

// Everything of fluid_3d_m / fluid_3d_c at one Gauss point that does not depend on the node pair.
//   al/yl: nodal acceleration / velocity+pressure (al[a][0..2], yl[a][0..3]); ym: nodal mesh velocity or null;
//   Nxx: physical second derivatives at THIS Gauss point, NxxL: those of the LAST Gauss point (continuity quirk).
template <int ENON, class NodeT = FluidNode>
SVB_HD void fluid_gen_gauss_point(const FluidDmn& dm, double dt, double af, double am, double gam_t, double w, const double ks[3][3],
                                  const double N[], const double Nx[][3], const double Nxx[][6], const double NxxL[][6],
                                  const double al[][3], const double yl[][4], const double bfl[][3], const double (*ym)[3],
                                  FluidGP& q, NodeT nd[], double uF = 0.0, const double* uV = nullptr)
{
  const double ctM = 1.0, ctC = 36.0;
  const double rho = dm.rho, Kd = dm.Kd;
  const double T1 = af * gam_t * dt;
  q.w = w; q.rho = rho; q.amd = am / T1; q.wl = w * T1; q.wr = w * rho;
  double ud[3] = {-dm.f[0], -dm.f[1], -dm.f[2]}, u[3] = {0, 0, 0}, ux[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, p = 0.0, px[3] = {0, 0, 0};
  double yv[ENON][3];
#pragma unroll
  for (int a = 0; a < ENON; a++) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
      yv[a][i] = yl[a][i];
      ud[i] += N[a] * (al[a][i] - bfl[a][i]);
      u[i] += N[a] * yl[a][i];
#pragma unroll
      for (int k = 0; k < 3; k++) ux[k][i] += Nx[a][k] * yl[a][i];      // ux[k][i] = d u_i / d x_k
    }
    p += N[a] * yl[a][3];
#pragma unroll
    for (int k = 0; k < 3; k++) px[k] += Nx[a][k] * yl[a][3];
  }
  q.divU = ux[0][0] + ux[1][1] + ux[2][2];
  if (ym != nullptr)
#pragma unroll
    for (int a = 0; a < ENON; a++)
#pragma unroll
      for (int i = 0; i < 3; i++) u[i] -= N[a] * ym[a][i];
  double es[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) es[i][j] = ux[i][j] + ux[j][i];
  q.es[0] = es[0][0]; q.es[1] = es[1][1]; q.es[2] = es[2][2]; q.es[3] = es[0][1]; q.es[4] = es[1][2]; q.es[5] = es[0][2];
  double gam = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) gam += es[i][j] * es[i][j];
  gam = sqrt(0.5 * gam);
  double mu, mu_g;
  viscosity(dm, gam, mu, mu_g);
  mu_g = is_zero(gam) ? 0.0 : mu_g / gam;
  q.mu = mu; q.mu_g = mu_g; q.muKd = mu * Kd;
  q.muKdT = q.muKd + uF;
  const double uV0 = uV ? uV[0] : 0.0, uV1 = uV ? uV[1] : 0.0, uV2 = uV ? uV[2] : 0.0;

  double gx[3], gxc[3];
  second_derivative_terms<ENON>(Nxx, yv, es, q.d2u2, gx);
  second_derivative_terms<ENON>(NxxL, yv, es, q.d2u2_c, gxc);
#pragma unroll
  for (int k = 0; k < 3; k++) { q.mu_x[k] = mu_g * gx[k]; q.mu_x_c[k] = mu_g * gxc[k]; }

  double kT = 4.0 * (ctM / dt) * (ctM / dt);
  kT += (Kd * mu / rho) * (Kd * mu / rho);
  kT += uF * uF;                                  // fluid.cpp:2006-2008
  double kU = 0.0, kS = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) { kU += u[i] * u[j] * ks[i][j]; kS += ks[i][j] * ks[i][j]; }
  kS = ctC * kS * (mu / rho) * (mu / rho);
  const double tauM = 1.0 / (rho * sqrt(kT + kU + kS));
  q.tauM = tauM;
  double up[3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const double rVj = ud[j] + u[0] * ux[0][j] + u[1] * ux[1][j] + u[2] * ux[2][j];
    const double rS = q.mu_x[0] * es[0][j] + q.mu_x[1] * es[1][j] + q.mu_x[2] * es[2][j] + mu * q.d2u2[j];
    const double rSc = q.mu_x_c[0] * es[0][j] + q.mu_x_c[1] * es[1][j] + q.mu_x_c[2] * es[2][j] + mu * q.d2u2_c[j];
    const double uVj = (j == 0) ? uV0 : (j == 1 ? uV1 : uV2);
    q.rB[j] = uF * u[j] - uVj;
    up[j] = -tauM * (rho * rVj + px[j] - rS + mu * Kd * u[j] + uF * u[j] - uVj);            // fluid.cpp:2042-2047
    q.up_c[j] = -tauM * (rho * rVj + px[j] - rSc + mu * Kd * u[j] + uF * u[j] - uVj);        // fluid.cpp:1689-1694
    q.up[j] = up[j];
    q.u[j] = u[j];
  }
  const double eps = 2.220446049250313e-16;
  q.tauC = 1.0 / (tauM * (ks[0][0] + ks[1][1] + ks[2][2]));
  double tauB = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) tauB += up[i] * up[j] * ks[i][j];
  if (is_zero(tauB)) tauB = eps;
  tauB = rho / sqrt(tauB);
  q.tauB = tauB;
  const double ua[3] = {u[0] + up[0], u[1] + up[1], u[2] + up[2]};
  const double pa = p - q.tauC * q.divU;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const double rVb = tauB * (up[0] * ux[0][j] + up[1] * ux[1][j] + up[2] * ux[2][j]);
#pragma unroll
    for (int i = 0; i < 3; i++) q.rM[i][j] = mu * es[i][j] - rho * up[j] * ua[i] + rVb * up[i] - (i == j ? pa : 0.0);
    q.rV[j] = ud[j] + ua[0] * ux[0][j] + ua[1] * ux[1][j] + ua[2] * ux[2][j];
  }
#pragma unroll
  for (int a = 0; a < ENON; a++) {
    FluidNode n;
    n.N = N[a];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      n.Nx[i] = Nx[a][i];
      n.esNx[i] = es[0][i] * Nx[a][0] + es[1][i] * Nx[a][1] + es[2][i] * Nx[a][2];
    }
    n.uNx = u[0] * Nx[a][0] + u[1] * Nx[a][1] + u[2] * Nx[a][2];
    n.upNx = up[0] * Nx[a][0] + up[1] * Nx[a][1] + up[2] * Nx[a][2];
    const double base = -rho * n.uNx - mu * Kd * N[a] - uF * N[a];                              // fluid.cpp:2126-2129, 1700-1703
    n.T1b = base + mu * (Nxx[a][0] + Nxx[a][1] + Nxx[a][2]) + q.mu_x[0] * Nx[a][0] + q.mu_x[1] * Nx[a][1] + q.mu_x[2] * Nx[a][2];
    n.T1b_c = base + mu * (NxxL[a][0] + NxxL[a][1] + NxxL[a][2]) + q.mu_x_c[0] * Nx[a][0] + q.mu_x_c[1] * Nx[a][1] + q.mu_x_c[2] * Nx[a][2];
    fluid_node_store(nd[a], n);
  }
}

// FluidNodeC -> FluidNode with the terms of the Gauss-point record (same expressions as above, same order of operations).
SVB_HD FluidNode fluid_node_expand(const FluidGP& q, const FluidNodeC& c)
{
  FluidNode n;
  n.N = c.N; n.Nx[0] = c.Nx[0]; n.Nx[1] = c.Nx[1]; n.Nx[2] = c.Nx[2]; n.T1b = c.T1b; n.T1b_c = c.T1b_c;
  const double e00 = q.es[0], e11 = q.es[1], e22 = q.es[2], e01 = q.es[3], e12 = q.es[4], e02 = q.es[5];
  n.esNx[0] = e00 * c.Nx[0] + e01 * c.Nx[1] + e02 * c.Nx[2];
  n.esNx[1] = e01 * c.Nx[0] + e11 * c.Nx[1] + e12 * c.Nx[2];
  n.esNx[2] = e02 * c.Nx[0] + e12 * c.Nx[1] + e22 * c.Nx[2];
  n.uNx = q.u[0] * c.Nx[0] + q.u[1] * c.Nx[1] + q.u[2] * c.Nx[2];
  n.upNx = q.up[0] * c.Nx[0] + q.up[1] * c.Nx[1] + q.up[2] * c.Nx[2];
  return n;
}

// lR(0..3, a) += ... (fluid.cpp:2108-2111, 2228-2235, 1726-1729)
SVB_HD void fluid_gen_residual(const FluidGP& q, const FluidNode& a, double lR[4])
{
#pragma unroll
  for (int j = 0; j < 3; j++)
    lR[j] += q.wr * a.N * q.rV[j] + q.w * (a.Nx[0] * q.rM[0][j] + a.Nx[1] * q.rM[1][j] + a.Nx[2] * q.rM[2][j]) +
             q.muKd * q.w * a.N * (q.u[j] + q.up[j]) + q.w * a.N * q.rB[j];
  lR[3] += q.w * (a.N * q.divU - (q.up_c[0] * a.Nx[0] + q.up_c[1] * a.Nx[1] + q.up_c[2] * a.Nx[2]));
}

// K(4 i + j) += block (a,b) at one Gauss point (fluid.cpp:2146-2224, 1733-1759); K is row-major 4x4.
SVB_HD void fluid_gen_block(const FluidGP& q, const FluidNode& a, const FluidNode& b, double K[16])
{
  const double rho = q.rho, mu = q.mu, wl = q.wl;
  const double NxNx = a.Nx[0] * b.Nx[0] + a.Nx[1] * b.Nx[1] + a.Nx[2] * b.Nx[2];
  const double uaNxa = a.uNx + a.upNx;
  const double rtu = rho * q.tauM * uaNxa;
  const double T1 = mu * NxNx + rho * q.amd * b.N * (a.N + rtu) + rho * a.N * (b.uNx + b.upNx) + q.tauB * a.upNx * b.upNx;
  const double dk = q.muKdT * b.N * a.N;
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      // updu[j][i][b] = mu_x[j] Nx(i,b) + d2u2[j] mu_g esNx(i,b) + delta_ij T1b
      const double updu = q.mu_x[j] * b.Nx[i] + q.d2u2[j] * q.mu_g * b.esNx[i] + (i == j ? b.T1b : 0.0);
      double T2;
      if (i == j) T2 = (mu + q.tauC) * a.Nx[i] * b.Nx[i] + a.esNx[i] * q.mu_g * b.esNx[i] - rtu * updu + T1 + dk;
      else T2 = mu * a.Nx[j] * b.Nx[i] + q.tauC * a.Nx[i] * b.Nx[j] + a.esNx[i] * q.mu_g * b.esNx[j] - rtu * updu;
      K[4 * i + j] += wl * T2;
    }
    K[4 * i + 3] -= wl * (a.Nx[i] * b.N - b.Nx[i] * rtu);
  }
  const double T1c = rho * q.amd * b.N;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    double T2 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const double updu_c = q.mu_x_c[j] * b.Nx[i] + q.d2u2_c[j] * q.mu_g * b.esNx[i] + (i == j ? b.T1b_c - T1c : 0.0);
      T2 += a.Nx[i] * updu_c;
    }
    K[12 + j] += wl * (a.N * b.Nx[j] - q.tauM * T2);
  }
  K[15] += wl * q.tauM * NxNx;
}

// ---- factored form of fluid_gen_block -----------------------------------------------------------------------------
// Everything of a tangent block that depends on (Gauss point, row node a) only is folded into a FluidRow once and reused
// for all column nodes b; a block then costs ~80 FMAs and reads 11 doubles (FluidNode_b):
//   K(i,j)  += U_j Nx_i,b + tC Nx_i,a Nx_j,b + mg esNx_i,a esNx_j,b + V_j esNx_i,b + delta_ij D
//      U_j = wl (mu Nx_j,a - rtu mu_x_j),  V_j = -wl rtu mu_g d2u2_j,  tC = wl tauC,  mg = wl mu_g,
//      D   = wl mu (Nx_a.Nx_b) + c0 N_b + c1 (uNx_b + upNx_b) + c2 upNx_b - wl rtu T1b_b
//      c0  = wl (rho amd (N_a + rtu) + muKd N_a),  c1 = wl rho N_a,  c2 = wl tauB upNx_a
//   K(i,3)  += -wl Nx_i,a N_b + wl rtu Nx_i,b
//   K(3,j)  += wl N_a Nx_j,b - wl tauM [ mu_x_c_j (Nx_a.Nx_b) + mu_g d2u2_c_j (Nx_a.esNx_b) + Nx_j,a (T1b_c_b - rho amd N_b) ]
//   K(3,3)  += wl tauM (Nx_a.Nx_b)
struct FluidRow {
  double U[3], V[3], tCNx[3], mgEs[3], wlNx[3], Nx[3];
  double wlmu, c0, c1, c2, wlrtu, wlNa, wltM, ramd;
  double Pc[3], Qc[3];       // wl tauM mu_x_c_j, wl tauM mu_g d2u2_c_j
};

SVB_HD void fluid_gen_row(const FluidGP& q, const FluidNode& a, FluidRow& r)
{
  const double wl = q.wl;
  const double rtu = q.rho * q.tauM * (a.uNx + a.upNx);
  r.wlrtu = wl * rtu;
  r.wlmu = wl * q.mu;
  r.c0 = wl * (q.rho * q.amd * (a.N + rtu) + q.muKdT * a.N);
  r.c1 = wl * q.rho * a.N;
  r.c2 = wl * q.tauB * a.upNx;
  r.wlNa = wl * a.N;
  r.wltM = wl * q.tauM;
  r.ramd = q.rho * q.amd;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    r.U[j] = wl * (q.mu * a.Nx[j] - rtu * q.mu_x[j]);
    r.V[j] = -wl * rtu * q.mu_g * q.d2u2[j];
    r.tCNx[j] = wl * q.tauC * a.Nx[j];
    r.mgEs[j] = wl * q.mu_g * a.esNx[j];
    r.wlNx[j] = wl * a.Nx[j];
    r.Nx[j] = a.Nx[j];
    r.Pc[j] = r.wltM * q.mu_x_c[j];
    r.Qc[j] = r.wltM * q.mu_g * q.d2u2_c[j];
  }
}

SVB_HD void fluid_gen_block_row(const FluidRow& r, const FluidNode& b, double K[16])
{
  const double NxNx = r.Nx[0] * b.Nx[0] + r.Nx[1] * b.Nx[1] + r.Nx[2] * b.Nx[2];
  const double NxEs = r.Nx[0] * b.esNx[0] + r.Nx[1] * b.esNx[1] + r.Nx[2] * b.esNx[2];
  const double D = r.wlmu * NxNx + r.c0 * b.N + r.c1 * (b.uNx + b.upNx) + r.c2 * b.upNx - r.wlrtu * b.T1b;
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++)
      K[4 * i + j] += r.U[j] * b.Nx[i] + r.tCNx[i] * b.Nx[j] + r.mgEs[i] * b.esNx[j] + r.V[j] * b.esNx[i] + (i == j ? D : 0.0);
    K[4 * i + 3] += r.wlrtu * b.Nx[i] - r.wlNx[i] * b.N;
  }
  const double tc = r.wltM * (b.T1b_c - r.ramd * b.N);
#pragma unroll
  for (int j = 0; j < 3; j++) K[12 + j] += r.wlNa * b.Nx[j] - (r.Pc[j] * NxNx + r.Qc[j] * NxEs + r.Nx[j] * tc);
  K[15] += r.wltM * NxNx;
}

}  // namespace svb
