// fluid_thood.cuh — Gauss-point algebra of the Navier-Stokes element on Taylor-Hood function spaces (mshType::nFs = 2: velocity on
// the mesh's quadratic element, pressure on its linear parent; no VMS extras).
//
// Reference: fluid::construct_fluid with vmsStab = false (Code/Source/solver/fluid.cpp:494-500, 596-748), fs::get_thood_fs
// (solver/fs.cpp:73-178), fluid_3d_m / fluid_3d_c with vmsFlag = false (fluid.cpp:1768-2237 / 1443-1760):
//   * momentum loop over the VELOCITY space's Gauss rule; the pressure and its gradient are interpolated with the pressure space's
//     shape functions evaluated at those points (Nq, Nqx over the first eNoNq element nodes);
//   * vmsFlag false: tauC = tauB = 0, ua = u, pa = p, uaNx = uNx — but tauM and the fine-scale velocity up stay (fluid.cpp:2070-2077);
//   * continuity loop over the PRESSURE space's Gauss rule with tauM = 0, up = 0, updu = 0 (fluid.cpp:1711-1715): what is left is
//     lR(3,a) += w Nq_a div u and lK(12+j,a,b) += wl Nq_a Nwx(j,b).
// Same layout of the work as fluid_gen.cuh (one record per Gauss point + a short record per element node); the routines below are the
// vmsFlag = false / two-space twins of fluid_gen_gauss_point, fluid_gen_residual, fluid_gen_row, fluid_gen_block_row.
#pragma once
#include "fluid_gen.cuh"

namespace svb {

// fluid_3d_m at one Gauss point of the velocity rule.  Nq[ENONQ], Nqx[ENONQ][3]: pressure space at this point.
template <int ENON, int ENONQ, class NodeT = FluidNode>
SVB_HD void thood_gauss_point_m(const FluidDmn& dm, double dt, double af, double am, double gam_t, double w, const double ks[3][3],
                                const double N[], const double Nx[][3], const double Nxx[][6], const double Nq[], const double Nqx[][3],
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
      for (int k = 0; k < 3; k++) ux[k][i] += Nx[a][k] * yl[a][i];
    }
  }
#pragma unroll
  for (int a = 0; a < ENONQ; a++) {                 // fluid.cpp:1862-1867: pressure space
    p += Nq[a] * yl[a][3];
#pragma unroll
    for (int k = 0; k < 3; k++) px[k] += Nqx[a][k] * yl[a][3];
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
  double gx[3];
  second_derivative_terms<ENON>(Nxx, yv, es, q.d2u2, gx);
#pragma unroll
  for (int k = 0; k < 3; k++) { q.mu_x[k] = mu_g * gx[k]; q.mu_x_c[k] = 0.0; q.d2u2_c[k] = 0.0; q.up_c[k] = 0.0; }
  double kT = 4.0 * (ctM / dt) * (ctM / dt);
  kT += (Kd * mu / rho) * (Kd * mu / rho);
  kT += uF * uF;
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
    const double uVj = (j == 0) ? uV0 : (j == 1 ? uV1 : uV2);
    q.rB[j] = uF * u[j] - uVj;
    up[j] = -tauM * (rho * rVj + px[j] - rS + mu * Kd * u[j] + uF * u[j] - uVj);
    q.up[j] = up[j];
    q.u[j] = u[j];
  }
  q.tauC = 0.0; q.tauB = 0.0;                       // fluid.cpp:2070-2077 (vmsFlag false): ua = u, pa = p
#pragma unroll
  for (int j = 0; j < 3; j++) {
#pragma unroll
    for (int i = 0; i < 3; i++) q.rM[i][j] = mu * es[i][j] - rho * up[j] * u[i] - (i == j ? p : 0.0);
    q.rV[j] = ud[j] + u[0] * ux[0][j] + u[1] * ux[1][j] + u[2] * ux[2][j];
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
    const double base = -rho * n.uNx - mu * Kd * N[a] - uF * N[a];
    n.T1b = base + mu * (Nxx[a][0] + Nxx[a][1] + Nxx[a][2]) + q.mu_x[0] * Nx[a][0] + q.mu_x[1] * Nx[a][1] + q.mu_x[2] * Nx[a][2];
    n.T1b_c = 0.0;
    fluid_node_store(nd[a], n);
  }
}

// lR(0..2, a) of fluid_3d_m (fluid.cpp:2108-2111, 2228-2235)
SVB_HD void thood_residual_m(const FluidGP& q, const FluidNode& a, double lR[4])
{
#pragma unroll
  for (int j = 0; j < 3; j++)
    lR[j] += q.wr * a.N * q.rV[j] + q.w * (a.Nx[0] * q.rM[0][j] + a.Nx[1] * q.rM[1][j] + a.Nx[2] * q.rM[2][j]) +
             q.muKd * q.w * a.N * (q.u[j] + q.up[j]) + q.w * a.N * q.rB[j];
}

// fluid_gen_row with uaNx = uNx (vmsFlag false, fluid.cpp:2117-2121)
SVB_HD void thood_row(const FluidGP& q, const FluidNode& a, FluidRow& r)
{
  const double wl = q.wl;
  const double rtu = q.rho * q.tauM * a.uNx;
  r.wlrtu = wl * rtu;
  r.wlmu = wl * q.mu;
  r.c0 = wl * (q.rho * q.amd * (a.N + rtu) + q.muKdT * a.N);
  r.c1 = wl * q.rho * a.N;
  r.c2 = 0.0;                                        // tauB = 0
  r.wlNa = wl * a.N;
  r.wltM = wl * q.tauM;
  r.ramd = q.rho * q.amd;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    r.U[j] = wl * (q.mu * a.Nx[j] - rtu * q.mu_x[j]);
    r.V[j] = -wl * rtu * q.mu_g * q.d2u2[j];
    r.tCNx[j] = 0.0;                                 // tauC = 0
    r.mgEs[j] = wl * q.mu_g * a.esNx[j];
    r.wlNx[j] = wl * a.Nx[j];
    r.Nx[j] = a.Nx[j];
    r.Pc[j] = 0.0;
    r.Qc[j] = 0.0;
  }
}

// velocity-velocity part of block (a, b) and, for a pressure node b (Nqb / Nqxb not null), the velocity-pressure column
// (fluid.cpp:2146-2224); the continuity rows are not touched by fluid_3d_m.
SVB_HD void thood_block_m(const FluidRow& r, const FluidNode& b, const double* Nqb, const double* Nqxb, double K[16])
{
  const double NxNx = r.Nx[0] * b.Nx[0] + r.Nx[1] * b.Nx[1] + r.Nx[2] * b.Nx[2];
  const double D = r.wlmu * NxNx + r.c0 * b.N + r.c1 * (b.uNx + b.upNx) - r.wlrtu * b.T1b;
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++)
      K[4 * i + j] += r.U[j] * b.Nx[i] + r.mgEs[i] * b.esNx[j] + r.V[j] * b.esNx[i] + (i == j ? D : 0.0);
    if (Nqb != nullptr) K[4 * i + 3] += r.wlrtu * Nqxb[i] - r.wlNx[i] * Nqb[0];
  }
}

}  // namespace svb
