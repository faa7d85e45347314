// fluid_elem.cuh — per-element VMS Navier-Stokes residual/tangent for linear tetrahedra, written for
// one CUDA thread per element.
//
// What it computes is what the reference computes in fluid::construct_fluid for a TET4 element
// (Code/Source/solver/fluid.cpp:545-749): nn::gnn once (solver/nn.cpp:862-899), then per Gauss
// point fluid_3d_m (fluid.cpp:1768-2237) and fluid_3d_c (fluid.cpp:1443-1760).  It is NOT a
// transcription: for a linear tetrahedron the shape-function gradients, the velocity/pressure
// gradients, the strain rate, the viscosity and the element metric are constant over the element
// and the second derivatives vanish, so the 16 (a,b) tangent blocks are not accumulated Gauss
// point by Gauss point.  Instead the Gauss loop only accumulates a small set of element scalars
// and per-node vectors ("moments" of the Gauss-point dependent stabilisation terms):
//
//   K_uu(i,j;a,b) = A1 Nx(j,a) Nx(i,b) + A2 Nx(i,a) Nx(j,b) + A3 esNx(i,a) esNx(j,b)
//                   + delta_ij ( D(a,b) + A1 gradNa.gradNb )
//   K_up(i;a,b)   = -Nx(i,a) Sb(b) + Nx(i,b) S2(a)
//   K_pu(j;a,b)   =  Nx(j,b) Sb(a) - Nx(j,a) S3(b)
//   K_pp(a,b)     =  Spp gradNa.gradNb
//
// with A1 = mu sum_g wl_g, A2 = sum_g wl_g tauC_g, A3 = (dmu/dgamma)/gamma sum_g wl_g,
// Sb(b) = sum_g wl_g N_b(g), S2(a) = sum_g wl_g rho tauM_g uaNx_a(g),
// S3(b) = sum_g wl_g tauM_g (T1_b(g) - rho amd N_b(g)), Spp = sum_g wl_g tauM_g and
// D(a,b) = sum_g of four rank-1 terms (see tet4_gauss below).  The 256 tangent entries are then
// emitted block by block from these ~60 numbers (tet4_block), which is what makes one element per
// thread fit in registers.  Results agree with the reference to FP64 round-off (different
// summation order), tested at 1e-12 relative.
#pragma once
#include "svb200_internal.h"

#ifndef SVB_HD
#ifdef __CUDACC__
#define SVB_HD __host__ __device__ __forceinline__
#else
#define SVB_HD inline
#endif
#endif

namespace svb {

// utils::is_zero(v) of the reference (Code/Source/solver/utils.cpp:141-160) for value2 = 0 reduces to
// |v| < 10*eps*max(|v|,eps)  <=>  |v| < 10*eps^2.
SVB_HD bool is_zero(double v)
{
  const double eps = 2.220446049250313e-16;
  return fabs(v) < 10.0 * eps * eps;
}

// fluid::get_viscosity (fluid.cpp:2240-2298).  `gam` may be modified (Casson clips it), exactly as the
// reference does through its by-reference argument; mu_g is d(mu)/d(gamma).
SVB_HD void viscosity(const FluidDmn& d, double& gam, double& mu, double& mu_g)
{
  if (d.viscType == SVB200_VISC_CONST) {
    mu = d.mu_i;
    mu_g = 0.0;
  } else if (d.viscType == SVB200_VISC_CY) {
    double T1 = 1.0 + pow(d.lam * gam, d.a);
    double T2 = pow(T1, (d.n - 1.0) / d.a);
    mu = d.mu_i + (d.mu_o - d.mu_i) * T2;
    T1 = T2 / T1;
    T2 = pow(d.lam, d.a) * pow(gam, d.a - 1.0) * T1;
    mu_g = (d.mu_o - d.mu_i) * (d.n - 1.0) * T2;
  } else {
    double mu_o;
    if (gam < d.lam) {
      mu_o = d.mu_o / sqrt(d.lam);
      gam = d.lam;
    } else {
      mu_o = d.mu_o / sqrt(gam);
    }
    mu = (d.mu_i + mu_o) * (d.mu_i + mu_o);
    mu_g = 2.0 * mu_o * (mu_o + d.mu_i) / gam;
  }
}

// Everything the block emitter needs, ~90 doubles.
struct Tet4Elem {
  double Nx[4][3];     // physical gradient of N_a
  double esNx[4][3];   // es . grad N_a (only used when A3 != 0)
  double D[4][4];
  double Sb[4], S2[4], S3[4];
  double A1, A2, A3, Spp;
  double lR[4][4];     // lR[a][i]
};

// Geometry + Gauss loop.  xl[a][i], yl[a][0..3] = (u,v,w,p), uc[a][i] = convective nodal velocity
// (= yl - mesh velocity when mvMsh), ab[a][i] = al - bfl.  Returns the Jacobian determinant.
SVB_HD double tet4_element(const FluidArgs& P, const FluidDmn& dm, const double xl[4][3], const double yl[4][4],
                           const double uc[4][3], const double ab[4][3], Tet4Elem& E)
{
  // ---- nn::gnn -------------------------------------------------------------------------------
  double xXi[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int k = 0; k < 3; k++) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < 4; a++) s += xl[a][i] * P.Nxi[0][a][k];
      xXi[i][k] = s;
    }
  const double Jac = xXi[0][0] * xXi[1][1] * xXi[2][2] + xXi[0][1] * xXi[1][2] * xXi[2][0] +
                     xXi[0][2] * xXi[1][0] * xXi[2][1] - xXi[0][0] * xXi[1][2] * xXi[2][1] -
                     xXi[0][1] * xXi[1][0] * xXi[2][2] - xXi[0][2] * xXi[1][1] * xXi[2][0];
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
  // metric ks = xiX^T xiX (symmetric): k00,k01,k02,k11,k12,k22
  const double k00 = xiX[0][0] * xiX[0][0] + xiX[1][0] * xiX[1][0] + xiX[2][0] * xiX[2][0];
  const double k01 = xiX[0][1] * xiX[0][0] + xiX[1][1] * xiX[1][0] + xiX[2][1] * xiX[2][0];
  const double k02 = xiX[0][2] * xiX[0][0] + xiX[1][2] * xiX[1][0] + xiX[2][2] * xiX[2][0];
  const double k11 = xiX[0][1] * xiX[0][1] + xiX[1][1] * xiX[1][1] + xiX[2][1] * xiX[2][1];
  const double k12 = xiX[0][1] * xiX[0][2] + xiX[1][1] * xiX[1][2] + xiX[2][1] * xiX[2][2];
  const double k22 = xiX[0][2] * xiX[0][2] + xiX[1][2] * xiX[1][2] + xiX[2][2] * xiX[2][2];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int i = 0; i < 3; i++)
      E.Nx[a][i] = P.Nxi[0][a][0] * xiX[0][i] + P.Nxi[0][a][1] * xiX[1][i] + P.Nxi[0][a][2] * xiX[2][i];

  // ---- element-constant kinematics (fluid.cpp:1827-1980) ----------------------------------------
  double ux[3][3], px[3];   // ux[i][j] = d u_j / d x_i
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < 4; a++) s += E.Nx[a][i] * yl[a][j];
      ux[i][j] = s;
    }
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < 4; a++) s += E.Nx[a][i] * yl[a][3];
    px[i] = s;
  }
  const double divU = ux[0][0] + ux[1][1] + ux[2][2];
  double es[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) es[i][j] = ux[i][j] + ux[j][i];
  double gam = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) gam += es[i][j] * es[i][j];
  gam = sqrt(0.5 * gam);
  double mu, mu_g;
  viscosity(dm, gam, mu, mu_g);
  mu_g = is_zero(gam) ? 0.0 : mu_g / gam;
  if (mu_g != 0.0) {
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int j = 0; j < 3; j++)
        E.esNx[a][j] = es[0][j] * E.Nx[a][0] + es[1][j] * E.Nx[a][1] + es[2][j] * E.Nx[a][2];
  }

  const double rho = dm.rho;
  const double T = P.af * P.gam * P.dt;
  const double amd = P.am / T;
  const double muKd = mu * dm.Kd;
  const double nu = mu / rho;
  double kT = 4.0 / (P.dt * P.dt) + (dm.Kd * nu) * (dm.Kd * nu);
  const double kS = 36.0 * (k00 * k00 + k11 * k11 + k22 * k22 + 2.0 * (k01 * k01 + k02 * k02 + k12 * k12)) * nu * nu;
  const double kTS = kT + kS;
  const double trK = k00 + k11 + k22;
  const double q1c = rho * amd + muKd;

  // ---- Gauss loop: accumulate moments --------------------------------------------------------------
  double RM[3][3], UP[3] = {0.0, 0.0, 0.0};
  double sPa = 0.0, sW = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) RM[i][j] = 0.0;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    E.Sb[a] = 0.0; E.S2[a] = 0.0; E.S3[a] = 0.0;
#pragma unroll
    for (int b = 0; b < 4; b++) { E.D[a][b] = 0.0; E.lR[a][b] = 0.0; }
  }
  double sA2 = 0.0, sPP = 0.0;

#pragma unroll 1
  for (int g = 0; g < 4; g++) {
    const double wJ = P.w[g] * Jac;
    const double wl = wJ * T;
    double u[3] = {0.0, 0.0, 0.0}, ud[3] = {-dm.f[0], -dm.f[1], -dm.f[2]}, p = 0.0;
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double Na = P.N[g][a];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        u[i] += Na * uc[a][i];
        ud[i] += Na * ab[a][i];
      }
      p += Na * yl[a][3];
    }
    const double kU = u[0] * u[0] * k00 + u[1] * u[1] * k11 + u[2] * u[2] * k22 +
                      2.0 * (u[0] * u[1] * k01 + u[0] * u[2] * k02 + u[1] * u[2] * k12);
    const double tauM = 1.0 / (rho * sqrt(kTS + kU));
    double up[3], ua[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double rV = ud[j] + u[0] * ux[0][j] + u[1] * ux[1][j] + u[2] * ux[2][j];
      up[j] = -tauM * (rho * rV + px[j] + muKd * u[j]);
    }
    const double tauC = 1.0 / (tauM * trK);
    double tauB = up[0] * up[0] * k00 + up[1] * up[1] * k11 + up[2] * up[2] * k22 +
                  2.0 * (up[0] * up[1] * k01 + up[0] * up[2] * k02 + up[1] * up[2] * k12);
    if (is_zero(tauB)) tauB = 2.220446049250313e-16;
    tauB = rho / sqrt(tauB);
#pragma unroll
    for (int i = 0; i < 3; i++) ua[i] = u[i] + up[i];
    const double pa = p - tauC * divU;
    sPa += wJ * pa;
    sW += wJ;
    sA2 += wl * tauC;
    sPP += wl * tauM;
    double Aj[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double rVb = tauB * (up[0] * ux[0][j] + up[1] * ux[1][j] + up[2] * ux[2][j]);
      const double rV2 = ud[j] + ua[0] * ux[0][j] + ua[1] * ux[1][j] + ua[2] * ux[2][j];
      Aj[j] = rho * rV2 + muKd * ua[j];
#pragma unroll
      for (int i = 0; i < 3; i++) RM[i][j] += wJ * (rVb * up[i] - rho * up[j] * ua[i]);
      UP[j] += wJ * up[j];
    }
    double P1[4], P2[4], P3[4], P4[4], Q1[4], Q2[4], Q3[4], Q4[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double Na = P.N[g][a];
      const double uNx = u[0] * E.Nx[a][0] + u[1] * E.Nx[a][1] + u[2] * E.Nx[a][2];
      const double upNx = up[0] * E.Nx[a][0] + up[1] * E.Nx[a][1] + up[2] * E.Nx[a][2];
      const double uaNx = uNx + upNx;
      const double c = rho * tauM * uaNx;
      P1[a] = wl * (Na + c);
      P2[a] = wl * rho * Na;
      P3[a] = wl * tauB * upNx;
      P4[a] = wl * c;
      Q1[a] = q1c * Na;
      Q2[a] = uaNx;
      Q3[a] = upNx;
      Q4[a] = rho * uNx;
      E.Sb[a] += wl * Na;
      E.S2[a] += P4[a];
      E.S3[a] -= wl * tauM * (Q4[a] + Q1[a]);
      const double wN = wJ * Na;
#pragma unroll
      for (int j = 0; j < 3; j++) E.lR[a][j] += wN * Aj[j];
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++)
        E.D[a][b] += P1[a] * Q1[b] + P2[a] * Q2[b] + P3[a] * Q3[b] + P4[a] * Q4[b];
  }

  // ---- finish residual ------------------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) RM[i][j] += mu * es[i][j] * sW;
    RM[i][i] -= sPa;
  }
  const double iT = 1.0 / T;
#pragma unroll
  for (int a = 0; a < 4; a++) {
#pragma unroll
    for (int j = 0; j < 3; j++)
      E.lR[a][j] += E.Nx[a][0] * RM[0][j] + E.Nx[a][1] * RM[1][j] + E.Nx[a][2] * RM[2][j];
    E.lR[a][3] = divU * (E.Sb[a] * iT) - (UP[0] * E.Nx[a][0] + UP[1] * E.Nx[a][1] + UP[2] * E.Nx[a][2]);
  }
  const double sWl = sW * T;
  E.A1 = mu * sWl;
  E.A2 = sA2;
  E.A3 = mu_g * sWl;
  E.Spp = sPP;
  return Jac;
}

// One 4x4 tangent block lK(:,a,b), row-major (entry 4*i+j), fluid.cpp:2146-2224 and :1733-1759.
SVB_HD void tet4_block(const Tet4Elem& E, int a, int b, double K[16])
{
  const double nn = E.Nx[a][0] * E.Nx[b][0] + E.Nx[a][1] * E.Nx[b][1] + E.Nx[a][2] * E.Nx[b][2];
  const double dd = E.D[a][b] + E.A1 * nn;
  const double q1[3] = {E.A1 * E.Nx[b][0], E.A1 * E.Nx[b][1], E.A1 * E.Nx[b][2]};
  const double q2[3] = {E.A2 * E.Nx[b][0], E.A2 * E.Nx[b][1], E.A2 * E.Nx[b][2]};
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double v = E.Nx[a][j] * q1[i] + E.Nx[a][i] * q2[j];
      if (i == j) v += dd;
      K[4 * i + j] = v;
    }
    K[4 * i + 3] = E.Nx[b][i] * E.S2[a] - E.Nx[a][i] * E.Sb[b];
    K[12 + i] = E.Nx[b][i] * E.Sb[a] - E.Nx[a][i] * E.S3[b];
  }
  K[15] = E.Spp * nn;
  if (E.A3 != 0.0) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const double t = E.A3 * E.esNx[a][i];
#pragma unroll
      for (int j = 0; j < 3; j++) K[4 * i + j] += t * E.esNx[b][j];
    }
  }
}


// =====================================================================================================
// Staged variant used by the grouped scatter (assemble_fluid.cu).  Same algebra as tet4_element, but
// organised for a small register footprint (3 CTAs of 128 threads per SM instead of 2):
//   * the element "record" (what tet4_block needs) is written to `rec` (shared memory on the device) as
//     soon as a field is final;
//   * the not-yet-final part of the record doubles as per-thread scratch for Gauss-point values
//     (u_g, ud_g, p_g -> up_g, tauM_g, tauB_g), so the nodal inputs die before the Gauss loops;
//   * the Gauss loop is split in two: loop 1 accumulates the residual moments, loop 2 the tangent
//     moments D, S2, S3 from the stored Gauss-point values.
// Record layout (doubles): Nx[4][3] | Sb[4] | S2[4] | S3[4] | D[4][4] | A1 A2 Spp | pad pad | (A3 | esNx[4][3]).
constexpr int REC_NEWT = 45, REC_NN = 59;   // odd strides: conflict-free lane-strided shared-memory access
constexpr int O_NX = 0, O_SB = 12, O_S2 = 16, O_S3 = 20, O_D = 24, O_A1 = 40, O_A2 = 41, O_SPP = 42, O_A3 = 45, O_ES = 46;

SVB_HD void tet4_element_staged(const FluidArgs& P, const FluidDmn& dm, const double xl[4][3], const double yl[4][4],
                                const double uc[4][3], const double ab[4][3], const bool NN, double* rec, double* lRout)
{
  double nx[4][3];
  double k00, k01, k02, k11, k12, k22, Jac;
  {
    double xXi[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int k = 0; k < 3; k++) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < 4; a++) s += xl[a][i] * P.Nxi[0][a][k];
        xXi[i][k] = s;
      }
    Jac = xXi[0][0] * xXi[1][1] * xXi[2][2] + xXi[0][1] * xXi[1][2] * xXi[2][0] + xXi[0][2] * xXi[1][0] * xXi[2][1] -
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
    k00 = xiX[0][0] * xiX[0][0] + xiX[1][0] * xiX[1][0] + xiX[2][0] * xiX[2][0];
    k01 = xiX[0][1] * xiX[0][0] + xiX[1][1] * xiX[1][0] + xiX[2][1] * xiX[2][0];
    k02 = xiX[0][2] * xiX[0][0] + xiX[1][2] * xiX[1][0] + xiX[2][2] * xiX[2][0];
    k11 = xiX[0][1] * xiX[0][1] + xiX[1][1] * xiX[1][1] + xiX[2][1] * xiX[2][1];
    k12 = xiX[0][1] * xiX[0][2] + xiX[1][1] * xiX[1][2] + xiX[2][1] * xiX[2][2];
    k22 = xiX[0][2] * xiX[0][2] + xiX[1][2] * xiX[1][2] + xiX[2][2] * xiX[2][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int i = 0; i < 3; i++) {
        nx[a][i] = P.Nxi[0][a][0] * xiX[0][i] + P.Nxi[0][a][1] * xiX[1][i] + P.Nxi[0][a][2] * xiX[2][i];
        rec[O_NX + 3 * a + i] = nx[a][i];
      }
  }

  // ---- element-constant kinematics ---------------------------------------------------------------------
  double ux[3][3], px[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < 4; a++) s += nx[a][i] * yl[a][j];
      ux[i][j] = s;
    }
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < 4; a++) s += nx[a][i] * yl[a][3];
    px[i] = s;
  }
  const double divU = ux[0][0] + ux[1][1] + ux[2][2];
  double gam = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double es = ux[i][j] + ux[j][i];
      gam += es * es;
    }
  gam = sqrt(0.5 * gam);
  double mu, mu_g;
  viscosity(dm, gam, mu, mu_g);
  mu_g = is_zero(gam) ? 0.0 : mu_g / gam;
  if (NN) {
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int j = 0; j < 3; j++)
        rec[O_ES + 3 * a + j] = (mu_g != 0.0) ? (ux[0][j] + ux[j][0]) * nx[a][0] + (ux[1][j] + ux[j][1]) * nx[a][1] +
                                                    (ux[2][j] + ux[j][2]) * nx[a][2]
                                              : 0.0;
  }

  // ---- Gauss-point interpolations go to scratch; the nodal inputs die here --------------------------------
  double* S = rec + O_SB;   // 33 scratch doubles: S[8g+q]
#pragma unroll
  for (int g = 0; g < 4; g++) {
    double u[3] = {0.0, 0.0, 0.0}, ud[3] = {-dm.f[0], -dm.f[1], -dm.f[2]}, p = 0.0;
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double Na = P.N[g][a];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        u[i] += Na * uc[a][i];
        ud[i] += Na * ab[a][i];
      }
      p += Na * yl[a][3];
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      S[8 * g + i] = u[i];
      S[8 * g + 3 + i] = ud[i];
    }
    S[8 * g + 6] = p;
  }

  const double rho = dm.rho;
  const double T = P.af * P.gam * P.dt;
  const double amd = P.am / T;
  const double muKd = mu * dm.Kd;
  const double nu = mu / rho;
  const double kT = 4.0 / (P.dt * P.dt) + (dm.Kd * nu) * (dm.Kd * nu);
  const double kS = 36.0 * (k00 * k00 + k11 * k11 + k22 * k22 + 2.0 * (k01 * k01 + k02 * k02 + k12 * k12)) * nu * nu;
  const double kTS = kT + kS;
  const double trK = k00 + k11 + k22;

  // ---- loop 1: residual moments ---------------------------------------------------------------------------
  double RM[3][3], UP[3] = {0.0, 0.0, 0.0}, lR[4][3];
  double sPa = 0.0, sW = 0.0, sA2 = 0.0, sPP = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) RM[i][j] = 0.0;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int j = 0; j < 3; j++) lR[a][j] = 0.0;

#pragma unroll 1
  for (int g = 0; g < 4; g++) {
    const double wJ = P.w[g] * Jac;
    const double wl = wJ * T;
    const double u[3] = {S[8 * g], S[8 * g + 1], S[8 * g + 2]};
    const double ud[3] = {S[8 * g + 3], S[8 * g + 4], S[8 * g + 5]};
    const double p = S[8 * g + 6];
    const double kU = u[0] * u[0] * k00 + u[1] * u[1] * k11 + u[2] * u[2] * k22 +
                      2.0 * (u[0] * u[1] * k01 + u[0] * u[2] * k02 + u[1] * u[2] * k12);
    const double tauM = 1.0 / (rho * sqrt(kTS + kU));
    double up[3], ua[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double rV = ud[j] + u[0] * ux[0][j] + u[1] * ux[1][j] + u[2] * ux[2][j];
      up[j] = -tauM * (rho * rV + px[j] + muKd * u[j]);
    }
    const double tauC = 1.0 / (tauM * trK);
    double tauB = up[0] * up[0] * k00 + up[1] * up[1] * k11 + up[2] * up[2] * k22 +
                  2.0 * (up[0] * up[1] * k01 + up[0] * up[2] * k02 + up[1] * up[2] * k12);
    if (is_zero(tauB)) tauB = 2.220446049250313e-16;
    tauB = rho / sqrt(tauB);
#pragma unroll
    for (int i = 0; i < 3; i++) ua[i] = u[i] + up[i];
    const double pa = p - tauC * divU;
    sPa += wJ * pa;
    sW += wJ;
    sA2 += wl * tauC;
    sPP += wl * tauM;
    double Aj[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double rVb = tauB * (up[0] * ux[0][j] + up[1] * ux[1][j] + up[2] * ux[2][j]);
      const double rV2 = ud[j] + ua[0] * ux[0][j] + ua[1] * ux[1][j] + ua[2] * ux[2][j];
      Aj[j] = rho * rV2 + muKd * ua[j];
#pragma unroll
      for (int i = 0; i < 3; i++) RM[i][j] += wJ * (rVb * up[i] - rho * up[j] * ua[i]);
      UP[j] += wJ * up[j];
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double wN = wJ * P.N[g][a];
#pragma unroll
      for (int j = 0; j < 3; j++) lR[a][j] += wN * Aj[j];
    }
    S[8 * g + 3] = up[0]; S[8 * g + 4] = up[1]; S[8 * g + 5] = up[2];
    S[8 * g + 6] = tauM;
    S[8 * g + 7] = tauB;
  }

  // ---- finish the residual -----------------------------------------------------------------------------
  double Sb[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const double wl = P.w[g] * Jac * T;
#pragma unroll
    for (int a = 0; a < 4; a++) Sb[a] += wl * P.N[g][a];
  }
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) RM[i][j] += mu * (ux[i][j] + ux[j][i]) * sW;
    RM[i][i] -= sPa;
  }
  const double iT = 1.0 / T;
#pragma unroll
  for (int a = 0; a < 4; a++) {
#pragma unroll
    for (int j = 0; j < 3; j++)
      lRout[4 * a + j] = lR[a][j] + (nx[a][0] * RM[0][j] + nx[a][1] * RM[1][j] + nx[a][2] * RM[2][j]);
    lRout[4 * a + 3] = divU * (Sb[a] * iT) - (UP[0] * nx[a][0] + UP[1] * nx[a][1] + UP[2] * nx[a][2]);
  }
  const double sWl = sW * T;
  const double A1 = mu * sWl, A3 = mu_g * sWl;

  // ---- loop 2: tangent moments ---------------------------------------------------------------------------
  const double q1c = rho * amd + muKd;
  double D[4][4], S2[4] = {0.0, 0.0, 0.0, 0.0}, S3[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) D[a][b] = 0.0;
#pragma unroll 1
  for (int g = 0; g < 4; g++) {
    const double wl = P.w[g] * Jac * T;
    const double u[3] = {S[8 * g], S[8 * g + 1], S[8 * g + 2]};
    const double up[3] = {S[8 * g + 3], S[8 * g + 4], S[8 * g + 5]};
    const double tauM = S[8 * g + 6], tauB = S[8 * g + 7];
    double Q1[4], Q2[4], Q3[4], Q4[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double uNx = u[0] * nx[a][0] + u[1] * nx[a][1] + u[2] * nx[a][2];
      const double upNx = up[0] * nx[a][0] + up[1] * nx[a][1] + up[2] * nx[a][2];
      Q1[a] = q1c * P.N[g][a];
      Q2[a] = uNx + upNx;
      Q3[a] = upNx;
      Q4[a] = rho * uNx;
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double Na = P.N[g][a];
      const double c = rho * tauM * Q2[a];
      const double P1 = wl * (Na + c), P2 = wl * rho * Na, P3 = wl * tauB * Q3[a], P4 = wl * c;
      S2[a] += P4;
      S3[a] -= wl * tauM * (Q4[a] + Q1[a]);
#pragma unroll
      for (int b = 0; b < 4; b++) D[a][b] += P1 * Q1[b] + P2 * Q2[b] + P3 * Q3[b] + P4 * Q4[b];
    }
  }
#pragma unroll
  for (int a = 0; a < 4; a++) {
    rec[O_SB + a] = Sb[a];
    rec[O_S2 + a] = S2[a];
    rec[O_S3 + a] = S3[a];
#pragma unroll
    for (int b = 0; b < 4; b++) rec[O_D + 4 * a + b] = D[a][b];
  }
  rec[O_A1] = A1;
  rec[O_A2] = sA2;
  rec[O_SPP] = sPP;
  if (NN) rec[O_A3] = A3;
}

// K += lK(:,a,b) of the element whose record is r (tet4_block from the record).
SVB_HD void tet4_block_rec_add(const double* r, const bool NN, const int a, const int b, double K[16])
{
  const double xa[3] = {r[O_NX + 3 * a], r[O_NX + 3 * a + 1], r[O_NX + 3 * a + 2]};
  const double xb[3] = {r[O_NX + 3 * b], r[O_NX + 3 * b + 1], r[O_NX + 3 * b + 2]};
  const double A1 = r[O_A1], A2 = r[O_A2];
  const double nn = xa[0] * xb[0] + xa[1] * xb[1] + xa[2] * xb[2];
  const double dd = r[O_D + 4 * a + b] + A1 * nn;
  const double Sba = r[O_SB + a], Sbb = r[O_SB + b], S2a = r[O_S2 + a], S3b = r[O_S3 + b];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double q1 = A1 * xb[i];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double v = xa[j] * q1 + xa[i] * (A2 * xb[j]);
      if (i == j) v += dd;
      K[4 * i + j] += v;
    }
    K[4 * i + 3] += xb[i] * S2a - xa[i] * Sbb;
    K[12 + i] += xb[i] * Sba - xa[i] * S3b;
  }
  K[15] += r[O_SPP] * nn;
  if (NN) {
    const double A3 = r[O_A3];
    if (A3 != 0.0) {
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double t = A3 * r[O_ES + 3 * a + i];
#pragma unroll
        for (int j = 0; j < 3; j++) K[4 * i + j] += t * r[O_ES + 3 * b + j];
      }
    }
  }
}

// Both blocks of an edge from one set of loads: K1 += lK(:,a,b), K2 += lK(:,b,a).  The velocity-velocity
// parts are transposes of each other up to the diagonal term: W(i,j) = A1 xa_j xb_i + A2 xa_i xb_j.
SVB_HD void tet4_edge_rec_add(const double* r, const bool NN, const int a, const int b, double K1[16], double K2[16])
{
  const double xa[3] = {r[O_NX + 3 * a], r[O_NX + 3 * a + 1], r[O_NX + 3 * a + 2]};
  const double xb[3] = {r[O_NX + 3 * b], r[O_NX + 3 * b + 1], r[O_NX + 3 * b + 2]};
  const double A1 = r[O_A1], A2 = r[O_A2];
  double m[3][3];
#pragma unroll
  for (int p = 0; p < 3; p++)
#pragma unroll
    for (int q = 0; q < 3; q++) m[p][q] = xa[p] * xb[q];
  const double nn = m[0][0] + m[1][1] + m[2][2];
  const double dd1 = r[O_D + 4 * a + b] + A1 * nn;
  const double dd2 = r[O_D + 4 * b + a] + A1 * nn;
  const double Sba = r[O_SB + a], Sbb = r[O_SB + b];
  const double S2a = r[O_S2 + a], S2b = r[O_S2 + b], S3a = r[O_S3 + a], S3b = r[O_S3 + b];
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double W = A1 * m[j][i] + A2 * m[i][j];
      K1[4 * i + j] += W;
      K2[4 * j + i] += W;
    }
    K1[4 * i + i] += dd1;
    K2[4 * i + i] += dd2;
    K1[4 * i + 3] += xb[i] * S2a - xa[i] * Sbb;
    K2[4 * i + 3] += xa[i] * S2b - xb[i] * Sba;
    K1[12 + i] += xb[i] * Sba - xa[i] * S3b;
    K2[12 + i] += xa[i] * Sbb - xb[i] * S3a;
  }
  const double pp = r[O_SPP] * nn;
  K1[15] += pp;
  K2[15] += pp;
  if (NN) {
    const double A3 = r[O_A3];
    if (A3 != 0.0) {
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double ta = A3 * r[O_ES + 3 * a + i], tb = A3 * r[O_ES + 3 * b + i];
#pragma unroll
        for (int j = 0; j < 3; j++) {
          K1[4 * i + j] += ta * r[O_ES + 3 * b + j];
          K2[4 * i + j] += tb * r[O_ES + 3 * a + j];
        }
      }
    }
  }
}

// =====================================================================================================
// Accumulator form of the two routines above, used by the grouped scatter since round 2.  A thread that owns an
// edge sums the contributions of ~5 elements; instead of forming both 4x4 blocks per element (about 90 flops each way)
// it accumulates the BILINEAR pieces the blocks are made of and assembles the two blocks once at the end:
//   P(p,q) += A1 xa_p xb_q,  Q(p,q) += A2 xa_p xb_q     =>  K1_uu(i,j) = P(j,i) + Q(i,j),  K2_uu = K1_uu^T,
//   sum A1 (grad Na . grad Nb) = tr P                    (the diagonal term needs no extra work),
//   T1 += xb Sb_a, T2 += xa Sb_b, T3 += xb S2_a, T4 += xa S3_b, T5 += xa S2_b, T6 += xb S3_a
//                                                        =>  K1_up = T3 - T2, K1_pu = T1 - T4, K2_up = T5 - T1, K2_pu = T2 - T6,
// 48 flops per contribution instead of ~90; same terms, different association (agreement with the reference stays at 1e-15).
struct EdgeAcc {
  double P[9], Q[9], T[18], E[9];
  double dd1, dd2, pp;
};

SVB_HD void edge_acc_zero(EdgeAcc& A, const bool NN)
{
#pragma unroll
  for (int k = 0; k < 9; k++) { A.P[k] = 0.0; A.Q[k] = 0.0; }
#pragma unroll
  for (int k = 0; k < 18; k++) A.T[k] = 0.0;
  if (NN) {
#pragma unroll
    for (int k = 0; k < 9; k++) A.E[k] = 0.0;
  }
  A.dd1 = 0.0; A.dd2 = 0.0; A.pp = 0.0;
}

SVB_HD void tet4_edge_rec_acc(const double* r, const bool NN, const int a, const int b, EdgeAcc& A)
{
  const double xa[3] = {r[O_NX + 3 * a], r[O_NX + 3 * a + 1], r[O_NX + 3 * a + 2]};
  const double xb[3] = {r[O_NX + 3 * b], r[O_NX + 3 * b + 1], r[O_NX + 3 * b + 2]};
  const double A1 = r[O_A1], A2 = r[O_A2];
#pragma unroll
  for (int p = 0; p < 3; p++) {
    const double a1 = A1 * xa[p], a2 = A2 * xa[p];
#pragma unroll
    for (int q = 0; q < 3; q++) {
      A.P[3 * p + q] += a1 * xb[q];
      A.Q[3 * p + q] += a2 * xb[q];
    }
  }
  A.dd1 += r[O_D + 4 * a + b];
  A.dd2 += r[O_D + 4 * b + a];
  A.pp += r[O_SPP] * (xa[0] * xb[0] + xa[1] * xb[1] + xa[2] * xb[2]);
  const double Sba = r[O_SB + a], Sbb = r[O_SB + b];
  const double S2a = r[O_S2 + a], S2b = r[O_S2 + b], S3a = r[O_S3 + a], S3b = r[O_S3 + b];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    A.T[i] += xb[i] * Sba;
    A.T[3 + i] += xa[i] * Sbb;
    A.T[6 + i] += xb[i] * S2a;
    A.T[9 + i] += xa[i] * S3b;
    A.T[12 + i] += xa[i] * S2b;
    A.T[15 + i] += xb[i] * S3a;
  }
  if (NN) {
    const double A3 = r[O_A3];
    if (A3 != 0.0) {
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double t = A3 * r[O_ES + 3 * a + i];
#pragma unroll
        for (int j = 0; j < 3; j++) A.E[3 * i + j] += t * r[O_ES + 3 * b + j];
      }
    }
  }
}

// which = 0: lK(:,a,b); which = 1: lK(:,b,a) of the accumulated edge.
SVB_HD void edge_acc_block(const EdgeAcc& A, const bool NN, const int which, double K[16])
{
  const double trP = A.P[0] + A.P[4] + A.P[8];
  const double dd = (which == 0 ? A.dd1 : A.dd2) + trP;
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      // K1(i,j) = P(j,i) + Q(i,j);  K2(i,j) = K1(j,i)
      double v = (which == 0) ? A.P[3 * j + i] + A.Q[3 * i + j] : A.P[3 * i + j] + A.Q[3 * j + i];
      if (NN) v += (which == 0) ? A.E[3 * i + j] : A.E[3 * j + i];
      if (i == j) v += dd;
      K[4 * i + j] = v;
    }
    K[4 * i + 3] = (which == 0) ? A.T[6 + i] - A.T[3 + i] : A.T[12 + i] - A.T[i];
    K[12 + i] = (which == 0) ? A.T[i] - A.T[9 + i] : A.T[3 + i] - A.T[15 + i];
  }
  K[15] = A.pp;
}

// Diagonal block (a, a): W = (A1 + A2) xa xa^T is symmetric.
struct DiagAcc {
  double S[6];       // (00, 01, 02, 11, 12, 22) of sum (A1 + A2) xa xa^T
  double up[3], pu[3], E[9];
  double dd, pp;
};

SVB_HD void diag_acc_zero(DiagAcc& A, const bool NN)
{
#pragma unroll
  for (int k = 0; k < 6; k++) A.S[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 3; k++) { A.up[k] = 0.0; A.pu[k] = 0.0; }
  if (NN) {
#pragma unroll
    for (int k = 0; k < 9; k++) A.E[k] = 0.0;
  }
  A.dd = 0.0; A.pp = 0.0;
}

SVB_HD void tet4_diag_rec_acc(const double* r, const bool NN, const int a, DiagAcc& A)
{
  const double xa[3] = {r[O_NX + 3 * a], r[O_NX + 3 * a + 1], r[O_NX + 3 * a + 2]};
  const double A1 = r[O_A1], A2 = r[O_A2];
  const double c = A1 + A2;
  const double c0 = c * xa[0], c1 = c * xa[1], c2 = c * xa[2];
  A.S[0] += c0 * xa[0]; A.S[1] += c0 * xa[1]; A.S[2] += c0 * xa[2];
  A.S[3] += c1 * xa[1]; A.S[4] += c1 * xa[2]; A.S[5] += c2 * xa[2];
  const double nn = xa[0] * xa[0] + xa[1] * xa[1] + xa[2] * xa[2];
  A.dd += r[O_D + 5 * a] + A1 * nn;
  A.pp += r[O_SPP] * nn;
  const double Sba = r[O_SB + a];
  const double du = r[O_S2 + a] - Sba, dp = Sba - r[O_S3 + a];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    A.up[i] += xa[i] * du;
    A.pu[i] += xa[i] * dp;
  }
  if (NN) {
    const double A3 = r[O_A3];
    if (A3 != 0.0) {
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double t = A3 * r[O_ES + 3 * a + i];
#pragma unroll
        for (int j = 0; j < 3; j++) A.E[3 * i + j] += t * r[O_ES + 3 * a + j];
      }
    }
  }
}

SVB_HD void diag_acc_block(const DiagAcc& A, const bool NN, double K[16])
{
  const double S[3][3] = {{A.S[0], A.S[1], A.S[2]}, {A.S[1], A.S[3], A.S[4]}, {A.S[2], A.S[4], A.S[5]}};
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double v = S[i][j];
      if (NN) v += A.E[3 * i + j];
      if (i == j) v += A.dd;
      K[4 * i + j] = v;
    }
    K[4 * i + 3] = A.up[i];
    K[12 + i] = A.pu[i];
  }
  K[15] = A.pp;
}

}  // namespace svb
