// heat_elem.cuh — Gauss-point algebra of the two scalar (dof = 1) heat equations, __host__ __device__ so that the
// CPU suite can check it against the compiled reference (tests/hostmath/heat_host.cpp) before a GPU is involved.
//
//   heatS  heats::heats_3d   Code/Source/solver/heats.cpp:186-233   rho c dT/dt = div(k grad T) + s  in a solid
//   heatF  heatf::heatf_3d   Code/Source/solver/heatf.cpp:238-331   advection-diffusion with the fluid velocity of state
//          dofs 0..2, SUPG-like stabilisation tauM = 1 / sqrt(4/dt^2 + u.K.u + 3 nu^2 K:K) and the discontinuity-capturing
//          conductivity nu + |Td + u.grad T| / (2 sqrt(grad T . K . grad T))
#pragma once
#ifndef SVB_HD
#define SVB_HD __host__ __device__ __forceinline__
#endif

namespace svb {

struct HeatDmn {
  double rho, nu, s;     // solid_density (heatS only), conductivity, source_term
  int Id, active, pad0, pad1;
};

// nn::gnn for insd = 3 (Code/Source/solver/nn.cpp:862-899) with the metric ks = xiX^T xiX that heatF needs.
template <int ENON>
SVB_HD double gnn3_metric(const double Nxi[][3], const double xl[][3], double Nx[][3], double ks[3][3])
{
  double xXi[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, xiX[3][3];
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

// What one Gauss point leaves for the rows: lR(a) += w (N_a c0 + nu (Nx_a . Tx) - udNx_a Tp),
// lK(a,b) += wl (nu Nx_a . Nx_b + (N_a + tauM udNx_a)(N_b amd + udNx_b))     [heatS: tauM = 0, udNx = 0, Tp = 0]
struct HeatGP {
  double c0, nu, Tp, tauM, amd, Tx[3], u[3];
};

// Tl/Tdl: nodal temperature and its rate (yl(s,a), al(s,a)); ul: nodal convective velocity (heatF; already minus
// the mesh velocity when mvMsh).
template <int ENON, bool FLUID>
SVB_HD void heat_gauss_point(const HeatDmn& dm, double dt, double af, double am, double gam, const double N[], const double Nx[][3],
                             const double ks[3][3], const double Tl[], const double Tdl[], const double ul[][3], HeatGP& q)
{
  const double T1 = af * gam * dt;
  double Td = -dm.s, Tx[3] = {0, 0, 0}, u[3] = {0, 0, 0};
#pragma unroll
  for (int a = 0; a < ENON; a++) {
    Td += N[a] * Tdl[a];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      Tx[i] += Nx[a][i] * Tl[a];
      if (FLUID) u[i] += N[a] * ul[a][i];
    }
  }
  q.Tx[0] = Tx[0]; q.Tx[1] = Tx[1]; q.Tx[2] = Tx[2];
  if (!FLUID) {
    q.c0 = Td * dm.rho; q.nu = dm.nu; q.Tp = 0.0; q.tauM = 0.0; q.amd = am * dm.rho / T1;
    q.u[0] = q.u[1] = q.u[2] = 0.0;
    return;
  }
  double kU = 0.0, kS = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      kU += u[i] * u[j] * ks[i][j];
      kS += ks[i][j] * ks[i][j];
    }
  double nTx = ks[0][0] * Tx[0] * Tx[0] + ks[1][1] * Tx[1] * Tx[1] + ks[2][2] * Tx[2] * Tx[2] +
               (ks[0][1] + ks[1][0]) * Tx[0] * Tx[1] + (ks[0][2] + ks[2][0]) * Tx[0] * Tx[2] + (ks[1][2] + ks[2][1]) * Tx[1] * Tx[2];
  // utils::is_zero(x) (Code/Source/solver/utils.cpp:141-159) is |x| / max(|x|, eps) < 10 eps, i.e. |x| < 10 eps^2
  if (fabs(nTx) < 10.0 * 2.220446049250313e-16 * 2.220446049250313e-16) nTx = 2.220446049250313e-16;
  const double udTx = u[0] * Tx[0] + u[1] * Tx[1] + u[2] * Tx[2];
  const double r = Td + udTx;
  const double nu = dm.nu + 0.5 * fabs(r) / sqrt(nTx);
  const double tauM = 1.0 / sqrt(4.0 / (dt * dt) + kU + 3.0 * nu * nu * kS);
  q.c0 = r; q.nu = nu; q.tauM = tauM; q.Tp = -tauM * r; q.amd = am / T1;
  q.u[0] = u[0]; q.u[1] = u[1]; q.u[2] = u[2];
}

// Row of node a (shape value Na, gradient Nxa) of the element matrix / residual for one Gauss point.
template <int ENON>
SVB_HD void heat_row(const HeatGP& q, double w, double wl, double Na, const double Nxa[3], const double N[], const double Nx[][3],
                     double& lR, double lK[])
{
  const double udNa = q.u[0] * Nxa[0] + q.u[1] * Nxa[1] + q.u[2] * Nxa[2];
  lR += w * (Na * q.c0 + (Nxa[0] * q.Tx[0] + Nxa[1] * q.Tx[1] + Nxa[2] * q.Tx[2]) * q.nu - udNa * q.Tp);
  const double ta = Na + q.tauM * udNa;
#pragma unroll
  for (int b = 0; b < ENON; b++) {
    const double udNb = q.u[0] * Nx[b][0] + q.u[1] * Nx[b][1] + q.u[2] * Nx[b][2];
    lK[b] += wl * (q.nu * (Nxa[0] * Nx[b][0] + Nxa[1] * Nx[b][1] + Nxa[2] * Nx[b][2]) + ta * (N[b] * q.amd + udNb));
  }
}

}  // namespace svb
