// struct_host.cpp — TEST-ONLY host build of struct_elem.cuh: plain element loop using the device Gauss-point
// routines (pk2cc_voigt, make_Bm, make_DBm, struct_block, gnn3), checked against the reference oracle.
#include <cmath>
#include <cstring>
using std::fabs; using std::sqrt; using std::pow; using std::exp;
#define SVB_HD inline
#include "../../svmultiphysics_b200/csrc/struct_elem.cuh"

struct HostStructArgs {
  const int* IEN; const double* fN; const double *x, *Ag, *Yg, *Dg, *Bf;
  int eNoN, nEl, nG, tDof, dof, s, nFn;
  double dt, af, am, gam, beta;
  double w[8], N[8][8], Nxi[8][8][3];
  svb::StructDmn dm;
  const double* Ya;              // (nNo, 3) active tensions f, s, n or null
  svb::CannRow cann[16];
};

template <int ENON>
static int run(const HostStructArgs* P, const int* rowPtr, const int* colPtr, double* R, double* Val)
{
  using namespace svb;
  const int dof = P->dof, s0 = P->s;
  const double afu = P->af * P->beta * P->dt * P->dt;
  const double amd = P->am * P->dm.rho + P->af * P->gam * P->dt * P->dm.dmp;
  for (int e = 0; e < P->nEl; e++) {
    int n[ENON];
    double xl[ENON][3], q[ENON][3], dl[ENON][3], yl[ENON][3], fN[2][3] = {{0, 0, 0}, {0, 0, 0}};
    for (int a = 0; a < ENON; a++) {
      n[a] = P->IEN[ENON * e + a];
      for (int i = 0; i < 3; i++) {
        xl[a][i] = P->x[3 * n[a] + i];
        dl[a][i] = P->Dg[P->tDof * n[a] + s0 + i];
        yl[a][i] = P->Yg[P->tDof * n[a] + s0 + i];
        q[a][i] = P->dm.rho * (P->Ag[P->tDof * n[a] + s0 + i] - P->Bf[3 * n[a] + i]) + P->dm.dmp * P->Yg[P->tDof * n[a] + s0 + i];
      }
    }
    for (int k = 0; k < P->nFn && k < 2; k++) for (int i = 0; i < 3; i++) fN[k][i] = P->fN[(size_t)3 * P->nFn * e + 3 * k + i];
    double lR[ENON][3] = {}, lK[ENON][ENON][3][3] = {};
    for (int g = 0; g < P->nG; g++) {
      double Nx[ENON][3];
      const double Jac = gnn3<ENON>(P->Nxi[g], xl, Nx);
      const double w = P->w[g] * Jac;
      double F[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, ud[3] = {-P->dm.rho * P->dm.f[0], -P->dm.rho * P->dm.f[1], -P->dm.rho * P->dm.f[2]};
      for (int a = 0; a < ENON; a++)
        for (int i = 0; i < 3; i++) {
          ud[i] += P->N[g][a] * q[a][i];
          for (int j = 0; j < 3; j++) F[i][j] += Nx[a][j] * dl[a][i];
        }
      double S[3][3], Dm[6][6], ya[3] = {0, 0, 0};
      const bool act = P->Ya && P->dm.active;
      if (act) for (int a = 0; a < ENON; a++) for (int i = 0; i < 3; i++) ya[i] += P->N[g][a] * P->Ya[3 * n[a] + i];
      if (pk2cc_voigt(P->dm, F, fN, act ? ya : nullptr, P->cann, P->nFn, S, Dm)) return 2;
      const bool visc = P->dm.viscType != 0 && P->dm.visc_mu != 0.0;
      const double afv = P->af * P->gam * P->dt;
      ViscGP vgp;
      double V[ENON][9];
      if (visc) {
        double vx[3][3] = {}, Svis[3][3];
        for (int a = 0; a < ENON; a++)
          for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) vx[i][j] += Nx[a][j] * yl[a][i];
        visc_gauss_point(P->dm.viscType, P->dm.visc_mu, afu, afv, F, vx, Svis, vgp);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) S[i][j] += Svis[i][j];
        for (int a = 0; a < ENON; a++) visc_node(vgp, Nx[a], V[a]);
      }
      double Pk[3][3];
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Pk[i][j] = F[i][0] * S[0][j] + F[i][1] * S[1][j] + F[i][2] * S[2][j];
      double Bm[ENON][6][3], DBm[ENON][6][3], SNx[ENON][3];
      for (int a = 0; a < ENON; a++) {
        for (int i = 0; i < 3; i++) lR[a][i] += w * (P->N[g][a] * ud[i] + Nx[a][0] * Pk[i][0] + Nx[a][1] * Pk[i][1] + Nx[a][2] * Pk[i][2]);
        make_Bm(Nx[a], F, Bm[a]);
        make_DBm(Dm, Bm[a], DBm[a]);
        for (int i = 0; i < 3; i++) SNx[a][i] = Nx[a][0] * S[0][i] + Nx[a][1] * S[1][i] + Nx[a][2] * S[2][i];
      }
      for (int a = 0; a < ENON; a++)
        for (int b = 0; b < ENON; b++) {
          struct_block(lK[a][b], w, amd * P->N[g][a] * P->N[g][b], afu, SNx[a], Nx[b], Bm[a], DBm[b]);
          if (visc) visc_block(P->dm.viscType, w * vgp.c, afu, afv, &vgp.M[0][0], V[a], V[b], lK[a][b]);
        }
    }
    for (int a = 0; a < ENON; a++) {
      for (int i = 0; i < 3; i++) R[dof * n[a] + i] += lR[a][i];
      for (int b = 0; b < ENON; b++) {
        int sl = -1;
        for (int k = rowPtr[n[a]]; k < rowPtr[n[a] + 1]; k++) if (colPtr[k] == n[b]) { sl = k; break; }
        if (sl < 0) return 1;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Val[(size_t)dof * dof * sl + dof * i + j] += lK[a][b][i][j];
      }
    }
  }
  return 0;
}

extern "C" int hostmath_struct(const HostStructArgs* P, const int* rowPtr, const int* colPtr, double* R, double* Val)
{
  if (P->eNoN == 4) return run<4>(P, rowPtr, colPtr, R, Val);
  if (P->eNoN == 8) return run<8>(P, rowPtr, colPtr, R, Val);
  return 3;
}
extern "C" int hostmath_sizeof_structargs() { return (int)sizeof(HostStructArgs); }

// compute_pk2cc of the device algebra for one deformation gradient: F(3,3) row-major, fN = fibre | sheet -> S(3,3), Dm(6,6).
extern "C" int hostmath_pk2cc(const svb::StructDmn* dm, const double* F, const double* fN, double* S, double* Dm, const double* ya,
                              const svb::CannRow* cann, int nFn)
{
  double Fm[3][3], f[2][3], Sm[3][3], D[6][6];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Fm[i][j] = F[3 * i + j];
  for (int k = 0; k < 2; k++) for (int i = 0; i < 3; i++) f[k][i] = fN ? fN[3 * k + i] : 0.0;
  const int rc = svb::pk2cc_voigt(*dm, Fm, f, ya, cann, nFn, Sm, D);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) S[3 * i + j] = Sm[i][j];
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Dm[6 * i + j] = D[i][j];
  return rc;
}
