// tet4_host.cpp — TEST-ONLY host build of the closed-form linear-tetrahedron routines of struct_elem.cuh (tet4_moments,
// struct_tet4_residual / struct_tet4_block, lelas_tet4_stress / _residual / _block): the element loops of
// assemble_struct_tet4_kernel and assemble_mesh_tet4_kernel written as plain loops, checked against the R / Val the reference
// assembled with its Gauss loops (tests/golden/struct.npz, tests/golden/lelas.npz).
#include <cmath>
#include <cstring>
using std::fabs; using std::sqrt; using std::pow; using std::exp;
#define SVB_HD inline
#include "../../svmultiphysics_b200/csrc/struct_elem.cuh"

struct HostTet4Args {
  const int* IEN; const double* fN; const double *x, *Ag, *Yg, *Dg, *Bf, *Do;
  int nEl, tDof, dof, s, nFn, kind;       // kind: 0 struct_3d, 1 lElas equation, 2 mesh equation
  double dt, af, am, gam, beta;
  double w[8], N[8][8], Nxi[8][8][3];
  svb::StructDmn dm;                      // kind 1 / 2: elasticity modulus in C10, Poisson ratio in C01
  const double* Ya;                       // (nNo, 3) active tensions or null
  svb::CannRow cann[16];
};

static int scatter(const int* rowPtr, const int* colPtr, int dof, const int n[4], int a, int b, const double K[3][3], bool transposed, double* Val)
{
  int sl = -1;
  for (int k = rowPtr[n[a]]; k < rowPtr[n[a] + 1]; k++) if (colPtr[k] == n[b]) { sl = k; break; }
  if (sl < 0) return 1;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Val[(size_t)dof * dof * sl + dof * i + j] += transposed ? K[j][i] : K[i][j];
  return 0;
}

extern "C" int hostmath_tet4(const HostTet4Args* P, const int* rowPtr, const int* colPtr, double* R, double* Val)
{
  using namespace svb;
  const int dof = P->dof, s0 = P->s, tD = P->tDof;
  for (int e = 0; e < P->nEl; e++) {
    int n[4];
    double xl[4][3], dl[4][3], ql[4][3], fN[2][3] = {{0, 0, 0}, {0, 0, 0}};
    for (int a = 0; a < 4; a++) {
      n[a] = P->IEN[4 * e + a];
      for (int i = 0; i < 3; i++) {
        const double dol = (P->kind == 2) ? P->Do[(size_t)tD * n[a] + s0 + i] : 0.0;
        xl[a][i] = P->x[3 * n[a] + i] + dol;
        dl[a][i] = P->Dg[(size_t)tD * n[a] + s0 + i] - dol;
        if (P->kind == 0)
          ql[a][i] = P->dm.rho * (P->Ag[(size_t)tD * n[a] + s0 + i] - P->Bf[3 * n[a] + i]) + P->dm.dmp * P->Yg[(size_t)tD * n[a] + s0 + i];
        else
          ql[a][i] = P->Ag[(size_t)tD * n[a] + s0 + i] - (P->kind == 1 ? P->Bf[3 * n[a] + i] : 0.0);
      }
    }
    for (int k = 0; k < P->nFn && k < 2; k++) for (int i = 0; i < 3; i++) fN[k][i] = P->fN[(size_t)3 * P->nFn * e + 3 * k + i];
    double Nx[4][3];
    const double Jac = gnn3<4>(P->Nxi[0], xl, Nx);
    Tet4Mom q;
    tet4_moments(P->w, &P->N[0][0], 8, (P->kind == 2) ? 1.0 : Jac, q);
    if (P->kind == 0) {
      const double afu = P->af * P->beta * P->dt * P->dt;
      const double amd = P->am * P->dm.rho + P->af * P->gam * P->dt * P->dm.dmp;
      double F[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, S[3][3], Dm[6][6], Pk[3][3];
      for (int a = 0; a < 4; a++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) F[i][j] += Nx[a][j] * dl[a][i];
      // active tensions at the weighted mean of the Gauss points (S, Dm affine in them), as assemble_struct_tet4_kernel does
      double ya[3] = {0, 0, 0};
      const bool act = P->Ya && P->dm.active;
      if (act) {
        for (int a = 0; a < 4; a++) for (int i = 0; i < 3; i++) ya[i] += q.m1[a] * P->Ya[3 * n[a] + i];
        for (int i = 0; i < 3; i++) ya[i] /= q.W;
      }
      if (pk2cc_voigt(P->dm, F, fN, act ? ya : nullptr, P->cann, P->nFn, S, Dm)) return 2;
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Pk[i][j] = F[i][0] * S[0][j] + F[i][1] * S[1][j] + F[i][2] * S[2][j];
      for (int a = 0; a < 4; a++) {
        double r[3];
        struct_tet4_residual(P->dm, q, a, Nx[a], Pk, ql, r);
        for (int i = 0; i < 3; i++) R[dof * n[a] + i] += r[i];
      }
      const double wafu = q.W * afu;
      for (int b = 0; b < 4; b++) {
        double Bmb[6][3], DBmb[6][3], SNb[3];
        make_Bm(Nx[b], F, Bmb);
        make_DBm(Dm, Bmb, DBmb);
        for (int i = 0; i < 3; i++) SNb[i] = S[i][0] * Nx[b][0] + S[i][1] * Nx[b][1] + S[i][2] * Nx[b][2];
        for (int a = 0; a <= b; a++) {
          double Bma[6][3], K[3][3];
          make_Bm(Nx[a], F, Bma);
          struct_tet4_block(wafu, amd * q.M2[a][b], Nx[a], SNb, Bma, DBmb, K);
          if (scatter(rowPtr, colPtr, dof, n, a, b, K, false, Val)) return 1;
          if (a != b && scatter(rowPtr, colPtr, dof, n, b, a, K, true, Val)) return 1;
        }
      }
    } else {
      const double elM = P->dm.C10, nu = P->dm.C01, rho = P->dm.rho;
      const double lambda = elM * nu / (1.0 + nu) / (1.0 - 2.0 * nu), mu = elM * 0.5 / (1.0 + nu), lDm = lambda / mu;
      const double T1c = P->af * P->beta * P->dt * P->dt, amd = P->am / T1c * rho;
      double S[6];
      lelas_tet4_stress(lambda, mu, Nx, dl, S);
      for (int a = 0; a < 4; a++) {
        double r[3];
        lelas_tet4_residual(rho, P->dm.f, q, a, Nx[a], S, ql, r);
        for (int i = 0; i < 3; i++) R[dof * n[a] + i] += r[i];
      }
      const double c0 = T1c * amd, c1 = T1c * mu * q.W;
      for (int b = 0; b < 4; b++)
        for (int a = 0; a <= b; a++) {
          double K[3][3];
          lelas_tet4_block(c0 * q.M2[a][b], c1, lDm, Nx[a], Nx[b], K);
          if (scatter(rowPtr, colPtr, dof, n, a, b, K, false, Val)) return 1;
          if (a != b && scatter(rowPtr, colPtr, dof, n, b, a, K, true, Val)) return 1;
        }
    }
  }
  return 0;
}
extern "C" int hostmath_sizeof_tet4args() { return (int)sizeof(HostTet4Args); }
