// heat_host.cpp — TEST-ONLY host build of heat_elem.cuh: plain element loop over the device Gauss-point routines
// (gnn3_metric, heat_gauss_point, heat_row), checked against the R / Val the reference assembled (tests/golden/heat.npz).
#include <cmath>
#include <cstring>
using std::fabs; using std::sqrt;
#define SVB_HD inline
#include "../../svmultiphysics_b200/csrc/heat_elem.cuh"

struct HostHeatArgs {
  const int* IEN; const double *x, *Ag, *Yg;
  int eNoN, nEl, nG, tDof, s, mvMsh, fluid, pad;
  double dt, af, am, gam;
  double w[8], N[8][8], Nxi[8][8][3];
  svb::HeatDmn dm;
};

template <int ENON, bool FLUID>
static int run(const HostHeatArgs* P, const int* rowPtr, const int* colPtr, double* R, double* Val)
{
  using namespace svb;
  const double T1 = P->af * P->gam * P->dt;
  for (int e = 0; e < P->nEl; e++) {
    int n[ENON];
    double xl[ENON][3], Tl[ENON], Tdl[ENON], ul[ENON][3];
    for (int a = 0; a < ENON; a++) {
      n[a] = P->IEN[ENON * e + a];
      const double* y = P->Yg + (size_t)P->tDof * n[a];
      for (int i = 0; i < 3; i++) {
        xl[a][i] = P->x[3 * n[a] + i];
        ul[a][i] = FLUID ? (y[i] - (P->mvMsh ? y[4 + i] : 0.0)) : 0.0;
      }
      Tl[a] = y[P->s];
      Tdl[a] = P->Ag[(size_t)P->tDof * n[a] + P->s];
    }
    double lR[ENON] = {}, lK[ENON][ENON] = {};
    double Nx[ENON][3], ks[3][3], Jac = 1.0;
    for (int g = 0; g < P->nG; g++) {
      if (g == 0 || ENON != 4) Jac = gnn3_metric<ENON>(P->Nxi[g], xl, Nx, ks);
      const double w = P->w[g] * Jac;
      HeatGP q;
      heat_gauss_point<ENON, FLUID>(P->dm, P->dt, P->af, P->am, P->gam, P->N[g], Nx, ks, Tl, Tdl, ul, q);
      for (int a = 0; a < ENON; a++) heat_row<ENON>(q, w, w * T1, P->N[g][a], Nx[a], P->N[g], Nx, lR[a], lK[a]);
    }
    for (int a = 0; a < ENON; a++) {
      R[n[a]] += lR[a];
      for (int b = 0; b < ENON; b++) {
        int sl = -1;
        for (int k = rowPtr[n[a]]; k < rowPtr[n[a] + 1]; k++) if (colPtr[k] == n[b]) { sl = k; break; }
        if (sl < 0) return 1;
        Val[sl] += lK[a][b];
      }
    }
  }
  return 0;
}

extern "C" int hostmath_heat(const HostHeatArgs* P, const int* rowPtr, const int* colPtr, double* R, double* Val)
{
  if (P->eNoN == 4) return P->fluid ? run<4, true>(P, rowPtr, colPtr, R, Val) : run<4, false>(P, rowPtr, colPtr, R, Val);
  if (P->eNoN == 8) return P->fluid ? run<8, true>(P, rowPtr, colPtr, R, Val) : run<8, false>(P, rowPtr, colPtr, R, Val);
  return 3;
}
extern "C" int hostmath_sizeof_heatargs() { return (int)sizeof(HostHeatArgs); }
