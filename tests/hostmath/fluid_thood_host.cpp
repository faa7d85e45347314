// fluid_thood_host.cpp — TEST-ONLY host build of fluid_thood.cuh: plain element loop over the two Gauss loops of construct_fluid on a
// Taylor-Hood mesh (velocity rule: momentum; pressure rule: continuity), checked against tests/golden/fluid_thood.npz.
#include <cmath>
#include <cstring>
using std::fabs; using std::sqrt; using std::pow; using std::exp; using std::cos;
#define SVB_HD inline
#include "../../svmultiphysics_b200/csrc/fluid_thood.cuh"

struct HostThoodArgs {
  const int* IEN; const double *x, *Ag, *Yg, *Bf;
  const double *w, *N, *Nxi, *Nxi2;                  // velocity space, velocity rule: w[nG], N[nG][eNoN], Nxi[nG][eNoN][3], Nxi2[nG][eNoN][6]
  const double *Nq1, *Nqxi1;                         // pressure space at the velocity rule: [nG][eNoNq], [nG][eNoNq][3]
  const double *w2, *Nw2, *Nwxi2, *Nq2, *Nqxi2;      // pressure rule: w2[nG2], Nw2[nG2][eNoN], Nwxi2[nG2][eNoN][3], Nq2[nG2][eNoNq], Nqxi2[nG2][eNoNq][3]
  int eNoN, eNoNq, nEl, nG, nG2, tDof, mvMsh, lShpFq;
  double dt, af, am, gam;
  svb::FluidDmn dm;
  const double* uris;        // URIS valves: (nNo, nUris, 5) or null
  int nUris;
  svb200_uris urisP[SVB200_MAX_URIS];
};

template <int ENON, int ENONQ>
static int run(const HostThoodArgs* P, const int* rowPtr, const int* colPtr, double* R, double* Val)
{
  using namespace svb;
  const double T1 = P->af * P->gam * P->dt;
  for (int e = 0; e < P->nEl; e++) {
    int n[ENON];
    double xl[ENON][3], al[ENON][3], yl[ENON][4], bfl[ENON][3], ym[ENON][3];
    for (int a = 0; a < ENON; a++) {
      n[a] = P->IEN[ENON * e + a];
      for (int i = 0; i < 3; i++) {
        xl[a][i] = P->x[3 * n[a] + i];
        al[a][i] = P->Ag[P->tDof * n[a] + i];
        bfl[a][i] = P->Bf[3 * n[a] + i];
        ym[a][i] = P->mvMsh ? P->Yg[P->tDof * n[a] + 4 + i] : 0.0;
      }
      for (int i = 0; i < 4; i++) yl[a][i] = P->Yg[P->tDof * n[a] + i];
    }
    static double lK[ENON][ENON][16];
    double lR[ENON][4] = {};
    std::memset(lK, 0, sizeof(lK));
    for (int g = 0; g < P->nG; g++) {
      double Nx[ENON][3], Nxx[ENON][6], xiX[3][3], ks[3][3], Nqx[ENONQ][3], xq[3][3], kq[3][3];
      const double Jac = gnn3_full<ENON>(reinterpret_cast<const double(*)[3]>(P->Nxi + (size_t)g * ENON * 3), xl, Nx, xiX, ks);
      if (is_zero(Jac)) return 4;
      gn_nxx3<ENON>(reinterpret_cast<const double(*)[6]>(P->Nxi2 + (size_t)g * ENON * 6), xl, xiX, Nx, Nxx);
      const int gq = P->lShpFq ? 0 : g;
      gnn3_full<ENONQ>(reinterpret_cast<const double(*)[3]>(P->Nqxi1 + (size_t)gq * ENONQ * 3), xl, Nqx, xq, kq);
      const double* Nq = P->Nq1 + (size_t)g * ENONQ;
      FluidGP q;
      FluidNode nd[ENON];
      double uF = 0.0, uV[3] = {0.0, 0.0, 0.0};
      if (P->uris) uris_factor<ENON>(P->uris, P->nUris, P->urisP, P->N + (size_t)g * ENON, n, uF, uV);
      thood_gauss_point_m<ENON, ENONQ>(P->dm, P->dt, P->af, P->am, P->gam, P->w[g] * Jac, ks, P->N + (size_t)g * ENON, Nx, Nxx, Nq, Nqx,
                                       al, yl, bfl, P->mvMsh ? ym : nullptr, q, nd, uF, uV);
      for (int a = 0; a < ENON; a++) {
        thood_residual_m(q, nd[a], lR[a]);
        FluidRow row;
        thood_row(q, nd[a], row);
        for (int b = 0; b < ENON; b++) thood_block_m(row, nd[b], b < ENONQ ? Nq + b : nullptr, b < ENONQ ? Nqx[b] : nullptr, lK[a][b]);
      }
    }
    for (int g = 0; g < P->nG2; g++) {
      double Nx[ENON][3], xiX[3][3], ks[3][3], Nqx[ENONQ][3];
      const double Jw = gnn3_full<ENON>(reinterpret_cast<const double(*)[3]>(P->Nwxi2 + (size_t)g * ENON * 3), xl, Nx, xiX, ks);
      const int gq = P->lShpFq ? 0 : g;
      const double Jq = gnn3_full<ENONQ>(reinterpret_cast<const double(*)[3]>(P->Nqxi2 + (size_t)gq * ENONQ * 3), xl, Nqx, xiX, ks);
      if (is_zero(Jw) || is_zero(Jq)) return 4;
      const double Jac = (g == 0 || !P->lShpFq) ? Jq : Jw;
      const double w = P->w2[g] * Jac, wl = w * T1;
      double divU = 0.0;
      for (int b = 0; b < ENON; b++) divU += Nx[b][0] * yl[b][0] + Nx[b][1] * yl[b][1] + Nx[b][2] * yl[b][2];
      for (int a = 0; a < ENONQ; a++) {
        const double Nqa = P->Nq2[(size_t)g * ENONQ + a];
        lR[a][3] += w * (Nqa * divU);
        for (int b = 0; b < ENON; b++)
          for (int j = 0; j < 3; j++) lK[a][b][12 + j] += wl * Nqa * Nx[b][j];
      }
    }
    for (int a = 0; a < ENON; a++) {
      for (int i = 0; i < 4; i++) R[4 * n[a] + i] += lR[a][i];
      for (int b = 0; b < ENON; b++) {
        int sl = -1;
        for (int k = rowPtr[n[a]]; k < rowPtr[n[a] + 1]; k++) if (colPtr[k] == n[b]) { sl = k; break; }
        if (sl < 0) return 1;
        for (int i = 0; i < 16; i++) Val[(size_t)16 * sl + i] += lK[a][b][i];
      }
    }
  }
  return 0;
}

extern "C" int hostmath_fluid_thood(const HostThoodArgs* P, const int* rowPtr, const int* colPtr, double* R, double* Val)
{
  if (P->eNoN == 10 && P->eNoNq == 4) return run<10, 4>(P, rowPtr, colPtr, R, Val);
  if (P->eNoN == 27 && P->eNoNq == 8) return run<27, 8>(P, rowPtr, colPtr, R, Val);
  if (P->eNoN == 20 && P->eNoNq == 8) return run<20, 8>(P, rowPtr, colPtr, R, Val);
  return 3;
}
extern "C" int hostmath_sizeof_thoodargs() { return (int)sizeof(HostThoodArgs); }
