// fluid_gen_host.cpp — TEST-ONLY host build of fluid_gen.cuh: plain element loop over any 3-D element with
// nG == eNoN Gauss points (HEX8, TET4) using the device Gauss-point routines, checked against the reference.
#include <cmath>
#include <cstring>
using std::fabs; using std::sqrt; using std::pow; using std::exp; using std::cos;
#define SVB_HD inline
#include "../../svmultiphysics_b200/csrc/fluid_gen.cuh"

struct HostFluidGenArgs {
  const int* IEN; const double *x, *Ag, *Yg, *Bf;
  int eNoN, nEl, nG, tDof, mvMsh, factored;   // factored: tangent through fluid_gen_row / fluid_gen_block_row
  double dt, af, am, gam;
  double w[8], N[8][8], Nxi[8][8][3], Nxi2[8][8][6];
  svb::FluidDmn dm;
  const double* uris;        // URIS valves: (nNo, nUris, 5) = |sdf|, |scaffold udf|, valve velocity; or null
  int nUris;
  svb200_uris urisP[SVB200_MAX_URIS];
};

template <int ENON>
static int run(const HostFluidGenArgs* P, const int* rowPtr, const int* colPtr, double* R, double* Val)
{
  using namespace svb;
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
    // physical second derivatives of the LAST Gauss point (what fluid_3d_c sees at every Gauss point)
    double NxxL[ENON][6];
    {
      double Nx[ENON][3], xiX[3][3], ks[3][3];
      const double Jac = gnn3_full<ENON>(P->Nxi[P->nG - 1], xl, Nx, xiX, ks);
      if (is_zero(Jac)) return 4;
      gn_nxx3<ENON>(P->Nxi2[P->nG - 1], xl, xiX, Nx, NxxL);
    }
    double lR[ENON][4] = {}, lK[ENON][ENON][16] = {};
    for (int g = 0; g < P->nG; g++) {
      double Nx[ENON][3], Nxx[ENON][6], xiX[3][3], ks[3][3];
      const double Jac = gnn3_full<ENON>(P->Nxi[g], xl, Nx, xiX, ks);
      if (is_zero(Jac)) return 4;
      gn_nxx3<ENON>(P->Nxi2[g], xl, xiX, Nx, Nxx);
      FluidGP q;
      FluidNode nd[ENON];
      double uF = 0.0, uV[3] = {0.0, 0.0, 0.0};
      if (P->uris) uris_factor<ENON>(P->uris, P->nUris, P->urisP, P->N[g], n, uF, uV);
      fluid_gen_gauss_point<ENON>(P->dm, P->dt, P->af, P->am, P->gam, P->w[g] * Jac, ks, P->N[g], Nx, Nxx, NxxL, al, yl, bfl,
                                  P->mvMsh ? ym : nullptr, q, nd, uF, uV);
      for (int a = 0; a < ENON; a++) {
        fluid_gen_residual(q, nd[a], lR[a]);
        if (P->factored) {
          FluidRow row;
          fluid_gen_row(q, nd[a], row);
          for (int b = 0; b < ENON; b++) fluid_gen_block_row(row, nd[b], lK[a][b]);
        } else {
          for (int b = 0; b < ENON; b++) fluid_gen_block(q, nd[a], nd[b], lK[a][b]);
        }
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

extern "C" int hostmath_fluid_gen(const HostFluidGenArgs* P, const int* rowPtr, const int* colPtr, double* R, double* Val)
{
  if (P->nG != P->eNoN) return 3;
  if (P->eNoN == 4) return run<4>(P, rowPtr, colPtr, R, Val);
  if (P->eNoN == 8) return run<8>(P, rowPtr, colPtr, R, Val);
  return 3;
}
extern "C" int hostmath_sizeof_fluidgenargs() { return (int)sizeof(HostFluidGenArgs); }


// ---- any element of nn_elem_props.h (nG != eNoN allowed): tables passed by pointer ------------------------------------------
struct HostFluidAnyArgs {
  const int* IEN; const double *x, *Ag, *Yg, *Bf;
  const double *w, *N, *Nxi, *Nxi2;           // w[nG], N[nG][eNoN], Nxi[nG][eNoN][3], Nxi2[nG][eNoN][6]
  int eNoN, nEl, nG, tDof, mvMsh, lShpF;      // lShpF: derivatives from Gauss point 0 only (TET4, WDG: nn_elem_props.h)
  double dt, af, am, gam;
  svb::FluidDmn dm;
};

template <int ENON>
static int run_any(const HostFluidAnyArgs* P, const int* rowPtr, const int* colPtr, double* R, double* Val)
{
  using namespace svb;
  const int nG = P->nG;
  auto Nxi = [&](int g) { return reinterpret_cast<const double(*)[3]>(P->Nxi + (size_t)g * ENON * 3); };
  auto Nxi2 = [&](int g) { return reinterpret_cast<const double(*)[6]>(P->Nxi2 + (size_t)g * ENON * 6); };
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
    double NxxL[ENON][6];
    {
      double Nx[ENON][3], xiX[3][3], ks[3][3];
      const int gl = P->lShpF ? 0 : nG - 1;
      const double Jac = gnn3_full<ENON>(Nxi(gl), xl, Nx, xiX, ks);
      if (is_zero(Jac)) return 4;
      gn_nxx3<ENON>(Nxi2(gl), xl, xiX, Nx, NxxL);
    }
    static double lK[ENON][ENON][16];
    double lR[ENON][4] = {};
    std::memset(lK, 0, sizeof(lK));
    for (int g = 0; g < nG; g++) {
      double Nx[ENON][3], Nxx[ENON][6], xiX[3][3], ks[3][3];
      const int gd = P->lShpF ? 0 : g;
      const double Jac = gnn3_full<ENON>(Nxi(gd), xl, Nx, xiX, ks);
      if (is_zero(Jac)) return 4;
      gn_nxx3<ENON>(Nxi2(gd), xl, xiX, Nx, Nxx);
      FluidGP q;
      FluidNode nd[ENON];
      fluid_gen_gauss_point<ENON>(P->dm, P->dt, P->af, P->am, P->gam, P->w[g] * Jac, ks, P->N + (size_t)g * ENON, Nx, Nxx, NxxL, al, yl,
                                  bfl, P->mvMsh ? ym : nullptr, q, nd);
      for (int a = 0; a < ENON; a++) {
        fluid_gen_residual(q, nd[a], lR[a]);
        FluidRow row;
        fluid_gen_row(q, nd[a], row);
        for (int b = 0; b < ENON; b++) fluid_gen_block_row(row, nd[b], lK[a][b]);
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

extern "C" int hostmath_fluid_any(const HostFluidAnyArgs* P, const int* rowPtr, const int* colPtr, double* R, double* Val)
{
  switch (P->eNoN) {
    case 4: return run_any<4>(P, rowPtr, colPtr, R, Val);
    case 6: return run_any<6>(P, rowPtr, colPtr, R, Val);
    case 8: return run_any<8>(P, rowPtr, colPtr, R, Val);
    case 10: return run_any<10>(P, rowPtr, colPtr, R, Val);
    case 20: return run_any<20>(P, rowPtr, colPtr, R, Val);
    case 27: return run_any<27>(P, rowPtr, colPtr, R, Val);
  }
  return 3;
}
extern "C" int hostmath_sizeof_fluidanyargs() { return (int)sizeof(HostFluidAnyArgs); }
