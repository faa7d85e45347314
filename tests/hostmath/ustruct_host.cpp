// ustruct_host.cpp — TEST-ONLY host build of ustruct_elem.cuh: plain element loop over the device Gauss-point routines,
// checked against the R / Val / Kd the reference assembled (tests/golden/ustruct.npz).
#include <cmath>
#include <cstring>
using std::fabs; using std::sqrt; using std::pow; using std::exp;
#define SVB_HD inline
#include "../../svmultiphysics_b200/csrc/ustruct_elem.cuh"

struct HostUstructArgs {
  const int* IEN; const double* fN; const double *x, *Ag, *Yg, *Dg, *Bf;
  int eNoN, nEl, nG, tDof, s, nFn;
  double dt, af, am, gam;
  double w[8], N[8][8], Nxi[8][8][3];
  svb::UstructDmn dm;
  const double* Ya;              // (nNo, 3) active tensions or null
  svb::CannRow cann[16];
};

template <int ENON>
static int run(const HostUstructArgs* P, const int* rowPtr, const int* colPtr, double* R, double* Val, double* Kd)
{
  using namespace svb;
  const int s0 = P->s, tD = P->tDof;
  const double af = P->af * P->gam * P->dt, am = P->am;
  for (int e = 0; e < P->nEl; e++) {
    int n[ENON];
    double xl[ENON][3], ql[ENON][3], vl[ENON][3], dl[ENON][3], pl[ENON], pdl[ENON], fN[2][3] = {{0, 0, 0}, {0, 0, 0}};
    for (int a = 0; a < ENON; a++) {
      n[a] = P->IEN[ENON * e + a];
      for (int i = 0; i < 3; i++) {
        xl[a][i] = P->x[3 * n[a] + i];
        ql[a][i] = P->Ag[(size_t)tD * n[a] + s0 + i] - P->Bf[3 * n[a] + i];
        vl[a][i] = P->Yg[(size_t)tD * n[a] + s0 + i];
        dl[a][i] = P->Dg[(size_t)tD * n[a] + s0 + i];
      }
      pl[a] = P->Yg[(size_t)tD * n[a] + s0 + 3];
      pdl[a] = P->Ag[(size_t)tD * n[a] + s0 + 3];
    }
    for (int k = 0; k < P->nFn && k < 2; k++) for (int i = 0; i < 3; i++) fN[k][i] = P->fN[(size_t)3 * P->nFn * e + 3 * k + i];
    double lR[ENON][4] = {}, lK[ENON][ENON][16] = {}, lKd[ENON][ENON][12] = {};
    for (int g = 0; g < P->nG; g++) {
      UGP q;
      ViscGP gu, gv;
      const bool visc = P->dm.st.viscType != 0 && P->dm.st.visc_mu != 0.0;
      double ya[3] = {0, 0, 0};
      const bool act = P->Ya && P->dm.st.active;
      if (act) for (int a = 0; a < ENON; a++) for (int i = 0; i < 3; i++) ya[i] += P->N[g][a] * P->Ya[3 * n[a] + i];
      if (ustruct_gauss_point<ENON>(P->dm, P->dt, P->af, P->am, P->gam, P->w[g], P->N[g], P->Nxi[g], xl, ql, vl, dl, pl, pdl, fN, q,
                                    visc ? &gu : nullptr, visc ? &gv : nullptr, act ? ya : nullptr, P->cann, P->nFn)) return 2;
      UNode nd[ENON];
      double Bm[ENON][6][3], DBm[ENON][6][3];
      for (int a = 0; a < ENON; a++) {
        ustruct_node(q, P->N[g][a], P->Nxi[g][a], nd[a]);
        make_Bm(nd[a].Nx, q.F, Bm[a]);
        make_DBm(q.Dm, Bm[a], DBm[a]);
        ustruct_resid(q, nd[a], lR[a]);
      }
      for (int a = 0; a < ENON; a++)
        for (int b = 0; b < ENON; b++) {
          ustruct_block(q, af, am, nd[a], nd[b], Bm[a], DBm[b], lK[a][b], lKd[a][b]);
          if (visc) ustruct_visc_block(P->dm.st.viscType, q, af, am, gu, gv, nd[a].Nx, nd[b].Nx, lK[a][b], lKd[a][b]);
        }
    }
    for (int a = 0; a < ENON; a++) {
      for (int i = 0; i < 4; i++) R[4 * n[a] + i] += lR[a][i];
      for (int b = 0; b < ENON; b++) {
        int sl = -1;
        for (int k = rowPtr[n[a]]; k < rowPtr[n[a] + 1]; k++) if (colPtr[k] == n[b]) { sl = k; break; }
        if (sl < 0) return 1;
        for (int i = 0; i < 16; i++) Val[(size_t)16 * sl + i] += lK[a][b][i];
        for (int i = 0; i < 12; i++) Kd[(size_t)12 * sl + i] += lKd[a][b][i];
      }
    }
  }
  return 0;
}

// The closed-form TET4 path (ustruct_tet4_setup / _resid / _block) as a plain element loop.
static int run_tet4(const HostUstructArgs* P, const int* rowPtr, const int* colPtr, double* R, double* Val, double* Kd)
{
  using namespace svb;
  const int s0 = P->s, tD = P->tDof;
  const double af = P->af * P->gam * P->dt, am = P->am;
  for (int e = 0; e < P->nEl; e++) {
    int n[4];
    double xl[4][3], ql[4][3], vl[4][3], dl[4][3], pl[4], pdl[4], fN[2][3] = {{0, 0, 0}, {0, 0, 0}};
    for (int a = 0; a < 4; a++) {
      n[a] = P->IEN[4 * e + a];
      for (int i = 0; i < 3; i++) {
        xl[a][i] = P->x[3 * n[a] + i];
        ql[a][i] = P->Ag[(size_t)tD * n[a] + s0 + i] - P->Bf[3 * n[a] + i];
        vl[a][i] = P->Yg[(size_t)tD * n[a] + s0 + i];
        dl[a][i] = P->Dg[(size_t)tD * n[a] + s0 + i];
      }
      pl[a] = P->Yg[(size_t)tD * n[a] + s0 + 3];
      pdl[a] = P->Ag[(size_t)tD * n[a] + s0 + 3];
    }
    for (int k = 0; k < P->nFn && k < 2; k++) for (int i = 0; i < 3; i++) fN[k][i] = P->fN[(size_t)3 * P->nFn * e + 3 * k + i];
    UTet4Const C; UTet4Mom M; double Dm[6][6], Je;
    double ya[3] = {0, 0, 0};
    const bool act = P->Ya && P->dm.st.active;
    if (act) {
      double wsum = 0.0;
      for (int g = 0; g < 4; g++) {
        wsum += P->w[g];
        for (int a = 0; a < 4; a++) for (int i = 0; i < 3; i++) ya[i] += P->w[g] * P->N[g][a] * P->Ya[3 * n[a] + i];
      }
      for (int i = 0; i < 3; i++) ya[i] /= wsum;
    }
    if (ustruct_tet4_setup(P->dm, af, am, P->w, &P->N[0][0], 8, P->Nxi[0], xl, ql, vl, dl, pl, pdl, fN, C, M, Dm, &Je,
                           act ? ya : nullptr, P->cann, P->nFn)) return 2;
    for (int a = 0; a < 4; a++) {
      double r[4];
      ustruct_tet4_resid(C, M, &P->N[0][0], 8, a, r);
      for (int i = 0; i < 4; i++) R[4 * n[a] + i] += r[i];
    }
    for (int b = 0; b < 4; b++) {
      double Bmb[6][3], DBmb[6][3];
      make_Bm(C.Nx[b], C.F, Bmb);
      make_DBm(Dm, Bmb, DBmb);
      for (int a = 0; a < 4; a++) {
        double Bma[6][3], K[16], Kdd[12];
        make_Bm(C.Nx[a], C.F, Bma);
        ustruct_tet4_block(C, M, &P->N[0][0], 8, af, am, a, b, Bma, DBmb, K, Kdd);
        int sl = -1;
        for (int k = rowPtr[n[a]]; k < rowPtr[n[a] + 1]; k++) if (colPtr[k] == n[b]) { sl = k; break; }
        if (sl < 0) return 1;
        for (int i = 0; i < 16; i++) Val[(size_t)16 * sl + i] += K[i];
        for (int i = 0; i < 12; i++) Kd[(size_t)12 * sl + i] += Kdd[i];
      }
    }
  }
  return 0;
}

extern "C" int hostmath_ustruct_tet4(const HostUstructArgs* P, const int* rowPtr, const int* colPtr, double* R, double* Val, double* Kd)
{
  if (P->eNoN != 4 || P->nG != 4) return 3;
  return run_tet4(P, rowPtr, colPtr, R, Val, Kd);
}

extern "C" int hostmath_ustruct(const HostUstructArgs* P, const int* rowPtr, const int* colPtr, double* R, double* Val, double* Kd)
{
  if (P->eNoN == 4) return run<4>(P, rowPtr, colPtr, R, Val, Kd);
  if (P->eNoN == 8) return run<8>(P, rowPtr, colPtr, R, Val, Kd);
  return 3;
}
extern "C" int hostmath_sizeof_ustructargs() { return (int)sizeof(HostUstructArgs); }
