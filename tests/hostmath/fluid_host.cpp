// fluid_host.cpp — TEST-ONLY host build of the device element routines (fluid_elem.cuh).
// Lets the CPU-only test suite check the hoisted tet4 element algebra against the reference oracle
// without a GPU.  Never linked into libsvb200.so; the product has no CPU path.
#include <cmath>
#include <cstring>
#include <vector>
using std::fabs; using std::sqrt; using std::pow;
#define SVB_HD inline
#include "../../svmultiphysics_b200/csrc/fluid_elem.cuh"

extern "C" int hostmath_fluid_tet4(const svb::FluidArgs* P, int nNo, const int* rowPtr, const int* colPtr,
                                   double* R, double* Val)
{
  using namespace svb;
  for (int e = P->e0; e < P->e1; e++) {
    int n[4];
    double xl[4][3], yl[4][4], uc[4][3], ab[4][3];
    for (int a = 0; a < 4; a++) {
      n[a] = P->IEN[4*e + a];
      for (int i = 0; i < 3; i++) {
        xl[a][i] = P->x[3*n[a] + i] + (P->ale ? P->Dg[P->tDof*n[a] + 4 + i] : 0.0);
        ab[a][i] = P->Ag[P->tDof*n[a] + i] - P->Bf[3*n[a] + i];
        uc[a][i] = P->Yg[P->tDof*n[a] + i] - (P->mvMsh ? P->Yg[P->tDof*n[a] + 4 + i] : 0.0);
      }
      for (int i = 0; i < 4; i++) yl[a][i] = P->Yg[P->tDof*n[a] + i];
    }
    int iD = 0;
    for (int d = 0; d < P->nDmn; d++) { iD = d; if (P->dmn[d].Id == -1) break; if (P->eId && ((P->eId[e] >> P->dmn[d].Id) & 1)) break; }
    if (!P->dmn[iD].isFluid) continue;
    Tet4Elem E;
    tet4_element(*P, P->dmn[iD], xl, yl, uc, ab, E);
    for (int a = 0; a < 4; a++) {
      for (int i = 0; i < 4; i++) R[4*n[a] + i] += E.lR[a][i];
      for (int b = 0; b < 4; b++) {
        double K[16];
        tet4_block(E, a, b, K);
        int s = -1;
        for (int k = rowPtr[n[a]]; k < rowPtr[n[a]+1]; k++) if (colPtr[k] == n[b]) { s = k; break; }
        if (s < 0) return 1;
        for (int i = 0; i < 16; i++) Val[16*(size_t)s + i] += K[i];
      }
    }
  }
  return 0;
}
// Same loop through the staged element routine + record-based block emitter used by the grouped scatter.
extern "C" int hostmath_fluid_tet4_staged(const svb::FluidArgs* P, int nNo, const int* rowPtr, const int* colPtr,
                                          double* R, double* Val)
{
  using namespace svb;
  bool NN = false;
  for (int d = 0; d < P->nDmn; d++) NN |= (P->dmn[d].viscType != SVB200_VISC_CONST);
  for (int e = P->e0; e < P->e1; e++) {
    int n[4];
    double xl[4][3], yl[4][4], uc[4][3], ab[4][3];
    for (int a = 0; a < 4; a++) {
      n[a] = P->IEN[4*e + a];
      for (int i = 0; i < 3; i++) {
        xl[a][i] = P->x[3*n[a] + i] + (P->ale ? P->Dg[P->tDof*n[a] + 4 + i] : 0.0);
        ab[a][i] = P->Ag[P->tDof*n[a] + i] - P->Bf[3*n[a] + i];
        uc[a][i] = P->Yg[P->tDof*n[a] + i] - (P->mvMsh ? P->Yg[P->tDof*n[a] + 4 + i] : 0.0);
      }
      for (int i = 0; i < 4; i++) yl[a][i] = P->Yg[P->tDof*n[a] + i];
    }
    int iD = 0;
    for (int d = 0; d < P->nDmn; d++) { iD = d; if (P->dmn[d].Id == -1) break; if (P->eId && ((P->eId[e] >> P->dmn[d].Id) & 1)) break; }
    if (!P->dmn[iD].isFluid) continue;
    double rec[REC_NN], lR[16];
    for (int i = 0; i < REC_NN; i++) rec[i] = 1e300;   // poison: every field read must have been written
    tet4_element_staged(*P, P->dmn[iD], xl, yl, uc, ab, NN, rec, lR);
    auto find = [&](int row, int col) { for (int k = rowPtr[row]; k < rowPtr[row+1]; k++) if (colPtr[k] == col) return k; return -1; };
    for (int a = 0; a < 4; a++) {
      for (int i = 0; i < 4; i++) R[4*n[a] + i] += lR[4*a + i];
      for (int b = a; b < 4; b++) {
        double K1[16], K2[16];
        for (int i = 0; i < 16; i++) K1[i] = K2[i] = 0.0;
        if (a == b) tet4_block_rec_add(rec, NN, a, a, K1);
        else tet4_edge_rec_add(rec, NN, a, b, K1, K2);   // both blocks of the edge, as the grouped kernel does
        const int s1 = find(n[a], n[b]), s2 = find(n[b], n[a]);
        if (s1 < 0 || s2 < 0) return 1;
        for (int i = 0; i < 16; i++) Val[16*(size_t)s1 + i] += K1[i];
        if (a != b) for (int i = 0; i < 16; i++) Val[16*(size_t)s2 + i] += K2[i];
      }
    }
  }
  return 0;
}
// Same element records, but the tangent goes through the accumulator form of the grouped kernel's phase 3 (EdgeAcc / DiagAcc):
// one accumulator per DISTINCT CSR diagonal block / mesh edge over all the elements that touch it, assembled into 4x4 blocks once.
#include <map>
extern "C" int hostmath_fluid_tet4_acc(const svb::FluidArgs* P, int nNo, const int* rowPtr, const int* colPtr,
                                       double* R, double* Val)
{
  using namespace svb;
  bool NN = false;
  for (int d = 0; d < P->nDmn; d++) NN |= (P->dmn[d].viscType != SVB200_VISC_CONST);
  auto find = [&](int row, int col) { for (int k = rowPtr[row]; k < rowPtr[row+1]; k++) if (colPtr[k] == col) return k; return -1; };
  std::map<int, EdgeAcc> edges;     // keyed by the slot of (lo, hi), lo = the node with the smaller id
  std::map<int, int> partner;
  std::map<int, DiagAcc> diags;
  for (int e = P->e0; e < P->e1; e++) {
    int n[4];
    double xl[4][3], yl[4][4], uc[4][3], ab[4][3];
    for (int a = 0; a < 4; a++) {
      n[a] = P->IEN[4*e + a];
      for (int i = 0; i < 3; i++) {
        xl[a][i] = P->x[3*n[a] + i] + (P->ale ? P->Dg[P->tDof*n[a] + 4 + i] : 0.0);
        ab[a][i] = P->Ag[P->tDof*n[a] + i] - P->Bf[3*n[a] + i];
        uc[a][i] = P->Yg[P->tDof*n[a] + i] - (P->mvMsh ? P->Yg[P->tDof*n[a] + 4 + i] : 0.0);
      }
      for (int i = 0; i < 4; i++) yl[a][i] = P->Yg[P->tDof*n[a] + i];
    }
    int iD = 0;
    for (int d = 0; d < P->nDmn; d++) { iD = d; if (P->dmn[d].Id == -1) break; if (P->eId && ((P->eId[e] >> P->dmn[d].Id) & 1)) break; }
    if (!P->dmn[iD].isFluid) continue;
    double rec[REC_NN], lR[16];
    for (int i = 0; i < REC_NN; i++) rec[i] = 1e300;
    tet4_element_staged(*P, P->dmn[iD], xl, yl, uc, ab, NN, rec, lR);
    for (int a = 0; a < 4; a++) {
      for (int i = 0; i < 4; i++) R[4*n[a] + i] += lR[4*a + i];
      const int sd = find(n[a], n[a]);
      if (sd < 0) return 1;
      if (!diags.count(sd)) diag_acc_zero(diags[sd], NN);
      tet4_diag_rec_acc(rec, NN, a, diags[sd]);
      for (int b = a + 1; b < 4; b++) {
        const int lo = n[a] < n[b] ? a : b, hi = n[a] < n[b] ? b : a;      // same orientation for every element sharing the edge
        const int s1 = find(n[lo], n[hi]), s2 = find(n[hi], n[lo]);
        if (s1 < 0 || s2 < 0) return 1;
        if (!edges.count(s1)) { edge_acc_zero(edges[s1], NN); partner[s1] = s2; }
        tet4_edge_rec_acc(rec, NN, lo, hi, edges[s1]);
      }
    }
  }
  double K[16];
  for (auto& kv : diags) { diag_acc_block(kv.second, NN, K); for (int i = 0; i < 16; i++) Val[16*(size_t)kv.first + i] += K[i]; }
  for (auto& kv : edges) {
    edge_acc_block(kv.second, NN, 0, K); for (int i = 0; i < 16; i++) Val[16*(size_t)kv.first + i] += K[i];
    edge_acc_block(kv.second, NN, 1, K); for (int i = 0; i < 16; i++) Val[16*(size_t)partner[kv.first] + i] += K[i];
  }
  return 0;
}
extern "C" int hostmath_sizeof_fluidargs() { return (int)sizeof(svb::FluidArgs); }
