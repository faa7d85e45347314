// flop_count.cpp — counting-type instantiation of the TET4 VMS fluid restatement (TEST INFRASTRUCTURE ONLY).
//
// SURVEY.md 8(d) fixes the numerator of the assembly roofline as the ALGORITHMIC flop count per TET4 element
// (hand count: one nn::gnn + 4 Gauss points x (fluid_3d_m + fluid_3d_c) = 11.6 kflop, add/sub/mul/div/sqrt = 1 each,
// Newtonian viscosity, K_darcy = 0) and asks for an instantiation of the oracle over a counting scalar to pin it.
// This file includes oracle/fluid_gp.inc — the same text sv_oracle.c compiles with REAL = double and that reproduces
// the compiled reference bit for bit — with REAL = Counted, runs ONE element the way construct_fluid does, and prints
//   executed   : what the restatement executes (gnn in both Gauss loops' first point, the shared front part of
//                fluid_3d_m / fluid_3d_c evaluated twice per Gauss point, multiplications by structural zeros included)
//   per routine: gnn, fluid_3d_m, fluid_3d_c separately, so that "one gnn + 4 x (m + c)" can be formed.
// Build + run: make -C oracle flops
#include <cmath>
#include <cstdio>
#include <cstring>
#include "../include/svb200.h"

static long long g_flops = 0;      // every operation executed
static long long g_nz = 0;         // operations whose operands are not exact zeros ("structurally-zero terms removed":
                                   // x*0, 0/x, x+0, x-0 are what mu_g = 0, K_darcy = 0, the off-diagonal updu ... leave behind)
struct Counted {
  double v;
  Counted() : v(0.0) {}
  Counted(double x) : v(x) {}
  Counted(int x) : v(x) {}
};
static inline void tally(double a, double b, char op)
{
  g_flops++;
  const bool trivial = (op == '/') ? (a == 0.0) : (a == 0.0 || b == 0.0);
  if (!trivial) g_nz++;
}
#define BINOP(op, ch)                                                                                  \
  static inline Counted operator op(const Counted& a, const Counted& b) { tally(a.v, b.v, ch); return Counted(a.v op b.v); } \
  static inline Counted operator op(const Counted& a, double b) { tally(a.v, b, ch); return Counted(a.v op b); }           \
  static inline Counted operator op(double a, const Counted& b) { tally(a, b.v, ch); return Counted(a op b.v); }
BINOP(+, '+') BINOP(-, '-') BINOP(*, '*') BINOP(/, '/')
#undef BINOP
static inline Counted& operator+=(Counted& a, const Counted& b) { tally(a.v, b.v, '+'); a.v += b.v; return a; }
static inline Counted operator-(const Counted& a) { return Counted(-a.v); }     // sign flip: not counted
static inline bool operator<(const Counted& a, const Counted& b) { return a.v < b.v; }
static inline bool operator<(const Counted& a, double b) { return a.v < b; }
static inline Counted sqrt(const Counted& a) { g_flops++; g_nz++; return Counted(std::sqrt(a.v)); }
static inline Counted pow(const Counted& a, const Counted& b) { g_flops++; g_nz += (a.v != 0.0); return Counted(std::pow(a.v, b.v)); }
static inline Counted pow(const Counted& a, double b) { g_flops++; g_nz += (a.v != 0.0); return Counted(std::pow(a.v, b)); }
static inline Counted pow(double a, const Counted& b) { g_flops++; g_nz += (a != 0.0); return Counted(std::pow(a, b.v)); }
static inline Counted fabs(const Counted& a) { return Counted(std::fabs(a.v)); }
static inline Counted fmax(const Counted& a, const Counted& b) { return Counted(a.v > b.v ? a.v : b.v); }
static inline Counted fmax(const Counted& a, double b) { return Counted(a.v > b ? a.v : b); }

#define MAXE 8
#define REAL Counted
#include "fluid_gp.inc"
#undef REAL

int main()
{
  // one TET4 element, reference tables of nn_elem_gip.h:214-226 / nn.cpp:174, generic state
  const int eNoN = 4, nG = 4, tDof = 4;
  const double s = (5.0 + 3.0 * std::sqrt(5.0)) / 20.0, t = (5.0 - std::sqrt(5.0)) / 20.0;
  double xi[4][3];
  for (int g = 0; g < 4; g++) for (int k = 0; k < 3; k++) xi[g][k] = t;
  for (int g = 0; g < 3; g++) xi[g][g] = s;
  Counted N[4][4], Nxi[12], w[4];
  const double nx[3][4] = {{1, 0, 0, -1}, {0, 1, 0, -1}, {0, 0, 1, -1}};
  for (int a = 0; a < 4; a++) for (int k = 0; k < 3; k++) Nxi[k + 3 * a] = nx[k][a];
  for (int g = 0; g < 4; g++) {
    w[g] = 1.0 / 24.0;
    for (int a = 0; a < 3; a++) N[g][a] = xi[g][a];
    N[g][3] = 1.0 - xi[g][0] - xi[g][1] - xi[g][2];
  }
  const double xs[4][3] = {{1.0, 0.1, 0.0}, {0.2, 1.1, 0.1}, {0.0, 0.3, 0.9}, {0.1, 0.0, 0.05}};
  Counted xl[12], al[16], yl[16], bfl[12];
  for (int a = 0; a < 4; a++) {
    for (int i = 0; i < 3; i++) { xl[i + 3 * a] = xs[a][i]; bfl[i + 3 * a] = 0.01 * (a + i); }
    for (int i = 0; i < 4; i++) { al[i + tDof * a] = 0.1 * (a - i) + 0.05; yl[i + tDof * a] = 1.0 + 0.3 * a - 0.2 * i; }
  }
  svb200_eqparams eq;
  std::memset(&eq, 0, sizeof eq);
  eq.dt = 1e-3; eq.af = 2.0 / 3.0; eq.am = 5.0 / 6.0; eq.gam = 2.0 / 3.0; eq.tDof = 4; eq.dof = 4;
  svb200_dmnparams dm;
  std::memset(&dm, 0, sizeof dm);
  dm.rho = 1.06; dm.mu_i = 0.04; dm.viscType = SVB200_VISC_CONST; dm.K_darcy = 0.0;

  Counted lR[16], lK[256], Nwx[12], ks[9], Jac;
  struct Tally { long long all = 0, nz = 0; } gnn, m, c, common, wj;
  auto reset = [] { g_flops = 0; g_nz = 0; };
  auto add = [](Tally& t) { t.all += g_flops; t.nz += g_nz; };
  reset(); gnn3(eNoN, Nxi, xl, Nwx, &Jac, ks); add(gnn);
  for (int g = 0; g < nG; g++) {
    reset(); const Counted wg = w[g] * Jac; add(wj);
    // lR / lK start from zero in the first Gauss point only: measure in the LAST one, where every accumulator is non-zero
    reset(); fluid_3d_m(&eq, &dm, eNoN, wg, ks, N[g], Nwx, al, yl, bfl, tDof, lR, lK); if (g == nG - 1) add(m);
    reset(); fluid_3d_c(&eq, &dm, eNoN, wg, ks, N[g], Nwx, al, yl, bfl, tDof, lR, lK); if (g == nG - 1) add(c);
    if (g == nG - 1) { GP q; reset(); gauss_point_common(&eq, &dm, eNoN, wg, ks, N[g], Nwx, al, yl, bfl, tDof, &q); add(common); }
  }
  const long long executed = 2 * gnn.all + nG * (m.all + c.all) + wj.all;
  const long long algorithmic = gnn.nz + nG * (m.nz + c.nz) + wj.nz;
  const long long hoisted = gnn.nz + nG * (m.nz + c.nz - common.nz) + wj.nz;
  std::printf("{\"element\": \"TET4\", \"viscosity\": \"Newtonian\", \"K_darcy\": 0, \"nG\": %d,\n"
              " \"executed_by_restatement\": {\"gnn\": %lld, \"fluid_3d_m_per_gauss_point\": %lld, \"fluid_3d_c_per_gauss_point\": %lld, "
              "\"front_part_shared_by_m_and_c\": %lld, \"per_element_2gnn_plus_nG_times_m_plus_c\": %lld},\n"
              " \"structural_zeros_removed\": {\"gnn\": %lld, \"fluid_3d_m_per_gauss_point\": %lld, \"fluid_3d_c_per_gauss_point\": %lld, "
              "\"front_part_shared_by_m_and_c\": %lld, \"per_element_gnn_plus_nG_times_m_plus_c\": %lld, "
              "\"per_element_front_part_evaluated_once\": %lld},\n"
              " \"survey_hand_count\": {\"gnn\": 227, \"fluid_3d_m\": 2133, \"fluid_3d_c\": 721, \"per_element\": 11600}}\n",
              nG, gnn.all, m.all, c.all, common.all, executed, gnn.nz, m.nz, c.nz, common.nz, algorithmic, hoisted);
  return 0;
}
