// ref_harness_genalpha.cpp — flat C entry points around the reference's own generalised-alpha time integration.
//
// TEST INFRASTRUCTURE ONLY (compiled into oracle/_ref/libsvref.so by oracle/Makefile; never linked into the product).
// Drives, unmodified and where they lie under /root/reference:
//
//   Integrator::predictor              solver/Integrator.cpp:393-643
//   Integrator::initiator              solver/Integrator.cpp:662-750
//   Integrator::corrector              solver/Integrator.cpp:774-1070  (state update, FSI copy, convergence flags)
//   set_bc::set_bc_dir / set_bc_dir_l  solver/set_bc.cpp:901-1122 / 1125-1182
//
// through a hand-filled Simulation / ComMod (the way the reference's unit tests fill a MockComMod).  initiator() and
// corrector() are private members of Integrator; the harness reaches them with the usual test-only `#define private public`
// around the reference's header (the class layout is unchanged).  The device kernels of svmultiphysics_b200/csrc/genalpha.cu
// are compared with these bit for bit (tests/test_gpu_genalpha.py, fixtures tests/golden/genalpha.npz).
#include "ComMod.h"
#include "CepMod.h"
#include "Simulation.h"
#define private public
#include "Integrator.h"
#undef private
#include "set_bc.h"
#include "consts.h"
#include "utils.h"

#include "../include/svb200.h"

#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

namespace {
thread_local std::string g_err;

struct GaCase {
  Simulation sim;
  std::unique_ptr<Integrator> integ;
};

consts::EquationType ga_phys(int p)
{
  switch (p) {
    case SVB200_PHYS_FLUID: return consts::EquationType::phys_fluid;
    case SVB200_PHYS_STRUCT: return consts::EquationType::phys_struct;
    case SVB200_PHYS_FSI: return consts::EquationType::phys_FSI;
    case SVB200_PHYS_MESH: return consts::EquationType::phys_mesh;
    case SVB200_PHYS_LELAS: return consts::EquationType::phys_lElas;
    case SVB200_PHYS_HEATS: return consts::EquationType::phys_heatS;
    case SVB200_PHYS_HEATF: return consts::EquationType::phys_heatF;
    case SVB200_PHYS_USTRUCT: return consts::EquationType::phys_ustruct;
  }
  throw std::runtime_error("[ref_harness_genalpha] unknown physics");
}

template <class F> int guarded(F&& f)
{
  try { f(); return 0; }
  catch (const std::exception& ex) { g_err = ex.what(); return SVB200_ERR_NUMERIC; }
  catch (...) { g_err = "unknown exception"; return SVB200_ERR_NUMERIC; }
}

Array<double>* sol_array(Integrator& I, int which, int k)
{
  auto& s = I.get_solutions();
  auto& st = which == SVB200_SOL_OLD ? s.old : (which == SVB200_SOL_CURRENT ? s.current : s.intermediate);
  return k == 0 ? &st.get_acceleration() : (k == 1 ? &st.get_velocity() : &st.get_displacement());
}
}  // namespace

extern "C" {

const char* svref_ga_last_error(void) { return g_err.c_str(); }

/// A Simulation with nEq equations (rows s..e, gen-alpha coefficients of svb200_eqtime) on tnNo nodes and its Integrator.
/// old = (Ao, Yo, Do) as given; current starts as a copy of old (Integrator::initialize_arrays).
void* svref_ga_create(int tDof, int tnNo, int nEq, const svb200_eqtime* eqs, double dt, int dFlag, int sstEq,
                      const double* Ao, const double* Yo, const double* Do, int maxBc)
{
  GaCase* c = nullptr;
  int rc = guarded([&] {
    c = new GaCase();
    auto& cm = c->sim.com_mod;
    cm.nsd = 3; cm.nsymd = 6;
    cm.tDof = tDof; cm.tnNo = tnNo; cm.dt = dt; cm.time = 0.0;
    cm.dFlag = dFlag != 0; cm.sstEq = sstEq != 0;
    cm.nMsh = 0; cm.nFacesLS = 0;
    cm.nEq = nEq;
    cm.eq = std::vector<eqType>(nEq);
    for (int i = 0; i < nEq; i++) {
      auto& eq = cm.eq[i];
      eq.phys = ga_phys(eqs[i].phys);
      eq.s = eqs[i].s; eq.e = eqs[i].e; eq.dof = eqs[i].e - eqs[i].s + 1;
      eq.af = eqs[i].af; eq.am = eqs[i].am; eq.gam = eqs[i].gam; eq.beta = eqs[i].beta;
      eq.nDmn = 1; eq.dmn.resize(1); eq.dmn[0].phys = eq.phys; eq.dmn[0].Id = -1;
      eq.nBc = 0;
      eq.bc = std::vector<bcType>(maxBc);      // filled by svref_ga_add_dir_bc (growing the vector would need the VTK-reading
                                               // BoundaryCondition.cpp for bcType's copy constructor)
      eq.itr = 0; eq.minItr = 1; eq.maxItr = 5; eq.tol = 1e-12; eq.coupled = true; eq.ok = false;
    }
    if (sstEq) { cm.Ad.resize(3, tnNo); cm.Rd.resize(3, tnNo); }
    // predictor() stores the (here: zero) active tension of struct / ustruct / FSI equations per node (Integrator.cpp:578-611)
    auto& cem = c->sim.cep_mod.cem;
    cem.Ya_f.resize(tnNo); cem.Ya_s.resize(tnNo); cem.Ya_n.resize(tnNo);
    SolutionStates sol;
    const double* src[3] = {Ao, Yo, Do};
    Array<double>* dst[3] = {&sol.old.get_acceleration(), &sol.old.get_velocity(), &sol.old.get_displacement()};
    for (int k = 0; k < 3; k++) {
      dst[k]->resize(tDof, tnNo);
      if (src[k]) std::memcpy(dst[k]->data(), src[k], sizeof(double) * tDof * tnNo);
    }
    c->integ = std::make_unique<Integrator>(&c->sim, std::move(sol));
  });
  if (rc) { delete c; return nullptr; }
  return c;
}

void svref_ga_destroy(void* h) { delete static_cast<GaCase*>(h); }

int svref_ga_set(void* h, int which, const double* A, const double* Y, const double* D)
{
  auto& c = *static_cast<GaCase*>(h);
  return guarded([&] {
    const double* src[3] = {A, Y, D};
    for (int k = 0; k < 3; k++)
      if (src[k]) { auto* a = sol_array(*c.integ, which, k); std::memcpy(a->data(), src[k], sizeof(double) * a->size()); }
  });
}

int svref_ga_get(void* h, int which, double* A, double* Y, double* D)
{
  auto& c = *static_cast<GaCase*>(h);
  return guarded([&] {
    double* dst[3] = {A, Y, D};
    for (int k = 0; k < 3; k++)
      if (dst[k]) { auto* a = sol_array(*c.integ, which, k); std::memcpy(dst[k], a->data(), sizeof(double) * a->size()); }
  });
}

int svref_ga_set_ad(void* h, const double* Ad)
{
  auto& c = *static_cast<GaCase*>(h);
  return guarded([&] {
    auto& cm = c.sim.com_mod;
    cm.Ad.resize(3, cm.tnNo);
    std::memcpy(cm.Ad.data(), Ad, sizeof(double) * 3 * cm.tnNo);
  });
}

int svref_ga_get_ad(void* h, double* Ad)
{
  auto& c = *static_cast<GaCase*>(h);
  return guarded([&] { std::memcpy(Ad, c.sim.com_mod.Ad.data(), sizeof(double) * c.sim.com_mod.Ad.size()); });
}

/// Two-domain equation (fluid Id 0 + solid Id 1, like construct_fsi): node flags -> com_mod.dmnId for all_fun::is_domain.
int svref_ga_set_solid_nodes(void* h, int iEq, int solid_phys, const int* is_solid)
{
  auto& c = *static_cast<GaCase*>(h);
  return guarded([&] {
    auto& cm = c.sim.com_mod;
    auto& eq = cm.eq.at(iEq);
    eq.nDmn = 2; eq.dmn.resize(2);
    eq.dmn[0].phys = consts::EquationType::phys_fluid; eq.dmn[0].Id = 0;
    eq.dmn[1].phys = ga_phys(solid_phys); eq.dmn[1].Id = 1;
    cm.dmnId.resize(cm.tnNo);
    for (int a = 0; a < cm.tnNo; a++) cm.dmnId(a) = is_solid[a] ? 2 : 1;     // bit Id of the domain the node belongs to
  });
}

int svref_ga_predictor(void* h)
{
  auto& c = *static_cast<GaCase*>(h);
  return guarded([&] { c.integ->predictor(); });
}

int svref_ga_initiator(void* h, int cEq)
{
  auto& c = *static_cast<GaCase*>(h);
  return guarded([&] {
    c.sim.com_mod.cEq = cEq;
    c.integ->initiator(c.integ->get_solutions());
  });
}

/// Integrator::corrector for equation cEq with the linear solver's increment R(dof, tnNo) (and com_mod.Rd(3, tnNo) for ustruct).
int svref_ga_corrector(void* h, int cEq, const double* R, const double* Rd)
{
  auto& c = *static_cast<GaCase*>(h);
  return guarded([&] {
    auto& cm = c.sim.com_mod;
    cm.cEq = cEq;
    auto& eq = cm.eq.at(cEq);
    cm.dof = eq.dof;
    cm.R.resize(eq.dof, cm.tnNo);
    std::memcpy(cm.R.data(), R, sizeof(double) * eq.dof * cm.tnNo);
    if (Rd) { cm.Rd.resize(3, cm.tnNo); std::memcpy(cm.Rd.data(), Rd, sizeof(double) * 3 * cm.tnNo); }
    eq.FSILS.RI.iNorm = 1.0; eq.iNorm = 1.0;       // the convergence bookkeeping at the end of corrector() is not compared
    c.integ->corrector();
  });
}

/// One strongly imposed Dirichlet BC of equation iEq on a face (node list gN, optional normals nV(3,n)): steady value g with
/// spatial profile gx(n), direction flags eDrn(3), impD = impose on (Y, D) instead of (A, Y).  Then set_bc::set_bc_dir.
int svref_ga_add_dir_bc(void* h, int iEq, int n, const int* gN, const int* eDrn, int impD, double g, const double* gx, const double* nV)
{
  auto& c = *static_cast<GaCase*>(h);
  return guarded([&] {
    auto& cm = c.sim.com_mod;
    if (cm.nMsh == 0) { cm.msh.emplace_back(); cm.nMsh = 1; }
    auto& m = cm.msh[0];
    m.fa.emplace_back();
    m.nFa = (int)m.fa.size();
    auto& fa = m.fa.back();
    fa.nNo = n; fa.iM = 0; fa.name = "dir" + std::to_string(m.nFa - 1);
    fa.gN.resize(n);
    for (int a = 0; a < n; a++) fa.gN(a) = gN[a];
    fa.nV.resize(3, n);
    if (nV) std::memcpy(fa.nV.data(), nV, sizeof(double) * 3 * n);
    auto& eq = cm.eq.at(iEq);
    if (eq.nBc >= (int)eq.bc.size()) throw std::runtime_error("[ref_harness_genalpha] more BCs than svref_ga_create reserved");
    auto& bc = eq.bc[eq.nBc++];
    bc.iM = 0; bc.iFa = m.nFa - 1;
    bc.bType = utils::ibset(0, consts::iBC_Dir);
    bc.bType = utils::ibset(bc.bType, consts::iBC_std);
    if (impD) bc.bType = utils::ibset(bc.bType, consts::iBC_impD);
    bc.weakDir = false;
    bc.g = g;
    bc.eDrn.resize(3);
    for (int i = 0; i < 3; i++) bc.eDrn(i) = eDrn[i];
    bc.gx.resize(n);
    for (int a = 0; a < n; a++) bc.gx(a) = gx ? gx[a] : 1.0;
  });
}

int svref_ga_set_bc_dir(void* h)
{
  auto& c = *static_cast<GaCase*>(h);
  return guarded([&] { set_bc::set_bc_dir(c.sim.com_mod, c.integ->get_solutions()); });
}

}  // extern "C"
