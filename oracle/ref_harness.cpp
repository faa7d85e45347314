// ref_harness.cpp — flat C entry points around the UNMODIFIED reference hot path.
//
// TEST INFRASTRUCTURE ONLY.  This file is compiled (by oracle/Makefile, target `ref`) together with
// the reference's own sources where they lie under /root/reference into oracle/_ref/libsvref.so.
// It fills a reference `ComMod` by hand for a synthetic mesh (the way the reference's unit tests
// do with MockComMod, tests/unitTests/test_common.h:78-96) and calls the reference routines:
//
//   nn::select_ele / fs::init_fs_msh            solver/nn.cpp:1302, solver/fs.cpp:271
//   lhsa_ns::add_col (the loop of lhsa_ns::lhsa) solver/lhsa.cpp:13-54,155-166,352-380
//   fsils_commu_create / fsils_lhs_create       linear_solver/commu.cpp:17, linear_solver/lhs.cpp:30
//   fsils_bc_create                             linear_solver/bc.cpp:18
//   fluid::construct_fluid                      solver/fluid.cpp:480
//   struct_ns::construct_dsolid                 solver/sv_struct.cpp:184
//   fsi::construct_fsi, mesh::construct_mesh    solver/fsi.cpp:24, solver/mesh.cpp:22
//   fsi_linear_solver::fsils_solve              linear_solver/solve.cpp:23
//   spar_mul::fsils_spar_mul_vv                 linear_solver/spar_mul.cpp:164
//   nn::get_gip / get_gnn (faces), nn::gnnb     solver/nn.cpp:455,500,911
//   fluid::b_fluid, l_elas::b_l_elas            solver/fluid.cpp:21, solver/l_elas.cpp:21
//   heats::construct_heats, heatf::construct_heatf   solver/heats.cpp:30, solver/heatf.cpp:52
//   ustruct::construct_usolid, ustruct::ustruct_r     solver/ustruct.cpp:203, 1742
//
// Parameter structs are shared with the product ABI (include/svb200.h) so that parity tests feed
// both sides the same bytes.  Nothing here is linked into libsvb200.so.

#include "ComMod.h"
#include "CepMod.h"
#include "SolutionStates.h"
#include "FsilsLinearAlgebra.h"
#include "fluid.h"
#include "sv_struct.h"
#include "fsi.h"
#include "mesh.h"
#include "fs.h"
#include "nn.h"
#include "lhsa.h"
#include "l_elas.h"
#include "heats.h"
#include "heatf.h"
#include "ustruct.h"
#include "all_fun.h"
#include "utils.h"
#include "fsils_api.hpp"
#include "commu.h"
#include "lhs.h"
#include "spar_mul.h"
#include "consts.h"

#include "../include/svb200.h"

#include <chrono>
#include <cstring>
#include <string>
#include <stdexcept>

namespace {

thread_local std::string g_err;

struct RefCase {
  ComMod com_mod;
  CepMod cep_mod;
  SolutionStates sol;
  FsilsLinearAlgebra* la = nullptr;
  bool graph_built = false;
  double last_assemble_s = 0.0;
  double last_solve_s = 0.0;
  // Optional plug-in backend (svref_set_backend): a LinearAlgebra implementation other than FsilsLinearAlgebra plus
  // the early-out hook a maintainer adds at the top of eq_assem::global_eq_assem (solver/eq_assem.cpp:397).
  LinearAlgebra* backend = nullptr;
  int (*assem_hook)(void*, void*, const void*, const void*) = nullptr;
  void (*backend_download)(void*, int, double*) = nullptr;
  void (*backend_thood_val_rc)(void*) = nullptr;            // B200LinearAlgebra::thood_val_rc(), INTEGRATION.md
  void (*backend_ustruct_r)(void*, void*) = nullptr;       // B200LinearAlgebra::ustruct_r(ComMod&), INTEGRATION.md
  // Multi-rank runs (oracle/ref_build/mpi_stub.cpp with SVREF_MPI_SIZE > 1): this rank's local -> global node map
  int gnNo = -1;
  std::vector<int> ltg;
};

// A do-nothing active-stress model: the hot path only tests the pointer (the model itself runs once per time step,
// solver/active_stress.cpp:50-66, outside the path).
struct HarnessActiveStress : public ActiveStress {
  HarnessActiveStress() : ActiveStress(0) {}
  std::unique_ptr<ActiveStressModelParameters> get_parameters() const override { return nullptr; }
  void read_model_specific_parameters(const ActiveStressModelParameters&) override {}
  void distribute_model_specific_parameters(const CmMod&, const cmType&) override {}
  void init_local(Vector<double>&) const override {}
  void advance_time_step_local(const double, const double, const double, const double, const double, Vector<double>&) const override {}
  double compute_active_tension_local(const Vector<double>&) const override { return 0.0; }
};

consts::EquationType to_phys(int p)
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
  throw std::runtime_error("[ref_harness] unknown physics");
}

void fill_domain(dmnType& d, const svb200_dmnparams& p)
{
  using namespace consts;
  d.Id = p.Id;
  d.phys = to_phys(p.phys);
  d.prop[PhysicalProperyType::fluid_density] = p.rho;
  d.prop[PhysicalProperyType::solid_density] = p.rho;
  d.prop[PhysicalProperyType::f_x] = p.f[0];
  d.prop[PhysicalProperyType::f_y] = p.f[1];
  d.prop[PhysicalProperyType::f_z] = p.f[2];
  d.prop[PhysicalProperyType::inverse_darcy_permeability] = p.K_darcy;
  d.prop[PhysicalProperyType::backflow_stab] = p.backflow_stab;
  d.prop[PhysicalProperyType::damping] = p.dmp;
  d.prop[PhysicalProperyType::elasticity_modulus] = p.E;
  d.prop[PhysicalProperyType::poisson_ratio] = p.nu;
  d.prop[PhysicalProperyType::conductivity] = p.conductivity;
  d.prop[PhysicalProperyType::source_term] = p.source_term;
  d.prop[PhysicalProperyType::ctau_M] = p.ctau_M;
  d.prop[PhysicalProperyType::ctau_C] = p.ctau_C;
  switch (p.viscType) {
    case SVB200_VISC_CONST: d.fluid_visc.viscType = FluidViscosityModelType::viscType_Const; break;
    case SVB200_VISC_CY: d.fluid_visc.viscType = FluidViscosityModelType::viscType_CY; break;
    case SVB200_VISC_CASSON: d.fluid_visc.viscType = FluidViscosityModelType::viscType_Cass; break;
  }
  d.fluid_visc.mu_i = p.mu_i; d.fluid_visc.mu_o = p.mu_o; d.fluid_visc.lam = p.lam;
  d.fluid_visc.a = p.a; d.fluid_visc.n = p.n;
  switch (p.isoType) {
    case SVB200_ISO_NHK: d.stM.isoType = ConstitutiveModelType::stIso_nHook; break;
    case SVB200_ISO_MR: d.stM.isoType = ConstitutiveModelType::stIso_MR; break;
    case SVB200_ISO_GUCCIONE: d.stM.isoType = ConstitutiveModelType::stIso_Gucci; break;
    case SVB200_ISO_STVK: d.stM.isoType = ConstitutiveModelType::stIso_StVK; break;
    case SVB200_ISO_HGO: d.stM.isoType = ConstitutiveModelType::stIso_HGO; break;
    case SVB200_ISO_HO: d.stM.isoType = ConstitutiveModelType::stIso_HO; break;
    case SVB200_ISO_HO_MA: d.stM.isoType = ConstitutiveModelType::stIso_HO_ma; break;
    case SVB200_ISO_CANN: {
      // what set_material_props.h:155-175 does with the <Add_row> entries of a Constitutive_model type="CANN"
      d.stM.isoType = ConstitutiveModelType::stArtificialNeuralNet;
      auto& t = d.stM.paramTable;
      t.num_rows = p.cann_rows;
      t.invariant_indices.resize(t.num_rows);
      t.activation_functions.resize(t.num_rows, 3);
      t.weights.resize(t.num_rows, 3);
      for (int r = 0; r < t.num_rows; r++) {
        t.invariant_indices(r) = p.cann_inv[r];
        for (int k = 0; k < 3; k++) { t.activation_functions(r, k) = p.cann_act[r][k]; t.weights(r, k) = p.cann_w[r][k]; }
      }
    } break;
  }
  // dmn.active_stress != nullptr is all struct_3d / ustruct_3d look at (sv_struct.cpp:277, ustruct.cpp:294); the nodal
  // tensions themselves are cep_mod.cem.Ya_f / Ya_s / Ya_n (svref_set_active_tension)
  if (p.active_stress) d.active_stress = std::make_shared<HarnessActiveStress>();
  else d.active_stress.reset();
  switch (p.volType) {
    case SVB200_VOL_NONE: d.stM.volType = ConstitutiveModelType::stVol_NA; break;
    case SVB200_VOL_QUAD: d.stM.volType = ConstitutiveModelType::stVol_Quad; break;
    case SVB200_VOL_ST91: d.stM.volType = ConstitutiveModelType::stVol_ST91; break;
    case SVB200_VOL_M94: d.stM.volType = ConstitutiveModelType::stVol_M94; break;
  }
  d.stM.Kpen = p.Kpen; d.stM.C10 = p.C10; d.stM.C01 = p.C01;
  d.stM.bff = p.bff; d.stM.bss = p.bss; d.stM.bfs = p.bfs;
  d.stM.a = p.st_a; d.stM.b = p.st_b; d.stM.aff = p.aff; d.stM.ass = p.ass; d.stM.afs = p.afs; d.stM.kap = p.kap; d.stM.khs = p.khs;
  if (p.solid_visc_mu != 0.0) {
    d.solid_visc.viscType = (p.solidViscType == SVB200_SOLID_VISC_POTENTIAL) ? SolidViscosityModelType::viscType_Potential
                                                                              : SolidViscosityModelType::viscType_Newtonian;
    d.solid_visc.mu = p.solid_visc_mu;
  } else {
    d.solid_visc.viscType = SolidViscosityModelType::viscType_NA;
  }
}

void fill_eq(RefCase& c, const svb200_eqparams& e, const svb200_dmnparams* dmn, int nDmn)
{
  auto& cm = c.com_mod;
  if (cm.eq.size() == 0) { cm.eq = std::vector<eqType>(1); cm.nEq = 1; }
  auto& eq = cm.eq[0];
  cm.cEq = 0;
  cm.dt = e.dt;
  cm.dof = e.dof;
  cm.tDof = e.tDof;
  cm.mvMsh = (e.mvMsh != 0);
  // prestress equation (com_mod.pstEq): accumulators as initialize.cpp:674-679 sizes them
  cm.pstEq = (e.reserved & SVB200_EQ_PRESTRESS) != 0;
  cm.nsymd = 6;
  if (cm.pstEq && cm.pSa.size() != cm.tnNo) {
    cm.pSn.resize(6, cm.tnNo); cm.pSa.resize(cm.tnNo);
    cm.pSn = 0.0; cm.pSa = 0.0;
  }
  eq.phys = to_phys(e.phys);
  eq.af = e.af; eq.am = e.am; eq.gam = e.gam; eq.beta = e.beta;
  eq.dof = e.dof; eq.s = e.s; eq.e = e.s + e.dof - 1;
  eq.nDmn = nDmn;
  eq.dmn.resize(nDmn);
  for (int i = 0; i < nDmn; i++) fill_domain(eq.dmn[i], dmn[i]);
  if (!c.la) c.la = new FsilsLinearAlgebra();
  eq.linear_algebra = c.backend ? c.backend : c.la;
  eq.linear_algebra_preconditioner = consts::PreconditionerType::PREC_FSILS;
}

template <class F> int guarded(F&& f)
{
  try { f(); return 0; }
  catch (const std::exception& ex) { g_err = ex.what(); return SVB200_ERR_NUMERIC; }
  catch (...) { g_err = "unknown exception"; return SVB200_ERR_NUMERIC; }
}

} // namespace

extern "C" {

const char* svref_last_error(void) { return g_err.c_str(); }

void* svref_create(void)
{
  auto c = new RefCase();
  c->com_mod.nsd = 3;
  c->com_mod.nsymd = 6;
  return c;
}

void svref_destroy(void* h) { delete static_cast<RefCase*>(h); }

/// Run the same harness through another LinearAlgebra plug-in (the product's C++ host layer B200LinearAlgebra):
/// `la` is a LinearAlgebra*, `hook` the early-out of global_eq_assem, `download` fetches R / Val from the backend.
int svref_set_backend(void* h, void* la, void* hook, void* download)
{
  auto& c = *static_cast<RefCase*>(h);
  c.backend = static_cast<LinearAlgebra*>(la);
  c.assem_hook = reinterpret_cast<int (*)(void*, void*, const void*, const void*)>(hook);
  c.backend_download = reinterpret_cast<void (*)(void*, int, double*)>(download);
  return 0;
}

int svref_set_backend_thood_val_rc(void* h, void* fn)
{
  static_cast<RefCase*>(h)->backend_thood_val_rc = reinterpret_cast<void (*)(void*)>(fn);
  return 0;
}

int svref_set_backend_ustruct_r(void* h, void* fn)
{
  static_cast<RefCase*>(h)->backend_ustruct_r = reinterpret_cast<void (*)(void*, void*)>(fn);
  return 0;
}

/// com_mod.x(3,nNo); also sizes Bf to zero.
int svref_set_coords(void* h, int nNo, const double* x)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& cm = c.com_mod;
    cm.tnNo = nNo;
    cm.x.resize(3, nNo);
    std::memcpy(cm.x.data(), x, sizeof(double)*3*nNo);
    cm.Bf.resize(3, nNo);
  });
}

/// Append one mesh; the reference fills its own Gauss/shape tables (select_ele, init_fs_msh).
int svref_add_mesh(void* h, int eNoN, int nEl, const int* IEN, const int* eId, int nFn, const double* fN)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& cm = c.com_mod;
    cm.msh.emplace_back();
    cm.nMsh = (int)cm.msh.size();
    auto& m = cm.msh.back();
    m.eNoN = eNoN; m.nEl = nEl; m.gnEl = nEl; m.nNo = cm.tnNo; m.gnNo = cm.tnNo;
    m.lShl = false; m.lFib = false;
    m.nFs = 1;     // P1-P1 / Q1-Q1 with VMS stabilisation (vmsStab = true, solver/fluid.cpp:496)
    m.IEN.resize(eNoN, nEl);
    std::memcpy(m.IEN.data(), IEN, sizeof(int)*eNoN*nEl);
    if (eId) { m.eId.resize(nEl); std::memcpy(m.eId.data(), eId, sizeof(int)*nEl); }
    m.nFn = nFn;
    if (nFn > 0 && fN) { m.fN.resize(3*nFn, nEl); std::memcpy(m.fN.data(), fN, sizeof(double)*3*nFn*nEl); }
    nn::select_ele(cm, m);
    fs::init_fs_msh(cm, m);
  });
}

/// Switch mesh iM to Taylor-Hood function spaces (mshType::nFs = 2, what read_msh does for <Use_taylor_hood_type_basis>): velocity on
/// the mesh's own (quadratic) element, pressure on its linear parent (fs::set_thood_fs, fs.cpp:336-390).
int svref_set_mesh_thood(void* h, int iM)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& m = c.com_mod.msh.at(iM);
    m.nFs = 2;
    fs::init_fs_msh(c.com_mod, m);
  });
}

/// The four function-space tables construct_fluid uses for a Taylor-Hood mesh (fs::get_thood_fs, fs.cpp:73-178):
///   loop 1 (momentum, the velocity space's Gauss rule, nG1 points): pressure shape functions Nq1(eNoNq, nG1), Nqxi1(3, eNoNq, nG1)
///   loop 2 (continuity, the pressure space's rule, nG2 points): w2(nG2), Nw2(eNoN, nG2), Nwxi2(3, eNoN, nG2), Nq2(eNoNq, nG2), Nqxi2(3, eNoNq, nG2)
/// dims = {eNoNq, nG1, nG2, lShpF of the velocity space, lShpF of the pressure space}; null pointers are skipped.
int svref_get_thood_tables(void* h, int iM, int* dims, double* Nq1, double* Nqxi1, double* w2, double* Nw2, double* Nwxi2, double* Nq2,
                           double* Nqxi2)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& m = c.com_mod.msh.at(iM);
    if (m.nFs != 2) throw std::runtime_error("[ref_harness] mesh has no Taylor-Hood function spaces");
    std::array<fsType, 2> f1, f2;
    fs::get_thood_fs(c.com_mod, f1, m, false, 1);
    fs::get_thood_fs(c.com_mod, f2, m, false, 2);
    dims[0] = f1[1].eNoN; dims[1] = f1[0].nG; dims[2] = f2[1].nG; dims[3] = f1[0].lShpF ? 1 : 0; dims[4] = f1[1].lShpF ? 1 : 0;
    auto cp = [](double* dst, const double* src, size_t n) { if (dst) std::memcpy(dst, src, sizeof(double) * n); };
    cp(Nq1, f1[1].N.data(), (size_t)f1[1].eNoN * f1[1].nG);
    cp(Nqxi1, f1[1].Nx.data(), (size_t)3 * f1[1].eNoN * f1[1].nG);
    cp(w2, f2[1].w.data(), (size_t)f2[1].nG);
    cp(Nw2, f2[0].N.data(), (size_t)f2[0].eNoN * f2[0].nG);
    cp(Nwxi2, f2[0].Nx.data(), (size_t)3 * f2[0].eNoN * f2[0].nG);
    cp(Nq2, f2[1].N.data(), (size_t)f2[1].eNoN * f2[1].nG);
    cp(Nqxi2, f2[1].Nx.data(), (size_t)3 * f2[1].eNoN * f2[1].nG);
  });
}

/// fs::thood_val_rc (fs.cpp:394-466, called by Integrator::step after the assembly): pressure rows of the edge nodes.
int svref_thood_val_rc(void* h)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    if (c.backend) {        // the patch of Integrator::step: R / Val live on the device
      if (!c.backend_thood_val_rc) throw std::runtime_error("[ref_harness] backend without thood_val_rc hook");
      c.backend_thood_val_rc(c.backend);
      return;
    }
    fs::thood_val_rc(c.com_mod);
  });
}

int svref_get_mesh_tables(void* h, int iM, int* nG, double* w, double* N, double* Nx)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& m = c.com_mod.msh.at(iM);
    *nG = m.nG;
    if (w) std::memcpy(w, m.w.data(), sizeof(double)*m.nG);
    if (N) std::memcpy(N, m.N.data(), sizeof(double)*m.eNoN*m.nG);
    if (Nx) std::memcpy(Nx, m.Nx.data(), sizeof(double)*3*m.eNoN*m.nG);
  });
}

/// Second derivatives of the shape functions on the reference element, fs[0].Nxx(6,eNoN,nG) (fs::init_fs_msh):
/// what construct_fluid hands to nn::gn_nxx (solver/fluid.cpp:648-650).
int svref_get_mesh_nxx(void* h, int iM, double* Nxx)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& m = c.com_mod.msh.at(iM);
    auto& f = m.fs.at(0);
    if (f.Nxx.size() != 6*m.eNoN*m.nG) throw std::runtime_error("[ref_harness] fs[0].Nxx has an unexpected size");
    std::memcpy(Nxx, f.Nxx.data(), sizeof(double)*6*m.eNoN*m.nG);
  });
}

/// The element loop of lhsa_ns::lhsa (solver/lhsa.cpp:155-166) through the reference's own add_col,
/// then its compaction (:352-380), then fsils_commu_create + fsils_lhs_create as initialize() does.
int svref_build_graph(void* h, int nFaces, int* nnz_out)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& cm = c.com_mod;
    int tnNo = cm.tnNo;
    int max_enon = 0;
    for (auto& m : cm.msh) max_enon = std::max(max_enon, m.eNoN);
    int mnnzeic = 10*max_enon;
    Array<int> uInd(mnnzeic, tnNo);
    uInd = -1;
    for (int iM = 0; iM < (int)cm.msh.size(); iM++) {
      auto& m = cm.msh[iM];
      for (int e = 0; e < m.nEl; e++) {
        for (int a = 0; a < m.eNoN; a++) {
          int rowN = m.IEN(a,e);
          for (int b = 0; b < m.eNoN; b++) {
            lhsa_ns::add_col(tnNo, rowN, m.IEN(b,e), mnnzeic, uInd);
          }
          // extra connections of the twin node across a RIS surface (lhsa.cpp:168-193)
          if (cm.risFlag) {
            for (int iProj = 0; iProj < cm.ris.nbrRIS; iProj++) {
              int jMRIS;
              if (cm.ris.lst(0,0,iProj) == iM) jMRIS = 1;
              else if (cm.ris.lst(1,0,iProj) == iM) jMRIS = 0;
              else continue;
              std::array<int, 2> mapIdx;
              utils::find_loc(cm.grisMapList[iProj].map, rowN, mapIdx);
              if (mapIdx[0] == -1) continue;
              const int rowNR = cm.grisMapList[iProj].map(jMRIS, mapIdx[1]);
              if (rowNR == -1) continue;
              for (int b = 0; b < m.eNoN; b++) lhsa_ns::add_col(tnNo, rowNR, m.IEN(b,e), mnnzeic, uInd);
            }
          }
        }
      }
    }
    int nnz = 0;
    for (int r = 0; r < tnNo; r++) {
      if (uInd(0,r) == -1) throw std::runtime_error("isolated node " + std::to_string(r));
      for (int i = 0; i < mnnzeic; i++) if (uInd(i,r) != -1) nnz++;
    }
    cm.colPtr.resize(nnz);
    cm.rowPtr.resize(tnNo+1);
    int j = 0;
    cm.rowPtr(0) = 0;
    for (int r = 0; r < tnNo; r++) {
      for (int i = 0; i < mnnzeic; i++) if (uInd(i,r) != -1) cm.colPtr(j++) = uInd(i,r);
      cm.rowPtr(r+1) = j;
    }
    cm.idMap.resize(tnNo);
    for (int a = 0; a < tnNo; a++) cm.idMap[a] = a;
    cm.ltg.resize(tnNo);
    for (int a = 0; a < tnNo; a++) cm.ltg[a] = c.ltg.empty() ? a : c.ltg.at(a);
    const int gnNo = c.ltg.empty() ? tnNo : c.gnNo;
    cm.cm.new_cm(MPI_COMM_WORLD);                 // np() > 1 makes all_fun::commu exchange (all_fun.cpp:96)

    fsi_linear_solver::FSILS_commuType communicator;
    fsi_linear_solver::fsils_commu_create(communicator, MPI_COMM_WORLD);
    fsi_linear_solver::fsils_lhs_create(cm.lhs, communicator, gnNo, tnNo, nnz, cm.ltg, cm.rowPtr, cm.colPtr, nFaces);
    c.graph_built = true;
    *nnz_out = nnz;
  });
}

/// Multi-rank: the local -> global node map of this rank's partition (com_mod.ltg) and the global node count; call before
/// svref_build_graph.  fsils_lhs_create then derives lhs.map, mynNo and the shared-node lists cS[] itself (lhs.cpp:30-348).
int svref_set_partition(void* h, int gnNo, int nNo, const int* ltg)
{
  auto& c = *static_cast<RefCase*>(h);
  c.gnNo = gnNo;
  c.ltg.assign(ltg, ltg + nNo);
  return 0;
}

/// What fsils_lhs_create built: mynNo, nReq and map(nNo) (host node -> FSILS position).
int svref_get_lhs(void* h, int* mynNo, int* nReq, int* map)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& lhs = c.com_mod.lhs;
    *mynNo = lhs.mynNo; *nReq = lhs.nReq;
    if (map) for (int a = 0; a < lhs.nNo; a++) map[a] = lhs.map(a);
  });
}

/// Shared-node list i: neighbour rank, length and (if ptr != NULL) the FSILS positions, lhs.cS[i].{iP,n,ptr}.
int svref_get_lhs_req(void* h, int i, int* iP, int* n, int* ptr)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& cs = c.com_mod.lhs.cS.at(i);
    *iP = cs.iP; *n = cs.n;
    if (ptr) for (int k = 0; k < cs.n; k++) ptr[k] = cs.ptr(k);
  });
}

/// all_fun::commu(com_mod, com_mod.R) of Integrator::step (Integrator.cpp:124-129).
int svref_commu_R(void* h)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] { all_fun::commu(c.com_mod, c.com_mod.R); });
}

/// MPI_Barrier over the ranks of a multi-rank run (no-op for one rank): bench.py brackets its timed steps with it.
int svref_barrier(void*)
{
  return guarded([&] { MPI_Barrier(MPI_COMM_WORLD); });
}

int svref_get_graph(void* h, int* rowPtr, int* colPtr)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& cm = c.com_mod;
    std::memcpy(rowPtr, cm.rowPtr.data(), sizeof(int)*(cm.tnNo+1));
    std::memcpy(colPtr, cm.colPtr.data(), sizeof(int)*cm.colPtr.size());
  });
}

int svref_set_face(void* h, int faIn, int bGrp, int face_dof, int nNo, const int* glob, const double* val)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    Vector<int> gNodes(nNo);
    for (int a = 0; a < nNo; a++) gNodes(a) = glob[a];
    Array<double> sVl(face_dof, nNo);
    if (val) std::memcpy(sVl.data(), val, sizeof(double)*face_dof*nNo);
    auto t = (bGrp == SVB200_BC_DIR) ? fsi_linear_solver::BcType::BC_TYPE_Dir : fsi_linear_solver::BcType::BC_TYPE_Neu;
    fsi_linear_solver::fsils_bc_create(c.com_mod.lhs, faIn, nNo, face_dof, t, gNodes, sVl);
  });
}

/// Capping surface of a coupled face: what CoupledBoundaryCondition leaves in lhs.face[faIn].{has_cap, cap_glob, cap_val, cap_valM}
/// (linear_solver/fils_struct.hpp:131-143); glob are host node ids (negative = not on this rank), mapped like bc.cpp:55-58.
int svref_set_face_cap(void* h, int faIn, int n, const int* glob, const double* val)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& lhs = c.com_mod.lhs;
    auto& fa = lhs.face.at(faIn);
    fa.has_cap = true;
    fa.cap_glob.resize(n);
    fa.cap_val.resize(fa.dof, n);
    fa.cap_valM.resize(fa.dof, n);
    for (int a = 0; a < n; a++) {
      fa.cap_glob(a) = glob[a] < 0 ? -1 : lhs.map(glob[a]);
      for (int i = 0; i < fa.dof; i++) { fa.cap_val(i,a) = val[(size_t)a*fa.dof + i]; fa.cap_valM(i,a) = 0.0; }
    }
  });
}

/// ls_alloc (solver/ls.cpp:24-40): fresh zero R(dof,tnNo), Val(dof*dof,nnz).
int svref_alloc(void* h, int dof)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& cm = c.com_mod;
    cm.dof = dof;
    cm.R.resize(dof, cm.tnNo);
    if (c.backend) {
      // ls_alloc (solver/ls.cpp:24-40) through the plug-in interface
      if (cm.eq.size() == 0) { cm.eq = std::vector<eqType>(1); cm.nEq = 1; }
      cm.cEq = 0;
      cm.eq[0].linear_algebra = c.backend;
      c.backend->alloc(cm, cm.eq[0]);
      return;
    }
    cm.Val.resize(dof*dof, cm.lhs.nnz);
    cm.R = 0.0;
    cm.Val = 0.0;
    if (cm.Kd.size() != 0) cm.Kd = 0.0;          // Integrator.cpp:106-109
    // Integrator::initiator zeroes the prestress accumulators once per Newton iteration (Integrator.cpp:745-748)
    if (cm.pSa.size() != 0) { cm.pSn = 0.0; cm.pSa = 0.0; }
  });
}

int svref_set_state(void* h, int tDof, const double* Ag, const double* Yg, const double* Dg, const double* Bf)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& cm = c.com_mod;
    int n = cm.tnNo;
    cm.tDof = tDof;
    auto& A = c.sol.intermediate.get_acceleration();
    auto& Y = c.sol.intermediate.get_velocity();
    auto& D = c.sol.intermediate.get_displacement();
    if (A.nrows() != tDof || A.ncols() != n) { A.resize(tDof, n); Y.resize(tDof, n); D.resize(tDof, n); }
    if (Ag) std::memcpy(A.data(), Ag, sizeof(double)*tDof*n);
    if (Yg) std::memcpy(Y.data(), Yg, sizeof(double)*tDof*n);
    if (Dg) std::memcpy(D.data(), Dg, sizeof(double)*tDof*n);
    if (Bf) std::memcpy(cm.Bf.data(), Bf, sizeof(double)*3*n);
    // Old displacement (used by the mesh equation, solver/mesh.cpp) mirrors Dg in the harness.
    auto& Do = c.sol.old.get_displacement();
    if (Do.nrows() != tDof || Do.ncols() != n) Do.resize(tDof, n);
    if (Dg) std::memcpy(Do.data(), Dg, sizeof(double)*tDof*n);
  });
}

/// com_mod.pS0(nsymd, tnNo); null clears it (pS0.size() == 0 switches the prestress terms off, sv_struct.cpp:273).
int svref_set_prestress(void* h, const double* pS0)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& cm = c.com_mod;
    cm.nsymd = 6;
    if (!pS0) { cm.pS0.resize(0, 0); return; }
    cm.pS0.resize(6, cm.tnNo);
    std::memcpy(cm.pS0.data(), pS0, sizeof(double) * 6 * cm.tnNo);
  });
}

int svref_get_prestress(void* h, double* pSn, double* pSa)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& cm = c.com_mod;
    if (cm.pSa.size() != cm.tnNo) throw std::runtime_error("[ref_harness] no prestress equation assembled");
    std::memcpy(pSn, cm.pSn.data(), sizeof(double) * 6 * cm.tnNo);
    std::memcpy(pSa, cm.pSa.data(), sizeof(double) * cm.tnNo);
  });
}

/// cep_mod.cem.Ya_f / Ya_s / Ya_n (solver/CepMod.h:205-217), nNo values each; Ya_s / Ya_n may be null (zero).
int svref_set_active_tension(void* h, const double* Ya_f, const double* Ya_s, const double* Ya_n)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    const int n = c.com_mod.tnNo;
    auto& cem = c.cep_mod.cem;
    cem.Ya_f.resize(n); cem.Ya_s.resize(n); cem.Ya_n.resize(n);
    for (int a = 0; a < n; a++) {
      cem.Ya_f[a] = Ya_f[a];
      cem.Ya_s[a] = Ya_s ? Ya_s[a] : 0.0;
      cem.Ya_n[a] = Ya_n ? Ya_n[a] : 0.0;
    }
  });
}

/// Fitted RIS: com_mod.risFlag, ris.nbrRIS / clsFlg / lst and grisMapList[] as ris::ris_read_msh leaves them for lhsa (extra CSR
/// connections, lhsa.cpp:168-193 — call BEFORE svref_build_graph) and for ris::doassem_ris (ris.cpp:269-349; the reference's own
/// ris.cpp is compiled into libsvref.so).  maps: for projection p, 2 * nMap[p] node ids, map(0,0), map(1,0), map(0,1), ...;
/// meshes(2, nProj): the mesh index of face 0 and of face 1 of each projection (RIS.lst(0,0,p), RIS.lst(1,0,p)).
int svref_set_ris(void* h, int nProj, const int* nMap, const int* maps, const int* closed, const int* meshes)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& cm = c.com_mod;
    cm.risFlag = nProj > 0;
    cm.ris.nbrRIS = nProj;
    cm.ris.clsFlg.assign(nProj, false);
    cm.ris.lst.resize(2, 2, std::max(nProj, 1));
    cm.grisMapList.clear();
    cm.grisMapList.resize(nProj);
    size_t off = 0;
    for (int p = 0; p < nProj; p++) {
      cm.ris.clsFlg[p] = closed[p] != 0;
      cm.ris.lst(0, 0, p) = meshes[2*p]; cm.ris.lst(1, 0, p) = meshes[2*p + 1];
      cm.ris.lst(0, 1, p) = 0; cm.ris.lst(1, 1, p) = 0;
      cm.grisMapList[p].map.resize(2, nMap[p]);
      for (int j = 0; j < nMap[p]; j++) {
        cm.grisMapList[p].map(0, j) = maps[off + 2*j];
        cm.grisMapList[p].map(1, j) = maps[off + 2*j + 1];
      }
      off += 2 * (size_t)nMap[p];
    }
  });
}

/// URIS valves: com_mod.urisFlag / urisActFlag / nUris / uris[] as uris::uris_read_msh and the time loop leave them for the
/// element routines (only the members uris::eval_uris_ris_factors_quadrature reads).  scal(3, nUris) = resistance, sdf_deps,
/// sdf_deps_close; flags(6, nUris) = clsFlg, cnt, DxOpen.nslices, DxClose.nslices, scaffold_flag, include_uris_velocity;
/// sdf / udf: (nNo, nUris), vel: (3, nNo, nUris).  nUris = 0 switches the valves off.
int svref_set_uris(void* h, int nUris, const double* scal, const int* flags, const double* sdf, const double* udf, const double* vel)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& cm = c.com_mod;
    const int n = cm.tnNo;
    cm.nUris = nUris;
    cm.urisFlag = cm.urisActFlag = (nUris > 0);
    cm.uris.clear();
    cm.uris.resize(nUris);
    for (int v = 0; v < nUris; v++) {
      auto& u = cm.uris[v];
      u.resistance = scal[3*v]; u.sdf_deps = scal[3*v + 1]; u.sdf_deps_close = scal[3*v + 2];
      u.clsFlg = flags[6*v] != 0; u.cnt = flags[6*v + 1];
      u.DxOpen.resize(1, 1, flags[6*v + 2]); u.DxClose.resize(1, 1, flags[6*v + 3]);
      u.scaffold_flag = flags[6*v + 4] != 0; u.include_uris_velocity = flags[6*v + 5] != 0;
      u.sdf.resize(n);
      for (int a = 0; a < n; a++) u.sdf(a) = sdf[(size_t)v*n + a];
      if (u.scaffold_flag) { u.scaffold_udf.resize(n); for (int a = 0; a < n; a++) u.scaffold_udf(a) = udf[(size_t)v*n + a]; }
      if (u.include_uris_velocity) {
        u.valve_velocity_fluid.resize(3, n);
        for (int a = 0; a < n; a++) for (int i = 0; i < 3; i++) u.valve_velocity_fluid(i, a) = vel[((size_t)v*n + a)*3 + i];
      }
    }
  });
}

int svref_set_old_disp(void* h, int tDof, const double* Do_in)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& Do = c.sol.old.get_displacement();
    int n = c.com_mod.tnNo;
    if (Do.nrows() != tDof || Do.ncols() != n) Do.resize(tDof, n);
    std::memcpy(Do.data(), Do_in, sizeof(double)*tDof*n);
  });
}

/// The switch of eq_assem::global_eq_assem (solver/eq_assem.cpp:397-449) for the in-scope physics.
int svref_assemble(void* h, int iM, const svb200_eqparams* e, const svb200_dmnparams* dmn, int nDmn)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    fill_eq(c, *e, dmn, nDmn);
    auto& cm = c.com_mod;
    auto& m = cm.msh.at(iM);
    auto t0 = std::chrono::steady_clock::now();
    if (c.assem_hook && c.assem_hook(&cm, &c.cep_mod, &m, &c.sol)) {
      c.last_assemble_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      return;
    }
    switch (e->phys) {
      case SVB200_PHYS_FLUID: fluid::construct_fluid(cm, m, c.sol); break;
      case SVB200_PHYS_STRUCT: struct_ns::construct_dsolid(cm, c.cep_mod, m, c.sol); break;
      case SVB200_PHYS_FSI: {
        // FSI with velocity-pressure solids (com_mod.sstEq, fsi.cpp:243-262): Kd as initialize.cpp:668-670 sizes it; it is zeroed
        // once per Newton iteration (Integrator.cpp:106-109), here by svref_alloc, so that the meshes of one equation add up
        bool ust = false;
        for (int d = 0; d < nDmn; d++) ust |= (dmn[d].phys == SVB200_PHYS_USTRUCT);
        if (ust) {
          if (cm.Kd.nrows() != 12 || cm.Kd.ncols() != cm.lhs.nnz) { cm.Kd.resize(12, cm.lhs.nnz); cm.Kd = 0.0; }
          if (cm.idMap.size() != cm.tnNo) { cm.idMap.resize(cm.tnNo); for (int a = 0; a < cm.tnNo; a++) cm.idMap(a) = a; }
        }
        fsi::construct_fsi(cm, c.cep_mod, m, c.sol);
      } break;
      case SVB200_PHYS_MESH: mesh::construct_mesh(cm, c.cep_mod, m, c.sol); break;
      case SVB200_PHYS_LELAS: l_elas::construct_l_elas(cm, m, c.sol); break;
      case SVB200_PHYS_HEATS: heats::construct_heats(cm, m, c.sol); break;
      case SVB200_PHYS_HEATF: heatf::construct_heatf(cm, m, c.sol); break;
      case SVB200_PHYS_USTRUCT: {
        // what Integrator::step does around the assembly of a ustruct equation (solver/Integrator.cpp:106-109,
        // initialize.cpp:668-670): Kd((nsd+1)*nsd, nnz) zeroed; idMap is the identity without Taylor-Hood elements
        if (cm.Kd.nrows() != 12 || cm.Kd.ncols() != cm.lhs.nnz) cm.Kd.resize(12, cm.lhs.nnz);
        cm.Kd = 0.0;
        if (cm.idMap.size() != cm.tnNo) { cm.idMap.resize(cm.tnNo); for (int a = 0; a < cm.tnNo; a++) cm.idMap(a) = a; }
        ustruct::construct_usolid(cm, c.cep_mod, m, c.sol);
      } break;
      default: throw std::runtime_error("[ref_harness] physics not supported");
    }
    c.last_assemble_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  });
}

int svref_get(void* h, int what, double* dst)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& cm = c.com_mod;
    if (c.backend && c.backend_download) { c.backend_download(c.backend, what, dst); return; }
    if (what == SVB200_ARRAY_R) std::memcpy(dst, cm.R.data(), sizeof(double)*cm.R.size());
    else if (what == SVB200_ARRAY_VAL) std::memcpy(dst, cm.Val.data(), sizeof(double)*cm.Val.size());
    else if (what == SVB200_ARRAY_KD) std::memcpy(dst, cm.Kd.data(), sizeof(double)*cm.Kd.size());
    else throw std::runtime_error("[ref_harness] bad array id");
  });
}

int svref_put(void* h, int what, int dof, const double* src)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& cm = c.com_mod;
    cm.dof = dof;
    if (what == SVB200_ARRAY_R) {
      cm.R.resize(dof, cm.tnNo);
      std::memcpy(cm.R.data(), src, sizeof(double)*cm.R.size());
    } else if (what == SVB200_ARRAY_VAL) {
      cm.Val.resize(dof*dof, cm.lhs.nnz);
      std::memcpy(cm.Val.data(), src, sizeof(double)*cm.Val.size());
    } else throw std::runtime_error("[ref_harness] bad array id");
  });
}

/// fsils_ls_create defaults are bypassed: every parameter comes from `ls` (as read_ls does after XML).
int svref_solve(void* h, int dof, int ls_type, int prec, const svb200_lsparams* ls, int nFaces, const int* incL,
    const double* res, double* R_out, svb200_lsresult* out)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    using namespace fsi_linear_solver;
    auto& cm = c.com_mod;
    FSILS_lsType fls;
    LinearSolverType t;
    switch (ls_type) {
      case SVB200_LS_NS: t = LinearSolverType::LS_TYPE_NS; break;
      case SVB200_LS_GMRES: t = LinearSolverType::LS_TYPE_GMRES; break;
      case SVB200_LS_CG: t = LinearSolverType::LS_TYPE_CG; break;
      case SVB200_LS_BICGS: t = LinearSolverType::LS_TYPE_BICGS; break;
      default: throw std::runtime_error("bad ls_type");
    }
    fsils_ls_create(fls, t);
    auto cp = [](FSILS_subLsType& d, const svb200_sublsparams& s) {
      d.mItr = s.mItr; d.sD = s.sD; d.relTol = s.relTol; d.absTol = s.absTol;
    };
    cp(fls.RI, ls->RI); cp(fls.GM, ls->GM); cp(fls.CG, ls->CG);
    Vector<int> incLv(nFaces);
    Vector<double> resv(nFaces);
    for (int i = 0; i < nFaces; i++) { incLv(i) = incL ? incL[i] : 1; resv(i) = res ? res[i] : 0.0; }
    const auto ptype = (prec == SVB200_PREC_RCS) ? consts::PreconditionerType::PREC_RCS : consts::PreconditionerType::PREC_FSILS;
    auto t0 = std::chrono::steady_clock::now();
    if (c.backend) {
      // ls_solve (solver/ls.cpp:42-54): lEq.linear_algebra->solve(com_mod, lEq, incL, res)
      auto& eq = cm.eq.at(0);
      eq.FSILS = fls;
      eq.linear_algebra = c.backend;
      eq.linear_algebra_preconditioner = ptype;
      eq.linear_algebra->set_preconditioner(ptype);
      cm.dof = dof;
      eq.linear_algebra->solve(cm, eq, incLv, resv);
      fls = eq.FSILS;
    } else {
      fsils_solve(cm.lhs, fls, dof, cm.R, cm.Val, ptype, incLv, resv);
    }
    c.last_solve_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (R_out) std::memcpy(R_out, cm.R.data(), sizeof(double)*cm.R.size());
    if (out) {
      auto co = [](svb200_sublsresult& d, const FSILS_subLsType& s) {
        d.success = s.success; d.itr = s.itr; d.iNorm = s.iNorm; d.fNorm = s.fNorm; d.dB = s.dB; d.callD = s.callD;
      };
      co(out->RI, fls.RI); co(out->GM, fls.GM); co(out->CG, fls.CG);
      out->Resm = fls.Resm; out->Resc = fls.Resc;
      out->hist_n = 0;
    }
  });
}

/// A boundary face of mesh iM: connectivity IENb(eNoNb,nElb) in global node ids and the parent element gE(nElb).
/// What nn::select_eleb does (solver/nn.cpp:1354-1402) with the reference's own Gauss/shape routines.
int svref_add_face(void* h, int iM, int eNoNb, int nElb, const int* IENb, const int* gE)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& m = c.com_mod.msh.at(iM);
    m.fa.emplace_back();
    m.nFa = (int)m.fa.size();
    auto& fa = m.fa.back();
    fa.iM = iM; fa.eNoN = eNoNb; fa.nEl = nElb; fa.gnEl = nElb; fa.name = "face" + std::to_string(m.nFa - 1);
    if (eNoNb == 3) { fa.eType = consts::ElementType::TRI3; fa.nG = 3; }      // nn_elem_props.h (face props)
    else if (eNoNb == 4) { fa.eType = consts::ElementType::QUD4; fa.nG = 4; }
    else throw std::runtime_error("[ref_harness] face type not supported");
    fa.IEN.resize(eNoNb, nElb);
    std::memcpy(fa.IEN.data(), IENb, sizeof(int)*eNoNb*nElb);
    fa.gE.resize(nElb);
    std::memcpy(fa.gE.data(), gE, sizeof(int)*nElb);
    fa.w = Vector<double>(fa.nG);
    fa.xi = Array<double>(2, fa.nG);
    nn::get_gip(nullptr, fa);
    fa.N = Array<double>(fa.eNoN, fa.nG);
    fa.Nx = Array3<double>(2, fa.eNoN, fa.nG);
    for (int g = 0; g < fa.nG; g++) nn::get_gnn(nullptr, g, fa);
  });
}

int svref_get_face_tables(void* h, int iM, int iFa, int* nG, double* w, double* N, double* Nx)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& fa = c.com_mod.msh.at(iM).fa.at(iFa);
    *nG = fa.nG;
    if (w) std::memcpy(w, fa.w.data(), sizeof(double)*fa.nG);
    if (N) std::memcpy(N, fa.N.data(), sizeof(double)*fa.eNoN*fa.nG);
    if (Nx) std::memcpy(Nx, fa.Nx.data(), sizeof(double)*2*fa.eNoN*fa.nG);
  });
}

/// The loop of eq_assem::b_assem_neu_bc (solver/eq_assem.cpp:31-149) around the reference's own nn::gnnb,
/// fluid::b_fluid / l_elas::b_l_elas and FsilsLinearAlgebra::assemble; hg(tnNo) as set_bc_neu_l builds it.
int svref_assemble_neu(void* h, int iM, int iFa, const svb200_eqparams* e, const svb200_dmnparams* dmn, int nDmn,
    const double* hg)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    fill_eq(c, *e, dmn, nDmn);
    auto& cm = c.com_mod;
    auto& msh = cm.msh.at(iM);
    auto& lFa = msh.fa.at(iFa);
    auto& eq = cm.eq[0];
    const int nsd = 3, dof = cm.dof, tDof = cm.tDof, eNoN = lFa.eNoN;
    const auto& Yg = c.sol.intermediate.get_velocity();
    for (int el = 0; el < lFa.nEl; el++) {
      int Ec = lFa.gE(el);
      cm.cDmn = all_fun::domain(cm, msh, cm.cEq, Ec);
      auto cPhys = eq.dmn[cm.cDmn].phys;
      Vector<int> ptr(eNoN);
      Vector<double> N(eNoN), hl(eNoN);
      Array<double> yl(tDof,eNoN), lR(dof,eNoN);
      Array3<double> lK(dof*dof,eNoN,eNoN);
      for (int a = 0; a < eNoN; a++) {
        int Ac = lFa.IEN(a,el);
        ptr(a) = Ac;
        hl(a) = hg[Ac];
        for (int i = 0; i < tDof; i++) yl(i,a) = Yg(i,Ac);
      }
      for (int g = 0; g < lFa.nG; g++) {
        Vector<double> nV(nsd);
        auto Nx = lFa.Nx.rslice(g);
        nn::gnnb(cm, lFa, el, g, nsd, nsd-1, eNoN, Nx, nV, c.sol, consts::MechanicalConfigurationType::reference);
        double Jac = utils::norm(nV);
        nV = nV / Jac;
        double w = lFa.w(g)*Jac;
        N = lFa.N.col(g);
        double hh = 0.0;
        Vector<double> y(tDof);
        for (int a = 0; a < eNoN; a++) {
          hh = hh + N(a)*hl(a);
          y = y + N(a)*yl.col(a);
        }
        if (cPhys == consts::EquationType::phys_fluid) fluid::b_fluid(cm, eNoN, w, N, y, hh, nV, lR, lK);
        else l_elas::b_l_elas(cm, eNoN, w, N, hh, nV, lR);
      }
      eq.linear_algebra->assemble(cm, eNoN, ptr, lK, lR);
    }
  });
}

/// KU = K U with the reference SpMV on the current com_mod.Val.
int svref_spmv(void* h, int dof, const double* U, double* KU)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& cm = c.com_mod;
    Array<double> u(dof, cm.tnNo), ku(dof, cm.tnNo);
    std::memcpy(u.data(), U, sizeof(double)*dof*cm.tnNo);
    spar_mul::fsils_spar_mul_vv(cm.lhs, cm.lhs.rowPtr, cm.lhs.colPtr, dof, cm.Val, u, ku);
    std::memcpy(KU, ku.data(), sizeof(double)*dof*cm.tnNo);
  });
}

/// ustruct::ustruct_r (solver/ustruct.cpp:1742-1845) for Newton iteration `itr` (1-based like eq.itr): R -= Kd Rd / am with
/// Rd = amg Ad - Yg; eq and domain parameters are those of the last svref_assemble.
int svref_ustruct_r(void* h, int itr, const double* Ad)
{
  auto& c = *static_cast<RefCase*>(h);
  return guarded([&] {
    auto& cm = c.com_mod;
    const int n = cm.tnNo;
    cm.Ad.resize(3, n);
    std::memcpy(cm.Ad.data(), Ad, sizeof(double)*3*n);
    cm.Rd.resize(3, n);
    cm.Rd = 0.0;
    cm.eq[0].itr = itr;
    // all_fun::is_domain (solver/all_fun.cpp:1059-1089) with a single domain needs nothing else; with several (FSI) it reads
    // com_mod.dmnId, built from the element domain ids as read_msh.cpp:1504-1519 does.  rowPtr/colPtr: svref_build_graph
    if (cm.eq[0].nDmn > 1) {
      cm.dmnId.resize(n);
      for (int a = 0; a < n; a++) cm.dmnId(a) = 0;
      for (auto& m : cm.msh)
        if (m.eId.size() != 0)
          for (int e = 0; e < m.nEl; e++)
            for (int a = 0; a < m.eNoN; a++) cm.dmnId(m.IEN(a, e)) |= m.eId(e);
    }
    if (c.backend) {
      // the patch of Integrator::step: the plug-in runs ustruct_r on the device-resident R and Kd
      if (!c.backend_ustruct_r) throw std::runtime_error("[ref_harness] backend without ustruct_r hook");
      c.backend_ustruct_r(c.backend, &cm);
      return;
    }
    ustruct::ustruct_r(cm, c.sol);
  });
}

int svref_last_timing(void* h, double* assemble_s, double* solve_s)
{
  auto& c = *static_cast<RefCase*>(h);
  if (assemble_s) *assemble_s = c.last_assemble_s;
  if (solve_s) *solve_s = c.last_solve_s;
  return 0;
}

} // extern "C"
