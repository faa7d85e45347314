// B200LinearAlgebra.cpp — see B200LinearAlgebra.h.  Host C++ over the C ABI of libsvb200.so; the only place where
// the reference's ComMod / eqType / mshType / FSILS_lhsType objects are translated into plain arrays.
#include "B200LinearAlgebra.h"

#include "consts.h"
#include "fils_struct.hpp"
#include "fs.h"

#include <algorithm>
#include <array>
#include <cstring>
#include <stdexcept>
#include <string>

namespace {

int to_phys(consts::EquationType p)
{
  using consts::EquationType;
  switch (p) {
    case EquationType::phys_fluid: return SVB200_PHYS_FLUID;
    case EquationType::phys_struct: return SVB200_PHYS_STRUCT;
    case EquationType::phys_FSI: return SVB200_PHYS_FSI;
    case EquationType::phys_mesh: return SVB200_PHYS_MESH;
    case EquationType::phys_lElas: return SVB200_PHYS_LELAS;
    case EquationType::phys_heatS: return SVB200_PHYS_HEATS;
    case EquationType::phys_heatF: return SVB200_PHYS_HEATF;
    case EquationType::phys_ustruct: return SVB200_PHYS_USTRUCT;
    default: return -1;
  }
}

double prop_or(const dmnType& d, consts::PhysicalProperyType k, double dflt = 0.0)
{
  auto it = d.prop.find(k);
  return it == d.prop.end() ? dflt : it->second;
}

}  // namespace

namespace b200 {

svb200_eqparams eq_params(const ComMod& com_mod, const eqType& eq, const mshType& lM, int scatter)
{
  svb200_eqparams e{};
  e.dt = com_mod.dt;
  e.af = eq.af; e.am = eq.am; e.gam = eq.gam; e.beta = eq.beta;
  e.phys = to_phys(eq.phys);
  if (e.phys < 0) throw std::runtime_error("[B200LinearAlgebra] this equation's physics is not on the device path");
  e.dof = com_mod.dof;
  e.tDof = com_mod.tDof;
  e.s = eq.s;
  e.mvMsh = com_mod.mvMsh ? 1 : 0;
  e.vmsStab = (lM.nFs == 1) ? 1 : 0;        // Code/Source/solver/fluid.cpp:496-500
  e.scatter = scatter;
  if (com_mod.pstEq) e.reserved |= SVB200_EQ_PRESTRESS;     // prestress equation: pSn / pSa accumulated on the device
  return e;
}

std::vector<svb200_dmnparams> domain_params(const eqType& eq)
{
  using namespace consts;
  std::vector<svb200_dmnparams> out(eq.nDmn);
  for (int i = 0; i < eq.nDmn; i++) {
    const dmnType& d = eq.dmn[i];
    svb200_dmnparams p{};
    p.Id = d.Id;
    p.phys = to_phys(d.phys);
    if (p.phys < 0) throw std::runtime_error("[B200LinearAlgebra] domain physics is not on the device path");
    const bool solid = (d.phys == EquationType::phys_struct || d.phys == EquationType::phys_ustruct);
    // l_elas_3d (mesh and linear-elasticity equations) reads solid_density like struct_3d (l_elas.cpp:275)
    // heats_3d reads solid_density as well (heats.cpp:204); heatf_3d uses no density
    const bool solid_rho = solid || d.phys == EquationType::phys_mesh || d.phys == EquationType::phys_lElas ||
                           d.phys == EquationType::phys_heatS;
    p.rho = prop_or(d, solid_rho ? PhysicalProperyType::solid_density : PhysicalProperyType::fluid_density);
    p.f[0] = prop_or(d, PhysicalProperyType::f_x);
    p.f[1] = prop_or(d, PhysicalProperyType::f_y);
    p.f[2] = prop_or(d, PhysicalProperyType::f_z);
    p.K_darcy = prop_or(d, PhysicalProperyType::inverse_darcy_permeability);
    p.backflow_stab = prop_or(d, PhysicalProperyType::backflow_stab);
    p.dmp = prop_or(d, PhysicalProperyType::damping);
    p.E = prop_or(d, PhysicalProperyType::elasticity_modulus);
    p.nu = prop_or(d, PhysicalProperyType::poisson_ratio);
    p.conductivity = prop_or(d, PhysicalProperyType::conductivity);
    p.source_term = prop_or(d, PhysicalProperyType::source_term);
    p.ctau_M = prop_or(d, PhysicalProperyType::ctau_M);
    p.ctau_C = prop_or(d, PhysicalProperyType::ctau_C);
    switch (d.fluid_visc.viscType) {
      case FluidViscosityModelType::viscType_CY: p.viscType = SVB200_VISC_CY; break;
      case FluidViscosityModelType::viscType_Cass: p.viscType = SVB200_VISC_CASSON; break;
      default: p.viscType = SVB200_VISC_CONST; break;
    }
    p.mu_i = d.fluid_visc.mu_i; p.mu_o = d.fluid_visc.mu_o; p.lam = d.fluid_visc.lam;
    p.a = d.fluid_visc.a; p.n = d.fluid_visc.n;
    if (solid) {
      switch (d.stM.isoType) {
        case ConstitutiveModelType::stIso_nHook: p.isoType = SVB200_ISO_NHK; break;
        case ConstitutiveModelType::stIso_MR: p.isoType = SVB200_ISO_MR; break;
        case ConstitutiveModelType::stIso_Gucci: p.isoType = SVB200_ISO_GUCCIONE; break;
        case ConstitutiveModelType::stIso_StVK: p.isoType = SVB200_ISO_STVK; break;
        case ConstitutiveModelType::stIso_HGO: p.isoType = SVB200_ISO_HGO; break;
        case ConstitutiveModelType::stIso_HO: p.isoType = SVB200_ISO_HO; break;
        case ConstitutiveModelType::stIso_HO_ma: p.isoType = SVB200_ISO_HO_MA; break;
        case ConstitutiveModelType::stArtificialNeuralNet: {
          // the CANN parameter table (set_material_props.h:155-175)
          p.isoType = SVB200_ISO_CANN;
          const auto& t = d.stM.paramTable;
          if (t.num_rows < 1 || t.num_rows > SVB200_CANN_MAX_ROWS)
            throw std::runtime_error("[B200LinearAlgebra] CANN parameter table with more than 16 rows");
          p.cann_rows = t.num_rows;
          for (int r = 0; r < t.num_rows; r++) {
            p.cann_inv[r] = t.invariant_indices(r);
            for (int k = 0; k < 3; k++) { p.cann_act[r][k] = t.activation_functions(r, k); p.cann_w[r][k] = t.weights(r, k); }
          }
        } break;
        default: throw std::runtime_error("[B200LinearAlgebra] isochoric constitutive model not implemented on the device");
      }
      p.active_stress = (d.active_stress != nullptr) ? 1 : 0;
      switch (d.stM.volType) {
        case ConstitutiveModelType::stVol_Quad: p.volType = SVB200_VOL_QUAD; break;
        case ConstitutiveModelType::stVol_ST91: p.volType = SVB200_VOL_ST91; break;
        case ConstitutiveModelType::stVol_M94: p.volType = SVB200_VOL_M94; break;
        default: p.volType = SVB200_VOL_NONE; break;
      }
      p.Kpen = d.stM.Kpen; p.C10 = d.stM.C10; p.C01 = d.stM.C01;
      p.bff = d.stM.bff; p.bss = d.stM.bss; p.bfs = d.stM.bfs;
      p.st_a = d.stM.a; p.st_b = d.stM.b; p.aff = d.stM.aff; p.ass = d.stM.ass; p.afs = d.stM.afs; p.kap = d.stM.kap; p.khs = d.stM.khs;
      if (d.solid_visc.viscType != SolidViscosityModelType::viscType_NA) {
        p.solid_visc_mu = d.solid_visc.mu;
        p.solidViscType = (d.solid_visc.viscType == SolidViscosityModelType::viscType_Potential) ? SVB200_SOLID_VISC_POTENTIAL
                                                                                                 : SVB200_SOLID_VISC_NEWTONIAN;
      }
    }
    out[i] = p;
  }
  return out;
}

svb200_lsparams ls_params(const fsi_linear_solver::FSILS_lsType& ls)
{
  auto cp = [](const fsi_linear_solver::FSILS_subLsType& s) {
    svb200_sublsparams d{};
    d.mItr = s.mItr; d.sD = s.sD; d.relTol = s.relTol; d.absTol = s.absTol;
    return d;
  };
  svb200_lsparams p{};
  p.RI = cp(ls.RI); p.GM = cp(ls.GM); p.CG = cp(ls.CG);
  return p;
}

bool global_eq_assem(ComMod& com_mod, CepMod& cep_mod, const mshType& lM, const SolutionStates& solutions)
{
  auto& eq = com_mod.eq[com_mod.cEq];
  auto* la = dynamic_cast<B200LinearAlgebra*>(eq.linear_algebra);
  if (!la) return false;
  if (to_phys(eq.phys) < 0) return false;          // stokes, shells, CEP, ... stay on the host loop (and its assemble())
  if (eq.phys == consts::EquationType::phys_FSI) {
    // construct_fsi assembles fluid, struct and ustruct domains (fsi.cpp:203-262) and throws for lElas (:236); any other domain is
    // skipped there.  Fail loudly instead of assembling a system without such a domain's stiffness and residual.
    for (int d = 0; d < eq.nDmn; d++) {
      const auto ph = eq.dmn[d].phys;
      if (ph != consts::EquationType::phys_fluid && ph != consts::EquationType::phys_struct && ph != consts::EquationType::phys_ustruct)
        throw std::runtime_error("[B200LinearAlgebra] FSI with a domain that is neither fluid, struct nor ustruct is not "
                                 "implemented on the device; use the fsils linear algebra for this equation");
    }
  }
  // nodal active tensions of the electromechanics coupling (sv_struct.cpp:277-281, ustruct.cpp:294-298): they change once per
  // time step (active_stress.cpp:50-66), the copy is 3 doubles per node
  bool active = false;
  for (int d = 0; d < eq.nDmn; d++) active |= (eq.dmn[d].active_stress != nullptr);
  if (active) la->set_active_tension(cep_mod);
  // fitted RIS: ris::doassem_ris (fluid.cpp:750-754, fsi.cpp:349-353) — on the device as a row operation on the assembled rows of
  // every mesh (csrc/ris.cu); the plan follows RIS.clsFlg
  if (com_mod.risFlag && (eq.phys == consts::EquationType::phys_fluid || eq.phys == consts::EquationType::phys_FSI))
    la->set_ris(com_mod);
  // URIS valves (construct_fluid, fluid.cpp:622-672; the fluid elements of construct_fsi, fsi.cpp:170-216): the signed distance
  // function and the valve velocity move with the valve, so they are handed over at every assembly
  if (com_mod.urisFlag && (eq.phys == consts::EquationType::phys_fluid || eq.phys == consts::EquationType::phys_FSI))
    la->set_uris(com_mod);
  la->assemble_mesh(com_mod, lM, solutions);
  return true;
}

}  // namespace b200

// ---------------------------------------------------------------------------------------------------------

B200LinearAlgebra::B200LinearAlgebra()
{
  // consts::LinearAlgebraType has no `b200` member in the unmodified reference (INTEGRATION.md adds it).
  interface_type = consts::LinearAlgebraType::none;
  assembly_type = consts::LinearAlgebraType::none;
  preconditioner_type = consts::PreconditionerType::PREC_FSILS;
}

B200LinearAlgebra::~B200LinearAlgebra() { finalize(); }

void B200LinearAlgebra::check(int rc) const
{
  // the reference signals every failure on this path with std::runtime_error (e.g. linear_solver/gmres.cpp:500-502)
  if (rc != SVB200_OK) throw std::runtime_error(std::string("[B200LinearAlgebra] ") + svb200_last_error());
}

void B200LinearAlgebra::check_options(const consts::PreconditionerType prec_cond_type, const consts::LinearAlgebraType atype)
{
  // the two FSILS preconditioners (consts::fsils_preconditioners), as FsilsLinearAlgebra::check_options
  if (prec_cond_type != consts::PreconditionerType::PREC_FSILS && prec_cond_type != consts::PreconditionerType::PREC_RCS &&
      prec_cond_type != consts::PreconditionerType::PREC_NONE) {
    throw std::runtime_error("[svMultiPhysics] ERROR: b200 linear algebra can't use '" +
        consts::preconditioner_type_to_name.at(prec_cond_type) + "' for a preconditioner.");
  }
  if (atype != consts::LinearAlgebraType::none) {
    throw std::runtime_error("[svMultiPhysics] ERROR: b200 linear algebra assembles on the device; no other assembly type can be set.");
  }
}

void B200LinearAlgebra::set_assembly(consts::LinearAlgebraType atype)
{
  if (atype == consts::LinearAlgebraType::none) return;
  throw std::runtime_error("[B200LinearAlgebra] ERROR: Can't set b200 linear algebra to use '" +
      LinearAlgebra::type_to_name.at(atype) + "' for assembly.");
}

void B200LinearAlgebra::set_preconditioner(consts::PreconditionerType prec_type)
{
  if (consts::fsils_preconditioners.count(prec_type) == 0) {
    throw std::runtime_error("[B200LinearAlgebra] ERROR: b200 linear algebra can't use '" +
        consts::preconditioner_type_to_name.at(prec_type) + "' for a preconditioner.");
  }
  preconditioner_type = prec_type;
}

void B200LinearAlgebra::initialize(ComMod& com_mod, eqType& lEq)
{
  (void)lEq;
  if (ctx) return;
  // add_eq_linear_algebra (Code/Source/solver/main.cpp:44-53) runs before lhsa / fsils_lhs_create
  // (initialize.cpp:620,641), so only the device context is created here; the structure goes up on the first alloc().
  check(svb200_create(&ctx, device));
  (void)com_mod;
}

void B200LinearAlgebra::finalize()
{
  if (ctx) { svb200_destroy(ctx); ctx = nullptr; }
  structure_uploaded = false;
}

void B200LinearAlgebra::upload_structure(ComMod& com_mod)
{
  auto& lhs = com_mod.lhs;
  const int tnNo = com_mod.tnNo;
  if (com_mod.cm.np() > 1) {
    // one MPI rank = one partition = one B200: fsils_commu_create's communicator becomes an NCCL communicator
    char id[128] = {0};
    if (com_mod.cm.idcm() == 0) check(svb200_comm_unique_id(id));
    MPI_Bcast(id, 128, MPI_CHAR, 0, com_mod.cm.com());
    check(svb200_comm_init(ctx, com_mod.cm.np(), com_mod.cm.idcm(), id));
  }
  std::vector<int> nrank, ncount, nptr;
  for (int i = 0; i < lhs.nReq; i++) {
    auto& c = lhs.cS[i];
    nrank.push_back(c.iP);
    ncount.push_back(c.n);
    for (int k = 0; k < c.n; k++) nptr.push_back(c.ptr(k));
  }
  if ((int)com_mod.rowPtr.size() != tnNo + 1) throw std::runtime_error("[B200LinearAlgebra] com_mod.rowPtr is not built (lhsa)");
  check(svb200_set_graph(ctx, tnNo, lhs.nnz, com_mod.rowPtr.data(), com_mod.colPtr.data(), lhs.mynNo, lhs.map.data(),
                         lhs.nReq, nrank.data(), ncount.data(), nptr.data()));
  inv_map.assign(tnNo, -1);
  for (int a = 0; a < tnNo; a++) inv_map[lhs.map(a)] = a;

  for (int iM = 0; iM < com_mod.nMsh; iM++) {
    auto& m = com_mod.msh[iM];
    check(svb200_set_mesh(ctx, iM, m.eNoN, m.nEl, m.IEN.data(), m.eId.size() ? m.eId.data() : nullptr,
                          m.nFn, (m.nFn > 0 && m.fN.size()) ? m.fN.data() : nullptr, m.nG, m.w.data(), m.N.data(), m.Nx.data()));
    // second derivatives of the shape functions for nn::gn_nxx (fluid on non-linear elements, fluid.cpp:648-650)
    if (!m.fs.empty() && m.fs[0].Nxx.size() == 6 * m.eNoN * m.nG) check(svb200_set_mesh_nxx(ctx, iM, m.fs[0].Nxx.data()));
    // Taylor-Hood function spaces (nFs = 2): the four tables construct_fluid builds with fs::get_thood_fs (fluid.cpp:567, 690)
    if (m.nFs == 2) {
      std::array<fsType, 2> f1, f2;
      fs::get_thood_fs(com_mod, f1, m, false, 1);
      fs::get_thood_fs(com_mod, f2, m, false, 2);
      check(svb200_set_mesh_thood(ctx, iM, f1[1].eNoN, f2[1].nG, f1[1].lShpF ? 1 : 0, f1[1].N.data(), f1[1].Nx.data(), f2[1].w.data(),
                                  f2[0].N.data(), f2[0].Nx.data(), f2[1].N.data(), f2[1].Nx.data()));
    }
  }
  check(svb200_set_coords(ctx, com_mod.x.data()));
  structure_uploaded = true;
}

void B200LinearAlgebra::upload_faces(ComMod& com_mod)
{
  auto& lhs = com_mod.lhs;
  check(svb200_set_num_faces(ctx, lhs.nFaces));
  std::vector<int> glob;
  for (int f = 0; f < lhs.nFaces; f++) {
    auto& fa = lhs.face[f];
    glob.resize(fa.nNo);
    for (int a = 0; a < fa.nNo; a++) glob[a] = inv_map[fa.glob(a)];     // FSILS order -> host order (bc.cpp:55-58)
    const int grp = (fa.bGrp == fsi_linear_solver::BcType::BC_TYPE_Dir) ? SVB200_BC_DIR : SVB200_BC_NEU;
    const int dofF = fa.dof > 0 ? fa.dof : 1;
    check(svb200_set_face(ctx, f, grp, dofF, fa.nNo, glob.data(), fa.nNo ? fa.val.data() : nullptr, fa.sharedFlag ? 2 : 0));
    if (fa.has_cap) {
      // capping surface of a coupled BC (fils_struct.hpp:131-143): cap_glob is in FSILS order, negative = not on this rank
      const int nc = fa.cap_glob.size();
      std::vector<int> cg(nc);
      for (int a = 0; a < nc; a++) cg[a] = fa.cap_glob(a) < 0 ? -1 : inv_map[fa.cap_glob(a)];
      check(svb200_set_face_cap(ctx, f, nc, cg.data(), nc ? fa.cap_val.data() : nullptr));
    }
  }
}

/// ls_alloc (Code/Source/solver/ls.cpp:24-40): R and Val are zeroed ON THE DEVICE; com_mod.Val is never allocated on the
/// host (3.2 GB per rank at 10 M tets), com_mod.R keeps its size because the corrector reads the increment from it.
void B200LinearAlgebra::alloc(ComMod& com_mod, eqType& lEq)
{
  if (!ctx) initialize(com_mod, lEq);
  if (!structure_uploaded) upload_structure(com_mod);
  alloc_dof = com_mod.dof;
  stage_rows.clear(); stage_R.clear(); stage_krows.clear(); stage_kcols.clear(); stage_K.clear();
  check(svb200_alloc(ctx, alloc_dof));
}

/// Per-element contributions the host still computes (surface integrals b_assem_neu_bc, coupled BCs ...) are staged as a
/// COO list in the host's node numbering — what lhsa_ns::do_assem (Code/Source/solver/lhsa.cpp:70-114) would add — and
/// flushed to the device R / Val before the solve.
void B200LinearAlgebra::assemble(ComMod& com_mod, const int num_elem_nodes, const Vector<int>& eqN,
    const Array3<double>& lK, const Array<double>& lR)
{
  const int dof = com_mod.dof;
  const int d2 = dof * dof;
  for (int a = 0; a < num_elem_nodes; a++) {
    const int rowN = eqN(a);
    if (rowN == -1) continue;
    stage_rows.push_back(rowN);
    for (int i = 0; i < dof; i++) stage_R.push_back(lR(i, a));
    for (int b = 0; b < num_elem_nodes; b++) {
      const int colN = eqN(b);
      if (colN == -1) continue;
      stage_krows.push_back(rowN);
      stage_kcols.push_back(colN);
      for (int i = 0; i < d2; i++) stage_K.push_back(lK(i, a, b));
    }
  }
}

void B200LinearAlgebra::flush_host_contrib(int dof)
{
  if (stage_rows.empty() && stage_krows.empty()) return;
  check(svb200_add_host_contrib(ctx, dof, (int)stage_rows.size(), stage_rows.data(), stage_R.data(),
                                (int)stage_krows.size(), stage_krows.data(), stage_kcols.data(), stage_K.data()));
  stage_rows.clear(); stage_R.clear(); stage_krows.clear(); stage_kcols.clear(); stage_K.clear();
}

void B200LinearAlgebra::assemble_mesh(ComMod& com_mod, const mshType& lM, const SolutionStates& solutions)
{
  if (!ctx || !structure_uploaded || alloc_dof != com_mod.dof)
    throw std::runtime_error("[B200LinearAlgebra] assemble_mesh before ls_alloc");
  auto& eq = com_mod.eq[com_mod.cEq];
  const int iM = (int)(&lM - com_mod.msh.data());
  if (iM < 0 || iM >= com_mod.nMsh) throw std::runtime_error("[B200LinearAlgebra] mesh is not a member of com_mod.msh");
  const auto& Ag = solutions.intermediate.get_acceleration();
  const auto& Yg = solutions.intermediate.get_velocity();
  const auto& Dg = solutions.intermediate.get_displacement();
  const int tDof = com_mod.tDof;
  check(svb200_set_state(ctx, tDof, Ag.data(), Yg.data(), Dg.size() ? Dg.data() : nullptr,
                         com_mod.Bf.size() ? com_mod.Bf.data() : nullptr));
  if (eq.phys == consts::EquationType::phys_mesh) {
    const auto& Do = solutions.old.get_displacement();            // Code/Source/solver/mesh.cpp:60-75
    check(svb200_set_old_disp(ctx, tDof, Do.data()));
  }
  // nodal prestress (sv_struct.cpp:271-274, l_elas.cpp:90-92, fsi.cpp:127-129): pS0 changes once per time step
  // (Integrator.cpp:422-424), 6 doubles per node
  const bool solidEq = eq.phys == consts::EquationType::phys_struct || eq.phys == consts::EquationType::phys_lElas ||
                       eq.phys == consts::EquationType::phys_FSI;
  if (solidEq) check(svb200_set_prestress(ctx, com_mod.pS0.size() ? com_mod.pS0.data() : nullptr));
  svb200_eqparams e = b200::eq_params(com_mod, eq, lM, scatter);
  std::vector<svb200_dmnparams> d = b200::domain_params(eq);
  check(svb200_assemble(ctx, iM, &e, d.data(), (int)d.size()));
  if (com_mod.pstEq && (eq.phys == consts::EquationType::phys_struct || eq.phys == consts::EquationType::phys_lElas)) {
    // the running sums of this Newton iteration (zeroed by ls_alloc) back into com_mod for Integrator::corrector
    // (Integrator.cpp:912-924: commu, division by pSa)
    if (com_mod.pSn.nrows() != 6 || com_mod.pSn.ncols() != com_mod.tnNo) com_mod.pSn.resize(6, com_mod.tnNo);
    if (com_mod.pSa.size() != com_mod.tnNo) com_mod.pSa.resize(com_mod.tnNo);
    check(svb200_get_prestress(ctx, com_mod.pSn.data(), com_mod.pSa.data()));
  }
}

void B200LinearAlgebra::set_active_tension(const CepMod& cep_mod)
{
  const auto& cem = cep_mod.cem;
  if (cem.Ya_f.size() == 0) throw std::runtime_error("[B200LinearAlgebra] active stress: cep_mod.cem.Ya_f is empty");
  check(svb200_set_active_tension(ctx, cem.Ya_f.data(), cem.Ya_s.size() ? cem.Ya_s.data() : nullptr,
                                  cem.Ya_n.size() ? cem.Ya_n.data() : nullptr));
}

void B200LinearAlgebra::set_ris(const ComMod& com_mod)
{
  const int nP = com_mod.ris.nbrRIS;
  std::vector<int> closed(nP);
  for (int p = 0; p < nP; p++) closed[p] = com_mod.ris.clsFlg[p] ? 1 : 0;
  if (closed == ris_state) return;                  // the surfaces did not change state since the plan was built
  std::vector<int> nMap(nP), maps;
  for (int p = 0; p < nP; p++) {
    const auto& mp = com_mod.grisMapList[p].map;    // (2, n), column-major
    nMap[p] = mp.ncols();
    maps.insert(maps.end(), mp.data(), mp.data() + 2 * (size_t)mp.ncols());
  }
  check(svb200_set_ris(ctx, nP, nMap.data(), maps.data(), closed.data()));
  ris_state = closed;
}

void B200LinearAlgebra::set_uris(const ComMod& com_mod)
{
  const int nU = (com_mod.urisFlag && com_mod.urisActFlag) ? com_mod.nUris : 0;      // uris.cpp:1592-1594
  if (nU == 0) { check(svb200_set_uris(ctx, 0, nullptr, nullptr, nullptr, nullptr)); return; }
  if (nU > SVB200_MAX_URIS) throw std::runtime_error("[B200LinearAlgebra] more URIS valves than SVB200_MAX_URIS");
  const size_t n = (size_t)com_mod.tnNo;
  std::vector<svb200_uris> v(nU);
  std::vector<double> sdf(nU * n, 0.0), udf, vel;
  for (int i = 0; i < nU; i++) {
    const auto& u = com_mod.uris[i];
    // half-thickness now: ramp from the previous state's value to the current one over the DxClose / DxOpen steps (uris.cpp:1625-1649)
    const double d0 = u.clsFlg ? u.sdf_deps : u.sdf_deps_close, d1 = u.clsFlg ? u.sdf_deps_close : u.sdf_deps;
    const int steps = u.clsFlg ? u.DxClose.nslices() : u.DxOpen.nslices();
    double deps = d1;
    if (steps > 0 && u.cnt < steps) deps = (u.cnt <= 0) ? d0 : d0 + (static_cast<double>(u.cnt) / static_cast<double>(steps)) * (d1 - d0);
    v[i].resistance = u.resistance;
    v[i].sdf_deps = deps;
    v[i].scaffold_deps = u.sdf_deps_close;
    v[i].scaffold = u.scaffold_flag ? 1 : 0;
    v[i].include_velocity = u.include_uris_velocity ? 1 : 0;
    if ((size_t)u.sdf.size() != n) throw std::runtime_error("[B200LinearAlgebra] URIS: sdf is not sized to the fluid mesh nodes");
    std::memcpy(sdf.data() + i * n, u.sdf.data(), sizeof(double) * n);
    if (u.scaffold_flag) {
      if (udf.empty()) udf.assign(nU * n, 0.0);
      std::memcpy(udf.data() + i * n, u.scaffold_udf.data(), sizeof(double) * n);
    }
    if (u.include_uris_velocity) {
      if (vel.empty()) vel.assign(3 * nU * n, 0.0);
      std::memcpy(vel.data() + 3 * i * n, u.valve_velocity_fluid.data(), sizeof(double) * 3 * n);      // (nsd, tnNo) column-major
    }
  }
  check(svb200_set_uris(ctx, nU, v.data(), sdf.data(), udf.empty() ? nullptr : udf.data(), vel.empty() ? nullptr : vel.data()));
}

/// ustruct::ustruct_r (Code/Source/solver/ustruct.cpp:1742-1845), called where Integrator::step calls it
/// (Integrator.cpp:135-137): the device holds R and Kd, the host passes com_mod.Ad and the Newton iteration count.
void B200LinearAlgebra::ustruct_r(ComMod& com_mod)
{
  auto& eq = com_mod.eq[com_mod.cEq];
  if (eq.phys != consts::EquationType::phys_ustruct && eq.phys != consts::EquationType::phys_FSI) return;     // ustruct.cpp:1755-1757
  flush_host_contrib(alloc_dof);
  if (eq.phys == consts::EquationType::phys_FSI) {
    // only the nodes of a ustruct domain take part (all_fun::is_domain, ustruct.cpp:1776-1793)
    // all_fun.cpp:1059-1089: one domain -> its physics decides; several -> bit dmn.Id of com_mod.dmnId(node)
    std::vector<int32_t> flag(com_mod.tnNo, 0);
    if (eq.nDmn > 1 && com_mod.dmnId.size() == 0) throw std::runtime_error("Domain partitioning info is not provided.");
    for (int a = 0; a < com_mod.tnNo; a++)
      for (int d = 0; d < eq.nDmn; d++) {
        if (eq.dmn[d].phys != consts::EquationType::phys_ustruct) continue;
        if (eq.nDmn == 1 || ((com_mod.dmnId(a) >> eq.dmn[d].Id) & 1)) { flag[a] = 1; break; }
      }
    check(svb200_set_node_flags(ctx, flag.data()));
  }
  svb200_eqparams e = b200::eq_params(com_mod, eq, com_mod.msh[0], scatter);
  check(svb200_ustruct_r(ctx, &e, eq.itr, com_mod.Ad.data()));
}

/// fs::thood_val_rc (Code/Source/solver/fs.cpp:394-466), called where Integrator::step calls it (Integrator.cpp:140-147).
void B200LinearAlgebra::thood_val_rc()
{
  flush_host_contrib(alloc_dof);
  check(svb200_thood_val_rc(ctx));
}

void B200LinearAlgebra::commu_R()
{
  flush_host_contrib(alloc_dof);
  check(svb200_commu_R(ctx));
}

void B200LinearAlgebra::download(int what, double* dst)
{
  flush_host_contrib(alloc_dof);
  check(svb200_download(ctx, what, dst));
}

/// ls_solve -> fsils_solve (Code/Source/linear_solver/solve.cpp:23-166): on return com_mod.R holds the increment and
/// lEq.FSILS.{RI,GM,CG} the iteration counts / norms the Newton convergence test reads (Integrator.cpp:941-964).
void B200LinearAlgebra::solve(ComMod& com_mod, eqType& lEq, const Vector<int>& incL, const Vector<double>& res)
{
  using fsi_linear_solver::LinearSolverType;
  const int dof = com_mod.dof;
  flush_host_contrib(dof);
  upload_faces(com_mod);
  int type;
  switch (lEq.FSILS.LS_type) {
    case LinearSolverType::LS_TYPE_NS: type = SVB200_LS_NS; break;
    case LinearSolverType::LS_TYPE_GMRES: type = SVB200_LS_GMRES; break;
    case LinearSolverType::LS_TYPE_CG: type = SVB200_LS_CG; break;
    case LinearSolverType::LS_TYPE_BICGS: type = SVB200_LS_BICGS; break;
    default: throw std::runtime_error("FSILS: LS_type not defined");     // linear_solver/solve.cpp:141
  }
  svb200_lsparams p = b200::ls_params(lEq.FSILS);
  svb200_lsresult r{};
  if (com_mod.R.nrows() != dof || com_mod.R.ncols() != com_mod.tnNo) com_mod.R.resize(dof, com_mod.tnNo);
  const int prec = (lEq.linear_algebra_preconditioner == consts::PreconditionerType::PREC_RCS) ? SVB200_PREC_RCS : SVB200_PREC_FSILS;
  check(svb200_solve(ctx, dof, type, prec, &p, incL.size(), incL.data(), res.data(), com_mod.R.data(), &r));
  auto back = [](fsi_linear_solver::FSILS_subLsType& d, const svb200_sublsresult& s) {
    d.success = s.success != 0; d.itr = s.itr; d.iNorm = s.iNorm; d.fNorm = s.fNorm; d.dB = s.dB; d.callD = s.callD;
  };
  back(lEq.FSILS.RI, r.RI); back(lEq.FSILS.GM, r.GM); back(lEq.FSILS.CG, r.CG);
  lEq.FSILS.Resm = r.Resm; lEq.FSILS.Resc = r.Resc;
}
