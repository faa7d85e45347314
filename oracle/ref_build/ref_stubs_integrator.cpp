// Link-closure stubs for the generalised-alpha part of oracle/_ref/libsvref.so (test infrastructure only).
//
// oracle/Makefile compiles the reference's Integrator.cpp, Simulation.cpp, Parameters.cpp, set_bc.cpp, eq_assem.cpp, ls.cpp
// (and what they need) unmodified, so that the harness can drive the reference's own Integrator::predictor / initiator /
// corrector, set_bc::set_bc_dir, eq_assem::global_eq_assem and ls_ns::ls_alloc / ls_solve.  Those translation units
// also reference routines of subsystems that are outside the hot path (0D/1D coupling, RIS, CMM, shells, CEP, contact, VTK
// output ...).  They are never reached by the harness; each stub throws if it ever is.
#include "ComMod.h"
#include "CepMod.h"
#include "CmMod.h"
#include "Simulation.h"
#include "SolutionStates.h"
#include "LinearAlgebra.h"
#include "FsilsLinearAlgebra.h"
#include "BoundaryCondition.h"
#include "CoupledBoundaryCondition.h"
#include "svZeroD_interface.h"
#include "svOneD_interface.h"
#include "ris.h"
#include "post.h"
#include "output.h"
#include "contact.h"
#include "cmm.h"
#include "cep_ion.h"
#include "cep.h"
#include "bf.h"
#include "baf_ini.h"
#include "stokes.h"
#include "shells.h"

#include <stdexcept>
#include <string>

#define OUT_OF_SCOPE(what) throw std::runtime_error(std::string("[oracle] ") + what + " is outside the hot path (stub)")

namespace svZeroD { void calc_svZeroD(ComMod&, const CmMod&, char) { OUT_OF_SCOPE("svZeroD"); } }
namespace svOneD { void calc_svOneD(ComMod&, const CmMod&, char) { OUT_OF_SCOPE("svOneD"); } }
// (namespace ris: the reference's own ris.cpp is compiled since round 2 session 3)
namespace post {
void fib_stretch(const ComMod&, const int, const mshType&, const Array<double>&, Vector<double>&) { OUT_OF_SCOPE("post::fib_stretch"); }
void fib_stretch_rate(const ComMod&, const int, const mshType&, const SolutionStates&, Vector<double>&) { OUT_OF_SCOPE("post::fib_stretch_rate"); }
}
namespace output { void output_result(Simulation*, std::array<double,3>&, const int, const int) {} }
namespace contact { void construct_contact_pnlty(ComMod&, CmMod&, const SolutionStates&) { OUT_OF_SCOPE("contact"); } }
namespace cmm {
void cmm_b(ComMod&, const faceType&, const int, const Array<double>&, const Array<double>&, const Array<double>&, const Array<double>&,
    const Vector<double>&, const Vector<double>&, const Vector<int>&, const SolutionStates&) { OUT_OF_SCOPE("cmm"); }
void construct_cmm(ComMod&, const mshType&, const SolutionStates&) { OUT_OF_SCOPE("cmm"); }
}
namespace cep_ion { void cep_integ(Simulation*, const int, const int, SolutionStates&, const Vector<double>&) { OUT_OF_SCOPE("cep_ion"); } }
namespace cep {
void b_cep(ComMod&, const int, const double, const Vector<double>&, const double, Array<double>&) { OUT_OF_SCOPE("cep"); }
void construct_cep(ComMod&, CepMod&, const mshType&, const SolutionStates&) { OUT_OF_SCOPE("cep"); }
}
namespace bf { void set_bf(ComMod&, const SolutionStates&) { OUT_OF_SCOPE("bf"); } }
namespace baf_ini_ns { void bc_ini(const ComMod&, const CmMod&, bcType&, faceType&, const SolutionStates&) { OUT_OF_SCOPE("baf_ini"); } }
namespace stokes { void construct_stokes(ComMod&, const mshType&, const SolutionStates&) { OUT_OF_SCOPE("stokes"); } }
namespace shells { void construct_shell(ComMod&, const mshType&, const SolutionStates&) { OUT_OF_SCOPE("shells"); } }

// The factory would also pull in the PETSc / Trilinos adapters (see ref_stubs.cpp).
LinearAlgebra* LinearAlgebraFactory::create_interface(consts::LinearAlgebraType t)
{
  if (t == consts::LinearAlgebraType::fsils) return new FsilsLinearAlgebra();
  OUT_OF_SCOPE("LinearAlgebraFactory for a non-FSILS interface");
}

// Boundary-condition objects that read VTK files (BoundaryCondition.cpp / CoupledBoundaryCondition.cpp need VTK headers).
double BoundaryCondition::get_value(const std::string&, int) const { OUT_OF_SCOPE("BoundaryCondition"); }
bool BoundaryCondition::get_flag(const std::string&) const { OUT_OF_SCOPE("BoundaryCondition"); }
int BoundaryCondition::get_local_index(int) const { OUT_OF_SCOPE("BoundaryCondition"); }
void CoupledBoundaryCondition::compute_flowrates(ComMod&, const CmMod&, const SolutionStates&) { OUT_OF_SCOPE("CoupledBoundaryCondition"); }
void CoupledBoundaryCondition::compute_pressures(ComMod&, const CmMod&, const SolutionStates&) { OUT_OF_SCOPE("CoupledBoundaryCondition"); }
void CoupledBoundaryCondition::copy_cap_surface_to_linear_solver_face(ComMod&, fsi_linear_solver::FSILS_faceType&,
    consts::MechanicalConfigurationType, const SolutionStates&) const { OUT_OF_SCOPE("CoupledBoundaryCondition"); }
double CoupledBoundaryCondition::get_Qn() const { OUT_OF_SCOPE("CoupledBoundaryCondition"); }
double CoupledBoundaryCondition::get_pressure() const { OUT_OF_SCOPE("CoupledBoundaryCondition"); }
void CoupledBoundaryCondition::perturb_flowrate(double) { OUT_OF_SCOPE("CoupledBoundaryCondition"); }
CoupledBoundaryCondition::State CoupledBoundaryCondition::save_state() const { OUT_OF_SCOPE("CoupledBoundaryCondition"); }
void CoupledBoundaryCondition::restore_state(const State&) { OUT_OF_SCOPE("CoupledBoundaryCondition"); }
