// B200LinearAlgebra — the reference-side C++ host layer of the B200 engine.
//
// A fourth implementation of svMultiPhysics' linear-algebra plugin interface `class LinearAlgebra`
// (Code/Source/solver/LinearAlgebra.h:13-37; siblings FsilsLinearAlgebra.cpp, PetscLinearAlgebra.cpp,
// TrilinosLinearAlgebra.cpp) that forwards to the C ABI of libsvb200.so (include/svb200.h).  It is compiled
// AGAINST THE REFERENCE'S OWN HEADERS (ComMod.h, LinearAlgebra.h ...) and is what a maintainer drops into
// Code/Source/solver/ next to the other backends.  Besides the eight virtuals it carries the
// whole-mesh assembly hook `b200::global_eq_assem`, the early-out of eq_assem::global_eq_assem
// (Code/Source/solver/eq_assem.cpp:377-455): a per-element virtual assemble() cannot feed a GPU, so the element
// loops construct_fluid / construct_dsolid / construct_fsi / construct_mesh are replaced as a whole.
//
// No torch, no Python: ComMod in, ComMod out.
#ifndef B200_LINEAR_ALGEBRA_H
#define B200_LINEAR_ALGEBRA_H

#include "LinearAlgebra.h"
#include "ComMod.h"
#include "CepMod.h"
#include "SolutionStates.h"

#include <vector>

extern "C" {
#include "svb200.h"
}

class B200LinearAlgebra : public virtual LinearAlgebra {
  public:
    B200LinearAlgebra();
    ~B200LinearAlgebra();

    // ---- LinearAlgebra interface (same meaning and error behaviour as FsilsLinearAlgebra) -----------------
    virtual void alloc(ComMod& com_mod, eqType& lEq);
    virtual void assemble(ComMod& com_mod, const int num_elem_nodes, const Vector<int>& eqN,
        const Array3<double>& lK, const Array<double>& lR);
    virtual void check_options(const consts::PreconditionerType prec_cond_type, const consts::LinearAlgebraType assembly_type);
    virtual void initialize(ComMod& com_mod, eqType& lEq);
    virtual void set_assembly(consts::LinearAlgebraType assembly_type);
    virtual void set_preconditioner(consts::PreconditionerType prec_type);
    virtual void solve(ComMod& com_mod, eqType& lEq, const Vector<int>& incL, const Vector<double>& res);
    virtual void finalize();

    // ---- whole-mesh element assembly on the device (called by b200::global_eq_assem) ---------------------
    void assemble_mesh(ComMod& com_mod, const mshType& lM, const SolutionStates& solutions);
    /// cep_mod.cem.Ya_f / Ya_s / Ya_n -> device (domains with an active-stress model; sv_struct.cpp:277-281).
    void set_active_tension(const CepMod& cep_mod);
    /// com_mod.uris[] (URIS valves) -> svb200_set_uris: nodal |sdf| / scaffold udf / valve velocity and, per valve, the resistance
    /// and the half-thickness in effect (the open/close ramp of uris.cpp:1625-1649); removes them when urisActFlag is off.
    void set_uris(const ComMod& com_mod);
    /// com_mod.ris / grisMapList -> svb200_set_ris whenever a surface opened or closed (ris::doassem_ris on the device, csrc/ris.cu).
    void set_ris(const ComMod& com_mod);
    /// fs::thood_val_rc on the device-resident R / Val (Taylor-Hood meshes; patch of Integrator::step like ustruct_r).
    void thood_val_rc();
    /// all_fun::commu(com_mod, com_mod.R) of Integrator::step (Code/Source/solver/Integrator.cpp:124-129).
    void commu_R();
    /// ustruct::ustruct_r (Code/Source/solver/ustruct.cpp:1742) on the device-resident R and Kd.
    void ustruct_r(ComMod& com_mod);
    /// Debug / parity: device R(dof,tnNo) or Val(dof*dof,nnz) in the host's node / CSR slot order.
    void download(int what, double* dst);

    int device = 0;                                   ///< CUDA device of this rank (default: rank % visible GPUs)
    int scatter = SVB200_SCATTER_ATOMIC;              ///< svb200_scatter
    svb200_ctx* ctx = nullptr;

  private:
    std::vector<int> ris_state;                       // RIS.clsFlg the device plan was built for (empty: no plan)
    void upload_structure(ComMod& com_mod);           // graph, meshes, coordinates: once
    void upload_faces(ComMod& com_mod);               // lhs.face[]: before every solve (fsils_bc_update may change val)
    void flush_host_contrib(int dof);                 // surface terms the host assembled through assemble()
    void check(int rc) const;

    bool structure_uploaded = false;
    int alloc_dof = 0;
    std::vector<int> inv_map;                         // FSILS order -> host order (inverse of lhs.map)
    // COO staging of per-element assemble() calls (b_assem_neu_bc, set_bc_cpl ... stay on the host)
    std::vector<int> stage_rows, stage_krows, stage_kcols;
    std::vector<double> stage_R, stage_K;
};

namespace b200 {

/// Early-out for eq_assem::global_eq_assem (Code/Source/solver/eq_assem.cpp:397): returns true when the equation's
/// backend is B200LinearAlgebra and the whole mesh has been assembled on the device, false to let the host loop run.
bool global_eq_assem(ComMod& com_mod, CepMod& cep_mod, const mshType& lM, const SolutionStates& solutions);

/// eqType / dmnType -> the plain structs of the C ABI.
svb200_eqparams eq_params(const ComMod& com_mod, const eqType& eq, const mshType& lM, int scatter);
std::vector<svb200_dmnparams> domain_params(const eqType& eq);
svb200_lsparams ls_params(const fsi_linear_solver::FSILS_lsType& ls);

}  // namespace b200

#endif
