/* svb200.h — C ABI of the B200-native "assemble + FSILS solve" engine.
 *
 * This is the drop-in boundary for the one hot path of svMultiPhysics that this repository
 * accelerates: per Newton iteration, (1) element residual/tangent assembly, (2) the scatter into
 * the FSILS block-CSR system (rowPtr/colPtr/Val, R) and (3) the FSILS Krylov solve.
 *
 * Conventions (identical to the reference):
 *   - all arrays are column-major ("Fortran order"), 0-based, int32 indices, FP64 values;
 *   - nodal state arrays are (tDof, nNo): dof fastest, node slowest;
 *   - Val is (dof*dof, nnz): block k holds 16 (dof=4) contiguous doubles, entry dof*i+j
 *     (row-major inside the block), see Code/Source/solver/FsilsLinearAlgebra.cpp:35;
 *   - every pointer argument is a HOST pointer that is only borrowed for the duration of the call
 *     unless the function name ends in `_dev`.
 *
 * Every function returns 0 on success and a non-zero svb200_status on failure; the message is
 * available from svb200_last_error().  The library never falls back to a CPU path: when no CUDA
 * device is usable svb200_create() fails.
 *
 * Reference interfaces replaced (file:line are into /root/reference at the surveyed snapshot):
 *   LinearAlgebra::{alloc,assemble,solve}            Code/Source/solver/LinearAlgebra.h:22-29
 *   FsilsLinearAlgebra::{alloc,assemble,solve}       Code/Source/solver/FsilsLinearAlgebra.cpp:26-128
 *   eq_assem::global_eq_assem                        Code/Source/solver/eq_assem.cpp:377-455
 *   fluid::construct_fluid                           Code/Source/solver/fluid.cpp:480-762
 *   struct_ns::construct_dsolid                      Code/Source/solver/sv_struct.cpp:184-341
 *   lhsa_ns::lhsa / do_assem                         Code/Source/solver/lhsa.cpp:126-381 / 70-114
 *   fsi_linear_solver::fsils_lhs_create              Code/Source/linear_solver/lhs.cpp:30-348
 *   fsi_linear_solver::fsils_bc_create               Code/Source/linear_solver/bc.cpp:18-102
 *   fsi_linear_solver::fsils_solve                   Code/Source/linear_solver/solve.cpp:23-166
 *   fsi_linear_solver::fsils_commuv                  Code/Source/linear_solver/in_commu.cpp:84-143
 *   all_fun::commu                                   Code/Source/solver/all_fun.cpp:95-119
 */
#ifndef SVB200_H
#define SVB200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVB200_ABI_VERSION 6

#if defined(__GNUC__)
#define SVB200_API __attribute__((visibility("default")))
#else
#define SVB200_API
#endif

typedef struct svb200_ctx svb200_ctx;

typedef enum {
  SVB200_OK = 0,
  SVB200_ERR_INVALID = 1,     /* bad argument / call order */
  SVB200_ERR_CUDA = 2,        /* CUDA runtime error */
  SVB200_ERR_NCCL = 3,        /* NCCL error */
  SVB200_ERR_NUMERIC = 4,     /* what the reference throws std::runtime_error for: zero residual
                                 norm (linear_solver/gmres.cpp:500-502), non-positive Jacobian
                                 (solver/fluid.cpp:637-639) */
  SVB200_ERR_UNSUPPORTED = 5
} svb200_status;

/* Equation / domain physics: the subset of consts::EquationType on the hot path
 * (Code/Source/solver/consts.h). */
typedef enum {
  SVB200_PHYS_FLUID = 0,
  SVB200_PHYS_STRUCT = 1,
  SVB200_PHYS_FSI = 2,
  SVB200_PHYS_MESH = 3,
  SVB200_PHYS_LELAS = 4,    /* linear elasticity: l_elas::construct_l_elas + l_elas_3d (solver/l_elas.cpp:36-145, 249-365) */
  SVB200_PHYS_HEATS = 5,    /* heat conduction in a solid: heats::construct_heats + heats_3d (solver/heats.cpp:30-119, 186-233), dof = 1 */
  SVB200_PHYS_HEATF = 6,    /* advection-diffusion in a fluid: heatf::construct_heatf + heatf_3d (solver/heatf.cpp:52-139, 238-331), dof = 1 */
  SVB200_PHYS_USTRUCT = 7   /* mixed velocity-pressure solid: ustruct::construct_usolid + ustruct_3d_m/c + ustruct_do_assem
                               (solver/ustruct.cpp:203-400, 1165-1591, 629-871, 1595-1737), dof = 4 */
} svb200_phys;

/* consts::FluidViscosityModelType (solver/fluid.cpp:2254-2297). */
typedef enum { SVB200_VISC_CONST = 0, SVB200_VISC_CY = 1, SVB200_VISC_CASSON = 2 } svb200_visc;

/* consts::ConstitutiveModelType, isochoric part (solver/mat_models.cpp:435-581). */
typedef enum {
  SVB200_ISO_NHK = 0, SVB200_ISO_MR = 1, SVB200_ISO_GUCCIONE = 2, SVB200_ISO_STVK = 3,
  SVB200_ISO_HGO = 4,      /* Holzapfel-Gasser-Ogden, additive split (mat_models.cpp:469-511): C10, aff, bff, ass, bss, kap */
  SVB200_ISO_HO = 5,       /* Holzapfel-Ogden myocardium (:582-687): st_a, st_b, aff, bff, ass, bss, afs, bfs, khs */
  SVB200_ISO_HO_MA = 6,    /* Holzapfel-Ogden, modified anisotropy / full invariants (:689-773) */
  SVB200_ISO_CANN = 7      /* constitutive artificial neural network (:776-800, ArtificialNeuralNetMaterial.cpp): cann_* table */
} svb200_iso;
#define SVB200_CANN_MAX_ROWS 16
/* volumetric part (solver/mat_models.cpp:1441-1464). */
typedef enum { SVB200_VOL_NONE = 0, SVB200_VOL_QUAD = 1, SVB200_VOL_ST91 = 2, SVB200_VOL_M94 = 3 } svb200_vol;

/* consts::SolidViscosityModelType (solver/consts.h:466-471, solver/mat_models.cpp:1583-1762). */
typedef enum { SVB200_SOLID_VISC_NONE = 0, SVB200_SOLID_VISC_NEWTONIAN = 1, SVB200_SOLID_VISC_POTENTIAL = 2 } svb200_solid_visc;

/* fsi_linear_solver::LinearSolverType (linear_solver/fils_struct.hpp). */
typedef enum { SVB200_LS_NS = 0, SVB200_LS_GMRES = 1, SVB200_LS_CG = 2, SVB200_LS_BICGS = 3 } svb200_ls_type;
/* consts::PreconditionerType, the two FSILS ones (consts.h:421-432): the diagonal (Jacobi) preconditioner
 * precond_diag (linear_solver/precond.cpp:95-242) and the row-and-column max-norm scaling precond_rcs (:251-523). */
typedef enum { SVB200_PREC_FSILS = 0, SVB200_PREC_RCS = 1 } svb200_prec;
/* fsi_linear_solver::BcType. */
typedef enum { SVB200_BC_DIR = 0, SVB200_BC_NEU = 1 } svb200_bc_type;
/* Scatter mode of the element assembly. */
typedef enum {
  SVB200_SCATTER_ATOMIC = 0,  /* FP64 red.global.add; fastest, last-bit non-deterministic */
  SVB200_SCATTER_COLORED = 1  /* graph-coloured, conflict-free, bitwise reproducible */
} svb200_scatter;
/* What svb200_download / svb200_upload move. */
typedef enum { SVB200_ARRAY_R = 0, SVB200_ARRAY_VAL = 1, SVB200_ARRAY_W = 2,
               SVB200_ARRAY_KD = 3,  /* com_mod.Kd((nsd+1)*nsd, nnz), the displacement tangent of ustruct (solver/ustruct.cpp:1621) */
               SVB200_ARRAY_RD = 4   /* com_mod.Rd(nsd, nNo), the displacement residual svb200_ustruct_r leaves for the ustruct corrector */
} svb200_array;

/* Per-equation time-integration parameters (eqType af/am/gam/beta, ComMod dt/tDof/dof/mvMsh). */
typedef struct {
  double dt;
  double af, am, gam, beta;
  int32_t phys;      /* svb200_phys of the equation */
  int32_t dof;       /* unknowns per node of this equation (com_mod.dof) */
  int32_t tDof;      /* state dofs per node (com_mod.tDof) */
  int32_t s;         /* eq.s: first state dof of this equation */
  int32_t mvMsh;     /* com_mod.mvMsh: ALE convective velocity (solver/fluid.cpp:1909-1915) */
  int32_t vmsStab;   /* lM.nFs == 1 (solver/fluid.cpp:496-500); only 1 is supported */
  int32_t scatter;   /* svb200_scatter */
  int32_t reserved;  /* bit flags, 0 by default; SVB200_EQ_GENERAL_KERNEL: run a TET4 fluid mesh through the
                        per-Gauss-point kernel that serves HEX8 (cross-check of the specialised TET4 kernel) */
} svb200_eqparams;
#define SVB200_EQ_GENERAL_KERNEL 1
#define SVB200_EQ_PRESTRESS 2        /* com_mod.pstEq: struct / lElas assembly accumulates pSn, pSa (svb200_get_prestress) */

/* Per-domain material parameters (dmnType, stModelType, fluidViscModelType). */
typedef struct {
  int32_t Id;        /* dmn.Id; -1 = whole mesh, else bit index tested against eId (all_fun.cpp:122) */
  int32_t phys;      /* svb200_phys of the domain */
  double rho;        /* fluid_density or solid_density */
  double f[3];       /* f_x, f_y, f_z body force per unit mass */
  /* fluid */
  double K_darcy;    /* inverse_darcy_permeability */
  int32_t viscType;  /* svb200_visc */
  int32_t isoType;   /* svb200_iso */
  double mu_i, mu_o, lam, a, n;
  /* solid */
  int32_t volType;   /* svb200_vol */
  int32_t solidViscType;   /* svb200_solid_visc; used only when solid_visc_mu != 0 (0 is read as Newtonian then) */
  double Kpen;
  double C10, C01;
  double bff, bss, bfs;     /* Guccione exponents (stModelType bff/bss/bfs) */
  double dmp;               /* damping */
  double E, nu;             /* elasticity_modulus, poisson_ratio (mesh / linear elasticity) */
  double solid_visc_mu;     /* solid_visc.mu; 0 = no solid viscosity */
  double backflow_stab;     /* backflow stabilisation coefficient (fluid Neumann faces, solver/fluid.cpp:65) */
  /* stModelType a, b, aff, ass, afs, kap, khs (HGO / Holzapfel-Ogden; bff, bss, bfs above) — appended in ABI version 2 */
  double st_a, st_b, aff, ass, afs, kap, khs;
  /* appended in ABI version 3 */
  double conductivity, source_term;   /* heatS / heatF (solver/heats.cpp:202-204, heatf.cpp:259-260) */
  double ctau_M, ctau_C;              /* ustruct VMS constants (mat_models.cpp:1478-1479) */
  /* appended in ABI version 5 */
  int32_t active_stress;              /* dmn.active_stress != nullptr: the nodal active tensions of svb200_set_active_tension enter
                                         compute_pk2cc along the fibre / sheet / sheet-normal directions (sv_struct.cpp:277-281) */
  int32_t cann_rows;                  /* SVB200_ISO_CANN: rows of stM.paramTable (ArtificialNeuralNetMaterial.h) */
  int32_t cann_inv[SVB200_CANN_MAX_ROWS];      /* invariant_indices(row), 1..9 */
  int32_t cann_act[SVB200_CANN_MAX_ROWS][3];   /* activation_functions(row, 0..2) */
  double cann_w[SVB200_CANN_MAX_ROWS][3];      /* weights(row, 0..2) */
} svb200_dmnparams;

/* FSILS_subLsType inputs (linear_solver/fils_struct.hpp:198-242). */
typedef struct {
  int32_t mItr;
  int32_t sD;
  double relTol;
  double absTol;
} svb200_sublsparams;

typedef struct {
  svb200_sublsparams RI, GM, CG;
} svb200_lsparams;

/* FSILS_subLsType outputs. */
typedef struct {
  int32_t success;
  int32_t itr;
  double iNorm, fNorm, dB, callD;
} svb200_sublsresult;

typedef struct {
  svb200_sublsresult RI, GM, CG;
  int32_t Resm, Resc;
  /* residual history of the outermost Krylov loop: |err(i+1)| after each inner iteration
   * (what linear_solver/gmres.cpp prints under debug_gmres_v); at most hist_cap entries kept. */
  int32_t hist_n;
  int32_t hist_cap;
  double* hist;      /* caller-provided buffer of hist_cap doubles, or NULL */
} svb200_lsresult;

/* ---- lifecycle -------------------------------------------------------------------------- */
SVB200_API int svb200_abi_version(void);
SVB200_API const char* svb200_last_error(void);
SVB200_API int svb200_create(svb200_ctx** out, int device);
SVB200_API int svb200_destroy(svb200_ctx* ctx);

/* One process per GPU: rank/nranks of this context and the 128-byte ncclUniqueId generated by
 * svb200_comm_unique_id() on rank 0 and broadcast by the host (torch.distributed / MPI).
 * Replaces fsils_commu_create (linear_solver/commu.cpp:17-46). */
SVB200_API int svb200_comm_unique_id(void* id128);
SVB200_API int svb200_comm_init(svb200_ctx* ctx, int nranks, int rank, const void* id128);
/* Transport of the shared-node sums and scalar all-reduces: "p2p" = the library's own kernels storing into the peers'
 * HBM over NVLink / NVSwitch (CUDA IPC mailboxes; default), "nccl" = ncclSend/Recv + ncclAllReduce (SVB200_COMM=nccl or
 * when IPC is unavailable), "none" = single partition.  Valid after svb200_set_graph. */
SVB200_API const char* svb200_comm_transport(svb200_ctx* ctx);

/* ---- structure (once) ------------------------------------------------------------------- */
/* CSR graph in INPUT node order exactly as lhsa_ns::lhsa builds it (columns ascending per row).
 * map/mynNo and the shared-node lists are what fsils_lhs_create computes (linear_solver/lhs.cpp);
 * pass map=NULL, mynNo=nNo, nReq=0 for a single partition.  neigh_ptr is the concatenation of the
 * per-neighbour lists cS[i].ptr (FSILS-order local ids), neigh_n[i] entries each. */
SVB200_API int svb200_set_graph(svb200_ctx* ctx, int32_t nNo, int32_t nnz, const int32_t* rowPtr, const int32_t* colPtr,
                     int32_t mynNo, const int32_t* map,
                     int32_t nReq, const int32_t* neigh_rank, const int32_t* neigh_n, const int32_t* neigh_ptr);

/* lhsa_ns::lhsa (solver/lhsa.cpp:126-381) on the device: the sorted node-to-node adjacency of all
 * meshes as CSR (columns ascending per row, every node pair of every element, diagonal included).
 * begin -> add_mesh (once per mesh, IEN(eNoN,nEl) in caller node ids) -> finish (returns nnz) -> get.
 * The result is bit-identical to the reference's rowPtr/colPtr; pass it to svb200_set_graph. */
SVB200_API int svb200_lhsa_begin(svb200_ctx* ctx, int32_t nNo);
SVB200_API int svb200_lhsa_add_mesh(svb200_ctx* ctx, int32_t eNoN, int32_t nEl, const int32_t* IEN);
SVB200_API int svb200_lhsa_finish(svb200_ctx* ctx, int32_t* nnz);
SVB200_API int svb200_lhsa_get(svb200_ctx* ctx, int32_t* rowPtr, int32_t* colPtr);

/* One mesh (mshType): connectivity IEN(eNoN,nEl) in input node ids, optional domain bitmask
 * eId(nEl), optional fibres fN(3*nFn,nEl), and the reference-element tables of fs[0]:
 * w(nG), N(eNoN,nG), Nx(3,eNoN,nG).  eNoN = 4 (TET4) or 8 (HEX8). */
SVB200_API int svb200_set_mesh(svb200_ctx* ctx, int32_t iM, int32_t eNoN, int32_t nEl, const int32_t* IEN,
                    const int32_t* eId, int32_t nFn, const double* fN,
                    int32_t nG, const double* w, const double* N, const double* Nx);

/* Second parametric derivatives of the shape functions, fs[0].Nxx(6,eNoN,nG), Voigt order (00,11,22,01,12,02): what
 * construct_fluid hands to nn::gn_nxx (solver/fluid.cpp:648-650, solver/nn.cpp:1172-1283).  Needed for the fluid on
 * HEX8 meshes; identically zero (and optional) for TET4.  Call after svb200_set_mesh. */
SVB200_API int svb200_set_mesh_nxx(svb200_ctx* ctx, int32_t iM, const double* Nxx);

/* Reference coordinates com_mod.x(3,nNo). */
SVB200_API int svb200_set_coords(svb200_ctx* ctx, const double* x);

/* Linear-solver faces (fsils_bc_create): glob are INPUT-order node ids, val(face_dof,nNo).
 * sharedFlag: 0 = face lives on this partition only; 1 = shared between partitions, the library sums val over the
 * shared nodes (linear_solver/bc.cpp:70-102); 2 = shared and val is lhs.face[].val, already summed by the host. */
SVB200_API int svb200_set_num_faces(svb200_ctx* ctx, int32_t nFaces);
SVB200_API int svb200_set_face(svb200_ctx* ctx, int32_t faIn, int32_t bGrp, int32_t face_dof, int32_t nNo,
                    const int32_t* glob, const double* val, int32_t sharedFlag);

/* Capping surface of a coupled Neumann face (FSILS_faceType::has_cap / cap_glob / cap_val, linear_solver/fils_struct.hpp:131-143):
 * cap_glob are INPUT-order node ids (negative = the cap node is not on this partition, add_bc_mul.cpp:70), cap_val(face_dof,
 * cap_nNo) the cap's nodal normal integrals.  The cap adds coef * (cap_val W . X) to the face's flow-rate sum in add_bc_mul
 * (add_bc_mul.cpp:62-81, 102-111); precond_diag scales it with W like face.val (precond.cpp:229-237).  Call after svb200_set_face,
 * on every partition (cap_nNo = 0 where the partition holds no cap node). */
SVB200_API int svb200_set_face_cap(svb200_ctx* ctx, int32_t faIn, int32_t cap_nNo, const int32_t* cap_glob, const double* cap_val);

/* ---- per Newton iteration --------------------------------------------------------------- */
/* ls_alloc (solver/ls.cpp:24-40): R(dof,nNo) and Val(dof*dof,nnz) are (re)zeroed on the device. */
SVB200_API int svb200_alloc(svb200_ctx* ctx, int32_t dof);

/* Upload the generalised-alpha intermediate state (Ag,Yg,Dg)(tDof,nNo) and body force Bf(3,nNo).
 * NULL pointers leave the device copy unchanged (Dg, Bf may never be set: treated as zero). */
SVB200_API int svb200_set_state(svb200_ctx* ctx, int32_t tDof, const double* Ag, const double* Yg, const double* Dg,
                     const double* Bf);

/* Old displacement Do(tDof,nNo) (solutions.old), read by the mesh-motion equation (solver/mesh.cpp:22-135). */
SVB200_API int svb200_set_old_disp(svb200_ctx* ctx, int32_t tDof, const double* Do);
/* Nodal active tensions cep_mod.cem.Ya_f / Ya_s / Ya_n (solver/CepMod.h:205-217; filled once per time step by the active-stress
 * model, solver/active_stress.cpp), nNo doubles each in INPUT node order; Ya_s / Ya_n may be NULL (zero).  Read by the struct,
 * FSI-solid and ustruct kernels for domains with active_stress set (sv_struct.cpp:277-281, ustruct.cpp:294-298).  The reference
 * throws when Ya_s or Ya_n is positive for a model other than Guccione / HO / HO-ma (mat_models.cpp:334-340): checked at assembly. */
/* Nodal prestress com_mod.pS0 (nsymd = 6, nNo), Voigt order 11,22,33,12,23,31, INPUT node order; NULL removes it.  Interpolated to
 * the Gauss points and added to the 2nd Piola-Kirchhoff stress by the struct, FSI-solid and lElas kernels (sv_struct.cpp:635-680,
 * l_elas.cpp:321-338).  With SVB200_EQ_PRESTRESS in eq.reserved (com_mod.pstEq) the struct / lElas assembly also accumulates
 * pSn(:,A) += w N_a pSl and pSa(A) += w N_a (sv_struct.cpp:327-336); svb200_alloc zeroes the accumulators (Integrator::initiator,
 * Integrator.cpp:745-748) and svb200_get_prestress returns this partition's raw sums — the corrector's commu and division
 * (Integrator.cpp:912-924) stay with the caller. */
SVB200_API int svb200_set_prestress(svb200_ctx* ctx, const double* pS0);
SVB200_API int svb200_get_prestress(svb200_ctx* ctx, double* pSn, double* pSa);
SVB200_API int svb200_set_active_tension(svb200_ctx* ctx, const double* Ya_f, const double* Ya_s, const double* Ya_n);

/* Taylor-Hood function spaces of a fluid mesh (mshType::nFs = 2, <Use_taylor_hood_type_basis>): velocity on the mesh's own quadratic
 * element (svb200_set_mesh + svb200_set_mesh_nxx), pressure on its linear parent with eNoNq nodes = the first eNoNq nodes of every
 * element (fs::set_thood_fs, solver/fs.cpp:336-390: TET10 -> TET4, HEX20 / HEX27 -> HEX8).  The tables are those of fs::get_thood_fs
 * (fs.cpp:73-178), column-major like fsType holds them:
 *   momentum loop (the mesh's own rule, nG points):   Nq1(eNoNq, nG), Nqxi1(3, eNoNq, nG)       pressure space at those points
 *   continuity loop (the pressure space's rule, nG2): w2(nG2), Nw2(eNoN, nG2), Nwxi2(3, eNoN, nG2), Nq2(eNoNq, nG2), Nqxi2(3, eNoNq, nG2)
 * lShpF_q: the pressure space is linear (TET4): its gnn runs at Gauss point 0 only (fluid.cpp:620-626, 709-716).  A fluid equation on such
 * a mesh is assembled with eq.vmsStab = 0 (fluid.cpp:494-500): fluid_3d_m / fluid_3d_c with vmsFlag false; an FSI equation likewise
 * (fsi.cpp:47-50: fluid elements on the moved geometry, struct elements on the velocity space).  eNoNq = 0 returns the mesh to
 * equal-order spaces.  svb200_thood_val_rc is fs::thood_val_rc (fs.cpp:394-466; Integrator::step calls it after the assembly and the
 * boundary terms): for every node that is not a pressure node of its elements R(3, a) = 0 and the pressure-pressure entries of its row
 * become the identity. */
SVB200_API int svb200_set_mesh_thood(svb200_ctx* ctx, int32_t iM, int32_t eNoNq, int32_t nG2, int32_t lShpF_q, const double* Nq1,
                                     const double* Nqxi1, const double* w2, const double* Nw2, const double* Nwxi2, const double* Nq2,
                                     const double* Nqxi2);
SVB200_API int svb200_thood_val_rc(svb200_ctx* ctx);

/* Fitted resistive immersed surfaces (RIS): the coupling an OPEN surface adds, ris::doassem_ris (solver/ris.cpp:269-349), called per
 * element by construct_fluid (fluid.cpp:750-754) and construct_fsi (fsi.cpp:349-353): the residual row and the tangent row of every
 * element node listed in grisMapList[iProj].map are added a second time into the row of its twin across the surface, mapped columns
 * replaced by their twins.  On the device this is done on the ASSEMBLED rows of each mesh (csrc/ris.cu), inside svb200_assemble of a
 * fluid / FSI equation.  nProj projections (RIS.nbrRIS); maps holds, for projection p, 2 * nMap[p] node ids in the order of the
 * column-major grisMapList[p].map(2, n): map(0,0), map(1,0), map(0,1), ... (INPUT node order, every node with a twin: -1 is
 * rejected); closed[p] = RIS.clsFlg[p] (closed surfaces add nothing).  The CSR graph must contain the extra connections lhsa adds
 * when com_mod.risFlag is set (lhsa.cpp:168-193).  nProj = 0 removes the plan.  Call after svb200_set_graph and again whenever a
 * surface opens or closes.  Both nodes of a pair must belong to this partition (the plan is built from the local graph). */
SVB200_API int svb200_set_ris(svb200_ctx* ctx, int32_t nProj, const int32_t* nMap, const int32_t* maps, const int32_t* closed);

/* Unfitted resistive immersed surfaces (URIS valves): the penalty terms of fluid_3d_m / fluid_3d_c (solver/fluid.cpp:2006-2008,
 * 2042-2047, 2126-2129, 2166-2204, 2228-2234 and :1660-1703) with the per-Gauss-point factor of
 * uris::eval_uris_ris_factors_quadrature (solver/uris.cpp:1577-1673), used by construct_fluid (fluid.cpp:622-672) and the fluid
 * elements of construct_fsi (fsi.cpp:170-216):
 *   dist = sum_a N_a |sdf(A)|,  delta = (1 + cos(pi dist / deps)) / (2 deps^2) if dist < deps (and deps > 0), else 0
 *   factor = sum_valves resistance (delta + delta_scaffold),  velocity term = sum_valves resistance delta sum_a N_a v_valve(A).
 * One entry per valve (com_mod.uris[i]); sdf_deps is the half-thickness in effect at this step — the linear open/close ramp between
 * sdf_deps and sdf_deps_close over DxOpen / DxClose (uris.cpp:1625-1649) is scalar host logic and stays with the caller. */
#define SVB200_MAX_URIS 4
typedef struct {
  double resistance;          /* urisType::resistance */
  double sdf_deps;            /* effective half-thickness of the valve surface at this step */
  double scaffold_deps;       /* urisType::sdf_deps_close: thickness of the scaffold surface (read when scaffold != 0) */
  int32_t scaffold;           /* urisType::scaffold_flag */
  int32_t include_velocity;   /* urisType::include_uris_velocity */
} svb200_uris;
/* nUris = 0 removes the valves (com_mod.urisActFlag false).  sdf(nNo, nUris): urisType::sdf of valve i at sdf + i*nNo (signed or
 * not: the absolute value is used); scaffold_udf likewise (NULL if no valve has a scaffold); valve_vel(3, nNo, nUris):
 * urisType::valve_velocity_fluid (NULL if no valve includes its velocity).  INPUT node order.  The URIS terms live in the
 * per-Gauss-point kernel: on a TET4 mesh (atomic scatter) only the band of elements with a node inside a valve's or scaffold's
 * thickness runs through it, the rest stays with the closed-form kernel (exact there: the factor is zero); other meshes and the
 * deterministic scatter take the per-Gauss-point kernel throughout. */
SVB200_API int svb200_set_uris(svb200_ctx* ctx, int32_t nUris, const svb200_uris* valves, const double* sdf,
                               const double* scaffold_udf, const double* valve_vel);

/* global_eq_assem for mesh iM: element loop + scatter, R/Val stay on the device. */
SVB200_API int svb200_assemble(svb200_ctx* ctx, int32_t iM, const svb200_eqparams* eq,
                    const svb200_dmnparams* dmn, int32_t nDmn);

/* The assembly stage for HOST-resident state in one pipelined call: what svb200_set_state(Ag, Yg) + svb200_alloc(dof) +
 * svb200_assemble + svb200_commu_R + svb200_download(R) do one after the other (the sequence B200LinearAlgebra runs per Newton
 * iteration), with the transfers hidden behind the element kernel: the nodal state goes up in node chunks on a
 * copy stream and every chunk of element groups starts as soon as the nodes it reads have arrived; on a single partition the
 * finished residual rows come back on a third stream while later groups are assembled.  Ag, Yg are (tDof, nNo), R_out (dof, nNo)
 * or NULL, all in INPUT node order; page-lock them (svb200_host_register) for the copies to be asynchronous.  Pipelined for the
 * TET4 fluid equation with the atomic scatter; any other equation runs the plain sequence.  Call svb200_alloc(dof) once before. */
SVB200_API int svb200_assemble_host(svb200_ctx* ctx, int32_t iM, const svb200_eqparams* eq, const svb200_dmnparams* dmn, int32_t nDmn,
                         const double* Ag, const double* Yg, double* R_out);

/* Boundary face iFa of mesh iM (faceType): connectivity IENb(eNoNb,nElb) in input node ids, parent element
 * gE(nElb) (index into the mesh's elements), and the face reference-element tables w(nGb), N(eNoNb,nGb),
 * Nx(2,eNoNb,nGb).  eNoNb = 3 (TRI3 on TET4) or 4 (QUD4 on HEX8). */
/* Device timeline of the last pipelined svb200_assemble_host call, in ms after its start (CUDA events): [0] last upload chunk on the
 * device, [1] last element group done, [2] shared-node sum done (0 on a single partition), [3] streamed residual rows on the host,
 * [4] end of the call's device work; host wall clock after entry: [5] everything enqueued, [6] streams drained, [7] return.
 * ms8 holds 8 doubles. */
SVB200_API int svb200_last_host_stage(svb200_ctx* ctx, double* ms8);
SVB200_API int svb200_set_bface(svb200_ctx* ctx, int32_t iFa, int32_t iM, int32_t eNoNb, int32_t nElb, const int32_t* IENb,
                     const int32_t* gE, int32_t nGb, const double* w, const double* N, const double* Nx);
/* eq_assem::b_assem_neu_bc (solver/eq_assem.cpp:31-149) on the device: Neumann / traction face integral with
 * nn::gnnb normals (solver/nn.cpp:911-1117, reference configuration or x + Do(4..6) when mvMsh),
 * fluid::b_fluid incl. backflow stabilisation and its tangent (solver/fluid.cpp:21-108) on fluid domains,
 * l_elas::b_l_elas (solver/l_elas.cpp:21-32) otherwise.  hg(nNo) is the nodal traction magnitude set_bc_neu_l
 * builds (solver/set_bc.cpp:1489-1606), input node order; it is added into the device R / Val. */
SVB200_API int svb200_assemble_neu(svb200_ctx* ctx, int32_t iFa, const svb200_eqparams* eq, const svb200_dmnparams* dmn,
                        int32_t nDmn, const double* hg);

/* Host-assembled surface terms (Neumann/backflow faces): R(:,rows[k]) += R_add(:,k),
 * Val(:,slot(rows_k,cols_k)) += K_add(:,k). */
SVB200_API int svb200_add_host_contrib(svb200_ctx* ctx, int32_t dof, int32_t nR, const int32_t* rows, const double* R_add,
                            int32_t nK, const int32_t* krows, const int32_t* kcols, const double* K_add);

/* all_fun::commu(R): shared-node sum of the residual across partitions (no-op for one rank). */
SVB200_API int svb200_commu_R(svb200_ctx* ctx);

/* ustruct::ustruct_r (solver/ustruct.cpp:1742-1845, called from Integrator::step after commu(R), Integrator.cpp:135-137):
 * in the first Newton iteration of a time step (itr == 1, 1-based like eq.itr) R -= Kd (amg Ad - Yg(s..s+2)) / am with
 * amg = (gam - am)/(gam - 1); later iterations leave R unchanged.  Ad(3,nNo) is com_mod.Ad in INPUT node order; Kd is what the
 * last svb200_assemble(phys = USTRUCT) left on the device (download: SVB200_ARRAY_KD).
 * For an FSI equation with ustruct solids (eq->phys = SVB200_PHYS_FSI; fsi.cpp:243-262 fills Kd) only the nodes flagged by
 * svb200_set_node_flags take part, as all_fun::is_domain(..., phys_ustruct) decides in the reference (:1776-1793): pass the
 * membership in the ustruct domains there. */
SVB200_API int svb200_ustruct_r(svb200_ctx* ctx, const svb200_eqparams* eq, int32_t itr, const double* Ad);
/* com_mod.Ad(3,nNo) for the device-resident ustruct loop: uploaded once, then svb200_predictor scales it by (gam-1)/gam
 * (Integrator.cpp:627-628), svb200_ustruct_r(..., Ad = NULL) reads it and svb200_corrector(phys = USTRUCT) updates it together
 * with An, Yn, Dn (Integrator.cpp:826-846: dUl = Rd/am + R af gam dt/am, Ad -= dUl, Dn -= dUl gam dt). */
SVB200_API int svb200_set_ad(svb200_ctx* ctx, const double* Ad);
SVB200_API int svb200_get_ad(svb200_ctx* ctx, double* Ad);

/* fsils_solve: preconditions in place, runs the Krylov solver, writes the increment to R_out
 * (dof,nNo, INPUT node order; may be NULL to keep it on the device only). */
SVB200_API int svb200_solve(svb200_ctx* ctx, int32_t dof, int32_t ls_type, int32_t prec, const svb200_lsparams* ls,
                 int32_t nFaces, const int32_t* incL, const double* res,
                 double* R_out, svb200_lsresult* result);

/* ---- generalised-alpha time integration on the device (optional) ----------------
 * With these the Newton loop is device-resident: svb200_predictor once per time step, then per Newton
 * iteration svb200_initiator -> svb200_alloc/assemble/solve -> svb200_corrector; only norms cross PCIe.
 * The time-integration state is solutions.old (Ao,Yo,Do) and solutions.current (An,Yn,Dn), each (tDof,nNo)
 * (Code/Source/solver/SolutionStates.h:15-43); the intermediate state is the one svb200_set_state uploads. */
typedef enum { SVB200_SOL_OLD = 0, SVB200_SOL_CURRENT = 1, SVB200_SOL_INTERMEDIATE = 2 } svb200_sol;
/* svb200_eqtime.reserved bit: com_mod.sstEq (read_files.cpp:1486) for an FSI equation, i.e. its solids are ustruct domains: the
 * predictor / corrector then take the velocity-pressure branches (Integrator.cpp:626-630, 828-846) for the FSI rows, followed by the
 * usual copy to the mesh rows (:887-912).  A ustruct equation (phys = SVB200_PHYS_USTRUCT) always takes them. */
#define SVB200_EQTIME_SSTEQ 1
typedef struct {
  int32_t s, e;      /* eq.s, eq.e: first and last state row of the equation (inclusive) */
  int32_t phys;      /* svb200_phys */
  int32_t reserved;  /* SVB200_EQTIME_* bits */
  double af, am, gam, beta;
} svb200_eqtime;
/* Upload / download one solution triple; NULL pointers are skipped. */
SVB200_API int svb200_set_solution(svb200_ctx* ctx, int32_t tDof, int32_t which, const double* A, const double* Y, const double* D);
SVB200_API int svb200_get_solution(svb200_ctx* ctx, int32_t which, double* A, double* Y, double* D);
/* Integrator::predictor (solver/Integrator.cpp:393-643), state part: An = Ao (gam-1)/gam, Yn = Yo,
 * Dn = Do + Yn dt + An dt^2 (gam/2 - beta)/(gam-1) when dFlag (struct / mesh / FSI), else Dn = Do; for a ustruct equation
 * (phys = SVB200_PHYS_USTRUCT) Dn = Do and Ad is scaled instead (:626-630). */
SVB200_API int svb200_predictor(svb200_ctx* ctx, int32_t nEq, const svb200_eqtime* eqs, double dt, int32_t dFlag);
/* Integrator::initiator (solver/Integrator.cpp:662-750): Ag, Yg, Dg from old and current. */
SVB200_API int svb200_initiator(svb200_ctx* ctx, int32_t nEq, const svb200_eqtime* eqs);
/* Integrator::corrector (solver/Integrator.cpp:774-867): current -= increment left in R by svb200_solve.
 * mesh_s >= 0 with solid-node flags set (svb200_set_node_flags) also applies the FSI copy of :887-912. */
SVB200_API int svb200_corrector(svb200_ctx* ctx, const svb200_eqtime* eq, double dt, int32_t mesh_s);
SVB200_API int svb200_set_node_flags(svb200_ctx* ctx, const int32_t* is_solid_node);
/* set_bc::set_bc_dir (solver/set_bc.cpp:901-1067), the write: X(row0+i, nodes[k]) = val(i,k) for X = A, Y and/or D of
 * the CURRENT solution (NULL = leave).  The prescribed values are the host's (profiles, time functions). */
SVB200_API int svb200_set_dirichlet_rows(svb200_ctx* ctx, int32_t row0, int32_t nrow, int32_t n, const int32_t* nodes,
                              const double* valA, const double* valY, const double* valD);
/* set_bc::set_bc_dir, the part that follows the write for a velocity-pressure solid (ustruct; solver/set_bc.cpp:1046-1117), for
 * the listed nodes and the directions i of dir_mask (bit i; 7 = all, the "no eDrn set" case), j = eq.s + i:
 *   impD == 0:  Dn(j) = gam dt Yn(j) - (gam-1) dt Ad(i) + Do(j),  Ad(i) = Yn(j)
 *   impD != 0:  An(j) = (Yn(j) - Yo(j) + (gam-1) dt Ao(j)) / (gam dt),  Ad(i) = (Dn(j) - Do(j) + (gam-1) dt Ad(i)) / (gam dt)
 * Call after svb200_set_dirichlet_rows; Ad is the device-resident com_mod.Ad (svb200_set_ad). */
SVB200_API int svb200_dirichlet_ustruct(svb200_ctx* ctx, const svb200_eqtime* eq, double dt, int32_t n, const int32_t* nodes,
                             int32_t dir_mask, int32_t impD);
/* End of a time step: old = current (solver/main.cpp: solutions.old = solutions.current). */
SVB200_API int svb200_advance_time_step(svb200_ctx* ctx);

/* ---- debug / parity --------------------------------------------------------------------- */
/* R(dof,nNo) and Val(dof*dof,nnz) are returned in INPUT node order / input CSR slot order. */
SVB200_API int svb200_download(svb200_ctx* ctx, int32_t what, double* dst);
SVB200_API int svb200_upload(svb200_ctx* ctx, int32_t what, int32_t dof, const double* src);
/* The rows of R / W (dst(dof, n)) or the CSR rows of Val / Kd (dst(dof*dof | 12, sum of the row lengths), the rows of the
 * listed nodes one after the other, columns in the caller's order) of n INPUT-order nodes: what the parity gate of bench.py
 * compares with an oracle assembly of a sub-block of a 10 M-element partition without moving the 3 GB matrix. */
SVB200_API int svb200_download_rows(svb200_ctx* ctx, int32_t what, int32_t n, const int32_t* nodes, double* dst);

/* Stand-alone operators on the current device Val, for tests and the bench:
 * KU = K*U (+ halo sum), both (dof,nNo) host arrays in INPUT order. */
SVB200_API int svb200_spmv(svb200_ctx* ctx, int32_t dof, const double* U, double* KU);

/* Device-side timing of the last svb200_assemble / svb200_solve call in milliseconds
 * (CUDA events on the library's stream). */
SVB200_API int svb200_last_timing(svb200_ctx* ctx, double* assemble_ms, double* solve_ms);

/* Page-lock / unlock a caller buffer so that the H2D/D2H copies of svb200_set_state, svb200_solve and
 * svb200_download run at full PCIe rate (cudaHostRegister / cudaHostUnregister). */
SVB200_API int svb200_host_register(svb200_ctx* ctx, void* ptr, size_t bytes);
SVB200_API int svb200_host_unregister(svb200_ctx* ctx, void* ptr);

/* CUDA-event stopwatch on the library's stream: mark(0) ... mark(1), then elapsed gives the device
 * time in ms between the two marks (synchronises on the second). */
SVB200_API int svb200_timer_mark(svb200_ctx* ctx, int32_t which);
SVB200_API int svb200_timer_elapsed(svb200_ctx* ctx, double* ms);

/* Repeat the assembly kernel(s) / SpMV n times on resident data and return the average device
 * time per launch in ms (CUDA events on the launching stream); used by bench.py for the roofline. */
SVB200_API int svb200_bench_assemble(svb200_ctx* ctx, int32_t iM, const svb200_eqparams* eq, const svb200_dmnparams* dmn,
                          int32_t nDmn, int32_t reps, double* ms_per_launch);
SVB200_API int svb200_bench_spmv(svb200_ctx* ctx, int32_t dof, int32_t reps, double* ms_per_launch);
/* Rectangular-block products of the NS solver on the context's graph (fsils_spar_mul_vv / sv / vs / ss,
 * Code/Source/linear_solver/spar_mul.cpp:19-231) with a CALLER-supplied matrix K (R*C, nnz), U (C, nNo), KU (R, nNo) — host arrays,
 * single-partition contexts (internal = input order).  `variant` selects the lane mapping (-1: the default, 0: thread per (row, i));
 * svb200_spmv_rc_variants returns how many exist for a shape.  svb200_bench_spmv_rc times `reps` launches on resident zero data. */
SVB200_API int svb200_spmv_rc(svb200_ctx* ctx, int32_t R, int32_t C, int32_t variant, const double* K, const double* U, double* KU);
SVB200_API int svb200_spmv_rc_variants(int32_t R, int32_t C);
SVB200_API int svb200_bench_spmv_rc(svb200_ctx* ctx, int32_t R, int32_t C, int32_t variant, int32_t reps, double* ms_per_launch);
/* The Schur-complement operator of cgrad::schur (Code/Source/linear_solver/cgrad.cpp:77-84), SP = L p - Gt (G p), nsd = 3, with
 * caller-supplied L (nnz), Gt (3, nnz), P (nNo), GP (3, nNo): SP (nNo) and the fused <p, SP> over the owned rows.  variant -2: the
 * separate-array kernel (schur_sp_kernel), -1: default interleaved kernel, >= 0: a specific lane mapping.  Host arrays, single partition. */
SVB200_API int svb200_schur_sp(svb200_ctx* ctx, int32_t variant, const double* L, const double* Gt, const double* P, const double* GP,
                               double* SP, double* p_dot_sp);
SVB200_API int svb200_schur_sp_variants(void);
SVB200_API int svb200_bench_schur_sp(svb200_ctx* ctx, int32_t variant, int32_t reps, double* ms_per_launch);
/* Measured FP64 FMA peak (independent DFMA chains) in TFLOP/s. */
SVB200_API int svb200_measure_fp64_peak(svb200_ctx* ctx, double* tflops);
/* Number of CUDA kernels this library has launched on ctx since creation. */
SVB200_API int64_t svb200_launch_count(svb200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SVB200_H */
