// svb200_internal.h — device-side state behind the opaque svb200_ctx (include/svb200.h).
//
// Data layout in HBM (all FP64 / int32, column-major like the reference):
//   * nodes are stored in FSILS order (lhs.map of linear_solver/lhs.cpp:215-218): interface nodes
//     shared with lower ranks first, interior nodes, interface nodes shared with higher ranks last,
//     so "owned" dots run over [0,mynNo) and halo packs touch two contiguous-ish ends;
//   * the block-CSR matrix (rowPtr/colPtr/Val) is stored with rows in that same order, the columns
//     of a row kept in the input order, one dof x dof block = dof*dof contiguous doubles
//     (row-major inside the block, Code/Source/solver/FsilsLinearAlgebra.cpp:35);
//   * nodal state (x, Ag, Yg, Dg, Bf, R) is (rows, nNo), node-major AoS exactly like the reference.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/svb200.h"

namespace svb {

constexpr int MAX_DMN = 8;
constexpr int MAX_ENON = 8;
constexpr int MAX_NG = 8;
// the general fluid kernel reads its tables from a device buffer and also takes the quadratic elements
constexpr int MAX_ENON_ANY = 27;
constexpr int MAX_NG_ANY = 27;
constexpr int ASM_GROUP = 128;   // elements per CTA of the grouped (pre-reduced) scatter, group_sched.cu

// Pre-reduction plan of one mesh for the grouped scatter (group_sched.cu).
struct GroupSched {
  int nGrp = 0;
  long long total = 0;               // sum of nuniq
  int* d_nuniq = nullptr;            // (nGrp) distinct targets of the group
  int* d_uptr = nullptr;             // (nGrp+1) offsets into d_uent
  int2* d_uent = nullptr;            // {target, start | count << 16}
  int* d_upartner = nullptr;         // tangent plan: transposed slot of an edge, -1 for a diagonal block
  unsigned short* d_contrib = nullptr;   // (nGrp, ASM_GROUP*PER_EL) contribution ids, sorted by target
};

struct Mesh {
  int eNoN = 0, nEl = 0, nG = 0, nFn = 0;
  int* d_IEN = nullptr;      // (eNoN, nEl) internal node ids
  int* d_eId = nullptr;      // (nEl) or null
  double* d_fN = nullptr;    // (3*nFn, nEl) or null
  int* d_slot = nullptr;     // (eNoN*eNoN, nEl): CSR slot of pair (a,b) = entry a*eNoN+b
  int* d_color_perm = nullptr;   // element ids sorted by colour (coloured scatter)
  GroupSched schedK, schedR;     // grouped scatter plans: tangent blocks / residual rows (TET4 meshes)
  std::vector<int> color_off;    // offsets into d_color_perm per colour
  int* d_gcolor_perm = nullptr;  // TET4: ids of the 128-element groups of the grouped scatter, sorted by GROUP colour
  std::vector<int> gcolor_off;   // offsets into d_gcolor_perm per group colour (groups of one colour share no node)
  std::vector<int> grp_node_need;    // TET4: running max over groups 0..g of (highest CALLER node id of the group) + 1: the nodal state
                                     // of the caller's nodes [0, grp_node_need[g]) must be on the device before group g runs
  std::vector<int> grp_node_done;    // suffix min over groups g.. of the lowest caller node id: the residual rows of the caller's nodes
                                     // [0, grp_node_done[g]) are final once the groups before g have run (svb200_assemble_host)
  std::vector<long long> grp_need;   // TET4: running max over groups 0..g of (highest CSR slot the group adds to) + 1 — how far Val
                                     // must be zeroed before group g may run (overlapped zeroing, svb200_api.cu run_assemble)
  std::vector<double> w, N, Nx;  // host copies of the reference-element tables
  std::vector<double> Nxx;       // (6,eNoN,nG) second parametric derivatives (svb200_set_mesh_nxx), empty = all zero
  double* d_gtab = nullptr;      // tables in the layout of assemble_fluid_gen.cu
  // Taylor-Hood function spaces (svb200_set_mesh_thood): pressure space with th_eNoNq nodes and its own rule of th_nG2 points
  int th_eNoNq = 0, th_nG2 = 0, th_lShpFq = 0;
  double* d_thtab = nullptr;     // tables in the layout of assemble_fluid_thood.cu
  // URIS split launch (TET4): elements with a node inside a valve's or scaffold's thickness, as mask and as compact list
  mutable unsigned char* d_uris_mask = nullptr;
  mutable int* d_uris_list = nullptr;
  mutable int n_uris_el = 0;
  mutable long long uris_version = -1;
  mutable bool jac_checked = false;   // element Jacobians verified for the current reference coordinates (fluid, fixed mesh)
  bool set = false;
};

struct Face {
  int bGrp = 0, dof = 0, nNo = 0, shared = 0;
  bool set = false;
  int* d_glob = nullptr;     // internal node ids
  double* d_val = nullptr;   // (dof, nNo)
  double* d_valM = nullptr;  // (dof, nNo): val * W (precond_diag)
  double nS = 0.0;           // sum of val^2 over owned nodes (fsils_bc_create)
  // capping surface of a coupled face (fils_struct.hpp:131-143): cap nodes on this partition, their normal integrals
  int cap_n = 0;
  bool has_cap = false;
  int* d_cap_glob = nullptr;     // internal node ids
  double* d_cap_val = nullptr;   // (dof, cap_n)
  double* d_cap_valM = nullptr;  // (dof, cap_n): cap_val * W (precond.cpp:229-237)
  // per-solve flags (fsils_solve)
  bool incFlag = true, coupledFlag = false;
  double res = 0.0;
};

// Boundary face (faceType) for the surface integrals of assemble_bnd.cu.
struct BFace {
  int iM = -1, eNoNb = 0, nElb = 0, nGb = 0;
  int* d_IENb = nullptr;     // (eNoNb, nElb) internal node ids
  int* d_gE = nullptr;       // (nElb) parent element index in mesh iM
  std::vector<double> w, N, Nx;
  bool set = false;
};

struct Neighbor {
  int rank = 0, n = 0;
  int* d_ptr = nullptr;       // internal node ids shared with that rank (same order on both sides)
  std::vector<int> h_ptr;     // host copy (the fused exchange tables of comm.cu are built from it)
  double* d_send = nullptr;   // (maxdof, n)
  double* d_recv = nullptr;
};

// Kernel argument block of the fluid assembly kernels (passed by value as __grid_constant__).
struct FluidDmn {
  double rho, f[3], Kd;
  double mu_i, mu_o, lam, a, n;
  int viscType, Id, isFluid, pad;
};

struct FluidArgs {
  const int* IEN;
  const int* eId;
  const int* slot;
  const int* perm;      // optional element permutation (coloured scatter) or null
  // grouped scatter (group_sched.cu): tangent (K) and residual (R) plans
  const int* kU_ptr; const int2* kU_ent; const int* kU_partner; const unsigned short* kContrib;
  const int* rU_ptr; const int2* rU_ent; const unsigned short* rContrib;
  const double* x;
  const double* Ag;
  const double* Yg;
  const double* Bf;
  const double* Dg;     // displacement state; dofs 4..6 = mesh displacement (FSI / ALE geometry)
  double* R;
  double* Val;
  int e0, e1;           // element range [e0,e1) (indices into perm when perm != null)
  int tDof, mvMsh, nDmn, atomic;
  int ale, bfZero;      // ale: element geometry is x + Dg(4..6) (fsi::construct_fsi, fsi.cpp:140-146); bfZero: Bf is all zeros
  int* err;             // device error word: 1 + index of an element with a zero Jacobian (0 = none)
  const int* gperm;     // grouped kernel: CTA blockIdx.x works on group gperm[g0 + blockIdx.x] (null: group blockIdx.x)
  int g0, nGrpLaunch;   // first entry / number of entries of gperm in this launch (deterministic mode: one group colour)
  double dt, af, am, gam;
  double w[MAX_NG];
  double N[MAX_NG][MAX_ENON];        // N[g][a]
  double Nxi[MAX_NG][MAX_ENON][3];   // Nxi[g][a][k] = d N_a / d xi_k at Gauss point g
  FluidDmn dmn[MAX_DMN];
  // URIS valves (svb200_set_uris): per node and valve |sdf|, |scaffold udf|, valve velocity = 5 doubles, node-major
  const double* uris;
  int nUris;
  svb200_uris urisP[SVB200_MAX_URIS];
  // optional element mask: an element takes part iff emask == nullptr || emask[e] == emask_val (URIS split launch: the closed-form
  // TET4 kernel runs the elements away from the valves, the per-Gauss-point kernel the band around them)
  const unsigned char* emask;
  int emask_val;
};

}  // namespace svb

namespace svb {
// plan of the open resistive immersed surfaces (ris.cu)
struct RisPlan {
  int nNode = 0, nRApply = 0, buf_dof = 0;
  long long nEnt = 0, nApply = 0;
  int *d_ent = nullptr, *d_src = nullptr, *d_dst = nullptr, *d_node = nullptr, *d_rsrc = nullptr, *d_rdst = nullptr;
  double *d_S = nullptr, *d_C = nullptr, *d_RS = nullptr, *d_RC = nullptr;
};
}  // namespace svb

struct svb200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t tm0 = nullptr, tm1 = nullptr;   // svb200_timer_*
  int64_t launches = 0;

  // graph
  int nNo = 0, nnz = 0, mynNo = 0;
  bool has_map = false;
  std::vector<int> h_map;        // input -> internal node id
  std::vector<int> h_rowPtr_in;  // input CSR (kept for slot translation on download)
  std::vector<int> h_rowPtr;     // internal CSR row pointer
  std::vector<int> h_colPtr;     // internal CSR column ids (host copy: slot look-up of svb200_add_host_contrib)
  int* d_map = nullptr;
  int* d_rowPtr_in = nullptr;    // (nNo+1) caller-order row pointer (block up/downloads when a node map is active)
  int* d_rowPtr = nullptr;       // (nNo+1) internal
  int* d_colPtr = nullptr;       // (nnz) internal column ids
  int* d_diagPtr = nullptr;      // (nNo)
  int* d_tslot = nullptr;        // (nnz) slot of the transposed entry (NS solver: Gt), built on first use

  // coordinates and state (internal order)
  int tDof = 0;
  double* d_x = nullptr;
  double* d_Ag = nullptr;
  double* d_Yg = nullptr;
  double* d_Dg = nullptr;
  double* d_Do = nullptr;        // old displacement (mesh-motion equation; solutions.old)
  svb::RisPlan ris;              // open fitted RIS surfaces (svb200_set_ris)
  double* d_uris = nullptr;      // URIS nodal fields (5 nUris, nNo): per valve |sdf|, |scaffold udf|, valve velocity (svb200_set_uris)
  int nUris = 0;
  long long uris_version = 0;    // bumped by every svb200_set_uris: the per-mesh element masks follow it
  svb200_uris urisP[SVB200_MAX_URIS] = {};
  double* d_Ya = nullptr;        // nodal active tensions (3, nNo): Ya_f, Ya_s, Ya_n (svb200_set_active_tension)
  bool ya_sn_positive = false;
  double* d_pS0 = nullptr;       // nodal prestress com_mod.pS0 (6, nNo) (svb200_set_prestress)
  double* d_pSn = nullptr;       // pstEq accumulators: pSn (6, nNo) followed by pSa (nNo)   // any Ya_s or Ya_n > 0 (only the Guccione / HO / HO-ma models accept that)
  double* d_Ao = nullptr; double* d_Yo = nullptr;                          // solutions.old
  double* d_An = nullptr; double* d_Yn = nullptr; double* d_Dn = nullptr;  // solutions.current
  int* d_nodeflag = nullptr;     // per node: belongs to a solid domain (FSI corrector)
  int* d_err = nullptr;          // element-loop error word (zero Jacobian), see FluidArgs::err
  double* d_Bf = nullptr;
  bool bf_set = false;           // a body-force array was uploaded (else d_Bf is all zeros and the TET4 fluid kernel skips its gather)
  double* d_stage = nullptr;     // staging buffer for permuted uploads/downloads
  size_t stage_bytes = 0;

  // linear system
  int dof = 0;
  double* d_R = nullptr;
  double* d_Val = nullptr;
  size_t R_cap = 0, Val_cap = 0;   // capacities in doubles
  bool val_zero_pending = false;   // svb200_alloc deferred the zeroing of Val: the next consumer zeroes it (overlapped with the first
                                   // chunk of the TET4 fluid kernel, or in full before anything else touches Val)
  cudaStream_t zstream = nullptr;  // stream of the overlapped zeroing / of the H2D copies of svb200_assemble_host
  cudaStream_t dstream = nullptr;  // stream of the D2H copies of svb200_assemble_host
  cudaEvent_t zev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t hev[4] = {nullptr, nullptr, nullptr, nullptr};   // svb200_assemble_host timeline: upload, kernels, halo, streamed D2H done
  double host_stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};                     // ... in ms after the start of the call, and the end of the call
  cudaEvent_t pev[2][16] = {};     // svb200_assemble_host: [0][c] chunk c uploaded, [1][c] chunk c assembled
  // svb200_assemble_host on a partitioned mesh: the interface rows (changed by the shared-node sum after they were streamed back)
  std::vector<int> h_shared_caller;   // caller node ids of the interface rows
  int* d_shared_rows = nullptr;       // their internal ids
  long long* d_shared_off = nullptr;  // 0, 1, 2, ... (offsets for the row gather)
  int* d_shared_caller = nullptr;     // h_shared_caller on the device (zero-copy scatter into the caller's residual)
  double* h_shared_buf = nullptr;     // pinned (dof, nShared)
  bool shared_built = false;
  double* d_W = nullptr;           // (dof,nNo) preconditioner scaling
  double* d_Kd = nullptr;          // (12,nnz) displacement tangent of the ustruct equation (com_mod.Kd), assemble_ustruct.cu
  double* d_Ad = nullptr;          // (3,nNo) com_mod.Ad: time derivative of the displacement (ustruct)
  double* d_Rd = nullptr;          // (3,nNo) com_mod.Rd as svb200_ustruct_r left it (read by the ustruct corrector)
  size_t W_cap = 0;

  std::vector<svb::Mesh> mesh;
  std::vector<svb::Face> face;
  std::vector<svb::BFace> bface;
  double* d_hg = nullptr;          // (nNo) nodal traction of svb200_assemble_neu

  // Krylov workspace
  double* d_work = nullptr;
  size_t work_cap = 0;
  double* d_red = nullptr;         // partial-reduction scratch
  double* h_pinned = nullptr;      // pinned host scratch for scalar read-back
  double* d_cg = nullptr;          // device scalars of the Schur-complement CG (8 doubles)
  double* h_cg = nullptr;          // pinned: 2 status slots x 8 doubles + 8 for the initial values
  cudaEvent_t ev_cg[2] = {nullptr, nullptr};
  size_t red_cap = 0;

  // multi-GPU
  int nranks = 1, rank = 0;
  void* nccl_comm = nullptr;
  void* p2p = nullptr;             // peer-memory transport state (comm.cu), null = NCCL transport
  std::vector<svb::Neighbor> neigh;

  double last_assemble_ms = 0.0, last_solve_ms = 0.0;

  // lhsa scratch (svb200_lhsa_*)
  int lhsa_nNo = 0;
  unsigned long long* d_lhsa_keys = nullptr;
  size_t lhsa_n = 0, lhsa_cap = 0;
  std::vector<int> lhsa_rowPtr, lhsa_colPtr;
};

namespace svb {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int flush_val_zero(svb200_ctx* ctx);     // svb200_api.cu: zero Val now if svb200_alloc deferred it

#define SVB_CUDA(call)                                                             \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) return svb::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

#define SVB_REQUIRE(cond, msg)                 \
  do {                                         \
    if (!(cond)) {                             \
      svb::set_error(std::string("svb200: ") + (msg)); \
      return SVB200_ERR_INVALID;               \
    }                                          \
  } while (0)

// assemble_fluid.cu
int launch_assemble_fluid(svb200_ctx* ctx, const Mesh& m, const FluidArgs& args);
int launch_tet4_jacobian_check(svb200_ctx* ctx, const Mesh& m, const FluidArgs& args);
// assemble_struct.cu
int upload_fluid_gen_tables(svb200_ctx* ctx, Mesh& m);
int build_uris_element_mask(svb200_ctx* ctx, const Mesh& m);
int upload_thood_tables(svb200_ctx* ctx, Mesh& m, int eNoNq, int nG2, const double* Nq1, const double* Nqxi1, const double* w2,
                        const double* Nw2, const double* Nwxi2, const double* Nq2, const double* Nqxi2);
int run_assemble_fluid_thood(svb200_ctx* ctx, const Mesh& m, const FluidArgs& F);
int run_thood_val_rc(svb200_ctx* ctx);
int run_assemble_fluid_gen(svb200_ctx* ctx, const Mesh& m, const FluidArgs& F);
int run_assemble_struct(svb200_ctx* ctx, const Mesh& m, const svb200_eqparams* eq, const svb200_dmnparams* dmn, int nDmn);
int run_assemble_mesh(svb200_ctx* ctx, const Mesh& m, const svb200_eqparams* eq, const svb200_dmnparams* dmn, int nDmn);
// assemble_ustruct.cu
int run_assemble_ustruct(svb200_ctx* ctx, const Mesh& m, const svb200_eqparams* eq, const svb200_dmnparams* dmn, int nDmn);
// com_mod.sstEq for this equation's gen-alpha update: a ustruct equation, or an FSI equation whose solids are ustruct
static inline bool eqtime_is_sst(const svb200_eqtime& q)
{
  return q.phys == SVB200_PHYS_USTRUCT || (q.phys == SVB200_PHYS_FSI && (q.reserved & SVB200_EQTIME_SSTEQ));
}
bool ris_active(const svb200_ctx* ctx);
int ris_build_plan(svb200_ctx* ctx, int nProj, const int* nMap, const int* maps, const int* closed);
int ris_begin(svb200_ctx* ctx);
int ris_end(svb200_ctx* ctx);
int run_ustruct_r(svb200_ctx* ctx, const svb200_eqparams* eq, int itr, const double* d_Ad);
// assemble_heat.cu
int run_assemble_heat(svb200_ctx* ctx, const Mesh& m, const svb200_eqparams* eq, const svb200_dmnparams* dmn, int nDmn);
// assemble_bnd.cu
int run_assemble_neu(svb200_ctx* ctx, const BFace& f, const svb200_eqparams* eq, const svb200_dmnparams* dmn, int nDmn,
                     const double* d_hg);
// genalpha.cu
int launch_predictor(svb200_ctx* ctx, int nEq, const svb200_eqtime* eqs, double dt, int dFlag);
int launch_initiator(svb200_ctx* ctx, int nEq, const svb200_eqtime* eqs);
int launch_corrector(svb200_ctx* ctx, const svb200_eqtime* eq, double dt, int mesh_s, const int* d_flag);
int launch_set_rows(svb200_ctx* ctx, int row0, int nrow, int n, const int* d_nodes, const double* d_val, double* dst);
int launch_dirichlet_ustruct(svb200_ctx* ctx, const svb200_eqtime* eq, double dt, int n, const int* d_nodes, int dir_mask, int impD);
// group_sched.cu
int build_group_schedules(svb200_ctx* ctx, Mesh& m);
int build_group_slot_need(svb200_ctx* ctx, Mesh& m);
void free_group_sched(GroupSched& S);
// graph_kernels.cu
int launch_build_slot_map(svb200_ctx* ctx, Mesh& m);
int launch_find_diag(svb200_ctx* ctx);
int launch_permute_cols(svb200_ctx* ctx, int rows, int n, const int* d_map, const double* src, double* dst, bool inverse,
                        cudaStream_t stream = nullptr);
int launch_permute_row_blocks(svb200_ctx* ctx, int a0, int a1, int d2, const int* d_rowPtr_in, double* internal, double* staged,
                              bool to_caller);
int launch_gather_row_blocks(svb200_ctx* ctx, int n, int d2, const int* d_rows, bool csr, const long long* d_off, const double* src,
                             double* dst);
// fsils_kernels.cu
int launch_spmv(svb200_ctx* ctx, int dof, const double* Val, const double* U, double* KU);
int launch_dots(svb200_ctx* ctx, int n, int nvec, const double* const* d_vec_list, const double* base, size_t stride,
                const double* v, double* d_out);
// fsils_solvers.cu
int fsils_solve_device(svb200_ctx* ctx, int dof, int ls_type, int prec, const svb200_lsparams* ls, int nFaces,
                       const int* incL, const double* res, svb200_lsresult* result);
int fp64_peak(svb200_ctx* ctx, double* tflops);

}  // namespace svb
