// assemble_ustruct.cu — element loop + scatter of the mixed velocity-pressure solid.
//
// Replaces ustruct::construct_usolid, ustruct_3d_m, ustruct_3d_c and ustruct_do_assem
// (Code/Source/solver/ustruct.cpp:203-400, 1165-1591, 629-871, 1595-1737) for equal-order (VMS) TET4 and HEX8 meshes,
// and ustruct::ustruct_r (:1742-1845).  Outputs: R(4,nNo), Val(16,nnz) and the displacement tangent Kd(12,nnz).
//
// Mapping (like assemble_struct.cu): ENON lanes per element, two phases.  Phase A: lane g evaluates Gauss point g ONCE
// per element (gnn, F, F^-1, compute_pk2cc without the volumetric part, g_vol_pen, compute_tau, the strong residuals)
// and leaves a UGP (110 doubles; + two ViscGP sets in the solid-viscosity instantiation) in shared memory.  Phase B: lane a owns row a of the element matrices; column nodes
// are taken NB at a time so that the 28 NB accumulators (16 of lK + 12 of lKd per block) stay in registers, the Gauss
// loop is inside (the column-node terms of a pass are published once per (g, b) through shared memory), and the finished
// blocks go out through a per-warp transposition tile as coalesced adds.
#include <cstdlib>
#include <vector>
#define SVB_UTET_ROLL_B 1      // the (a, b) loops of the closed-form TET4 kernel stay rolled
#define SVB_UTET_ROLL_A 1
#include "svb200_internal.h"
#include "ustruct_elem.cuh"
#include "fsils_kernels.h"

namespace svb {

int fill_solid_extras(svb200_ctx* ctx, const svb200_dmnparams& p, StructDmn& o, CannRow* table, int* used);   // assemble_struct.cu

struct UstructArgs {
  const int* IEN;
  const int* eId;
  const int* slot;
  const int* perm;
  const double* fN;
  const double* x;
  const double* Ag;
  const double* Yg;
  const double* Dg;
  const double* Bf;
  const double* Ya;      // nodal active tensions (3, nNo) or nullptr (ustruct.cpp:294-298)
  double* R;
  double* Val;
  double* Kd;
  int* err;
  int e0, e1;
  int tDof, s, nFn, nDmn, nG, pad;
  double dt, af, am, gam;
  double w[MAX_NG];
  double N[MAX_NG][MAX_ENON];
  double Nxi[MAX_NG][MAX_ENON][3];
  UstructDmn dmn[MAX_DMN];
  int active[MAX_DMN];
  CannRow cann[MAX_CANN_ROWS];
};

constexpr int USTRUCT_THREADS = 64;
// Gauss-point record of the VISC instantiation: the two ViscGP sets of ustruct_gauss_point follow the UGP
struct UGPV {
  UGP g;
  ViscGP gu, gv;
};
// per-element stride in doubles: the elements of a warp (4 for HEX8, 8 for TET4) read the same UGP field at the same
// time, so the stride is padded to 4 (HEX8) / 2 (TET4) mod 16 to spread them over the banks
// What phase B needs of a column node b at one Gauss point: published once per (g, b) by lane g instead of being
// recomputed by every row lane a (DBm_b alone is 108 FMAs).
struct UCol {
  UNode n;
  double DBm[6][3];
};
constexpr int USTRUCT_NB = 1;           // column nodes per pass of phase B (2: 56 accumulators, spills)
constexpr int UCOL_LD = (int)(sizeof(UCol) / sizeof(double));
// Shared memory of one warp: the Gauss-point records of its 32/ENON elements (element stride padded to 4 (HEX8) / 2 (TET4)
// mod 16 doubles: the elements of a warp read the same field at the same time), then ONE area that holds the column-node
// records of the current pass and, once the pass has consumed them, the 32 x 29 transposition tile of the scatter.
__host__ __device__ constexpr int ustruct_el_ld(int enon, int ld)
{
  const int n = enon * ld, want = (enon == 8) ? 4 : 2;
  return n + ((want - (n % 16)) + 16) % 16;
}
__host__ __device__ constexpr int ustruct_warp_ld(int enon, int ld)
{
  const int epw = 32 / enon;
  const int cols = epw * enon * USTRUCT_NB * UCOL_LD, tile = 32 * 29;
  return epw * ustruct_el_ld(enon, ld) + (cols > tile ? cols : tile);
}
template <bool VISC> struct UGPSel { using type = UGP; };
template <> struct UGPSel<true> { using type = UGPV; };
__device__ __forceinline__ UGP& ugp_of(UGP& r) { return r; }
__device__ __forceinline__ UGP& ugp_of(UGPV& r) { return r.g; }

template <bool ATOMIC>
__device__ __forceinline__ void uadd(double* p, double v)
{
  if (ATOMIC) asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
  else *p += v;
}

template <int ENON, bool ATOMIC, bool VISC, bool CANN>
__global__ void __launch_bounds__(USTRUCT_THREADS)
assemble_ustruct_kernel(const __grid_constant__ UstructArgs P)
{
  using GP = typename UGPSel<VISC>::type;
  constexpr int EPW = 32 / ENON;
  constexpr int NB = USTRUCT_NB;
  constexpr int GP_LDD = (int)(sizeof(GP) / sizeof(double));
  constexpr int EL_LD = ustruct_el_ld(ENON, GP_LDD);
  constexpr int WARP_LD = ustruct_warp_ld(ENON, GP_LDD);
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a = lane % ENON, el = lane / ENON;
  double* wbase = sm + (size_t)warp * WARP_LD;
  GP* gp = reinterpret_cast<GP*>(wbase + (size_t)el * EL_LD);
  double* tile = wbase + (size_t)EPW * EL_LD;                 // aliases the column-node records (see above)
  UCol (*col)[NB] = reinterpret_cast<UCol (*)[NB]>(tile + (size_t)el * ENON * NB * UCOL_LD);

  const long long idx = (long long)P.e0 + ((long long)blockIdx.x * (USTRUCT_THREADS / 32) + warp) * EPW + el;
  bool active = idx < P.e1;
  int e = 0;
  if (active) e = P.perm ? P.perm[idx] : (int)idx;
  int iD = 0;
  if (active) {
    for (int d = 0; d < P.nDmn; d++) {
      iD = d;
      if (P.dmn[d].st.Id == -1) break;
      if (P.eId != nullptr && ((P.eId[e] >> P.dmn[d].st.Id) & 1)) break;
    }
    if (!P.active[iD]) active = false;
  }
  const UstructDmn& dm = P.dmn[iD];
  const double af = P.af * P.gam * P.dt, am = P.am;

  // ---- phase A: lane g = a evaluates Gauss point g (nG == ENON) ------------------------------------------------
  int node[ENON];
  if (active) {
    double xl[ENON][3], ql[ENON][3], vl[ENON][3], dl[ENON][3], pl[ENON], pdl[ENON];
#pragma unroll
    for (int b = 0; b < ENON; b++) {
      node[b] = P.IEN[(size_t)e * ENON + b];
      const size_t n = (size_t)node[b];
      const double* A = P.Ag + (size_t)P.tDof * n + P.s;
      const double* Y = P.Yg + (size_t)P.tDof * n + P.s;
      const double* D = P.Dg + (size_t)P.tDof * n + P.s;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        xl[b][i] = __ldg(P.x + 3 * n + i);
        ql[b][i] = __ldg(A + i) - __ldg(P.Bf + 3 * n + i);
        vl[b][i] = __ldg(Y + i);
        dl[b][i] = __ldg(D + i);
      }
      pl[b] = __ldg(Y + 3);
      pdl[b] = __ldg(A + 3);
    }
    double fN[2][3] = {{0, 0, 0}, {0, 0, 0}};
    if (P.fN != nullptr)
      for (int k = 0; k < P.nFn && k < 2; k++)
#pragma unroll
        for (int i = 0; i < 3; i++) fN[k][i] = __ldg(P.fN + (size_t)3 * P.nFn * e + 3 * k + i);
    const int g = a;
    // active tensions at the Gauss point (ustruct.cpp:955-957, 1265-1267)
    double ya[3] = {0.0, 0.0, 0.0};
    const bool act = CANN && (P.Ya != nullptr) && dm.st.active;
    if (act) {
#pragma unroll
      for (int b = 0; b < ENON; b++)
#pragma unroll
        for (int i = 0; i < 3; i++) ya[i] += P.N[g][b] * __ldg(P.Ya + 3 * (size_t)node[b] + i);
    }
    const double* yap = act ? ya : nullptr;
    UGP& q = ugp_of(gp[g]);               // written in place: a local copy would cost 880 B of stack per thread
    if constexpr (VISC) {
      // a domain of this launch without viscosity leaves c = 0 sets: its viscous blocks vanish
      gp[g].gu.c = 0.0; gp[g].gv.c = 0.0;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
          gp[g].gu.T[i][j] = gp[g].gu.A[i][j] = gp[g].gu.B[i][j] = gp[g].gu.M[i][j] = 0.0;
          gp[g].gv.T[i][j] = gp[g].gv.A[i][j] = gp[g].gv.B[i][j] = gp[g].gv.M[i][j] = 0.0;
        }
      ustruct_gauss_point<ENON, CANN>(dm, P.dt, P.af, P.am, P.gam, P.w[g], P.N[g], P.Nxi[g], xl, ql, vl, dl, pl, pdl, fN, q, &gp[g].gu, &gp[g].gv, yap, P.cann, P.nFn);
    } else {
      ustruct_gauss_point<ENON, CANN>(dm, P.dt, P.af, P.am, P.gam, P.w[g], P.N[g], P.Nxi[g], xl, ql, vl, dl, pl, pdl, fN, q, nullptr, nullptr, yap, P.cann, P.nFn);
    }
    // construct_usolid throws when utils::is_zero(Jac) (ustruct.cpp:312-314); q.w = w_g * Jac
    if (fabs(q.w) < fabs(P.w[g]) * 10.0 * 2.220446049250313e-16 * 2.220446049250313e-16) atomicMax(P.err, e + 1);
  }
  __syncwarp();

  // ---- phase B: lane a owns row a ------------------------------------------------------------------------------
  // Lanes without an element stay in the loops: the scatter below is a whole-warp operation.
  if (active) {
    int na = 0;
#pragma unroll
    for (int b = 0; b < ENON; b++)
      if (b == a) na = node[b];
    double lR[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
    for (int g = 0; g < ENON; g++) {
      UNode A;
      ustruct_node(ugp_of(gp[g]), P.N[g][a], P.Nxi[g][a], A);
      ustruct_resid(ugp_of(gp[g]), A, lR);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) uadd<ATOMIC>(P.R + (size_t)4 * na + i, lR[i]);
  }
  // Per-warp transposition tile of the scatter: every lane deposits its finished 16 + 12 doubles, then the warp adds them
  // with consecutive lanes on consecutive doubles (a half-warp = one 128-byte Val block) — lane-strided REDs run at a
  // third of the coalesced rate or worse (profiles/r1_microbench_fp64_red.txt: 174 vs 554 G adds/s, 42 vs 285 from DRAM).
  constexpr int TILE_LD = 29;
  const int* sl = P.slot + (size_t)e * ENON * ENON + a * ENON;
#pragma unroll 1
  for (int b0 = 0; b0 < ENON; b0 += NB) {
    // lane g = a publishes the column-node data of b0 .. b0+NB-1 at Gauss point g
    __syncwarp();
    if (active) {
      const UGP& q = ugp_of(gp[a]);
#pragma unroll
      for (int k = 0; k < NB; k++) {
        UCol& c = col[a][k];
        double Bmb[6][3];
        ustruct_node(q, P.N[a][b0 + k], P.Nxi[a][b0 + k], c.n);
        make_Bm(c.n.Nx, q.F, Bmb);
        make_DBm(q.Dm, Bmb, c.DBm);
      }
    }
    __syncwarp();
    double K[NB][16], Kd[NB][12];
#pragma unroll
    for (int k = 0; k < NB; k++) {
#pragma unroll
      for (int i = 0; i < 16; i++) K[k][i] = 0.0;
#pragma unroll
      for (int i = 0; i < 12; i++) Kd[k][i] = 0.0;
    }
    if (active) {
#pragma unroll 1
      for (int g = 0; g < ENON; g++) {
        const UGP& q = ugp_of(gp[g]);
        UNode A;
        double Bma[6][3];
        ustruct_node(q, P.N[g][a], P.Nxi[g][a], A);
        make_Bm(A.Nx, q.F, Bma);
#pragma unroll
        for (int k = 0; k < NB; k++) {
          const UCol& c = col[g][k];
          ustruct_block(q, af, am, A, c.n, Bma, c.DBm, K[k], Kd[k]);
          if constexpr (VISC)
            if (dm.st.viscType != SVB200_SOLID_VISC_NONE)
              ustruct_visc_block(dm.st.viscType, q, af, am, gp[g].gu, gp[g].gv, A.Nx, c.n.Nx, K[k], Kd[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NB; k++) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 16; i++) tile[lane * TILE_LD + i] = K[k][i];
#pragma unroll
      for (int i = 0; i < 12; i++) tile[lane * TILE_LD + 16 + i] = Kd[k][i];
      const int myslot = active ? sl[b0 + k] : -1;
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 16; it++) {             // 32 Val blocks x 16 doubles
        const int p = it * 32 + lane, src = p >> 4, i = p & 15;
        const int s_ = __shfl_sync(0xffffffffu, myslot, src);
        if (s_ >= 0) uadd<ATOMIC>(P.Val + (size_t)16 * s_ + i, tile[src * TILE_LD + i]);
      }
#pragma unroll
      for (int it = 0; it < 12; it++) {             // 32 Kd blocks x 12 doubles
        const int p = it * 32 + lane, src = p / 12, i = p - 12 * src;
        const int s_ = __shfl_sync(0xffffffffu, myslot, src);
        if (s_ >= 0) uadd<ATOMIC>(P.Kd + (size_t)12 * s_ + i, tile[src * TILE_LD + 16 + i]);
      }
    }
  }
}

// ---- linear tetrahedra: one thread per element, Gauss sums in closed form (ustruct_elem.cuh: ustruct_tet4_*) ---------------
// One compute_pk2cc, four Dm Bm_b and the scalars of the four Gauss points per element; the 16 blocks are assembled one at a
// time from them and leave through the per-warp transposition tile.  Dm lives in shared memory (one column per thread).  Without solid viscosity.
constexpr int UTET_THREADS = 128;
constexpr size_t UTET_SMEM = sizeof(double) * ((size_t)(UTET_THREADS / 32) * 32 * 29 + 36 * UTET_THREADS);

template <bool ATOMIC, bool CANN>
__global__ void __launch_bounds__(UTET_THREADS)
assemble_ustruct_tet4_kernel(const __grid_constant__ UstructArgs P)
{
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* tile = sm + (size_t)warp * 32 * 29;
  double* sDm = sm + (size_t)(UTET_THREADS / 32) * 32 * 29 + threadIdx.x;      // element r*6+c at sDm[(6 r + c) * UTET_THREADS]
  const long long idx = (long long)P.e0 + (long long)blockIdx.x * UTET_THREADS + threadIdx.x;
  bool active = idx < P.e1;
  int e = 0;
  if (active) e = P.perm ? P.perm[idx] : (int)idx;
  int iD = 0;
  if (active) {
    for (int d = 0; d < P.nDmn; d++) {
      iD = d;
      if (P.dmn[d].st.Id == -1) break;
      if (P.eId != nullptr && ((P.eId[e] >> P.dmn[d].st.Id) & 1)) break;
    }
    if (!P.active[iD]) active = false;
  }
  const UstructDmn& dm = P.dmn[iD];
  const double af = P.af * P.gam * P.dt, am = P.am;
  int sl[16];
#pragma unroll
  for (int k = 0; k < 16; k++) sl[k] = -1;
  UTet4Const C;
  UTet4Mom M;
  if (active) {
    int node[4];
    const int4 nn = __ldg(reinterpret_cast<const int4*>(P.IEN) + e);
    node[0] = nn.x; node[1] = nn.y; node[2] = nn.z; node[3] = nn.w;
    const int4* sp = reinterpret_cast<const int4*>(P.slot) + (size_t)e * 4;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int4 v = __ldg(sp + k);
      sl[4 * k] = v.x; sl[4 * k + 1] = v.y; sl[4 * k + 2] = v.z; sl[4 * k + 3] = v.w;
    }
    double xl[4][3], ql[4][3], vl[4][3], dl[4][3], pl[4], pdl[4];
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const size_t n = (size_t)node[b];
      const double* A = P.Ag + (size_t)P.tDof * n + P.s;
      const double* Y = P.Yg + (size_t)P.tDof * n + P.s;
      const double* D = P.Dg + (size_t)P.tDof * n + P.s;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        xl[b][i] = __ldg(P.x + 3 * n + i);
        ql[b][i] = __ldg(A + i) - __ldg(P.Bf + 3 * n + i);
        vl[b][i] = __ldg(Y + i);
        dl[b][i] = __ldg(D + i);
      }
      pl[b] = __ldg(Y + 3);
      pdl[b] = __ldg(A + 3);
    }
    double fN[2][3] = {{0, 0, 0}, {0, 0, 0}};
    if (P.fN != nullptr)
      for (int k = 0; k < P.nFn && k < 2; k++)
#pragma unroll
        for (int i = 0; i < 3; i++) fN[k][i] = __ldg(P.fN + (size_t)3 * P.nFn * e + 3 * k + i);
    // active tensions: Siso and Dm are affine in them and enter the Gauss sums with the weight alone, so one evaluation at
    // the weighted mean sum_g w_g ya_g / sum_g w_g reproduces the sum
    double ya[3] = {0.0, 0.0, 0.0};
    const bool act = CANN && (P.Ya != nullptr) && dm.st.active;
    if (act) {
      double wsum = 0.0;
#pragma unroll
      for (int g = 0; g < 4; g++) {
        wsum += P.w[g];
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
          for (int i = 0; i < 3; i++) ya[i] += P.w[g] * P.N[g][b] * __ldg(P.Ya + 3 * (size_t)node[b] + i);
      }
#pragma unroll
      for (int i = 0; i < 3; i++) ya[i] /= wsum;
    }
    double Dm[6][6], Je;
    ustruct_tet4_setup<CANN>(dm, af, am, P.w, &P.N[0][0], MAX_ENON, P.Nxi[0], xl, ql, vl, dl, pl, pdl, fN, C, M, Dm, &Je,
                       act ? ya : nullptr, P.cann, P.nFn);
    if (fabs(Je) < 10.0 * 2.220446049250313e-16 * 2.220446049250313e-16) atomicMax(P.err, e + 1);
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int c = 0; c < 6; c++) sDm[(6 * r + c) * UTET_THREADS] = Dm[r][c];
#pragma unroll
    for (int a = 0; a < 4; a++) {
      double r[4];
      ustruct_tet4_resid(C, M, &P.N[0][0], MAX_ENON, a, r);
#pragma unroll
      for (int i = 0; i < 4; i++) uadd<ATOMIC>(P.R + (size_t)4 * node[a] + i, r[i]);
    }
  }
  constexpr int TILE_LD = 29;
#ifdef SVB_UTET_ROLL_B
#pragma unroll 1
#else
#pragma unroll
#endif
  for (int b = 0; b < 4; b++) {
    double DBmb[6][3];
    if (active) {
      double Bmb[6][3];
      make_Bm(C.Nx[b], C.F, Bmb);
#pragma unroll
      for (int r = 0; r < 6; r++) {
        double d[6];
#pragma unroll
        for (int c = 0; c < 6; c++) d[c] = sDm[(6 * r + c) * UTET_THREADS];
#pragma unroll
        for (int j = 0; j < 3; j++)
          DBmb[r][j] = d[0] * Bmb[0][j] + d[1] * Bmb[1][j] + d[2] * Bmb[2][j] + d[3] * Bmb[3][j] + d[4] * Bmb[4][j] + d[5] * Bmb[5][j];
      }
    }
#ifdef SVB_UTET_ROLL_A
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int a = 0; a < 4; a++) {
      double K[16], Kd[12];
      if (active) {
        double Bma[6][3];
        make_Bm(C.Nx[a], C.F, Bma);
        ustruct_tet4_block(C, M, &P.N[0][0], MAX_ENON, af, am, a, b, Bma, DBmb, K, Kd);
      }
      __syncwarp();
      if (active) {
#pragma unroll
        for (int i = 0; i < 16; i++) tile[lane * TILE_LD + i] = K[i];
#pragma unroll
        for (int i = 0; i < 12; i++) tile[lane * TILE_LD + 16 + i] = Kd[i];
      }
      int myslot = -1;
#pragma unroll
      for (int k = 0; k < 16; k++)
        if (k == 4 * a + b) myslot = sl[k];
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 16; it++) {
        const int p = it * 32 + lane, src = p >> 4, i = p & 15;
        const int s_ = __shfl_sync(0xffffffffu, myslot, src);
        if (s_ >= 0) uadd<ATOMIC>(P.Val + (size_t)16 * s_ + i, tile[src * TILE_LD + i]);
      }
#pragma unroll
      for (int it = 0; it < 12; it++) {
        const int p = it * 32 + lane, src = p / 12, i = p - 12 * src;
        const int s_ = __shfl_sync(0xffffffffu, myslot, src);
        if (s_ >= 0) uadd<ATOMIC>(P.Kd + (size_t)12 * s_ + i, tile[src * TILE_LD + 16 + i]);
      }
    }
  }
}

// CANN: a domain of the launch uses the CANN constitutive model (the twin of the kernel with that branch compiled in, struct_elem.cuh)
template <bool CANN>
static int launch_ustruct_tet4(svb200_ctx* ctx, const UstructArgs& A, bool atomic)
{
  static bool configured = false;
  if (!configured) {
    SVB_CUDA(cudaFuncSetAttribute(assemble_ustruct_tet4_kernel<true, CANN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UTET_SMEM));
    SVB_CUDA(cudaFuncSetAttribute(assemble_ustruct_tet4_kernel<false, CANN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UTET_SMEM));
    configured = true;
  }
  const long long n = (long long)A.e1 - A.e0;
  if (n <= 0) return SVB200_OK;
  const unsigned blocks = (unsigned)((n + UTET_THREADS - 1) / UTET_THREADS);
  if (atomic) assemble_ustruct_tet4_kernel<true, CANN><<<blocks, UTET_THREADS, UTET_SMEM, ctx->stream>>>(A);
  else assemble_ustruct_tet4_kernel<false, CANN><<<blocks, UTET_THREADS, UTET_SMEM, ctx->stream>>>(A);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

template <int ENON, bool VISC, bool CANN>
static int launch_ustruct(svb200_ctx* ctx, const UstructArgs& A, bool atomic)
{
  using GP = typename UGPSel<VISC>::type;
  constexpr int EPB = (USTRUCT_THREADS / 32) * (32 / ENON);
  constexpr size_t smem = sizeof(double) * (size_t)(USTRUCT_THREADS / 32) * ustruct_warp_ld(ENON, (int)(sizeof(GP) / sizeof(double)));
  static bool configured = false;
  if (!configured) {
    SVB_CUDA(cudaFuncSetAttribute(assemble_ustruct_kernel<ENON, true, VISC, CANN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SVB_CUDA(cudaFuncSetAttribute(assemble_ustruct_kernel<ENON, false, VISC, CANN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const long long n = (long long)A.e1 - A.e0;
  if (n <= 0) return SVB200_OK;
  const unsigned blocks = (unsigned)((n + EPB - 1) / EPB);
  if (atomic) assemble_ustruct_kernel<ENON, true, VISC, CANN><<<blocks, USTRUCT_THREADS, smem, ctx->stream>>>(A);
  else assemble_ustruct_kernel<ENON, false, VISC, CANN><<<blocks, USTRUCT_THREADS, smem, ctx->stream>>>(A);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

int run_assemble_ustruct(svb200_ctx* ctx, const Mesh& m, const svb200_eqparams* eq, const svb200_dmnparams* dmn, int nDmn)
{
  SVB_REQUIRE(nDmn >= 1 && nDmn <= MAX_DMN, "svb200_assemble: between 1 and 8 domains are supported");
  SVB_REQUIRE(eq->dof == 4 && ctx->dof == 4, "svb200_assemble: the ustruct equation has dof = 4 (call svb200_alloc(4))");
  SVB_REQUIRE(eq->tDof == ctx->tDof && ctx->d_Dg && ctx->d_Yg && ctx->d_Ag, "svb200_assemble: state not set or tDof mismatch");
  SVB_REQUIRE(eq->s >= 0 && eq->s + 4 <= eq->tDof, "svb200_assemble: eq.s out of range");
  SVB_REQUIRE(eq->vmsStab == 1 && m.th_eNoNq == 0, "svb200_assemble: only equal-order (VMS-stabilised) ustruct elements are supported (no Taylor-Hood ustruct)");
  SVB_REQUIRE(ctx->d_x, "svb200_assemble: coordinates not set");
  SVB_REQUIRE(m.eNoN == 4 || m.eNoN == 8, "svb200_assemble: ustruct is implemented for TET4 and HEX8 meshes");
  SVB_REQUIRE(m.nG == m.eNoN, "svb200_assemble: the ustruct kernel expects nG == eNoN (TET4: 4, HEX8: 8 Gauss points)");
  // Kd((nsd+1)*nsd, nnz): allocated with the first ustruct assembly, zeroed there and by every later svb200_alloc
  // (Integrator::step zeroes it together with R and Val, solver/Integrator.cpp:106-109)
  if (!ctx->d_Kd) {
    SVB_CUDA(cudaMalloc(&ctx->d_Kd, sizeof(double) * 12 * std::max<size_t>((size_t)ctx->nnz, 1)));
    SVB_CUDA(cudaMemsetAsync(ctx->d_Kd, 0, sizeof(double) * 12 * (size_t)ctx->nnz, ctx->stream));
  }
  UstructArgs A;
  memset(&A, 0, sizeof(A));
  A.IEN = m.d_IEN; A.eId = m.d_eId; A.slot = m.d_slot; A.perm = nullptr; A.fN = m.d_fN;
  A.x = ctx->d_x; A.Ag = ctx->d_Ag; A.Yg = ctx->d_Yg; A.Dg = ctx->d_Dg; A.Bf = ctx->d_Bf;
  A.R = ctx->d_R; A.Val = ctx->d_Val; A.Kd = ctx->d_Kd;
  A.Ya = ctx->d_Ya;
  int cann_used = 0;
  if (!ctx->d_err) {
    SVB_CUDA(cudaMalloc(&ctx->d_err, sizeof(int)));
    SVB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
  }
  A.err = ctx->d_err;
  A.e0 = 0; A.e1 = m.nEl;
  A.tDof = eq->tDof; A.s = eq->s; A.nFn = m.nFn; A.nDmn = nDmn; A.nG = m.nG;
  A.dt = eq->dt; A.af = eq->af; A.am = eq->am; A.gam = eq->gam;
  for (int g = 0; g < m.nG; g++) {
    A.w[g] = m.w[g];
    for (int a = 0; a < m.eNoN; a++) {
      A.N[g][a] = m.N[(size_t)g * m.eNoN + a];
      for (int k = 0; k < 3; k++) A.Nxi[g][a][k] = m.Nx[((size_t)g * m.eNoN + a) * 3 + k];
    }
  }
  bool whole = false, visc = false;
  for (int d = 0; d < nDmn; d++) {
    StructDmn& o = A.dmn[d].st;
    o.rho = dmn[d].rho;
    for (int k = 0; k < 3; k++) o.f[k] = dmn[d].f[k];
    o.dmp = dmn[d].dmp;
    o.Kpen = dmn[d].Kpen; o.C10 = dmn[d].C10; o.C01 = dmn[d].C01;
    o.bff = dmn[d].bff; o.bss = dmn[d].bss; o.bfs = dmn[d].bfs;
    o.st_a = dmn[d].st_a; o.st_b = dmn[d].st_b; o.aff = dmn[d].aff; o.ass = dmn[d].ass; o.afs = dmn[d].afs;
    o.kap = dmn[d].kap; o.khs = dmn[d].khs;
    o.isoType = dmn[d].isoType; o.volType = dmn[d].volType;
    o.visc_mu = dmn[d].solid_visc_mu;
    o.viscType = SVB200_SOLID_VISC_NONE;
    if (dmn[d].phys == SVB200_PHYS_USTRUCT && dmn[d].solid_visc_mu != 0.0) {
      o.viscType = (dmn[d].solidViscType == SVB200_SOLID_VISC_POTENTIAL) ? SVB200_SOLID_VISC_POTENTIAL : SVB200_SOLID_VISC_NEWTONIAN;
      visc = true;
    }
    o.Id = dmn[d].Id;
    o.isStruct = A.active[d] = (dmn[d].phys == SVB200_PHYS_USTRUCT);
    A.dmn[d].E = dmn[d].E; A.dmn[d].nu = dmn[d].nu; A.dmn[d].ctM = dmn[d].ctau_M; A.dmn[d].ctC = dmn[d].ctau_C;
    SVB_REQUIRE(o.Id >= -1 && o.Id < 31, "svb200_assemble: domain Id out of range");
    if (A.active[d]) {
      SVB_REQUIRE(o.isoType >= SVB200_ISO_NHK && o.isoType <= SVB200_ISO_CANN, "svb200_assemble: constitutive model not implemented");
      int rce = fill_solid_extras(ctx, dmn[d], o, A.cann, &cann_used);
      if (rce) return rce;
      const bool fibres = (m.nFn == 2 && m.d_fN);
      if ((o.isoType == SVB200_ISO_GUCCIONE || o.isoType == SVB200_ISO_HGO || o.isoType == SVB200_ISO_HO || o.isoType == SVB200_ISO_HO_MA) && !fibres) {
        set_error("[compute_pk2cc] Min fiber directions not defined for this material model.");
        return SVB200_ERR_INVALID;
      }
    }
    whole |= (o.Id == -1);
  }
  if (!whole && !m.d_eId) { set_error("eId is not allocated"); return SVB200_ERR_INVALID; }
  const bool atomic = (eq->scatter == SVB200_SCATTER_ATOMIC);
  // linear tets without solid viscosity: closed-form kernel (SVB200_STRUCT_GENERAL=1 keeps the general one: cross-check)
  static const bool force_general = (getenv("SVB200_STRUCT_GENERAL") != nullptr && atoi(getenv("SVB200_STRUCT_GENERAL")) != 0);
  bool cannM = cann_used > 0;       // the full-featured twin: CANN model or active stress in a domain of this launch
  for (int d = 0; d < nDmn; d++) cannM |= (A.active[d] && A.dmn[d].st.active);
  auto launch = [&](const UstructArgs& B) {
    if (m.eNoN == 4 && !visc && !force_general) return cannM ? launch_ustruct_tet4<true>(ctx, B, atomic) : launch_ustruct_tet4<false>(ctx, B, atomic);
    if (visc) {      // viscosity + CANN: the CANN twin exists for the plain kernels only; take it out of the hot build
      if (m.eNoN == 8) return launch_ustruct<8, true, true>(ctx, B, atomic);
      return launch_ustruct<4, true, true>(ctx, B, atomic);
    }
    if (m.eNoN == 8) return cannM ? launch_ustruct<8, false, true>(ctx, B, atomic) : launch_ustruct<8, false, false>(ctx, B, atomic);
    return cannM ? launch_ustruct<4, false, true>(ctx, B, atomic) : launch_ustruct<4, false, false>(ctx, B, atomic);
  };
  int rc = SVB200_OK;
  if (atomic) rc = launch(A);
  else {
    A.perm = m.d_color_perm;
    for (size_t c = 0; c + 1 < m.color_off.size() && rc == SVB200_OK; c++) {
      A.e0 = m.color_off[c];
      A.e1 = m.color_off[c + 1];
      rc = launch(A);
    }
  }
  if (rc) return rc;
  int e = 0;
  SVB_CUDA(cudaMemcpyAsync(&e, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (e != 0) {
    SVB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
    set_error("[construct_usolid] Jacobian for element " + std::to_string(e - 1) + " is < 0.");
    return SVB200_ERR_NUMERIC;
  }
  return SVB200_OK;
}

// ---- ustruct::ustruct_r (ustruct.cpp:1742-1845) -----------------------------------------------------------------
// Rd = amg Ad - Yg(s..s+2);  KU = Kd Rd (4x3 blocks);  halo sum;  R -= KU / am.   Only in the first Newton iteration.
__global__ void __launch_bounds__(256)
ustruct_rd_kernel(int nNo, int tDof, int s, double amg, const double* __restrict__ Ad, const double* __restrict__ Yg,
                  const int* __restrict__ flag, double* __restrict__ Rd)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * nNo) return;
  const int n = t / 3, i = t % 3;
  // FSI: nodes outside the ustruct domains keep the zero Integrator::step wrote (Integrator.cpp:106-109)
  Rd[t] = (flag && !flag[n]) ? 0.0 : amg * Ad[t] - Yg[(size_t)tDof * n + s + i];
}

__global__ void __launch_bounds__(256)
ustruct_kd_spmv_kernel(int nNo, const int* __restrict__ rowPtr, const int* __restrict__ colPtr, const double* __restrict__ Kd,
                       const double* __restrict__ Rd, const int* __restrict__ flag, double* __restrict__ KU)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nNo * 4) return;
  const int row = (int)(t >> 2), i = (int)(t & 3);
  double acc = 0.0;
  if (flag && !flag[row]) { KU[t] = 0.0; return; }
  for (int k = rowPtr[row]; k < rowPtr[row + 1]; k++) {
    const double* v = Kd + (size_t)k * 12 + 3 * i;
    const double* u = Rd + (size_t)colPtr[k] * 3;
    acc += v[0] * u[0] + v[1] * u[1] + v[2] * u[2];
  }
  KU[t] = acc;
}

__global__ void __launch_bounds__(256)
ustruct_r_update_kernel(size_t n, double ami, const double* __restrict__ KU, double* __restrict__ R)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) R[t] -= ami * KU[t];
}

int run_ustruct_r(svb200_ctx* ctx, const svb200_eqparams* eq, int itr, const double* d_Ad)
{
  const int n = ctx->nNo;
  if (n == 0) return SVB200_OK;
  if (!ctx->d_Rd) SVB_CUDA(cudaMalloc(&ctx->d_Rd, sizeof(double) * 3 * (size_t)n));
  if (itr > 1) {                            // Rd = 0: nothing to add (ustruct.cpp:1773-1775)
    SVB_CUDA(cudaMemsetAsync(ctx->d_Rd, 0, sizeof(double) * 3 * (size_t)n, ctx->stream));
    return SVB200_OK;
  }
  double* KU = nullptr;
  SVB_CUDA(cudaMalloc(&KU, sizeof(double) * 4 * (size_t)n));
  double* Rd = ctx->d_Rd;
  const double amg = (eq->gam - eq->am) / (eq->gam - 1.0), ami = 1.0 / eq->am;
  const int* flag = (eq->phys == SVB200_PHYS_FSI) ? ctx->d_nodeflag : nullptr;     // is_domain(..., phys_ustruct) of an FSI equation
  ustruct_rd_kernel<<<(3 * n + 255) / 256, 256, 0, ctx->stream>>>(n, ctx->tDof, eq->s, amg, d_Ad, ctx->d_Yg, flag, Rd);
  ustruct_kd_spmv_kernel<<<(unsigned)(((long long)n * 4 + 255) / 256), 256, 0, ctx->stream>>>(n, ctx->d_rowPtr, ctx->d_colPtr, ctx->d_Kd, Rd, flag, KU);
  ctx->launches += 2;
  int rc = halo_sum(ctx, 4, KU);
  if (!rc) {
    ustruct_r_update_kernel<<<(unsigned)(((size_t)n * 4 + 255) / 256), 256, 0, ctx->stream>>>((size_t)n * 4, ami, KU, ctx->d_R);
    ctx->launches++;
  }
  cudaError_t ce = cudaStreamSynchronize(ctx->stream);
  cudaFree(KU);
  if (rc) return rc;
  SVB_CUDA(ce);
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

}  // namespace svb
