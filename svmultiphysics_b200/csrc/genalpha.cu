// genalpha.cu — generalised-alpha state updates on the device, so that a Newton
// iteration needs no nodal array to cross PCIe: only norms and the solver report do.
//
// Replaces Integrator::predictor (Code/Source/solver/Integrator.cpp:393-643; the state part for the equations
// on the path), Integrator::initiator (:662-750), the state update of Integrator::corrector (:774-867, :887-912)
// and the strong-Dirichlet write of set_bc::set_bc_dir (Code/Source/solver/set_bc.cpp:901-1067).
// All kernels stream (tDof, nNo) arrays once; products and sums are issued as separate roundings
// (__dmul_rn/__dadd_rn) in the reference's evaluation order, so results are bit-identical to it.
#include "svb200_internal.h"

namespace svb {

struct TimeEqs {
  int nEq;
  int s[8], e[8], phys[8];
  int sst[8];        // velocity-pressure solid update (com_mod.sstEq): a ustruct equation, or an FSI equation with SVB200_EQTIME_SSTEQ
  double af[8], am[8], gam[8], beta[8];
};

__device__ __forceinline__ int eq_of_row(const TimeEqs& Q, int j)
{
  for (int i = 0; i < Q.nEq; i++)
    if (j >= Q.s[i] && j <= Q.e[i]) return i;
  return -1;
}

// An = Ao (gam-1)/gam ; Yn = Yo ; Dn = Do + Yn dt + An dt^2 (gam/2 - beta)/(gam - 1)  (dFlag) or Dn = Do.
__global__ void predictor_kernel(int tDof, long long n, const __grid_constant__ TimeEqs Q, double dt, int dFlag,
                                 const double* __restrict__ Ao, const double* __restrict__ Yo, const double* __restrict__ Do,
                                 double* __restrict__ An, double* __restrict__ Yn, double* __restrict__ Dn)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int q = eq_of_row(Q, (int)(t % tDof));
  if (q < 0) return;
  const double coef = (Q.gam[q] - 1.0) / Q.gam[q];
  const double an = __dmul_rn(Ao[t], coef);
  const double yn = Yo[t];
  An[t] = an;
  Yn[t] = yn;
  // ustruct / FSI with ustruct solids (sstEq, Integrator.cpp:626-630): the displacement rows are integrated through Ad by the
  // corrector, Dn = Do here; the mesh equation of such an FSI case keeps the usual update (:632-635)
  if (dFlag && !Q.sst[q]) {
    const double cD = dt * dt * (0.5 * Q.gam[q] - Q.beta[q]) / (Q.gam[q] - 1.0);
    Dn[t] = __dadd_rn(__dadd_rn(Do[t], __dmul_rn(yn, dt)), __dmul_rn(an, cD));
  } else {
    Dn[t] = Do[t];
  }
}

// Ag = Ao (1-am) + An am ; Yg = Yo (1-af) + Yn af ; Dg = Do (1-af) + Dn af.
__global__ void initiator_kernel(int tDof, long long n, const __grid_constant__ TimeEqs Q,
                                 const double* __restrict__ Ao, const double* __restrict__ Yo, const double* __restrict__ Do,
                                 const double* __restrict__ An, const double* __restrict__ Yn, const double* __restrict__ Dn,
                                 double* __restrict__ Ag, double* __restrict__ Yg, double* __restrict__ Dg)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int q = eq_of_row(Q, (int)(t % tDof));
  if (q < 0) return;
  const double c0 = 1.0 - Q.am[q], c1 = Q.am[q], c2 = 1.0 - Q.af[q], c3 = Q.af[q];
  Ag[t] = __dadd_rn(__dmul_rn(Ao[t], c0), __dmul_rn(An[t], c1));
  Yg[t] = __dadd_rn(__dmul_rn(Yo[t], c2), __dmul_rn(Yn[t], c3));
  Dg[t] = __dadd_rn(__dmul_rn(Do[t], c2), __dmul_rn(Dn[t], c3));
}

// An(s+i,a) -= R(i,a) ; Yn -= R gam dt ; Dn -= R beta dt^2   for i = 0..e-s  (R = the solver's increment, (dof,nNo)).
__global__ void corrector_kernel(int tDof, int dof, int s, int nrow, long long nNo, double c0, double c1,
                                 const double* __restrict__ R, double* __restrict__ An, double* __restrict__ Yn,
                                 double* __restrict__ Dn)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nNo * nrow) return;
  const long long a = t / nrow;
  const int i = (int)(t % nrow);
  const double r = R[a * dof + i];
  const long long k = a * tDof + s + i;
  An[k] = __dadd_rn(An[k], -r);
  Yn[k] = __dadd_rn(Yn[k], -__dmul_rn(r, c0));
  Dn[k] = __dadd_rn(Dn[k], -__dmul_rn(r, c1));
}

// ustruct (Integrator.cpp:826-846): An(s+i) -= R(i), Yn(s+i) -= R(i) gam dt for the 4 rows; for the 3 velocity rows
// dUl = Rd(i)/am + R(i) af gam dt / am, Ad(i) -= dUl, Dn(s+i) -= dUl gam dt.
__global__ void corrector_ustruct_kernel(int tDof, int s, long long nNo, double c0, double c2, double c3, const double* __restrict__ R,
                                         const double* __restrict__ Rd, double* __restrict__ An, double* __restrict__ Yn,
                                         double* __restrict__ Dn, double* __restrict__ Ad)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nNo * 4) return;
  const long long a = t >> 2;
  const int i = (int)(t & 3);
  const double r = R[t];
  const long long k = a * tDof + s + i;
  An[k] = __dadd_rn(An[k], -r);
  Yn[k] = __dadd_rn(Yn[k], -__dmul_rn(r, c0));
  if (i < 3) {
    const double dUl = __dadd_rn(__dmul_rn(Rd[a * 3 + i], c2), __dmul_rn(r, c3));
    Ad[a * 3 + i] = __dadd_rn(Ad[a * 3 + i], -dUl);
    Dn[k] = __dadd_rn(Dn[k], -__dmul_rn(dUl, c0));
  }
}

__global__ void scale_kernel(long long n, double c, double* __restrict__ x)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) x[t] = __dmul_rn(x[t], c);
}

// FSI without explicit geometric coupling (Integrator.cpp:887-912): on solid nodes the mesh-equation rows
// s1.. take the values of the FSI rows 0..: An(i+s1,Ac) = An(i,Ac), same for Yn, Dn.
__global__ void corrector_fsi_copy_kernel(int tDof, int s1, int nrow, long long nNo, const int* __restrict__ flag,
                                          double* __restrict__ An, double* __restrict__ Yn, double* __restrict__ Dn)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nNo * nrow) return;
  const long long a = t / nrow;
  const int i = (int)(t % nrow);
  if (!flag[a]) return;
  An[a * tDof + s1 + i] = An[a * tDof + i];
  Yn[a * tDof + s1 + i] = Yn[a * tDof + i];
  Dn[a * tDof + s1 + i] = Dn[a * tDof + i];
}

// dst(row0 + i, nodes[k]) = val(i, k), i < nrow (set_bc_dir's writes of the prescribed values).
__global__ void set_rows_kernel(int tDof, int row0, int nrow, int n, const int* __restrict__ nodes,
                                const double* __restrict__ val, double* __restrict__ dst)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * nrow) return;
  const int k = t / nrow, i = t % nrow;
  dst[(size_t)nodes[k] * tDof + row0 + i] = val[t];
}

// set_bc::set_bc_dir for a velocity-pressure solid (ustruct; set_bc.cpp:1046-1117), after the prescribed values were written:
//   impD == 0:  Dn(j) = c1 Yn(j) - c2 Ad(i) + Do(j) ;  Ad(i) = Yn(j)
//   impD != 0:  An(j) = c1i (Yn(j) - Yo(j) + c2 Ao(j)) ;  Ad(i) = c1i (Dn(j) - Do(j) + c2 Ad(i))
// with j = s + i, c1 = gam dt, c1i = 1 / c1, c2 = (gam - 1) dt, for the directions i of dir_mask.
__global__ void dirichlet_ustruct_kernel(int tDof, int s, int n, const int* __restrict__ nodes, int dir_mask, int impD, double c1,
                                         double c1i, double c2, const double* __restrict__ Ao, const double* __restrict__ Yo,
                                         const double* __restrict__ Do, double* __restrict__ An, const double* __restrict__ Yn,
                                         double* __restrict__ Dn, double* __restrict__ Ad)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * 3) return;
  const int k = t / 3, i = t % 3;
  if (!((dir_mask >> i) & 1)) return;
  const size_t a = (size_t)nodes[k];
  const size_t j = a * tDof + s + i;
  double* ad = Ad + a * 3 + i;
  if (impD) {
    An[j] = __dmul_rn(c1i, __dadd_rn(__dadd_rn(Yn[j], -Yo[j]), __dmul_rn(c2, Ao[j])));
    *ad = __dmul_rn(c1i, __dadd_rn(__dadd_rn(Dn[j], -Do[j]), __dmul_rn(c2, *ad)));
  } else {
    Dn[j] = __dadd_rn(__dadd_rn(__dmul_rn(c1, Yn[j]), -__dmul_rn(c2, *ad)), Do[j]);
    *ad = Yn[j];
  }
}

int launch_dirichlet_ustruct(svb200_ctx* ctx, const svb200_eqtime* eq, double dt, int n, const int* d_nodes, int dir_mask, int impD)
{
  if (n == 0) return SVB200_OK;
  const double c1 = eq->gam * dt, c1i = 1.0 / c1, c2 = (eq->gam - 1.0) * dt;
  dirichlet_ustruct_kernel<<<(n * 3 + 255) / 256, 256, 0, ctx->stream>>>(ctx->tDof, eq->s, n, d_nodes, dir_mask, impD, c1, c1i, c2,
                                                                        ctx->d_Ao, ctx->d_Yo, ctx->d_Do, ctx->d_An, ctx->d_Yn,
                                                                        ctx->d_Dn, ctx->d_Ad);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

static int fill(const svb200_eqtime* eqs, int nEq, TimeEqs& Q)
{
  SVB_REQUIRE(eqs && nEq >= 1 && nEq <= 8, "gen-alpha: between 1 and 8 equations");
  Q.nEq = nEq;
  for (int i = 0; i < nEq; i++) {
    Q.s[i] = eqs[i].s; Q.e[i] = eqs[i].e; Q.phys[i] = eqs[i].phys;
    Q.sst[i] = eqtime_is_sst(eqs[i]);
    Q.af[i] = eqs[i].af; Q.am[i] = eqs[i].am; Q.gam[i] = eqs[i].gam; Q.beta[i] = eqs[i].beta;
    SVB_REQUIRE(eqs[i].s >= 0 && eqs[i].e >= eqs[i].s, "gen-alpha: bad equation row range");
  }
  return SVB200_OK;
}

static inline unsigned nblk(long long n) { return (unsigned)((n + 255) / 256); }

int launch_predictor(svb200_ctx* ctx, int nEq, const svb200_eqtime* eqs, double dt, int dFlag)
{
  TimeEqs Q;
  int rc = fill(eqs, nEq, Q);
  if (rc) return rc;
  const long long n = (long long)ctx->tDof * ctx->nNo;
  if (n == 0) return SVB200_OK;
  predictor_kernel<<<nblk(n), 256, 0, ctx->stream>>>(ctx->tDof, n, Q, dt, dFlag, ctx->d_Ao, ctx->d_Yo, ctx->d_Do, ctx->d_An,
                                                     ctx->d_Yn, ctx->d_Dn);
  ctx->launches++;
  if (dFlag && ctx->d_Ad)
    for (int i = 0; i < nEq; i++)
      if (eqtime_is_sst(eqs[i])) {                       // Ad = Ad (gam-1)/gam, Integrator.cpp:627-628
        const long long m = 3LL * ctx->nNo;
        scale_kernel<<<nblk(m), 256, 0, ctx->stream>>>(m, (eqs[i].gam - 1.0) / eqs[i].gam, ctx->d_Ad);
        ctx->launches++;
      }
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

int launch_initiator(svb200_ctx* ctx, int nEq, const svb200_eqtime* eqs)
{
  TimeEqs Q;
  int rc = fill(eqs, nEq, Q);
  if (rc) return rc;
  const long long n = (long long)ctx->tDof * ctx->nNo;
  if (n == 0) return SVB200_OK;
  initiator_kernel<<<nblk(n), 256, 0, ctx->stream>>>(ctx->tDof, n, Q, ctx->d_Ao, ctx->d_Yo, ctx->d_Do, ctx->d_An, ctx->d_Yn,
                                                     ctx->d_Dn, ctx->d_Ag, ctx->d_Yg, ctx->d_Dg);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

int launch_corrector(svb200_ctx* ctx, const svb200_eqtime* eq, double dt, int mesh_s, const int* d_flag)
{
  const int nrow = eq->e - eq->s + 1;
  const long long n = (long long)ctx->nNo * nrow;
  if (n == 0) return SVB200_OK;
  const double c0 = eq->gam * dt, c1 = eq->beta * dt * dt;
  if (eqtime_is_sst(*eq)) {
    // Integrator.cpp:828-846, the same update for a ustruct equation and for an FSI equation with ustruct solids (all nodes)
    const double c2 = 1.0 / eq->am, c3 = eq->af * c0 * c2;
    corrector_ustruct_kernel<<<nblk((long long)ctx->nNo * 4), 256, 0, ctx->stream>>>(ctx->tDof, eq->s, ctx->nNo, c0, c2, c3, ctx->d_R,
                                                                                      ctx->d_Rd, ctx->d_An, ctx->d_Yn, ctx->d_Dn, ctx->d_Ad);
  } else {
    corrector_kernel<<<nblk(n), 256, 0, ctx->stream>>>(ctx->tDof, ctx->dof, eq->s, nrow, ctx->nNo, c0, c1, ctx->d_R, ctx->d_An,
                                                       ctx->d_Yn, ctx->d_Dn);
  }
  ctx->launches++;
  if (mesh_s >= 0 && d_flag) {
    const long long m = (long long)ctx->nNo * 3;
    corrector_fsi_copy_kernel<<<nblk(m), 256, 0, ctx->stream>>>(ctx->tDof, mesh_s, 3, ctx->nNo, d_flag, ctx->d_An, ctx->d_Yn,
                                                                 ctx->d_Dn);
    ctx->launches++;
  }
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

int launch_set_rows(svb200_ctx* ctx, int row0, int nrow, int n, const int* d_nodes, const double* d_val, double* dst)
{
  if (n * nrow == 0) return SVB200_OK;
  set_rows_kernel<<<nblk((long long)n * nrow), 256, 0, ctx->stream>>>(ctx->tDof, row0, nrow, n, d_nodes, d_val, dst);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

}  // namespace svb
