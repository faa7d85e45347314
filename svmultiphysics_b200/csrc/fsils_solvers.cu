// fsils_solvers.cu — host drivers of the FSILS Krylov solvers running on device-resident data.
//
//   fsils_solve_device   fsi_linear_solver::fsils_solve   Code/Source/linear_solver/solve.cpp:23-166
//   precond_diag_device  precond::precond_diag            linear_solver/precond.cpp:95-242
//   gmres_device         gmres::gmres_v / gmres_s         linear_solver/gmres.cpp:425-609 / 257-412
//   cg_device            cgrad::cgrad_v / cgrad_s         linear_solver/cgrad.cpp:139-219 / 225-305
//   bicgs_device         bicgs::bicgsv / bicgss           linear_solver/bicgs.cpp:22-120 / 123-218
//   add_bc_mul_device    add_bc_mul::add_bc_mul           linear_solver/add_bc_mul.cpp:26-124
//
// The algorithms keep the reference's exact sequence (classical Gram-Schmidt with the Pythagoras
// norm, Givens updates, restart and early-return semantics, iteration counting) because the parity
// contract is on the residual HISTORY, not only on the answer.  The host only sees O(sD) scalars
// per iteration (the Hessenberg column); vectors and the matrix never leave HBM.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <limits>
#include "svb200_internal.h"
#include "fsils_kernels.h"

namespace svb {

#define SVB_TRY(call)                  \
  do {                                 \
    int rc__ = (call);                 \
    if (rc__ != SVB200_OK) return rc__; \
  } while (0)

static double now_s()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int ensure_work(svb200_ctx* ctx, size_t doubles)
{
  if (doubles > ctx->work_cap) {
    if (ctx->d_work) cudaFree(ctx->d_work);
    ctx->d_work = nullptr;
    ctx->work_cap = 0;
    SVB_CUDA(cudaMalloc(&ctx->d_work, sizeof(double) * doubles));
    ctx->work_cap = doubles;
  }
  return SVB200_OK;
}

// d_scal: small device scratch for reduction results; h_pinned: its pinned host mirror.
constexpr int SCAL_N = 1024;

static int ensure_scalars(svb200_ctx* ctx)
{
  if (!ctx->h_pinned) SVB_CUDA(cudaMallocHost(&ctx->h_pinned, sizeof(double) * SCAL_N));
  if (!ctx->h_cg) SVB_CUDA(cudaMallocHost(&ctx->h_cg, sizeof(double) * 24));
  if (!ctx->d_cg) SVB_CUDA(cudaMalloc(&ctx->d_cg, sizeof(double) * 8));
  for (int k = 0; k < 2; k++)
    if (!ctx->ev_cg[k]) SVB_CUDA(cudaEventCreateWithFlags(&ctx->ev_cg[k], cudaEventDisableTiming));
  return SVB200_OK;
}

// Copy n scalars from the device to pinned host memory and wait for them.
static int fetch(svb200_ctx* ctx, const double* d_src, int n, double* h_dst)
{
  SVB_CUDA(cudaMemcpyAsync(h_dst, d_src, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SVB200_OK;
}

// Global dot over owned nodes (dot::fsils_dot_v) and norm (norm::fsi_ls_normv).
static int dot_owned(svb200_ctx* ctx, int dof, const double* a, const double* b, double* d_scal, double* out)
{
  SVB_TRY(multi_dot(ctx, (long long)ctx->mynNo * dof, 1, a, 0, b, d_scal, true));
  SVB_TRY(fetch(ctx, d_scal, 1, ctx->h_pinned));
  *out = ctx->h_pinned[0];
  return SVB200_OK;
}

static int norm_owned(svb200_ctx* ctx, int dof, const double* a, double* d_scal, double* out)
{
  SVB_TRY(dot_owned(ctx, dof, a, a, d_scal, out));
  *out = std::sqrt(*out);
  return SVB200_OK;
}

// ---- coupled Neumann faces: Y += coef * valM (valM . X) ------------------------------------------
// Two-stage, fixed-order reduction (bitwise reproducible): FACE_DOT_BLOCKS partial sums, then one block adds them.
// (A single CTA walking the whole face took 154 us per call on the 14 k-node outlet of C2 — a quarter of the NS solve.)
constexpr int FACE_DOT_BLOCKS = 64;
__global__ void __launch_bounds__(256)
face_dot_kernel(int fnNo, int fdof, int nd, int dof, int mynNo, int shared, const int* __restrict__ glob,
                const double* __restrict__ valM, const double* __restrict__ X, double* __restrict__ part)
{
  __shared__ double red[256];
  double s = 0.0;
  const int total = fnNo * nd;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int a = t / nd, i = t % nd;
    const int Ac = glob[a];
    if (!shared || Ac < mynNo) s += valM[(size_t)a * fdof + i] * X[(size_t)Ac * dof + i];
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

__global__ void face_dot_stage2_kernel(int nblocks, const double* __restrict__ part, double* __restrict__ out)
{
  __shared__ double red[FACE_DOT_BLOCKS];
  red[threadIdx.x] = threadIdx.x < nblocks ? part[threadIdx.x] : 0.0;
  __syncthreads();
  for (int o = FACE_DOT_BLOCKS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

// Y += valM * S with S = coef * (valM . X) (+ coef * (cap_valM . X) when the face has a cap, add_bc_mul.cpp:62-81)
__global__ void face_axpy_kernel(int fnNo, int fdof, int nd, int dof, double coef, const double* __restrict__ S, int has_cap,
                                 const int* __restrict__ glob, const double* __restrict__ valM, double* __restrict__ Y)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= fnNo * nd) return;
  const int a = t / nd, i = t % nd;
  double s = coef * S[0];
  if (has_cap) s = s + coef * S[1];
  Y[(size_t)glob[a] * dof + i] += valM[(size_t)a * fdof + i] * s;
}

// op: 0 = BCOP_TYPE_ADD (coef = res), 1 = BCOP_TYPE_PRE (coef = -res/(1+res*nS)).
int add_bc_mul_device(svb200_ctx* ctx, int op, int dof, const double* X, double* Y, double* d_scal)
{
  for (auto& f : ctx->face) {
    if (!f.set || !f.coupledFlag) continue;
    const int nd = std::min(f.dof, dof);
    const double coef = (op == 0) ? f.res : -f.res / (1.0 + f.res * f.nS);
    // partial sums go to the top of the scalar scratch (consumed at once by stage 2), the result to d_scal[0]
    double* part = d_scal + SCAL_N - FACE_DOT_BLOCKS;
    face_dot_kernel<<<FACE_DOT_BLOCKS, 256, 0, ctx->stream>>>(f.nNo, f.dof, nd, dof, ctx->mynNo, f.shared, f.d_glob, f.d_valM, X, part);
    face_dot_stage2_kernel<<<1, FACE_DOT_BLOCKS, 0, ctx->stream>>>(FACE_DOT_BLOCKS, part, d_scal);
    ctx->launches += 2;
    if (f.has_cap) {
      // the cap's flow-rate contribution: ALL cap nodes of this partition, not only the owned ones (add_bc_mul.cpp:67-79)
      face_dot_kernel<<<FACE_DOT_BLOCKS, 256, 0, ctx->stream>>>(f.cap_n, f.dof, nd, dof, ctx->mynNo, 0, f.d_cap_glob, f.d_cap_valM, X, part);
      face_dot_stage2_kernel<<<1, FACE_DOT_BLOCKS, 0, ctx->stream>>>(FACE_DOT_BLOCKS, part, d_scal + 1);
      ctx->launches += 2;
    }
    if (f.shared) SVB_TRY(allreduce_sum(ctx, d_scal, f.has_cap ? 2 : 1));
    const int n = f.nNo * nd;
    if (n > 0) {
      face_axpy_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(f.nNo, f.dof, nd, dof, coef, d_scal, f.has_cap ? 1 : 0, f.d_glob, f.d_valM, Y);
      ctx->launches++;
    }
    SVB_CUDA(cudaGetLastError());
  }
  return SVB200_OK;
}

// face.nS = |valM|^2 over owned nodes, first dof-1 components (gmres.cpp:22-62 bc_pre).
static int bc_pre_device(svb200_ctx* ctx, int dof, double* d_scal)
{
  for (auto& f : ctx->face) {
    if (!f.set || !f.coupledFlag) continue;
    const int nsd = dof - 1;
    std::vector<double> valM((size_t)f.dof * f.nNo);
    std::vector<int> glob(f.nNo);
    SVB_CUDA(cudaMemcpyAsync(valM.data(), f.d_valM, sizeof(double) * valM.size(), cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(cudaMemcpyAsync(glob.data(), f.d_glob, sizeof(int) * glob.size(), cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(cudaStreamSynchronize(ctx->stream));
    double nS = 0.0;
    for (int a = 0; a < f.nNo; a++) {
      if (f.shared && glob[a] >= ctx->mynNo) continue;
      for (int i = 0; i < std::min(nsd, f.dof); i++) nS += valM[(size_t)a * f.dof + i] * valM[(size_t)a * f.dof + i];
    }
    if (f.shared && ctx->nranks > 1) {
      ctx->h_pinned[0] = nS;
      SVB_CUDA(cudaMemcpyAsync(d_scal, ctx->h_pinned, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      SVB_TRY(allreduce_sum(ctx, d_scal, 1));
      SVB_TRY(fetch(ctx, d_scal, 1, ctx->h_pinned));
      nS = ctx->h_pinned[0];
    }
    f.nS = nS;
  }
  return SVB200_OK;
}

// K*U + halo sum (every fsils_spar_mul_* ends in fsils_commuv, spar_mul.cpp:230).
static int spmv_halo(svb200_ctx* ctx, int dof, const double* Val, const double* U, double* KU)
{
  if (dof == 4 && ctx->nranks > 1) {
    // interface rows + push, interior rows, wait-add: the exchange travels behind the interior rows (comm.cu)
    bool handled = false;
    SVB_TRY(spmv4_halo_fused(ctx, Val, U, KU, &handled));
    if (handled) return SVB200_OK;
  }
  SVB_TRY(launch_spmv(ctx, dof, Val, U, KU));
  SVB_TRY(halo_sum(ctx, dof, KU));
  return SVB200_OK;
}

// ---- precond_diag ----------------------------------------------------------------------------
static int precond_diag_device(svb200_ctx* ctx, int dof, double* Val, double* R, double* W)
{
  const long long n = (long long)ctx->nNo * dof;
  SVB_TRY(precond_extract_diag(ctx, dof, Val, W));
  SVB_TRY(halo_sum(ctx, dof, W));
  SVB_TRY(precond_invsqrt(ctx, dof, W));
  for (auto& f : ctx->face) {
    if (!f.set || !f.incFlag) continue;
    if (f.bGrp == SVB200_BC_DIR) SVB_TRY(precond_face_scale(ctx, f, dof, W));
  }
  SVB_TRY(precond_scale_matrix(ctx, dof, W, W, Val));
  SVB_TRY(hadamard(ctx, n, W, R, R));
  for (auto& f : ctx->face) {
    if (!f.set || !f.coupledFlag) continue;
    SVB_TRY(precond_face_valm(ctx, f, dof, W));
  }
  return SVB200_OK;
}

// ---- precond_rcs (precond.cpp:251-523) ---------------------------------------------------------------
// Row-and-column max-norm equilibration, at most 10 sweeps.  W2 (column scaling) is what fsils_solve multiplies the
// solution with afterwards (solve.cpp:157-159); W1 (row scaling) is applied to R here.  Reference behaviour kept:
// the shared-node exchange of the max norms is a SUM (fsils_commuv), the Dirichlet mask is renormalised through
// sign(Wr - 0.5), and face.valM of coupled Neumann faces is NOT recomputed by this preconditioner.
static int precond_rcs_device(svb200_ctx* ctx, int dof, double* Val, double* R, double* W1, double* W2, double* Wr,
                              double* Wc, double* d_scal)
{
  const long long n = (long long)ctx->nNo * dof;
  if (dof > 4) {
    set_error("precond_rcs: dof > 4 is not supported");
    return SVB200_ERR_UNSUPPORTED;
  }
  const int maxiter = 10;
  const double tol = 2.0;
  SVB_TRY(fill(ctx, n, W1, 1.0));
  SVB_TRY(fill(ctx, n, W2, 1.0));
  SVB_TRY(fill(ctx, n, Wr, 1.0));
  for (auto& f : ctx->face) {
    if (!f.set || !f.incFlag) continue;
    if (f.bGrp == SVB200_BC_DIR) SVB_TRY(precond_face_scale(ctx, f, dof, Wr));
  }
  SVB_TRY(halo_sum(ctx, dof, Wr));
  SVB_TRY(rcs_renorm(ctx, n, Wr));
  // kill the Dirichlet rows and columns, unit diagonal
  SVB_TRY(precond_scale_matrix(ctx, dof, Wr, Wr, Val));
  SVB_TRY(hadamard(ctx, n, Wr, R, R));
  SVB_TRY(rcs_diag_one(ctx, dof, Wr, Val));
  bool flag = true;
  int iter = 0;
  while (flag) {
    SVB_CUDA(cudaMemsetAsync(Wr, 0, sizeof(double) * n, ctx->stream));
    SVB_CUDA(cudaMemsetAsync(Wc, 0, sizeof(double) * n, ctx->stream));
    iter++;
    if (iter >= maxiter) flag = false;
    SVB_TRY(rcs_rowcol_max(ctx, dof, Val, Wr, Wc));
    SVB_TRY(halo_sum(ctx, dof, Wr));
    SVB_TRY(halo_sum(ctx, dof, Wc));
    SVB_TRY(rcs_dev_from_one(ctx, n, Wr, Wc, d_scal));
    SVB_TRY(fetch(ctx, d_scal, 2, ctx->h_pinned));
    if (ctx->h_pinned[0] < tol && ctx->h_pinned[1] < tol) flag = false;
    SVB_TRY(rcs_invsqrt_acc(ctx, n, Wr, W1));
    SVB_TRY(rcs_invsqrt_acc(ctx, n, Wc, W2));
    SVB_TRY(precond_scale_matrix(ctx, dof, Wr, Wc, Val));
    if (ctx->nranks > 1) {
      // MPI_Allgather of the flags + any() (precond.cpp:507-512)
      ctx->h_pinned[0] = flag ? 1.0 : 0.0;
      SVB_CUDA(cudaMemcpyAsync(d_scal, ctx->h_pinned, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      SVB_TRY(allreduce_sum(ctx, d_scal, 1));
      SVB_TRY(fetch(ctx, d_scal, 1, ctx->h_pinned));
      flag = ctx->h_pinned[0] > 0.0;
    }
  }
  SVB_TRY(hadamard(ctx, n, W1, R, R));
  return SVB200_OK;
}

// ---- GMRES -------------------------------------------------------------------------------------
// One routine for both reference variants:
//   ns_inner == false: gmres_v / gmres_s (gmres.cpp:425-609 / 257-412): R is overwritten by the solution, the
//                      tolerance comes from |R|, every SpMV is counted, early return when |R| <= absTol;
//   ns_inner == true : gmres (gmres.cpp:65-250), the momentum solve inside the NS solver: R is read only, X is
//                      the output, no SpMV on the first restart, eps fixed by the first restart, itr and callD
//                      ACCUMULATE over calls, and the BCOP_TYPE_PRE correction is applied when a coupled face
//                      exists (it is dead code in gmres_v, `flag = false`, gmres.cpp:437).
// u: workspace of (sD+1) vectors, d_scal: >= sD+16 doubles of device scratch.
static int gmres_core(svb200_ctx* ctx, int dof, const svb200_sublsparams& p, svb200_sublsresult& r, const double* Val,
                      double* R, double* X, double* u, double* d_scal, bool ns_inner, svb200_lsresult* full)
{
  const long long n = (long long)ctx->nNo * dof;
  const long long us = (n + 1) & ~1ll;     // 16-byte aligned stride between Krylov vectors (double2 kernels)
  const int sD = p.sD;
  double* d_h = d_scal + 8;
  double* d_hn = d_scal + 4;
  double* hp = ctx->h_pinned;
  std::vector<double> h((size_t)(sD + 1) * sD, 0.0), y(sD), c(sD), s(sD), err(sD + 1, 0.0);
  auto H = [&](int i, int j) -> double& { return h[(size_t)j * (sD + 1) + i]; };
  bool anyCoupled = false;
  for (auto& f : ctx->face) anyCoupled |= (f.set && f.coupledFlag);

  const double t0 = now_s();
  r.success = 0;
  double eps = 0.0;
  int last_i = 0;
  if (!ns_inner) {
    SVB_TRY(norm_owned(ctx, dof, R, d_scal, &eps));
    r.iNorm = eps;
    r.fNorm = eps;
    eps = std::max(p.absTol, p.relTol * eps);
    r.itr = 0;
    SVB_TRY(bc_pre_device(ctx, dof, d_scal));
    if (full) full->hist_n = 0;
    if (r.iNorm <= p.absTol) {
      r.callD = std::numeric_limits<double>::epsilon();
      r.dB = 0.0;
      r.success = 1;
      return SVB200_OK;     // R is left untouched, as in the reference (gmres.cpp:470-475)
    }
  }
  SVB_CUDA(cudaMemsetAsync(X, 0, sizeof(double) * n, ctx->stream));

  for (int l = 0; l < p.mItr; l++) {
    if (!ns_inner) {
      r.dB = r.fNorm;
      r.itr++;
    }
    if (l == 0) {
      // X = 0: K X (+ coupled-face term) is exactly zero, u0 = R.
      SVB_CUDA(cudaMemcpyAsync(u, R, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      SVB_TRY(spmv_halo(ctx, dof, Val, X, u));
      SVB_TRY(add_bc_mul_device(ctx, 0, dof, X, u, d_scal));
      if (ns_inner) r.itr++;
      SVB_TRY(axpby(ctx, n, 1.0, R, -1.0, u, u));
    }
    if (ns_inner && anyCoupled) SVB_TRY(add_bc_mul_device(ctx, 1, dof, u, u, d_scal));
    SVB_TRY(norm_owned(ctx, dof, u, d_scal, &err[0]));
    if (err[0] == 0.0) {
      set_error("FSILS: A zero matrix norm has been computed. This is probably caused by ill-posed boundary conditions.");
      return SVB200_ERR_NUMERIC;
    }
    if (ns_inner) {
      if (l == 0) {
        eps = err[0];
        r.iNorm = eps;
        r.fNorm = eps;
        eps = std::max(p.absTol, p.relTol * eps);
      }
      r.dB = r.fNorm;
    }
    SVB_TRY(axpby(ctx, n, 1.0 / err[0], u, 0.0, nullptr, u));

    for (int i = 0; i < sD; i++) {
      if (!ns_inner) r.itr++;
      last_i = i;
      double* ui = u + (size_t)us * i;
      double* ui1 = u + (size_t)us * (i + 1);
      SVB_TRY(spmv_halo(ctx, dof, Val, ui, ui1));
      SVB_TRY(add_bc_mul_device(ctx, 0, dof, ui, ui1, d_scal));
      if (ns_inner) {
        r.itr++;
        if (anyCoupled) SVB_TRY(add_bc_mul_device(ctx, 1, dof, ui1, ui1, d_scal));
      }
      SVB_TRY(multi_dot(ctx, (long long)ctx->mynNo * dof, i + 2, u, us, ui1, d_h, true));
      SVB_CUDA(cudaMemcpyAsync(hp, d_h, sizeof(double) * (i + 2), cudaMemcpyDeviceToHost, ctx->stream));
      SVB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
      SVB_TRY(cgs_update(ctx, n, i + 1, u, us, ui1, d_h, d_hn));
      SVB_CUDA(cudaEventSynchronize(ctx->ev1));

      for (int j = 0; j <= i + 1; j++) H(j, i) = hp[j];
      for (int j = 0; j <= i; j++) H(i + 1, i) = H(i + 1, i) - H(j, i) * H(j, i);
      H(i + 1, i) = std::sqrt(std::fabs(H(i + 1, i)));

      for (int j = 0; j <= i - 1; j++) {
        const double tmp = c[j] * H(j, i) + s[j] * H(j + 1, i);
        H(j + 1, i) = -s[j] * H(j, i) + c[j] * H(j + 1, i);
        H(j, i) = tmp;
      }
      const double tmp = std::sqrt(H(i, i) * H(i, i) + H(i + 1, i) * H(i + 1, i));
      c[i] = H(i, i) / tmp;
      s[i] = H(i + 1, i) / tmp;
      H(i, i) = tmp;
      H(i + 1, i) = 0.0;
      err[i + 1] = -s[i] * err[i];
      err[i] = c[i] * err[i];
      if (!ns_inner && full && full->hist && full->hist_n < full->hist_cap) full->hist[full->hist_n++] = std::fabs(err[i + 1]);
      if (std::fabs(err[i + 1]) < eps) {
        r.success = 1;
        break;
      }
    }
    if (last_i >= sD) last_i = sD - 1;
    for (int i = 0; i <= last_i; i++) y[i] = err[i];
    for (int j = last_i; j >= 0; j--) {
      for (int k = j + 1; k <= last_i; k++) y[j] = y[j] - H(j, k) * y[k];
      y[j] = y[j] / H(j, j);
    }
    Coefs cf;
    for (int j = 0; j <= last_i; j++) cf.c[j] = y[j];
    SVB_TRY(lincomb(ctx, n, last_i + 1, cf, u, us, X));
    r.fNorm = std::fabs(err[last_i + 1]);
    if (r.success) break;
  }
  if (!ns_inner) {
    SVB_CUDA(cudaMemcpyAsync(R, X, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    SVB_CUDA(cudaStreamSynchronize(ctx->stream));
    r.callD = now_s() - t0;
  } else {
    SVB_CUDA(cudaStreamSynchronize(ctx->stream));
    r.callD = now_s() - t0 + r.callD;
  }
  r.dB = 10.0 * std::log(r.fNorm / r.dB);
  return SVB200_OK;
}

static int gmres_device(svb200_ctx* ctx, int dof, const svb200_sublsparams& p, svb200_sublsresult& r, const double* Val,
                        double* R, svb200_lsresult* full)
{
  const long long n = (long long)ctx->nNo * dof;
  if (p.sD < 1 || p.sD > 250) {
    set_error("svb200: Krylov space dimension must be in [1,250]");
    return SVB200_ERR_INVALID;
  }
  const long long us = (n + 1) & ~1ll;
  SVB_TRY(ensure_work(ctx, (size_t)us * (p.sD + 2) + SCAL_N));
  double* u = ctx->d_work;
  double* X = u + (size_t)us * (p.sD + 1);
  double* d_scal = X + us;
  return gmres_core(ctx, dof, p, r, Val, R, X, u, d_scal, false, full);
}

// ---- CG ------------------------------------------------------------------------------------------
static int cg_device(svb200_ctx* ctx, int dof, const svb200_sublsparams& p, svb200_sublsresult& r, const double* Val,
                     double* R, svb200_lsresult* full)
{
  const long long n = (long long)ctx->nNo * dof;
  const long long ns = (n + 1) & ~1ll;
  SVB_TRY(ensure_work(ctx, (size_t)ns * 3 + SCAL_N));
  double* P = ctx->d_work;
  double* KP = P + ns;
  double* X = KP + ns;
  double* d_scal = X + ns;
  const double t0 = now_s();
  r.success = 0;
  SVB_TRY(norm_owned(ctx, dof, R, d_scal, &r.iNorm));
  const double tol = std::max(p.absTol, p.relTol * r.iNorm);
  const double eps = tol * tol;
  double errO = r.iNorm * r.iNorm;
  double err = errO;
  SVB_CUDA(cudaMemsetAsync(X, 0, sizeof(double) * n, ctx->stream));
  SVB_CUDA(cudaMemcpyAsync(P, R, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  int last_i = 0;
  if (full) full->hist_n = 0;
  for (int i = 0; i < p.mItr; i++) {
    last_i = i;
    if (err < eps) {
      r.success = 1;
      break;
    }
    errO = err;
    SVB_TRY(spmv_halo(ctx, dof, Val, P, KP));
    double pkp;
    SVB_TRY(dot_owned(ctx, dof, P, KP, d_scal, &pkp));
    const double alpha = errO / pkp;
    SVB_TRY(axpby(ctx, n, alpha, P, 1.0, X, X));
    SVB_TRY(axpby(ctx, n, -alpha, KP, 1.0, R, R));
    SVB_TRY(norm_owned(ctx, dof, R, d_scal, &err));
    err = err * err;
    if (full && full->hist && full->hist_n < full->hist_cap) full->hist[full->hist_n++] = std::sqrt(err);
    // P = (P + (errO/err) R) * (err/errO): two roundings like omp_sum_v followed by omp_mul_v
    SVB_TRY(axpby(ctx, n, errO / err, R, 1.0, P, P));
    SVB_TRY(axpby(ctx, n, err / errO, P, 0.0, nullptr, P));
  }
  SVB_CUDA(cudaMemcpyAsync(R, X, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  r.itr = last_i;
  r.fNorm = std::sqrt(err);
  r.callD = now_s() - t0;
  r.dB = (errO < std::numeric_limits<double>::epsilon()) ? 0.0 : 5.0 * std::log(err / errO);
  return SVB200_OK;
}

// ---- BiCGStab -----------------------------------------------------------------------------------
static int bicgs_device(svb200_ctx* ctx, int dof, const svb200_sublsparams& p, svb200_sublsresult& r, const double* Val,
                        double* R, svb200_lsresult* full)
{
  const long long n = (long long)ctx->nNo * dof;
  const long long ns = (n + 1) & ~1ll;
  SVB_TRY(ensure_work(ctx, (size_t)ns * 6 + SCAL_N));
  double* P = ctx->d_work;
  double* Rh = P + ns;
  double* X = Rh + ns;
  double* V = X + ns;
  double* S = V + ns;
  double* T = S + ns;
  double* d_scal = T + ns;
  const double t0 = now_s();
  r.success = 0;
  double err;
  SVB_TRY(norm_owned(ctx, dof, R, d_scal, &err));
  double errO = err;
  r.iNorm = err;
  const double eps = std::max(p.absTol, p.relTol * err);
  double rho = err * err;
  double beta = rho;
  (void)beta;
  SVB_CUDA(cudaMemsetAsync(X, 0, sizeof(double) * n, ctx->stream));
  SVB_CUDA(cudaMemcpyAsync(P, R, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  SVB_CUDA(cudaMemcpyAsync(Rh, R, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  int i_itr = 1;
  if (full) full->hist_n = 0;
  for (int i = 0; i < p.mItr; i++) {
    if (err < eps) {
      r.success = 1;
      break;
    }
    SVB_TRY(spmv_halo(ctx, dof, Val, P, V));
    double rhv;
    SVB_TRY(dot_owned(ctx, dof, Rh, V, d_scal, &rhv));
    const double alpha = rho / rhv;
    SVB_TRY(axpby(ctx, n, 1.0, R, -alpha, V, S));
    SVB_TRY(spmv_halo(ctx, dof, Val, S, T));
    double omega, ts;
    SVB_TRY(norm_owned(ctx, dof, T, d_scal, &omega));
    SVB_TRY(dot_owned(ctx, dof, T, S, d_scal, &ts));
    omega = ts / (omega * omega);
    // X = X + alpha*P + omega*S
    SVB_TRY(axpby(ctx, n, alpha, P, 1.0, X, X));
    SVB_TRY(axpby(ctx, n, omega, S, 1.0, X, X));
    SVB_TRY(axpby(ctx, n, 1.0, S, -omega, T, R));
    errO = err;
    SVB_TRY(norm_owned(ctx, dof, R, d_scal, &err));
    if (full && full->hist && full->hist_n < full->hist_cap) full->hist[full->hist_n++] = err;
    const double rhoO = rho;
    SVB_TRY(dot_owned(ctx, dof, R, Rh, d_scal, &rho));
    beta = rho * alpha / (rhoO * omega);
    // P = R + beta*(P - omega*V)
    SVB_TRY(axpby(ctx, n, 1.0, P, -omega, V, P));
    SVB_TRY(axpby(ctx, n, 1.0, R, beta, P, P));
    i_itr++;
  }
  SVB_CUDA(cudaMemcpyAsync(R, X, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  r.itr = i_itr - 1;
  r.fNorm = err;
  r.callD = now_s() - t0;
  r.dB = (errO < std::numeric_limits<double>::epsilon()) ? 0.0 : 10.0 * std::log(err / errO);
  return SVB200_OK;
}

int ns_solver_device(svb200_ctx* ctx, int dof, const svb200_lsparams* ls, svb200_lsresult* res, double* Val, double* R);

// ---- fsils_solve ---------------------------------------------------------------------------------
int fsils_solve_device(svb200_ctx* ctx, int dof, int ls_type, int prec, const svb200_lsparams* ls, int nFaces,
                       const int* incL, const double* res, svb200_lsresult* result)
{
  SVB_TRY(ensure_scalars(ctx));
  // face flags (solve.cpp:45-85)
  bool anyNeu = false;
  for (size_t f = 0; f < ctx->face.size(); f++) {
    auto& face = ctx->face[f];
    face.incFlag = true;
    if (incL && (int)f < nFaces && incL[f] == 0) face.incFlag = false;
    if (face.set && face.bGrp == SVB200_BC_NEU) anyNeu = true;
  }
  if (anyNeu && res == nullptr) {
    set_error("[fsils_solve] res is required for Neu surfaces");
    return SVB200_ERR_INVALID;
  }
  for (size_t f = 0; f < ctx->face.size(); f++) {
    auto& face = ctx->face[f];
    face.coupledFlag = false;
    if (!face.set || !face.incFlag) continue;
    if (face.bGrp == SVB200_BC_NEU && (int)f < nFaces && res[f] != 0.0) {
      face.res = res[f];
      face.coupledFlag = true;
    }
  }
  const size_t nW = (size_t)ctx->nNo * dof;
  // d_W = [Wc | W1 | Wr | Wc_sweep | 8 scalars]; the diagonal preconditioner only uses the first slice
  const size_t nWs = (nW + 1) & ~(size_t)1;
  const size_t needW = (prec == SVB200_PREC_RCS) ? 4 * nWs + 8 : nW;
  if (needW > ctx->W_cap) {
    if (ctx->d_W) cudaFree(ctx->d_W);
    ctx->d_W = nullptr;
    ctx->W_cap = 0;
    SVB_CUDA(cudaMalloc(&ctx->d_W, sizeof(double) * std::max<size_t>(needW, 1)));
    ctx->W_cap = needW;
  }
  svb200_lsresult local{};
  svb200_lsresult* out = result ? result : &local;
  double* hist = out->hist;
  int hist_cap = out->hist_cap;
  *out = svb200_lsresult{};
  out->hist = hist;
  out->hist_cap = hist ? hist_cap : 0;

  if (prec == SVB200_PREC_RCS) {
    double* W = ctx->d_W;
    SVB_TRY(precond_rcs_device(ctx, dof, ctx->d_Val, ctx->d_R, W + nWs, W, W + 2 * nWs, W + 3 * nWs, W + 4 * nWs));
  } else {
    SVB_TRY(precond_diag_device(ctx, dof, ctx->d_Val, ctx->d_R, ctx->d_W));
  }

  switch (ls_type) {
    case SVB200_LS_NS:
      SVB_TRY(ns_solver_device(ctx, dof, ls, out, ctx->d_Val, ctx->d_R));
      break;
    case SVB200_LS_GMRES:
      SVB_TRY(gmres_device(ctx, dof, ls->RI, out->RI, ctx->d_Val, ctx->d_R, out));
      break;
    case SVB200_LS_CG:
      SVB_TRY(cg_device(ctx, dof, ls->RI, out->RI, ctx->d_Val, ctx->d_R, out));
      break;
    case SVB200_LS_BICGS:
      SVB_TRY(bicgs_device(ctx, dof, ls->RI, out->RI, ctx->d_Val, ctx->d_R, out));
      break;
    default:
      set_error("FSILS: LS_type not defined");
      return SVB200_ERR_INVALID;
  }
  // R = Wc o R (solve.cpp:157-159)
  SVB_TRY(hadamard(ctx, (long long)nW, ctx->d_W, ctx->d_R, ctx->d_R));
  return SVB200_OK;
}

// ---- NS solver (Schur complement) -------------------------------------------------------------------
// ge::ge (linear_solver/ge.cpp:13-112): diagonally scaled Gaussian elimination with partial pivoting on the
// leading N x N part of the (nV x nV, column-major) Gram matrix; tiny, stays on the host.
static bool ge_host(int nV, int N, const std::vector<double>& A, std::vector<double>& B)
{
  auto a = [&](int i, int j) { return A[(size_t)j * nV + i]; };
  std::vector<double> W(std::max(N, 1));
  const double tol = std::numeric_limits<double>::denorm_min();
  const double eps = std::numeric_limits<double>::epsilon();
  for (int i = 0; i < N; i++) {
    if (std::fabs(a(i, i)) < tol) { std::fill(B.begin(), B.end(), 0.0); return false; }
    W[i] = 1.0 / std::sqrt(std::fabs(a(i, i)));
  }
  std::vector<double> Cm((size_t)std::max(N, 1) * (N + 1));
  auto C = [&](int i, int j) -> double& { return Cm[(size_t)j * N + i]; };
  for (int i = 0; i < N; i++) {
    for (int j = 0; j < N; j++) C(i, j) = W[i] * W[j] * a(i, j);
    C(i, N) = W[i] * B[i];
  }
  if (N <= 0) return false;
  if (N == 1) {
    B[0] = C(0, 1) / C(0, 0);
    B[0] = B[0] * W[0];
    return true;
  }
  if (N == 2) {
    const double pivot = C(0, 0) * C(1, 1) - C(1, 0) * C(0, 1);
    if (std::fabs(pivot) < eps) { std::fill(B.begin(), B.end(), 0.0); return false; }
    B[0] = (C(0, 2) * C(1, 1) - C(1, 2) * C(0, 1)) / pivot;
    B[1] = (C(1, 2) * C(0, 0) - C(0, 2) * C(1, 0)) / pivot;
    B[0] = W[0] * B[0];
    B[1] = W[1] * B[1];
    return true;
  }
  for (int m = 0; m < N - 1; m++) {
    int ipv = m;
    double pivot = std::fabs(C(m, m));
    for (int i = m + 1; i < N; i++)
      if (std::fabs(C(i, m)) > pivot) { ipv = i; pivot = std::fabs(C(i, m)); }
    if (std::fabs(pivot) < eps) { std::fill(B.begin(), B.end(), 0.0); return false; }
    if (ipv != m)
      for (int j = m; j < N + 1; j++) std::swap(C(m, j), C(ipv, j));
    for (int i = m + 1; i < N; i++) {
      const double saveEl = C(i, m) / C(m, m);
      C(i, m) = 0.0;
      for (int j = m + 1; j < N + 1; j++) C(i, j) = C(i, j) - saveEl * C(m, j);
    }
  }
  for (int j = N - 1; j >= 0; j--) {
    for (int i = j + 1; i < N; i++) C(j, N) = C(j, N) - C(j, i) * C(i, N);
    C(j, N) = C(j, N) / C(j, j);
  }
  for (int i = 0; i < N; i++) B[i] = W[i] * C(i, N);
  return true;
}

// cgrad::schur (cgrad.cpp:23-133): CG on S p = L p - Gt (G p); R(nNo) in/out.
// work: X, P, SP, DGP (nNo each) and GP (nsd*nNo).
static int schur_device(svb200_ctx* ctx, int nsd, const svb200_sublsparams& p, svb200_sublsresult& r, const double* Gt,
                        const double* mG, const double* mL, const double* DL, double* R, double* work, double* d_scal)
{
  const long long nNo = ctx->nNo;
  const long long nns = (nNo + 1) & ~1ll;
  double* X = work;
  double* P = X + nns;
  double* SP = P + nns;
  double* DGP = SP + nns;
  double* GP = DGP + nns;
  bool anyCoupled = false;
  for (auto& f : ctx->face) anyCoupled |= (f.set && f.coupledFlag);
  const double t0 = now_s();
  r.success = 0;
  SVB_TRY(norm_owned(ctx, 1, R, d_scal, &r.iNorm));
  const double tol = std::max(p.absTol, p.relTol * r.iNorm);
  const double eps = tol * tol;
  double errO = r.iNorm * r.iNorm;
  double err = errO;
  SVB_CUDA(cudaMemsetAsync(X, 0, sizeof(double) * nNo, ctx->stream));
  SVB_CUDA(cudaMemcpyAsync(P, R, sizeof(double) * nNo, cudaMemcpyDeviceToDevice, ctx->stream));
  int last_i = 0;
  static const bool host_loop = getenv("SVB200_SCHUR_HOST_LOOP") != nullptr;     // A/B knob: the round-1 host-driven loop
  if (!host_loop) {
    // Device-resident loop: alpha, beta and the stopping test are evaluated on the device (fsils_kernels.cu, `cg` scalars);
    // the host enqueues iteration i+1 while iteration i runs and reads the status one iteration late, so the GPU never
    // waits for a host round trip (two per iteration in the host-driven loop, with ~13 small launches between them).
    if (err < eps) {
      r.success = 1;          // cgrad.cpp:72-75 at i = 0
    } else if (p.mItr > 0) {
      double* cg = ctx->d_cg;
      double* h0 = ctx->h_cg + 16;
      h0[0] = err; h0[1] = err; h0[2] = 0.0; h0[3] = eps; h0[4] = 0.0; h0[5] = 0.0; h0[6] = err; h0[7] = 0.0;
      SVB_CUDA(cudaMemcpyAsync(cg, h0, sizeof(double) * 8, cudaMemcpyHostToDevice, ctx->stream));
      bool done = false;
      // nsd = 3: the interleaved { Gt, L } operator (schur_sp4_kernel); on a single partition the vector part of the iteration is
      // fused as well: <p,Sp> partials in the operator's epilogue, X / R update + <r,r> partials, P update — 6 launches per
      // iteration instead of 10 (SVB200_SCHUR_UNFUSED=1: the separate kernels)
      static const bool unfused = getenv("SVB200_SCHUR_UNFUSED") != nullptr;
      const bool sp4 = (nsd == 3 && DL != nullptr && !unfused);
      const bool fused = sp4 && ctx->nranks == 1;
      double *part_psp = nullptr, *part_rr = nullptr;
      if (fused) {
        const size_t need = 148 * 8 * 2 + 16;
        if (need > ctx->red_cap) {
          if (ctx->d_red) cudaFree(ctx->d_red);
          ctx->red_cap = need * 2;
          SVB_CUDA(cudaMalloc(&ctx->d_red, sizeof(double) * ctx->red_cap));
        }
        part_psp = ctx->d_red;
        part_rr = ctx->d_red + 148 * 8;
      }
      for (int i = 0; i < p.mItr && !done; i++) {
        SVB_TRY(spmv_rc(ctx, nsd, 1, mG, P, GP));
        SVB_TRY(halo_sum(ctx, nsd, GP));
        if (anyCoupled) SVB_TRY(add_bc_mul_device(ctx, 1, nsd, GP, GP, d_scal));
        if (fused) {
          int nparts = 0;
          SVB_TRY(schur_sp4(ctx, -1, DL, P, GP, SP, part_psp, &nparts));
          SVB_TRY(schur_cg_fused_tail(ctx, cg, nparts, part_psp, SP, P, X, R, part_rr));
        } else {
          if (sp4) SVB_TRY(schur_sp4(ctx, -1, DL, P, GP, SP, nullptr, nullptr));
          else SVB_TRY(schur_sp(ctx, nsd, mL, Gt, P, GP, SP));
          SVB_TRY(halo_sum(ctx, 1, SP));
          SVB_TRY(multi_dot(ctx, (long long)ctx->mynNo, 1, P, 0, SP, cg + 2, true));
          SVB_TRY(cg_step_kernels(ctx, 0, nNo, cg, P, SP, X, R, nullptr));
          SVB_TRY(multi_dot(ctx, (long long)ctx->mynNo, 1, R, 0, R, cg + 1, true));
          SVB_TRY(cg_step_kernels(ctx, 1, nNo, cg, nullptr, nullptr, nullptr, R, P));
          SVB_TRY(cg_step_kernels(ctx, 2, 1, cg, nullptr, nullptr, nullptr, nullptr, nullptr));
        }
        SVB_CUDA(cudaMemcpyAsync(ctx->h_cg + 8 * (i & 1), cg, sizeof(double) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        SVB_CUDA(cudaEventRecord(ctx->ev_cg[i & 1], ctx->stream));
        if (i >= 1) {
          SVB_CUDA(cudaEventSynchronize(ctx->ev_cg[(i - 1) & 1]));
          done = ctx->h_cg[8 * ((i - 1) & 1) + 4] != 0.0;
        }
      }
      SVB_CUDA(cudaMemcpyAsync(h0, cg, sizeof(double) * 8, cudaMemcpyDeviceToHost, ctx->stream));
      SVB_CUDA(cudaStreamSynchronize(ctx->stream));
      const int executed = (int)h0[5];
      // the reference tests err < eps at the TOP of the next iteration: converging in the very last allowed iteration
      // is not reported as success, and last_i stays mItr-1 (cgrad.cpp:70-76, 128)
      r.success = (h0[4] != 0.0 && executed < p.mItr) ? 1 : 0;
      last_i = r.success ? executed : p.mItr - 1;
      err = std::sqrt(h0[1]);      // cg[1] is the raw <r,r>; the reference squares the norm (cgrad.cpp:92-93)
      err = err * err;
      errO = h0[6];
    }
  } else
  for (int i = 0; i < p.mItr; i++) {
    last_i = i;
    if (err < eps) {
      r.success = 1;
      break;
    }
    errO = err;
    SVB_TRY(spmv_rc(ctx, nsd, 1, mG, P, GP));
    SVB_TRY(halo_sum(ctx, nsd, GP));
    if (anyCoupled) SVB_TRY(add_bc_mul_device(ctx, 1, nsd, GP, GP, d_scal));
    SVB_TRY(spmv_rc(ctx, 1, nsd, Gt, GP, DGP));
    SVB_TRY(halo_sum(ctx, 1, DGP));
    SVB_TRY(spmv_rc(ctx, 1, 1, mL, P, SP));
    SVB_TRY(halo_sum(ctx, 1, SP));
    SVB_TRY(axpby(ctx, nNo, -1.0, DGP, 1.0, SP, SP));
    double psp;
    SVB_TRY(dot_owned(ctx, 1, P, SP, d_scal, &psp));
    const double alpha = errO / psp;
    SVB_TRY(axpby(ctx, nNo, alpha, P, 1.0, X, X));
    SVB_TRY(axpby(ctx, nNo, -alpha, SP, 1.0, R, R));
    SVB_TRY(norm_owned(ctx, 1, R, d_scal, &err));
    err = err * err;
    SVB_TRY(axpby(ctx, nNo, errO / err, R, 1.0, P, P));
    SVB_TRY(axpby(ctx, nNo, err / errO, P, 0.0, nullptr, P));
  }
  SVB_CUDA(cudaMemcpyAsync(R, X, sizeof(double) * nNo, cudaMemcpyDeviceToDevice, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  r.fNorm = std::sqrt(err);
  r.callD = now_s() - t0 + r.callD;
  r.itr = r.itr + last_i;
  r.dB = (errO < std::numeric_limits<double>::epsilon()) ? 0.0 : 5.0 * std::log(err / errO);
  return SVB200_OK;
}

// ns_solver::ns_solver (ns_solver.cpp:140-472).
int ns_solver_device(svb200_ctx* ctx, int dof, const svb200_lsparams* ls, svb200_lsresult* res, double* Val, double* Ri)
{
  const long long nNo = ctx->nNo, nnz = ctx->nnz;
  const int nsd = dof - 1;
  if (nsd != 3 && nsd != 2) {
    set_error("svb200: the NS solver needs dof = 3 (2-D) or 4 (3-D)");
    return SVB200_ERR_INVALID;
  }
  const int iBmax = ls->RI.mItr;
  const int nB = 2 * iBmax;
  const int sD = ls->GM.sD;
  if (sD < 1 || sD > 250 || iBmax < 1 || iBmax > 100) {
    set_error("svb200: NS solver needs 1 <= Krylov dimension <= 250 and 1 <= Max_iterations <= 100");
    return SVB200_ERR_INVALID;
  }
  const long long nv = nsd * nNo;
  const long long nvs = (nv + 1) & ~1ll, nns = (nNo + 1) & ~1ll;   // 16-byte aligned strides
  // workspace layout (doubles)
  size_t need = 0;
  auto take = [&](size_t cnt) { size_t o = need; need += (cnt + 1) & ~(size_t)1; return o; };
  const size_t oRm = take(nv), oRmi = take(nv), oRc = take(nNo), oRci = take(nNo);
  const size_t oU = take((size_t)nvs * iBmax), oMU = take((size_t)nvs * nB), oP = take((size_t)nns * iBmax), oMP = take((size_t)nns * nB);
  const size_t oK = take((size_t)nnz * nsd * nsd), oG = take((size_t)nnz * nsd), oD = take((size_t)nnz * nsd), oL = take(nnz),
               oGt = take((size_t)nnz * nsd), oDL = take(nsd == 3 ? (size_t)nnz * 4 : 0);
  // scalar scratch: SCAL_N for the solvers + the Gram-matrix partial results, 2 (4 i + 5) doubles in outer iteration i
  const size_t nGram = (size_t)2 * (4 * iBmax + 5);
  const size_t oGm = take((size_t)nvs * (sD + 1)), oSch = take((size_t)nns * 4 + nvs), oScal = take(SCAL_N + nGram);
  SVB_TRY(ensure_work(ctx, need));
  double* W = ctx->d_work;
  double *Rm = W + oRm, *Rmi = W + oRmi, *Rc = W + oRc, *Rci = W + oRci, *U = W + oU, *MU = W + oMU, *P = W + oP, *MP = W + oMP;
  double *mK = W + oK, *mG = W + oG, *mD = W + oD, *mL = W + oL, *Gt = W + oGt, *gm_u = W + oGm, *sch = W + oSch, *d_scal = W + oScal;
  double* DL = nsd == 3 ? W + oDL : nullptr;
  double* d_gram = d_scal + SCAL_N;  // Gram-matrix partial results, sized from Max_iterations above
  if (!ctx->d_tslot) {
    SVB_CUDA(cudaMalloc(&ctx->d_tslot, sizeof(int) * std::max<long long>(nnz, 1)));
    SVB_TRY(build_transpose_slots(ctx, ctx->d_tslot));
  }
  svb200_sublsresult &RI = res->RI, &GM = res->GM, &CG = res->CG;
  const double t0 = now_s();

  SVB_TRY(ns_split(ctx, dof, Ri, Rmi, Rci));
  SVB_CUDA(cudaMemcpyAsync(Rm, Rmi, sizeof(double) * nv, cudaMemcpyDeviceToDevice, ctx->stream));
  SVB_CUDA(cudaMemcpyAsync(Rc, Rci, sizeof(double) * nNo, cudaMemcpyDeviceToDevice, ctx->stream));
  double nm, nc;
  SVB_TRY(norm_owned(ctx, nsd, Rm, d_scal, &nm));
  SVB_TRY(norm_owned(ctx, 1, Rc, d_scal, &nc));
  double eps = std::sqrt(nm * nm + nc * nc);
  RI.iNorm = eps;
  RI.fNorm = eps * eps;
  CG.callD = 0.0; GM.callD = 0.0;
  CG.itr = 0; GM.itr = 0;
  RI.success = 0;
  eps = std::max(ls->RI.absTol, ls->RI.relTol * eps);
  SVB_TRY(ns_depart(ctx, nsd, Val, ctx->d_tslot, mK, mG, mD, mL, Gt, DL));
  SVB_TRY(bc_pre_device(ctx, dof, d_scal));     // nS over the first nsd = dof-1 components (ns_solver.cpp:29-58)

  std::vector<double> A((size_t)nB * nB, 0.0), B(nB, 0.0), xB(nB, 0.0), oldxB(nB, 0.0);
  int iBB = 0, i_count = 0;
  for (int i = 0; i < iBmax; i++) {
    int iB = 2 * i;
    iBB = 2 * i + 1;
    RI.dB = RI.fNorm;
    i_count = i;
    double* Ui = U + (size_t)nvs * i;
    double* Pi = P + (size_t)nns * i;
    double* MU_iB = MU + (size_t)nvs * iB;
    double* MU_iBB = MU + (size_t)nvs * iBB;
    double* MP_iB = MP + (size_t)nns * iB;
    double* MP_iBB = MP + (size_t)nns * iBB;
    // U = K^-1 Rm
    SVB_TRY(gmres_core(ctx, nsd, ls->GM, GM, mK, Rm, Ui, gm_u, d_scal, true, nullptr));
    // P = Rc - D U ; P = S^-1 P
    SVB_TRY(spmv_rc(ctx, 1, nsd, mD, Ui, Pi));
    SVB_TRY(halo_sum(ctx, 1, Pi));
    SVB_TRY(axpby(ctx, nNo, 1.0, Rc, -1.0, Pi, Pi));
    SVB_TRY(schur_device(ctx, nsd, ls->CG, CG, Gt, mG, mL, DL, Pi, sch, d_scal));
    // MU(iB) = G P ; MU(iBB) = Rm - G P ; U = K^-1 MU(iBB)
    SVB_TRY(spmv_rc(ctx, nsd, 1, mG, Pi, MU_iB));
    SVB_TRY(halo_sum(ctx, nsd, MU_iB));
    SVB_TRY(axpby(ctx, nv, 1.0, Rm, -1.0, MU_iB, MU_iBB));
    SVB_TRY(gmres_core(ctx, nsd, ls->GM, GM, mK, MU_iBB, Ui, gm_u, d_scal, true, nullptr));
    // MU(iBB) = K U + bc(U) ; MP(iB) = L P ; MP(iBB) = D U
    SVB_TRY(launch_spmv(ctx, nsd, mK, Ui, MU_iBB));
    SVB_TRY(halo_sum(ctx, nsd, MU_iBB));
    SVB_TRY(add_bc_mul_device(ctx, 0, nsd, Ui, MU_iBB, d_scal));
    SVB_TRY(spmv_rc(ctx, 1, 1, mL, Pi, MP_iB));
    SVB_TRY(halo_sum(ctx, 1, MP_iB));
    SVB_TRY(spmv_rc(ctx, 1, nsd, mD, Ui, MP_iBB));
    SVB_TRY(halo_sum(ctx, 1, MP_iBB));
    // Gram matrix columns iB, iBB: [ <MU_j,MU_k> (j<=k), <MU_k,Rmi> ] and the same with MP / Rci
    int cnt = 0;
    for (int k = iB; k <= iBB; k++) {
      SVB_TRY(multi_dot(ctx, (long long)ctx->mynNo * nsd, k + 1, MU, nvs, MU + (size_t)nvs * k, d_gram + cnt));
      SVB_TRY(multi_dot(ctx, (long long)ctx->mynNo * nsd, 1, Rmi, 0, MU + (size_t)nvs * k, d_gram + cnt + k + 1));
      cnt += k + 2;
    }
    const int half = cnt;
    for (int k = iB; k <= iBB; k++) {
      SVB_TRY(multi_dot(ctx, (long long)ctx->mynNo, k + 1, MP, nns, MP + (size_t)nns * k, d_gram + cnt));
      SVB_TRY(multi_dot(ctx, (long long)ctx->mynNo, 1, Rci, 0, MP + (size_t)nns * k, d_gram + cnt + k + 1));
      cnt += k + 2;
    }
    if ((size_t)cnt > nGram || cnt > SCAL_N) {      // h_pinned holds SCAL_N doubles; 8 i + 10 <= 802 for i < 100
      set_error("svb200: NS Gram buffer overflow");
      return SVB200_ERR_INVALID;
    }
    SVB_TRY(allreduce_sum(ctx, d_gram, cnt));
    SVB_TRY(fetch(ctx, d_gram, cnt, ctx->h_pinned));
    {
      const double* g = ctx->h_pinned;
      int c = 0;
      for (int k = iB; k <= iBB; k++) {
        for (int j = 0; j <= k; j++) {
          const double v = g[c] + g[half + c];
          A[(size_t)k * nB + j] = v;
          A[(size_t)j * nB + k] = v;
          c++;
        }
        B[k] = g[c] + g[half + c];
        c++;
      }
    }
    xB = B;
    if (ge_host(nB, iBB + 1, A, xB)) {
      oldxB = xB;
    } else {
      // the reference throws on the master rank (ns_solver.cpp:346-348), which aborts the run
      set_error("FSILS: Singular matrix detected");
      return SVB200_ERR_NUMERIC;
    }
    double sum = 0.0;
    for (int q = 0; q <= iBB; q++) sum += xB[q] * B[q];
    RI.fNorm = RI.iNorm * RI.iNorm - sum;
    if (RI.fNorm < eps * eps) {
      RI.success = 1;
      break;
    }
    // Rm = Rmi - sum_j xB_j MU_j ; Rc = Rci - sum_j xB_j MP_j
    Coefs cf;
    for (int j = 0; j <= iBB; j++) cf.c[j] = -xB[j];
    SVB_CUDA(cudaMemcpyAsync(Rm, Rmi, sizeof(double) * nv, cudaMemcpyDeviceToDevice, ctx->stream));
    SVB_CUDA(cudaMemcpyAsync(Rc, Rci, sizeof(double) * nNo, cudaMemcpyDeviceToDevice, ctx->stream));
    SVB_TRY(lincomb(ctx, nv, iBB + 1, cf, MU, nvs, Rm));
    SVB_TRY(lincomb(ctx, nNo, iBB + 1, cf, MP, nns, Rc));
  }
  RI.itr = i_count;
  {
    Coefs cf;
    for (int j = 0; j <= iBB; j++) cf.c[j] = -xB[j];
    SVB_CUDA(cudaMemcpyAsync(Rc, Rci, sizeof(double) * nNo, cudaMemcpyDeviceToDevice, ctx->stream));
    SVB_TRY(lincomb(ctx, nNo, iBB + 1, cf, MP, nns, Rc));
  }
  double nrc;
  SVB_TRY(norm_owned(ctx, 1, Rc, d_scal, &nrc));
  res->Resc = static_cast<int>(100.0 * nrc * nrc / RI.fNorm);
  res->Resm = 100 - res->Resc;
  // solution: Rmi = sum_i xB(2i+1) U_i ; Rci = sum_i xB(2i) P_i
  {
    Coefs cu, cp;
    for (int i = 0; i <= RI.itr; i++) { cu.c[i] = xB[2 * i + 1]; cp.c[i] = xB[2 * i]; }
    SVB_CUDA(cudaMemsetAsync(Rmi, 0, sizeof(double) * nv, ctx->stream));
    SVB_CUDA(cudaMemsetAsync(Rci, 0, sizeof(double) * nNo, ctx->stream));
    SVB_TRY(lincomb(ctx, nv, RI.itr + 1, cu, U, nvs, Rmi));
    SVB_TRY(lincomb(ctx, nNo, RI.itr + 1, cp, P, nns, Rci));
  }
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  RI.callD = now_s() - t0;
  RI.dB = 5.0 * std::log(RI.fNorm / RI.dB);
  if (res->Resc < 0 || res->Resm < 0) {
    res->Resc = 0;
    res->Resm = 0;
    RI.dB = 0;
    RI.fNorm = 0.0;
  }
  RI.fNorm = std::sqrt(RI.fNorm);
  SVB_TRY(ns_merge(ctx, dof, Rmi, Rci, Ri));
  return SVB200_OK;
}

}  // namespace svb
