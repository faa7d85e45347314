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
  SVB_TRY(multi_dot(ctx, (long long)ctx->mynNo * dof, 1, a, 0, b, d_scal));
  SVB_TRY(allreduce_sum(ctx, d_scal, 1));
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
__global__ void face_dot_kernel(int fnNo, int fdof, int nd, int dof, int mynNo, int shared, const int* __restrict__ glob,
                                const double* __restrict__ valM, const double* __restrict__ X, double* __restrict__ out)
{
  __shared__ double red[256];
  double s = 0.0;
  for (int t = threadIdx.x; t < fnNo * nd; t += blockDim.x) {
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
  if (threadIdx.x == 0) *out = red[0];
}

__global__ void face_axpy_kernel(int fnNo, int fdof, int nd, int dof, double coef, const double* __restrict__ S,
                                 const int* __restrict__ glob, const double* __restrict__ valM, double* __restrict__ Y)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= fnNo * nd) return;
  const int a = t / nd, i = t % nd;
  Y[(size_t)glob[a] * dof + i] += valM[(size_t)a * fdof + i] * (coef * (*S));
}

// op: 0 = BCOP_TYPE_ADD (coef = res), 1 = BCOP_TYPE_PRE (coef = -res/(1+res*nS)).
int add_bc_mul_device(svb200_ctx* ctx, int op, int dof, const double* X, double* Y, double* d_scal)
{
  for (auto& f : ctx->face) {
    if (!f.set || !f.coupledFlag) continue;
    const int nd = std::min(f.dof, dof);
    const double coef = (op == 0) ? f.res : -f.res / (1.0 + f.res * f.nS);
    face_dot_kernel<<<1, 256, 0, ctx->stream>>>(f.nNo, f.dof, nd, dof, ctx->mynNo, f.shared, f.d_glob, f.d_valM, X, d_scal);
    ctx->launches++;
    if (f.shared) SVB_TRY(allreduce_sum(ctx, d_scal, 1));
    const int n = f.nNo * nd;
    if (n > 0) {
      face_axpy_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(f.nNo, f.dof, nd, dof, coef, d_scal, f.d_glob, f.d_valM, Y);
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
  SVB_TRY(precond_scale_matrix(ctx, dof, W, Val));
  SVB_TRY(hadamard(ctx, n, W, R, R));
  for (auto& f : ctx->face) {
    if (!f.set || !f.coupledFlag) continue;
    SVB_TRY(precond_face_valm(ctx, f, dof, W));
  }
  return SVB200_OK;
}

// ---- GMRES -------------------------------------------------------------------------------------
static int gmres_device(svb200_ctx* ctx, int dof, const svb200_sublsparams& p, svb200_sublsresult& r, const double* Val,
                        double* R, svb200_lsresult* full)
{
  const long long n = (long long)ctx->nNo * dof;
  const int sD = p.sD;
  if (sD < 1 || sD > 250) {
    set_error("svb200: Krylov space dimension must be in [1,250]");
    return SVB200_ERR_INVALID;
  }
  // workspace: u (sD+1 vectors) | X | scalars
  SVB_TRY(ensure_work(ctx, (size_t)n * (sD + 2) + SCAL_N));
  double* u = ctx->d_work;
  double* X = u + (size_t)n * (sD + 1);
  double* d_scal = X + n;
  double* d_h = d_scal + 8;        // Hessenberg column (sD+2)
  double* d_hn = d_scal + 4;
  double* hp = ctx->h_pinned;

  std::vector<double> h((size_t)(sD + 1) * sD, 0.0), y(sD), c(sD), s(sD), err(sD + 1, 0.0);
  auto H = [&](int i, int j) -> double& { return h[(size_t)j * (sD + 1) + i]; };

  const double t0 = now_s();
  r.success = 0;
  double eps;
  SVB_TRY(norm_owned(ctx, dof, R, d_scal, &eps));
  r.iNorm = eps;
  r.fNorm = eps;
  eps = std::max(p.absTol, p.relTol * eps);
  r.itr = 0;
  int last_i = 0;
  SVB_TRY(bc_pre_device(ctx, dof, d_scal));
  if (full) full->hist_n = 0;

  if (r.iNorm <= p.absTol) {
    r.callD = std::numeric_limits<double>::epsilon();
    r.dB = 0.0;
    r.success = 1;
    return SVB200_OK;     // R is left untouched, as in the reference (gmres.cpp:470-475)
  }
  SVB_CUDA(cudaMemsetAsync(X, 0, sizeof(double) * n, ctx->stream));

  for (int l = 0; l < p.mItr; l++) {
    r.dB = r.fNorm;
    r.itr++;
    if (l == 0) {
      // X = 0: K X (+ coupled-face term) is exactly zero, u0 = R.
      SVB_CUDA(cudaMemcpyAsync(u, R, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      SVB_TRY(spmv_halo(ctx, dof, Val, X, u));
      SVB_TRY(add_bc_mul_device(ctx, 0, dof, X, u, d_scal));
      SVB_TRY(axpby(ctx, n, 1.0, R, -1.0, u, u));
    }
    SVB_TRY(norm_owned(ctx, dof, u, d_scal, &err[0]));
    if (err[0] == 0.0) {
      set_error("FSILS: A zero matrix norm has been computed. This is probably caused by ill-posed boundary conditions.");
      return SVB200_ERR_NUMERIC;
    }
    SVB_TRY(axpby(ctx, n, 1.0 / err[0], u, 0.0, nullptr, u));

    for (int i = 0; i < sD; i++) {
      r.itr++;
      last_i = i;
      double* ui = u + (size_t)n * i;
      double* ui1 = u + (size_t)n * (i + 1);
      SVB_TRY(spmv_halo(ctx, dof, Val, ui, ui1));
      SVB_TRY(add_bc_mul_device(ctx, 0, dof, ui, ui1, d_scal));
      SVB_TRY(multi_dot(ctx, (long long)ctx->mynNo * dof, i + 2, u, n, ui1, d_h));
      SVB_TRY(allreduce_sum(ctx, d_h, i + 2));
      SVB_CUDA(cudaMemcpyAsync(hp, d_h, sizeof(double) * (i + 2), cudaMemcpyDeviceToHost, ctx->stream));
      SVB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
      SVB_TRY(cgs_update(ctx, n, i + 1, u, n, ui1, d_h, d_hn));
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
      if (full && full->hist && full->hist_n < full->hist_cap) full->hist[full->hist_n++] = std::fabs(err[i + 1]);
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
    SVB_TRY(lincomb(ctx, n, last_i + 1, cf, u, n, X));
    r.fNorm = std::fabs(err[last_i + 1]);
    if (r.success) break;
  }
  SVB_CUDA(cudaMemcpyAsync(R, X, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  r.callD = now_s() - t0;
  r.dB = 10.0 * std::log(r.fNorm / r.dB);
  return SVB200_OK;
}

// ---- CG ------------------------------------------------------------------------------------------
static int cg_device(svb200_ctx* ctx, int dof, const svb200_sublsparams& p, svb200_sublsresult& r, const double* Val,
                     double* R, svb200_lsresult* full)
{
  const long long n = (long long)ctx->nNo * dof;
  SVB_TRY(ensure_work(ctx, (size_t)n * 3 + SCAL_N));
  double* P = ctx->d_work;
  double* KP = P + n;
  double* X = KP + n;
  double* d_scal = X + n;
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
  SVB_TRY(ensure_work(ctx, (size_t)n * 6 + SCAL_N));
  double* P = ctx->d_work;
  double* Rh = P + n;
  double* X = Rh + n;
  double* V = X + n;
  double* S = V + n;
  double* T = S + n;
  double* d_scal = T + n;
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
int fsils_solve_device(svb200_ctx* ctx, int dof, int ls_type, const svb200_lsparams* ls, int nFaces, const int* incL,
                       const double* res, svb200_lsresult* result)
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
  if (nW > ctx->W_cap) {
    if (ctx->d_W) cudaFree(ctx->d_W);
    ctx->d_W = nullptr;
    SVB_CUDA(cudaMalloc(&ctx->d_W, sizeof(double) * std::max<size_t>(nW, 1)));
    ctx->W_cap = nW;
  }
  svb200_lsresult local{};
  svb200_lsresult* out = result ? result : &local;
  double* hist = out->hist;
  int hist_cap = out->hist_cap;
  *out = svb200_lsresult{};
  out->hist = hist;
  out->hist_cap = hist ? hist_cap : 0;

  SVB_TRY(precond_diag_device(ctx, dof, ctx->d_Val, ctx->d_R, ctx->d_W));

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

// Placeholder until the NS (Schur complement) solver lands: fail loudly, never fall back.
int ns_solver_device(svb200_ctx*, int, const svb200_lsparams*, svb200_lsresult*, double*, double*)
{
  set_error("svb200: LS type NS is not implemented yet");
  return SVB200_ERR_UNSUPPORTED;
}

}  // namespace svb
