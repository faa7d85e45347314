// fsils_kernels.h — launch wrappers of fsils_kernels.cu / comm.cu used by the solver drivers.
#pragma once
#include "svb200_internal.h"

namespace svb {

struct Coefs { double c[256]; };

// reduce_ranks: d_out is summed over the ranks as well (fused into the second stage on the p2p transport)
int multi_dot(svb200_ctx* ctx, long long n, int nvec, const double* ubase, long long stride, const double* v, double* d_out,
              bool reduce_ranks = false);
int cgs_update(svb200_ctx* ctx, long long n, int nprev, const double* ubase, long long stride, double* v, double* d_h, double* d_hn);
int lincomb(svb200_ctx* ctx, long long n, int ny, const Coefs& y, const double* ubase, long long stride, double* X);
int axpby(svb200_ctx* ctx, long long n, double a, const double* x, double b, const double* y, double* z);
int hadamard(svb200_ctx* ctx, long long n, const double* x, const double* y, double* z);
int precond_extract_diag(svb200_ctx* ctx, int dof, const double* Val, double* W);
int precond_invsqrt(svb200_ctx* ctx, int dof, double* W);
int precond_face_scale(svb200_ctx* ctx, const Face& f, int dof, double* W);
int precond_face_valm(svb200_ctx* ctx, const Face& f, int dof, const double* W);
int precond_scale_matrix(svb200_ctx* ctx, int dof, const double* Wr, const double* Wc, double* Val);
int fill(svb200_ctx* ctx, long long n, double* W, double v);
int rcs_renorm(svb200_ctx* ctx, long long n, double* W);
int rcs_diag_one(svb200_ctx* ctx, int dof, const double* Wr, double* Val);
int rcs_rowcol_max(svb200_ctx* ctx, int dof, const double* Val, double* Wr, double* Wc);
int rcs_dev_from_one(svb200_ctx* ctx, long long n, const double* Wr, const double* Wc, double* d_out);
int rcs_invsqrt_acc(svb200_ctx* ctx, long long n, double* W, double* Wacc);

int spmv_rc(svb200_ctx* ctx, int R, int C, const double* K, const double* U, double* KU);
int schur_sp(svb200_ctx* ctx, int nsd, const double* L, const double* D, const double* P, const double* GP, double* SP);
int cg_step_kernels(svb200_ctx* ctx, int which, long long n, double* cg, const double* P, const double* SP, double* X, double* R,
                    double* Pw);
int build_transpose_slots(svb200_ctx* ctx, int* d_tslot);
int ns_depart(svb200_ctx* ctx, int nsd, const double* Val, const int* d_tslot, double* mK, double* mG, double* mD, double* mL, double* Gt,
              double* DL);   // DL (4, nnz) = { Gt(0..2), L } interleaved, nsd = 3 only (else nullptr)
// spmv_lanegroup.cu
int spmv_rc_variant(svb200_ctx* ctx, int R, int C, int variant, const double* K, const double* U, double* KU);
int spmv_rc_num_variants(int R, int C);
int schur_sp4(svb200_ctx* ctx, int variant, const double* DL, const double* P, const double* GP, double* SP, double* part, int* nparts);
int schur_sp4_num_variants();
int schur_cg_fused_tail(svb200_ctx* ctx, double* cg, int npart_psp, double* part_psp, const double* SP, double* P, double* X, double* R,
                        double* part_rr);
int ns_split(svb200_ctx* ctx, int dof, const double* Ri, double* Rm, double* Rc);
int ns_merge(svb200_ctx* ctx, int dof, const double* Rm, const double* Rc, double* Ri);

// comm.cu: shared-node sums and scalar all-reduces (no-ops for a single partition)
int halo_sum(svb200_ctx* ctx, int dof, double* V);
int allreduce_sum(svb200_ctx* ctx, double* d_buf, int n);
int spmv4_halo_fused(svb200_ctx* ctx, const double* Val, const double* U, double* KU, bool* handled);
int dot_stage2_allreduce(svb200_ctx* ctx, int nblocks, int nvec, const double* d_part, double* d_out, bool* handled);
int launch_spmv4_rows(svb200_ctx* ctx, int row0, int nrows, const double* Val, const double* U, double* KU);

}  // namespace svb
