// assemble_bnd.cu — Neumann / traction boundary-face integrals on the device.
//
// Replaces eq_assem::b_assem_neu_bc (Code/Source/solver/eq_assem.cpp:31-149) with its callees nn::gnnb
// (Code/Source/solver/nn.cpp:911-1117; non-shell branch), fluid::b_fluid incl. the backflow stabilisation
// and its tangent (Code/Source/solver/fluid.cpp:21-108), l_elas::b_l_elas (Code/Source/solver/l_elas.cpp:21-32)
// and the do_assem scatter.  One thread per face element; faces are O(nEl^(2/3)) so this kernel is
// latency-trivial next to the volume assembly — what it buys is that no host-assembled patch has to be
// uploaded into R / Val inside the Newton iteration.
#include "svb200_internal.h"

namespace svb {

struct BndArgs {
  const int* IENb;     // (eNoNb, nElb) internal node ids
  const int* gE;       // (nElb) parent element
  const int* IEN;      // parent mesh connectivity (eNoN, nEl)
  const int* eId;
  const int* rowPtr;
  const int* colPtr;
  const double* x;
  const double* Yg;
  const double* Do;    // old displacement (moving mesh) or null
  const double* hg;    // (nNo) internal order
  double* R;
  double* Val;
  int* err;
  int nElb, eNoN, nGb, tDof, dof, mvMsh, nDmn, pad;
  double dt, af, gam;
  double w[4], N[4][4], Nx[4][4][2];   // [g][a], [g][a][i]
  struct { double rho, backflow; int Id, isFluid; } dmn[MAX_DMN];
};

template <int ENB>
__global__ void __launch_bounds__(128)
assemble_neu_kernel(const __grid_constant__ BndArgs P)
{
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.nElb) return;
  const int Ec = P.gE[e];
  int iD = 0;
  for (int d = 0; d < P.nDmn; d++) {
    iD = d;
    if (P.dmn[d].Id == -1) break;
    if (P.eId != nullptr && ((P.eId[Ec] >> P.dmn[d].Id) & 1)) break;
  }
  // nn::gnnb: face nodes first, then the element's other nodes; the first "other" node orients the normal
  int fn[ENB];
#pragma unroll
  for (int a = 0; a < ENB; a++) fn[a] = P.IENb[(size_t)e * ENB + a];
  unsigned used = 0;
  for (int a = 0; a < ENB; a++) {
    int found = -1;
    for (int b = 0; b < P.eNoN; b++)
      if (!((used >> b) & 1) && P.IEN[(size_t)Ec * P.eNoN + b] == fn[a]) { found = b; break; }
    if (found < 0) { atomicExch(P.err, 1); return; }
    used |= 1u << found;
  }
  int inner = -1;
  for (int b = 0; b < P.eNoN; b++)
    if (!((used >> b) & 1)) { inner = P.IEN[(size_t)Ec * P.eNoN + b]; break; }
  if (inner < 0) { atomicExch(P.err, 1); return; }
  auto coord = [&](int n, int i) {
    double v = __ldg(P.x + 3 * (size_t)n + i);
    if (P.mvMsh) v += __ldg(P.Do + (size_t)P.tDof * n + 4 + i);
    return v;
  };
  double lX[ENB][3], xin[3], hl[ENB], yl[ENB][3];
#pragma unroll
  for (int a = 0; a < ENB; a++) {
    hl[a] = __ldg(P.hg + fn[a]);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      lX[a][i] = coord(fn[a], i);
      const double* yp = P.Yg + (size_t)P.tDof * fn[a];
      yl[a][i] = __ldg(yp + i) - (P.mvMsh ? __ldg(yp + 4 + i) : 0.0);
    }
  }
#pragma unroll
  for (int i = 0; i < 3; i++) xin[i] = coord(inner, i);

  const bool fluid = P.dmn[iD].isFluid != 0;
  double lR[ENB][3], lKd[ENB][ENB];
#pragma unroll
  for (int a = 0; a < ENB; a++) {
#pragma unroll
    for (int i = 0; i < 3; i++) lR[a][i] = 0.0;
#pragma unroll
    for (int b = 0; b < ENB; b++) lKd[a][b] = 0.0;
  }
  for (int g = 0; g < P.nGb; g++) {
    double t0[3] = {0, 0, 0}, t1[3] = {0, 0, 0};   // xXi(:,0), xXi(:,1)
#pragma unroll
    for (int a = 0; a < ENB; a++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        t0[j] += P.Nx[g][a][0] * lX[a][j];
        t1[j] += P.Nx[g][a][1] * lX[a][j];
      }
    double n[3] = {t0[1] * t1[2] - t0[2] * t1[1], t0[2] * t1[0] - t0[0] * t1[2], t0[0] * t1[1] - t0[1] * t1[0]};
    const double sgn = n[0] * (lX[0][0] - xin[0]) + n[1] * (lX[0][1] - xin[1]) + n[2] * (lX[0][2] - xin[2]);
    if (sgn < 0.0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
    const double Jac = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    const double nV[3] = {n[0] / Jac, n[1] / Jac, n[2] / Jac};
    const double w = P.w[g] * Jac;
    double h = 0.0, u[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < ENB; a++) {
      h += P.N[g][a] * hl[a];
#pragma unroll
      for (int i = 0; i < 3; i++) u[i] += P.N[g][a] * yl[a][i];
    }
    double hc[3] = {h * nV[0], h * nV[1], h * nV[2]};
    if (fluid) {
      double udn = u[0] * nV[0] + u[1] * nV[1] + u[2] * nV[2];
      udn = 0.5 * P.dmn[iD].backflow * P.dmn[iD].rho * (udn - fabs(udn));
      const double wl = w * P.af * P.gam * P.dt;
#pragma unroll
      for (int i = 0; i < 3; i++) hc[i] += udn * u[i];
#pragma unroll
      for (int a = 0; a < ENB; a++)
#pragma unroll
        for (int b = 0; b < ENB; b++) lKd[a][b] -= wl * P.N[g][a] * P.N[g][b] * udn;
    }
#pragma unroll
    for (int a = 0; a < ENB; a++)
#pragma unroll
      for (int i = 0; i < 3; i++) lR[a][i] -= w * P.N[g][a] * hc[i];
  }
  // scatter (do_assem): residual rows 0..2, tangent entries (i,i), i < 3 of every face node pair
  const int dof = P.dof;
#pragma unroll
  for (int a = 0; a < ENB; a++) {
#pragma unroll
    for (int i = 0; i < 3; i++) atomicAdd(P.R + (size_t)dof * fn[a] + i, lR[a][i]);
    if (!fluid) continue;
#pragma unroll
    for (int b = 0; b < ENB; b++) {
      if (lKd[a][b] == 0.0) continue;
      int s = -1;
      for (int k = P.rowPtr[fn[a]]; k < P.rowPtr[fn[a] + 1]; k++)
        if (P.colPtr[k] == fn[b]) { s = k; break; }
      if (s < 0) { atomicExch(P.err, 1); return; }
      double* v = P.Val + (size_t)dof * dof * s;
      for (int i = 0; i < 3; i++) atomicAdd(v + (dof + 1) * i, lKd[a][b]);
    }
  }
}

int run_assemble_neu(svb200_ctx* ctx, const BFace& f, const svb200_eqparams* eq, const svb200_dmnparams* dmn, int nDmn,
                     const double* d_hg)
{
  SVB_REQUIRE(f.iM >= 0 && f.iM < (int)ctx->mesh.size() && ctx->mesh[f.iM].set, "svb200_assemble_neu: parent mesh not set");
  SVB_REQUIRE(nDmn >= 1 && nDmn <= MAX_DMN, "svb200_assemble_neu: between 1 and 8 domains are supported");
  SVB_REQUIRE(eq->dof == ctx->dof && eq->dof >= 3, "svb200_assemble_neu: dof mismatch with svb200_alloc");
  SVB_REQUIRE(eq->tDof == ctx->tDof && ctx->d_Yg && ctx->d_x, "svb200_assemble_neu: state / coordinates not set");
  SVB_REQUIRE(!eq->mvMsh || (ctx->d_Do && eq->tDof >= 7), "svb200_assemble_neu: moving mesh needs the old displacement");
  if (f.nElb == 0) return SVB200_OK;
  const Mesh& m = ctx->mesh[f.iM];
  BndArgs A;
  memset(&A, 0, sizeof(A));
  A.IENb = f.d_IENb; A.gE = f.d_gE; A.IEN = m.d_IEN; A.eId = m.d_eId; A.rowPtr = ctx->d_rowPtr; A.colPtr = ctx->d_colPtr;
  A.x = ctx->d_x; A.Yg = ctx->d_Yg; A.Do = ctx->d_Do; A.hg = d_hg; A.R = ctx->d_R; A.Val = ctx->d_Val;
  A.nElb = f.nElb; A.eNoN = m.eNoN; A.nGb = f.nGb; A.tDof = eq->tDof; A.dof = eq->dof; A.mvMsh = eq->mvMsh; A.nDmn = nDmn;
  A.dt = eq->dt; A.af = eq->af; A.gam = eq->gam;
  for (int g = 0; g < f.nGb; g++) {
    A.w[g] = f.w[g];
    for (int a = 0; a < f.eNoNb; a++) {
      A.N[g][a] = f.N[(size_t)g * f.eNoNb + a];
      for (int i = 0; i < 2; i++) A.Nx[g][a][i] = f.Nx[((size_t)g * f.eNoNb + a) * 2 + i];
    }
  }
  bool whole = false;
  for (int d = 0; d < nDmn; d++) {
    A.dmn[d].rho = dmn[d].rho; A.dmn[d].backflow = dmn[d].backflow_stab; A.dmn[d].Id = dmn[d].Id;
    A.dmn[d].isFluid = (dmn[d].phys == SVB200_PHYS_FLUID);
    whole |= (dmn[d].Id == -1);
  }
  if (!whole && !m.d_eId) { set_error("eId is not allocated"); return SVB200_ERR_INVALID; }
  int* d_err = nullptr;
  SVB_CUDA(cudaMalloc(&d_err, sizeof(int)));
  SVB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream));
  A.err = d_err;
  const int blocks = (f.nElb + 127) / 128;
  if (f.eNoNb == 3) assemble_neu_kernel<3><<<blocks, 128, 0, ctx->stream>>>(A);
  else assemble_neu_kernel<4><<<blocks, 128, 0, ctx->stream>>>(A);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  int herr = 0;
  SVB_CUDA(cudaMemcpyAsync(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_err);
  if (herr) {
    // same condition the reference raises in gnnb (nn.cpp:965-969)
    set_error("[svMultiPhysics::gnnb] ERROR: a face node could not be matched to a node in the volume mesh.");
    return SVB200_ERR_INVALID;
  }
  return SVB200_OK;
}

}  // namespace svb
