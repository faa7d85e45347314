// ris.cu — fitted resistive immersed surfaces: the coupling an OPEN surface adds to the assembled system.
//
// Reference: ris::doassem_ris (Code/Source/solver/ris.cpp:269-349), called per element from construct_fluid
// (fluid.cpp:750-754) and construct_fsi (fsi.cpp:349-353) after the element's own scatter.  For every open projection
// (RIS.clsFlg[iProj] false) and every element node a that is listed in grisMapList[iProj].map(2, n) — the node-to-node map of
// the two coincident faces — the element's residual row lR(:,a) and tangent row lK(:,a,b) are added a SECOND time, into the
// row of the twin node a' across the surface, with the columns b that are mapped themselves replaced by their twins:
//     R(:, a') += lR(:, a),      Val(:, slot(a', twin(b) or b)) += lK(:, a, b).
// Summed over the elements of one mesh this is a statement about assembled rows: with C = what THIS mesh's elements
// contributed to the rows of the mapped nodes,   R(a') += C_R(a),   Val(a', c(b)) += C(a, b)   — no element matrix is needed
// again.  So svb200_assemble brackets the element kernels of a fluid / FSI mesh:
//   1. ris_take_kernel   : S = rows of the mapped nodes as they are (earlier meshes, earlier RIS additions), rows zeroed;
//   2. the element kernels (any of them: closed-form TET4, general, FSI fluid + solid);
//   3. ris_put_kernel    : C = rows now (exactly this mesh's contributions: no cancellation), rows = S + C;
//   4. ris_apply_kernel  : Val[dst] += C[src], R(a') += C_R(a) along a plan built once on the host (svb200_set_ris): for every
//                          CSR entry of a mapped row its destination slot in the twin row (linear search in the row, like
//                          slot_map_kernel), FP64 atomics because a node may belong to two projections.
// The plan is per context (node pairs are global to the mesh set); entries whose destination does not exist in the CSR graph
// are an error at svb200_set_ris — the reference's binary search (ris.cpp:330-341) would not terminate on them.
#include <algorithm>
#include <vector>
#include "svb200_internal.h"

namespace svb {

// S[i] = V[ent[i]] (blocks of d2 doubles), V[ent[i]] = 0
__global__ void ris_take_kernel(long long n, int d2, const int* __restrict__ ent, double* __restrict__ V, double* __restrict__ S)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * d2) return;
  const long long i = t / d2;
  const int k = (int)(t - i * d2);
  double* p = V + (size_t)ent[i] * d2 + k;
  S[t] = *p;
  *p = 0.0;
}

// C[i] = V[ent[i]], V[ent[i]] = S[i] + C[i]
__global__ void ris_put_kernel(long long n, int d2, const int* __restrict__ ent, double* __restrict__ V, const double* __restrict__ S,
                               double* __restrict__ C)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * d2) return;
  const long long i = t / d2;
  const int k = (int)(t - i * d2);
  double* p = V + (size_t)ent[i] * d2 + k;
  const double c = *p;
  C[t] = c;
  *p = S[t] + c;
}

// V[dst[j]] += C[src[j]] (src indexes the entry list of the take / put kernels)
__global__ void ris_apply_kernel(long long n, int d2, const int* __restrict__ src, const int* __restrict__ dst, const double* __restrict__ C,
                                 double* __restrict__ V)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * d2) return;
  const long long j = t / d2;
  const int k = (int)(t - j * d2);
  atomicAdd(V + (size_t)dst[j] * d2 + k, C[(size_t)src[j] * d2 + k]);
}

static inline unsigned nblk(long long n) { return (unsigned)((n + 255) / 256); }

bool ris_active(const svb200_ctx* ctx) { return ctx->ris.nEnt > 0 || ctx->ris.nNode > 0; }

// svb200_set_ris: the plan.  maps: for projection p, 2 * nMap[p] node ids (caller numbering), map(0,j), map(1,j) interleaved
// like the column-major Array<int>(2, n) of grisMapList[p].map.
int ris_build_plan(svb200_ctx* ctx, int nProj, const int* nMap, const int* maps, const int* closed)
{
  RisPlan& P = ctx->ris;
  cudaFree(P.d_ent); cudaFree(P.d_src); cudaFree(P.d_dst); cudaFree(P.d_node); cudaFree(P.d_rsrc); cudaFree(P.d_rdst);
  cudaFree(P.d_S); cudaFree(P.d_C); cudaFree(P.d_RS); cudaFree(P.d_RC);
  P = RisPlan();
  if (nProj <= 0) return SVB200_OK;
  const int nNo = ctx->nNo;
  const std::vector<int>& rp = ctx->h_rowPtr;
  const std::vector<int>& cp = ctx->h_colPtr;
  // mapped nodes of the OPEN projections (internal ids), each once: their rows are what the take / put kernels bracket
  std::vector<int> nodeIdx(nNo, -1), nodes;
  size_t off = 0;
  std::vector<size_t> offs(nProj);
  for (int p = 0; p < nProj; p++) {
    offs[p] = off;
    SVB_REQUIRE(nMap[p] >= 0, "svb200_set_ris: negative map length");
    if (!closed[p])
      for (int j = 0; j < 2 * nMap[p]; j++) {
        const int a = maps[off + j];
        // ris.cpp:302-308 keeps a stale row for a node without a twin; such maps are rejected here
        SVB_REQUIRE(a >= 0 && a < nNo, "svb200_set_ris: every mapped node needs a twin (node id out of range or -1)");
        const int ai = ctx->h_map[a];
        if (nodeIdx[ai] < 0) { nodeIdx[ai] = (int)nodes.size(); nodes.push_back(ai); }
      }
    off += 2 * (size_t)nMap[p];
  }
  if (nodes.empty()) return SVB200_OK;
  // entry list: all CSR entries of the mapped rows; entBase[k] = first entry of nodes[k]
  std::vector<int> ent, entBase(nodes.size() + 1, 0);
  for (size_t k = 0; k < nodes.size(); k++) {
    entBase[k] = (int)ent.size();
    for (int q = rp[nodes[k]]; q < rp[nodes[k] + 1]; q++) ent.push_back(q);
  }
  entBase[nodes.size()] = (int)ent.size();
  std::vector<int> src, dst, rsrc, rdst, twin(nNo, -1);
  for (int p = 0; p < nProj; p++) {
    if (closed[p]) continue;
    const int* mp = maps + offs[p];
    for (int j = 0; j < nMap[p]; j++) {
      const int a0 = ctx->h_map[mp[2 * j]], a1 = ctx->h_map[mp[2 * j + 1]];
      twin[a0] = a1; twin[a1] = a0;
    }
    for (int j = 0; j < nMap[p]; j++)
      for (int side = 0; side < 2; side++) {
        const int a = ctx->h_map[mp[2 * j + side]], at = ctx->h_map[mp[2 * j + 1 - side]];
        const int k = nodeIdx[a];
        rsrc.push_back(k); rdst.push_back(at);
        for (int q = rp[a], i = entBase[k]; q < rp[a + 1]; q++, i++) {
          const int b = cp[q];
          const int c = twin[b] >= 0 ? twin[b] : b;          // ris.cpp:315-325
          int s = -1;
          for (int r = rp[at]; r < rp[at + 1]; r++)
            if (cp[r] == c) { s = r; break; }
          if (s < 0) {
            // an entry of row a that only the RIS connections of lhsa put there (lhsa.cpp:168-193) has no element contribution
            // and no counterpart in the twin's row: it carries zeros, skip it
            continue;
          }
          src.push_back(i); dst.push_back(s);
        }
      }
    for (int j = 0; j < nMap[p]; j++) { twin[ctx->h_map[mp[2 * j]]] = -1; twin[ctx->h_map[mp[2 * j + 1]]] = -1; }
  }
  P.nNode = (int)nodes.size(); P.nEnt = (long long)ent.size(); P.nApply = (long long)src.size(); P.nRApply = (int)rsrc.size();
  auto up = [&](int** d, const std::vector<int>& h) -> int {
    SVB_CUDA(cudaMalloc(d, sizeof(int) * std::max<size_t>(h.size(), 1)));
    SVB_CUDA(cudaMemcpyAsync(*d, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
    return SVB200_OK;
  };
  int rc;
  if ((rc = up(&P.d_ent, ent)) || (rc = up(&P.d_src, src)) || (rc = up(&P.d_dst, dst)) || (rc = up(&P.d_node, nodes)) ||
      (rc = up(&P.d_rsrc, rsrc)) || (rc = up(&P.d_rdst, rdst))) return rc;
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SVB200_OK;
}

static int ris_buffers(svb200_ctx* ctx, int dof)
{
  RisPlan& P = ctx->ris;
  if (P.buf_dof == dof) return SVB200_OK;
  cudaFree(P.d_S); cudaFree(P.d_C); cudaFree(P.d_RS); cudaFree(P.d_RC);
  P.d_S = P.d_C = P.d_RS = P.d_RC = nullptr;
  const size_t nv = (size_t)std::max<long long>(P.nEnt, 1) * dof * dof, nr = (size_t)std::max(P.nNode, 1) * dof;
  SVB_CUDA(cudaMalloc(&P.d_S, sizeof(double) * nv));
  SVB_CUDA(cudaMalloc(&P.d_C, sizeof(double) * nv));
  SVB_CUDA(cudaMalloc(&P.d_RS, sizeof(double) * nr));
  SVB_CUDA(cudaMalloc(&P.d_RC, sizeof(double) * nr));
  P.buf_dof = dof;
  return SVB200_OK;
}

// before the element kernels of one mesh
int ris_begin(svb200_ctx* ctx)
{
  RisPlan& P = ctx->ris;
  const int dof = ctx->dof, d2 = dof * dof;
  int rc = ris_buffers(ctx, dof);
  if (rc) return rc;
  ris_take_kernel<<<nblk(P.nEnt * d2), 256, 0, ctx->stream>>>(P.nEnt, d2, P.d_ent, ctx->d_Val, P.d_S);
  ris_take_kernel<<<nblk((long long)P.nNode * dof), 256, 0, ctx->stream>>>(P.nNode, dof, P.d_node, ctx->d_R, P.d_RS);
  ctx->launches += 2;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// after them
int ris_end(svb200_ctx* ctx)
{
  RisPlan& P = ctx->ris;
  const int dof = ctx->dof, d2 = dof * dof;
  ris_put_kernel<<<nblk(P.nEnt * d2), 256, 0, ctx->stream>>>(P.nEnt, d2, P.d_ent, ctx->d_Val, P.d_S, P.d_C);
  ris_put_kernel<<<nblk((long long)P.nNode * dof), 256, 0, ctx->stream>>>(P.nNode, dof, P.d_node, ctx->d_R, P.d_RS, P.d_RC);
  if (P.nApply) ris_apply_kernel<<<nblk(P.nApply * d2), 256, 0, ctx->stream>>>(P.nApply, d2, P.d_src, P.d_dst, P.d_C, ctx->d_Val);
  if (P.nRApply) ris_apply_kernel<<<nblk((long long)P.nRApply * dof), 256, 0, ctx->stream>>>(P.nRApply, dof, P.d_rsrc, P.d_rdst, P.d_RC, ctx->d_R);
  ctx->launches += 4;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

}  // namespace svb
