// comm.cu — cross-partition exchanges: one mesh partition per B200, one process per GPU, NCCL over
// NVLink 5 / NVSwitch.
//
//   halo_sum        fsils_commuv / fsils_commus   Code/Source/linear_solver/in_commu.cpp:84-143 / 22-75
//   allreduce_sum   fsils_bcast_v, the MPI_Allreduce of fsils_dot_v / fsi_ls_normv
//                   linear_solver/bcast.cpp:24-31, dot.cpp:35-60, norm.cpp:33-87
//
// A halo "sum" is a sparse neighbour all-reduce: every rank sends its partial values of the nodes it
// shares with neighbour iP and adds what it receives, neighbours visited in ascending rank order like
// the reference (in_commu.cpp:128-135), so shared nodes end up with the same total (up to the
// order of additions) everywhere.  Pack -> grouped ncclSend/ncclRecv -> unpack-add all run on the
// context's stream: no host round trip, no CPU fallback.
//
// NCCL is resolved with dlopen at svb200_comm_init time so that single-GPU users of libsvb200.so do
// not need libnccl at all.
#include <dlfcn.h>
#include <nccl.h>
#include "svb200_internal.h"
#include "fsils_kernels.h"

namespace svb {

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int nccl_fail(ncclResult_t r, const char* what)
{
  set_error(std::string("svb200: NCCL error in ") + what + ": " +
            (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "unknown"));
  return SVB200_ERR_NCCL;
}
#define SVB_NCCL(call)                                  \
  do {                                                  \
    ncclResult_t r__ = (call);                          \
    if (r__ != ncclSuccess) return nccl_fail(r__, #call); \
  } while (0)
}  // namespace

int nccl_load()
{
  if (g_nccl.handle) return SVB200_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) {
    set_error(std::string("svb200: cannot load libnccl.so.2: ") + dlerror());
    return SVB200_ERR_NCCL;
  }
#define LOAD(field, sym)                                                       \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.handle, sym);                        \
  if (!g_nccl.field) {                                                         \
    set_error(std::string("svb200: libnccl is missing symbol ") + sym);        \
    return SVB200_ERR_NCCL;                                                    \
  }
  LOAD(GetUniqueId, "ncclGetUniqueId")
  LOAD(CommInitRank, "ncclCommInitRank")
  LOAD(CommDestroy, "ncclCommDestroy")
  LOAD(AllReduce, "ncclAllReduce")
  LOAD(Send, "ncclSend")
  LOAD(Recv, "ncclRecv")
  LOAD(GroupStart, "ncclGroupStart")
  LOAD(GroupEnd, "ncclGroupEnd")
  LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
  return SVB200_OK;
}

int nccl_unique_id(void* id128)
{
  int rc = nccl_load();
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  SVB_NCCL(g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId*>(id128)));
  return SVB200_OK;
}

int nccl_init(svb200_ctx* ctx, int nranks, int rank, const void* id128)
{
  int rc = nccl_load();
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  SVB_CUDA(cudaSetDevice(ctx->device));
  SVB_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
  ctx->nccl_comm = comm;
  ctx->nranks = nranks;
  ctx->rank = rank;
  return SVB200_OK;
}

void nccl_destroy(svb200_ctx* ctx)
{
  if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
}

__global__ void halo_pack_kernel(int n, int dof, const int* __restrict__ ptr, const double* __restrict__ V,
                                 double* __restrict__ buf)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * dof) return;
  buf[t] = V[(size_t)ptr[t / dof] * dof + t % dof];
}

__global__ void halo_unpack_add_kernel(int n, int dof, const int* __restrict__ ptr, const double* __restrict__ buf,
                                       double* __restrict__ V)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * dof) return;
  V[(size_t)ptr[t / dof] * dof + t % dof] += buf[t];
}

int halo_sum(svb200_ctx* ctx, int dof, double* V)
{
  if (ctx->nranks <= 1 || ctx->neigh.empty()) return SVB200_OK;
  if (!ctx->nccl_comm) {
    set_error("svb200: graph has neighbour partitions but svb200_comm_init was not called");
    return SVB200_ERR_INVALID;
  }
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  for (auto& nb : ctx->neigh) {
    const int n = nb.n * dof;
    if (n == 0) continue;
    halo_pack_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(nb.n, dof, nb.d_ptr, V, nb.d_send);
    ctx->launches++;
  }
  SVB_NCCL(g_nccl.GroupStart());
  for (auto& nb : ctx->neigh) {
    const size_t n = (size_t)nb.n * dof;
    if (n == 0) continue;
    SVB_NCCL(g_nccl.Send(nb.d_send, n, ncclDouble, nb.rank, comm, ctx->stream));
    SVB_NCCL(g_nccl.Recv(nb.d_recv, n, ncclDouble, nb.rank, comm, ctx->stream));
  }
  SVB_NCCL(g_nccl.GroupEnd());
  for (auto& nb : ctx->neigh) {   // neighbours are stored in ascending rank order
    const int n = nb.n * dof;
    if (n == 0) continue;
    halo_unpack_add_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(nb.n, dof, nb.d_ptr, nb.d_recv, V);
    ctx->launches++;
  }
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

int allreduce_sum(svb200_ctx* ctx, double* d_buf, int n)
{
  if (ctx->nranks <= 1) return SVB200_OK;
  if (!ctx->nccl_comm) {
    set_error("svb200: multi-rank context without a communicator");
    return SVB200_ERR_INVALID;
  }
  SVB_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return SVB200_OK;
}

}  // namespace svb
