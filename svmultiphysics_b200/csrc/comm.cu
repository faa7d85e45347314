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
// Two transports, same semantics:
//   * "p2p" (default when CUDA IPC works between the ranks of the box): every rank owns a MAILBOX in its HBM that its
//     peers map through cudaIpcOpenMemHandle.  A halo sum is then ONE push kernel per neighbour that gathers the shared
//     rows of V and stores them straight into the neighbour's mailbox over NVLink / NVSwitch, followed by a release
//     flag; the receiving side's unpack kernel spins on that flag and adds.  The scalar all-reduce of a GMRES / CG
//     iteration is ONE single-CTA kernel: store my partial results into every peer's slot, signal, wait for all
//     peers, then sum the nranks slots in rank order — so every rank computes bit-identical sums, and there is no
//     NCCL launch (3 proxied kernels + host-side group bookkeeping per iteration before).  Mailboxes are double-buffered
//     by the parity of a sequence number: a rank cannot run more than one exchange ahead of a peer because it needs
//     that peer's data to get past its own wait.
//     Since round 2 the exchange is FUSED with the producers and consumers of an iteration: the SpMV of the
//     interface rows runs first and its epilogue stores every finished row straight into the mailboxes of all the ranks
//     sharing it (bsr_spmv4_bnd_push_kernel: one kernel = compute + NVLink transfer + flag), the interior rows follow
//     while the data travels, and ONE wait-add kernel folds all neighbours' contributions in ascending rank order; the
//     all-reduce of the i+2 Gram-Schmidt dot products is the tail of the dot kernel's second stage
//     (dot_stage2_allreduce_kernel).  An inner GMRES iteration has 6 launches on N ranks against 4 on one, whatever the
//     number of neighbours (2 push + 2 wait-add + 1 all-reduce launches PER NEIGHBOUR PAIR / iteration before).
//   * "nccl": grouped ncclSend/ncclRecv + ncclAllReduce (SVB200_COMM=nccl, or when IPC is unavailable).
//
// NCCL is resolved with dlopen at svb200_comm_init time so that single-GPU users of libsvb200.so do
// not need libnccl at all.
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>
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
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
  LOAD(AllGather, "ncclAllGather")
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

// ---- peer-memory transport ---------------------------------------------------------------------------------
constexpr int AR_MAX = 512;                       // doubles per scalar all-reduce
constexpr unsigned long long SPIN_LIMIT_NS = 20ull * 1000 * 1000 * 1000;   // give up (error flag) after 20 s

struct P2PNeighbor {
  size_t my_off = 0;       // offset (doubles) in MY mailbox of the two receive buffers for this neighbour
  size_t peer_off = 0;     // offset (doubles) in the NEIGHBOUR's mailbox of its receive buffers for me
  unsigned* d_count = nullptr;   // block counter of the push kernel
};

struct P2P {
  bool ready = false;
  int nranks = 0, rank = 0;
  double* base = nullptr;                 // my mailbox
  size_t doubles = 0;
  std::vector<double*> peer;              // mapped mailboxes, peer[rank] = base
  double** d_peer = nullptr;              // device copy of peer[]
  std::vector<P2PNeighbor> nb;            // same order as ctx->neigh
  unsigned long long halo_seq = 0, ar_seq = 0;
  // fused exchange tables (all neighbours in one launch)
  int nB = 0, nnb = 0;                    // distinct interface rows, neighbours
  int nLow = 0;                           // interface rows below mynNo
  bool contiguous = false;                // interface rows are exactly [0,nLow) U [mynNo,nNo) (FSILS order): interior = one range
  int* d_brow = nullptr;                  // (nB) interface rows, ascending
  int* d_bptr = nullptr;                  // (nB+1) offsets into d_bent
  int2* d_bent = nullptr;                 // {neighbour index j (ascending rank), position in j's shared list}
  double** d_rbuf = nullptr;              // (2*nnb) [parity][j]: the neighbour's receive buffer for me
  double** d_lbuf = nullptr;              // (2*nnb) [parity][j]: my receive buffer for neighbour j
  unsigned long long** d_rflag = nullptr; // (2*nnb) [parity][j]: the neighbour's flag for me
  unsigned long long** d_lflag = nullptr; // (2*nnb) [parity][j]: my flag for neighbour j
  unsigned* d_count = nullptr;            // block counter of the push kernels
  // fixed layout at the start of every mailbox (in doubles / 8-byte words):
  //   [0, 2R)            halo flags  [parity][src rank]
  //   [2R, 4R)           all-reduce flags [parity][src rank]
  //   [4R, 4R + 1)       error word
  //   [hdr, hdr + 2 R AR_MAX)   all-reduce slots [parity][src rank][AR_MAX]
  size_t off_arflag() const { return 2 * (size_t)nranks; }
  size_t off_err() const { return 4 * (size_t)nranks; }
  size_t off_slots() const { return 4 * (size_t)nranks + 8; }
  size_t off_data() const { return off_slots() + 2 * (size_t)nranks * AR_MAX; }
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Spin until *flag >= seq; on time-out raise the error word and return false.
__device__ __forceinline__ bool wait_flag(const unsigned long long* flag, unsigned long long seq, unsigned long long* err)
{
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(flag) < seq) {
    __nanosleep(64);
    if (global_ns() - t0 > SPIN_LIMIT_NS) { *err = 1ull; return false; }
  }
  return true;
}

// Gather the shared rows of V and store them into the neighbour's mailbox; the last block to finish publishes the flag.
__global__ void __launch_bounds__(256)
halo_push_kernel(int n, int dof, const int* __restrict__ ptr, const double* __restrict__ V, double* __restrict__ remote_buf,
                 unsigned long long* remote_flag, unsigned long long seq, unsigned* __restrict__ count)
{
  const int total = n * dof;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x)
    remote_buf[t] = V[(size_t)ptr[t / dof] * dof + t % dof];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(count, 1u);
    if (done == gridDim.x - 1) {
      *count = 0u;
      __threadfence_system();
      st_release_sys(remote_flag, seq);
    }
  }
}

__global__ void __launch_bounds__(256)
halo_wait_add_kernel(int n, int dof, const int* __restrict__ ptr, const double* __restrict__ buf, double* __restrict__ V,
                     const unsigned long long* flag, unsigned long long seq, unsigned long long* err)
{
  __shared__ int ok;
  if (threadIdx.x == 0) ok = wait_flag(flag, seq, err) ? 1 : 0;
  __syncthreads();
  if (!ok) return;
  const int total = n * dof;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x)
    V[(size_t)ptr[t / dof] * dof + t % dof] += __ldcg(buf + t);
}

// ---- all neighbours in one launch --------------------------------------------------------------------------------
// Tail of a push kernel: once every block has stored its rows (system-scope fence, block counter), the last block
// publishes the sequence number to the flags of all neighbours.
__device__ __forceinline__ void publish_flags(int nnb, unsigned long long* const* rflag, unsigned long long seq, unsigned* count)
{
  __shared__ int s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(count, 1u);
    s_last = (done == gridDim.x - 1);
    if (s_last) *count = 0u;
  }
  __syncthreads();
  if (s_last && (int)threadIdx.x < nnb) {
    __threadfence_system();
    st_release_sys(rflag[threadIdx.x], seq);
  }
}

// Interface rows of V -> the mailboxes of every rank sharing them.
__global__ void __launch_bounds__(256)
halo_push_all_kernel(int nB, int nnb, int dof, const int* __restrict__ brow, const int* __restrict__ bptr, const int2* __restrict__ bent,
                     double* const* __restrict__ rbuf, const double* __restrict__ V, unsigned long long* const* __restrict__ rflag,
                     unsigned long long seq, unsigned* __restrict__ count)
{
  const int total = nB * dof;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int b = t / dof, i = t - b * dof;
    const double v = V[(size_t)brow[b] * dof + i];
    for (int e = bptr[b]; e < bptr[b + 1]; e++) {
      const int2 q = bent[e];
      rbuf[q.x][(size_t)q.y * dof + i] = v;
    }
  }
  publish_flags(nnb, rflag, seq, count);
}

// V(row) += sum over the neighbours sharing the row, ascending rank (in_commu.cpp:128-135), after all of them have arrived.
__global__ void __launch_bounds__(256)
halo_wait_add_all_kernel(int nB, int nnb, int dof, const int* __restrict__ brow, const int* __restrict__ bptr,
                         const int2* __restrict__ bent, const double* const* __restrict__ lbuf, double* __restrict__ V,
                         const unsigned long long* const* __restrict__ lflag, unsigned long long seq, unsigned long long* err)
{
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  if ((int)threadIdx.x < nnb && !wait_flag(lflag[threadIdx.x], seq, err)) ok = 0;
  __syncthreads();
  if (!ok) return;
  const int total = nB * dof;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int b = t / dof, i = t - b * dof;
    double* p = V + (size_t)brow[b] * dof + i;
    double v = *p;
    for (int e = bptr[b]; e < bptr[b + 1]; e++) {
      const int2 q = bent[e];
      v += __ldcg(lbuf[q.x] + (size_t)q.y * dof + i);
    }
    *p = v;
  }
}

// dof = 4 SpMV of the INTERFACE rows whose epilogue is the halo push: 8 lanes per row like bsr_spmv4_kernel
// (fsils_kernels.cu); the lane holding component i of a finished row stores it to KU and, over NVLink, into the
// mailbox of every rank sharing the row; the last block raises the neighbours' flags.  spar_mul.cpp:164-231 +
// in_commu.cpp:104-121 in one kernel; the interior rows are multiplied while this data is in flight.
__global__ void __launch_bounds__(256)
bsr_spmv4_bnd_push_kernel(int nB, int nnb, const int* __restrict__ brow, const int* __restrict__ bptr, const int2* __restrict__ bent,
                          double* const* __restrict__ rbuf, unsigned long long* const* __restrict__ rflag, unsigned long long seq,
                          unsigned* __restrict__ count, const int* __restrict__ rowPtr, const int* __restrict__ colPtr,
                          const double* __restrict__ Val, const double* __restrict__ U, double* __restrict__ KU)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = t >> 3;
  const int l = t & 7;
  const int i = l >> 1, j0 = (l & 1) << 1;
  double acc = 0.0;
  int row = 0;
  if (b < nB) {
    row = brow[b];
    const int k0 = rowPtr[row], k1 = rowPtr[row + 1];
    const double2* V2 = reinterpret_cast<const double2*>(Val) + l;
    int k = k0;
    for (; k + 4 <= k1; k += 4) {          // 4 blocks in flight per lane, same summation order as bsr_spmv4_kernel
      const int c0 = __ldg(colPtr + k), c1 = __ldg(colPtr + k + 1), c2 = __ldg(colPtr + k + 2), c3 = __ldg(colPtr + k + 3);
      const double2 v0 = __ldcs(V2 + 8 * (size_t)k), v1 = __ldcs(V2 + 8 * (size_t)(k + 1));
      const double2 v2 = __ldcs(V2 + 8 * (size_t)(k + 2)), v3 = __ldcs(V2 + 8 * (size_t)(k + 3));
      const double2 u0 = *reinterpret_cast<const double2*>(U + 4 * (size_t)c0 + j0);
      const double2 u1 = *reinterpret_cast<const double2*>(U + 4 * (size_t)c1 + j0);
      const double2 u2 = *reinterpret_cast<const double2*>(U + 4 * (size_t)c2 + j0);
      const double2 u3 = *reinterpret_cast<const double2*>(U + 4 * (size_t)c3 + j0);
      acc += v0.x * u0.x + v0.y * u0.y;
      acc += v1.x * u1.x + v1.y * u1.y;
      acc += v2.x * u2.x + v2.y * u2.y;
      acc += v3.x * u3.x + v3.y * u3.y;
    }
    for (; k < k1; k++) {
      const int c0 = __ldg(colPtr + k);
      const double2 v0 = __ldcs(V2 + 8 * (size_t)k);
      const double2 u0 = *reinterpret_cast<const double2*>(U + 4 * (size_t)c0 + j0);
      acc += v0.x * u0.x + v0.y * u0.y;
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if (b < nB && (l & 1) == 0) {
    KU[4 * (size_t)row + i] = acc;
    for (int e = bptr[b]; e < bptr[b + 1]; e++) {
      const int2 q = bent[e];
      rbuf[q.x][4 * (size_t)q.y + i] = acc;
    }
  }
  publish_flags(nnb, rflag, seq, count);
}

// Second stage of multi_dot (fsils_kernels.cu) with the cross-rank sum as its tail: one CTA adds the per-CTA partials of
// the nvec dot products in the fixed order of multi_dot_stage2, stores them into every rank's slot, signals, waits for
// all ranks and adds the slots in rank order — every rank obtains bit-identical sums (dot.cpp:35-60 / bcast.cpp:24-31).
__global__ void __launch_bounds__(1024)
dot_stage2_allreduce_kernel(int nblocks, int nvec, const double* __restrict__ part, double* __restrict__ out, int nranks, int rank,
                            double* const* __restrict__ peer, size_t off_arflag, size_t off_slots, size_t off_err,
                            unsigned long long seq)
{
  const int bpar = (int)(seq & 1ull);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int j = warp; j < nvec; j += 32) {
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += part[(size_t)b * nvec + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane < nranks) peer[lane][off_slots + ((size_t)bpar * nranks + rank) * AR_MAX + j] = s;
    for (int r = 32 + lane; r < nranks; r += 32) peer[r][off_slots + ((size_t)bpar * nranks + rank) * AR_MAX + j] = s;
  }
  __threadfence_system();
  __syncthreads();
  if (t < nranks)
    st_release_sys(reinterpret_cast<unsigned long long*>(peer[t]) + off_arflag + (size_t)bpar * nranks + rank, seq);
  double* mine = peer[rank];
  __shared__ int ok;
  if (t == 0) ok = 1;
  __syncthreads();
  if (t < nranks) {
    if (!wait_flag(reinterpret_cast<const unsigned long long*>(mine) + off_arflag + (size_t)bpar * nranks + t, seq,
                   reinterpret_cast<unsigned long long*>(mine) + off_err))
      ok = 0;
  }
  __syncthreads();
  if (!ok) return;
  for (int j = t; j < nvec; j += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < nranks; r++) s += __ldcg(mine + off_slots + ((size_t)bpar * nranks + r) * AR_MAX + j);
    out[j] = s;
  }
}

// Scalar all-reduce (sum) of n <= AR_MAX doubles in one CTA: push to every mailbox, signal, wait, sum in rank order.
__global__ void __launch_bounds__(512)
allreduce_p2p_kernel(int n, int nranks, int rank, double* __restrict__ buf, double* const* __restrict__ peer, size_t off_arflag,
                     size_t off_slots, size_t off_err, unsigned long long seq)
{
  const int b = (int)(seq & 1ull);
  const int t = threadIdx.x;
  if (t < n) {
    const double v = buf[t];
    for (int r = 0; r < nranks; r++) peer[r][off_slots + ((size_t)b * nranks + rank) * AR_MAX + t] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (t < nranks)
    st_release_sys(reinterpret_cast<unsigned long long*>(peer[t]) + off_arflag + (size_t)b * nranks + rank, seq);
  double* mine = peer[rank];
  __shared__ int ok;
  if (t == 0) ok = 1;
  __syncthreads();
  if (t < nranks) {
    if (!wait_flag(reinterpret_cast<const unsigned long long*>(mine) + off_arflag + (size_t)b * nranks + t, seq,
                   reinterpret_cast<unsigned long long*>(mine) + off_err))
      ok = 0;
  }
  __syncthreads();
  if (!ok || t >= n) return;
  double s = 0.0;
  for (int r = 0; r < nranks; r++) s += __ldcg(mine + off_slots + ((size_t)b * nranks + r) * AR_MAX + t);
  buf[t] = s;
}

static P2P* p2p_of(svb200_ctx* ctx) { return static_cast<P2P*>(ctx->p2p); }

void p2p_destroy(svb200_ctx* ctx)
{
  P2P* p = p2p_of(ctx);
  if (!p) return;
  for (int r = 0; r < (int)p->peer.size(); r++)
    if (r != p->rank && p->peer[r]) cudaIpcCloseMemHandle(p->peer[r]);
  for (auto& nb : p->nb) cudaFree(nb.d_count);
  cudaFree(p->d_brow); cudaFree(p->d_bptr); cudaFree(p->d_bent); cudaFree(p->d_rbuf); cudaFree(p->d_lbuf);
  cudaFree(p->d_rflag); cudaFree(p->d_lflag); cudaFree(p->d_count);
  cudaFree(p->d_peer);
  cudaFree(p->base);
  delete p;
  ctx->p2p = nullptr;
}

// Collective over all ranks (call from svb200_set_graph / svb200_comm_init on every rank).  Any failure on any rank
// makes ALL ranks fall back to the NCCL transport.
int p2p_setup(svb200_ctx* ctx)
{
  p2p_destroy(ctx);
  if (ctx->nranks <= 1 || !ctx->nccl_comm) return SVB200_OK;
  const char* mode = getenv("SVB200_COMM");
  const bool want = !(mode && std::string(mode) == "nccl");
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  const int R = ctx->nranks;
  P2P* p = new P2P();
  p->nranks = R; p->rank = ctx->rank;
  p->nb.resize(ctx->neigh.size());
  size_t off = p->off_data();
  for (size_t j = 0; j < ctx->neigh.size(); j++) {
    p->nb[j].my_off = off;
    off += 2 * (size_t)4 * std::max(ctx->neigh[j].n, 1);
  }
  p->doubles = off;
  bool ok = want;
  cudaIpcMemHandle_t handle;
  memset(&handle, 0, sizeof(handle));
  if (ok) ok = cudaMalloc(&p->base, sizeof(double) * p->doubles) == cudaSuccess;
  if (ok) ok = cudaMemset(p->base, 0, sizeof(double) * p->doubles) == cudaSuccess;
  if (ok) ok = cudaIpcGetMemHandle(&handle, p->base) == cudaSuccess;
  cudaGetLastError();
  // record per rank: ok flag | IPC handle (64 B) | offset of my receive buffer for every source rank (-1 = none)
  const size_t rec = 8 + sizeof(handle) + sizeof(long long) * R;
  std::vector<unsigned char> mine(rec, 0), all(rec * R, 0);
  long long okw = ok ? 1 : 0;
  memcpy(mine.data(), &okw, 8);
  memcpy(mine.data() + 8, &handle, sizeof(handle));
  std::vector<long long> offs(R, -1);
  for (size_t j = 0; j < ctx->neigh.size(); j++) offs[ctx->neigh[j].rank] = (long long)p->nb[j].my_off;
  memcpy(mine.data() + 8 + sizeof(handle), offs.data(), sizeof(long long) * R);
  unsigned char *d_mine = nullptr, *d_all = nullptr;
  SVB_CUDA(cudaMalloc(&d_mine, rec));
  SVB_CUDA(cudaMalloc(&d_all, rec * R));
  SVB_CUDA(cudaMemcpyAsync(d_mine, mine.data(), rec, cudaMemcpyHostToDevice, ctx->stream));
  SVB_NCCL(g_nccl.AllGather(d_mine, d_all, rec, ncclChar, comm, ctx->stream));
  SVB_CUDA(cudaMemcpyAsync(all.data(), d_all, rec * R, cudaMemcpyDeviceToHost, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < R; r++) {
    long long f;
    memcpy(&f, all.data() + rec * r, 8);
    ok = ok && f == 1;
  }
  p->peer.assign(R, nullptr);
  if (ok) {
    for (int r = 0; r < R && ok; r++) {
      if (r == ctx->rank) { p->peer[r] = p->base; continue; }
      cudaIpcMemHandle_t h;
      memcpy(&h, all.data() + rec * r + 8, sizeof(h));
      void* q = nullptr;
      ok = cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
      p->peer[r] = static_cast<double*>(q);
    }
    cudaGetLastError();
  }
  // second agreement round: every rank must have mapped every mailbox
  double* d_flag = reinterpret_cast<double*>(d_mine);
  double okd = ok ? 1.0 : 0.0;
  SVB_CUDA(cudaMemcpyAsync(d_flag, &okd, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  SVB_NCCL(g_nccl.AllReduce(d_flag, d_flag, 1, ncclDouble, ncclMin, comm, ctx->stream));
  SVB_CUDA(cudaMemcpyAsync(&okd, d_flag, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_mine); cudaFree(d_all);
  ctx->p2p = p;
  if (okd != 1.0) {
    p2p_destroy(ctx);
    return SVB200_OK;     // NCCL transport
  }
  for (size_t j = 0; j < ctx->neigh.size(); j++) {
    long long o;
    memcpy(&o, all.data() + rec * ctx->neigh[j].rank + 8 + sizeof(handle) + sizeof(long long) * ctx->rank, sizeof(o));
    if (o < 0) {
      set_error("svb200: neighbour lists are not symmetric between partitions");
      p2p_destroy(ctx);
      return SVB200_ERR_INVALID;
    }
    p->nb[j].peer_off = (size_t)o;
    SVB_CUDA(cudaMalloc(&p->nb[j].d_count, sizeof(unsigned)));
    SVB_CUDA(cudaMemset(p->nb[j].d_count, 0, sizeof(unsigned)));
  }
  SVB_CUDA(cudaMalloc(&p->d_peer, sizeof(double*) * R));
  SVB_CUDA(cudaMemcpy(p->d_peer, p->peer.data(), sizeof(double*) * R, cudaMemcpyHostToDevice));
  {
    // fused exchange tables: per distinct interface row the list of (neighbour, position in its shared list), neighbours
    // in ascending rank order (ctx->neigh is sorted), and the mailbox / flag addresses of both parities
    const int nnb = (int)ctx->neigh.size();
    p->nnb = nnb;
    std::vector<std::pair<int, int2>> ent;      // (row, {j, pos})
    for (int j = 0; j < nnb; j++)
      for (int k = 0; k < ctx->neigh[j].n; k++) ent.push_back({ctx->neigh[j].h_ptr[k], make_int2(j, k)});
    std::stable_sort(ent.begin(), ent.end(), [](const std::pair<int, int2>& a, const std::pair<int, int2>& b) { return a.first < b.first; });
    std::vector<int> brow, bptr;
    std::vector<int2> bent(ent.size());
    for (size_t e = 0; e < ent.size(); e++) {
      if (e == 0 || ent[e].first != ent[e - 1].first) { brow.push_back(ent[e].first); bptr.push_back((int)e); }
      bent[e] = ent[e].second;
    }
    bptr.push_back((int)ent.size());
    p->nB = (int)brow.size();
    p->nLow = 0;
    for (int r : brow) p->nLow += (r < ctx->mynNo);
    p->contiguous = true;
    for (int b = 0; b < p->nB; b++) {
      const int expect = b < p->nLow ? b : ctx->mynNo + (b - p->nLow);
      if (brow[b] != expect) { p->contiguous = false; break; }
    }
    if (p->nB - p->nLow != ctx->nNo - ctx->mynNo) p->contiguous = false;
    std::vector<double*> rbuf(2 * std::max(nnb, 1)), lbuf(2 * std::max(nnb, 1));
    std::vector<unsigned long long*> rflag(2 * std::max(nnb, 1)), lflag(2 * std::max(nnb, 1));
    for (int b = 0; b < 2; b++)
      for (int j = 0; j < nnb; j++) {
        const auto& nb = ctx->neigh[j];
        const size_t half = (size_t)b * 4 * std::max(nb.n, 1);
        rbuf[b * nnb + j] = p->peer[nb.rank] + p->nb[j].peer_off + half;
        lbuf[b * nnb + j] = p->base + p->nb[j].my_off + half;
        rflag[b * nnb + j] = reinterpret_cast<unsigned long long*>(p->peer[nb.rank]) + (size_t)b * R + ctx->rank;
        lflag[b * nnb + j] = reinterpret_cast<unsigned long long*>(p->base) + (size_t)b * R + nb.rank;
      }
    auto up = [&](auto** d, const auto& h) -> int {
      using T = typename std::remove_reference<decltype(h)>::type::value_type;
      SVB_CUDA(cudaMalloc(d, sizeof(T) * std::max<size_t>(h.size(), 1)));
      if (!h.empty()) SVB_CUDA(cudaMemcpy(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
      return SVB200_OK;
    };
    int rc = up(&p->d_brow, brow);
    if (!rc) rc = up(&p->d_bptr, bptr);
    if (!rc) rc = up(&p->d_bent, bent);
    if (!rc) rc = up(&p->d_rbuf, rbuf);
    if (!rc) rc = up(&p->d_lbuf, lbuf);
    if (!rc) rc = up(&p->d_rflag, rflag);
    if (!rc) rc = up(&p->d_lflag, lflag);
    if (rc) return rc;
    SVB_CUDA(cudaMalloc(&p->d_count, sizeof(unsigned)));
    SVB_CUDA(cudaMemset(p->d_count, 0, sizeof(unsigned)));
  }
  p->ready = true;
  return SVB200_OK;
}

const char* comm_transport(svb200_ctx* ctx)
{
  if (ctx->nranks <= 1) return "none";
  return (p2p_of(ctx) && p2p_of(ctx)->ready) ? "p2p" : "nccl";
}

// Raise an error if a spin-wait of the peer-memory transport timed out since the last check.
int p2p_check(svb200_ctx* ctx)
{
  P2P* p = p2p_of(ctx);
  if (!p || !p->ready) return SVB200_OK;
  unsigned long long e = 0;
  SVB_CUDA(cudaMemcpyAsync(&e, reinterpret_cast<unsigned long long*>(p->base) + p->off_err(), sizeof(e), cudaMemcpyDeviceToHost, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (e != 0) {
    set_error("svb200: a peer-memory exchange timed out (a partner rank did not arrive within 20 s)");
    return SVB200_ERR_NCCL;
  }
  return SVB200_OK;
}

static const bool g_per_neighbour = getenv("SVB200_HALO_PER_NEIGHBOUR") != nullptr;   // A/B knob: the round-1 exchange

static int halo_wait_add_all(svb200_ctx* ctx, P2P* p, int dof, double* V, unsigned long long seq)
{
  const int b = (int)(seq & 1ull);
  const int blocks = std::max(1, std::min((p->nB * dof + 255) / 256, 148));
  halo_wait_add_all_kernel<<<blocks, 256, 0, ctx->stream>>>(p->nB, p->nnb, dof, p->d_brow, p->d_bptr, p->d_bent, p->d_lbuf + b * p->nnb, V,
                                                           p->d_lflag + b * p->nnb, seq,
                                                           reinterpret_cast<unsigned long long*>(p->base) + p->off_err());
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

static int halo_sum_p2p(svb200_ctx* ctx, int dof, double* V)
{
  P2P* p = p2p_of(ctx);
  const unsigned long long seq = ++p->halo_seq;
  const int b = (int)(seq & 1ull);
  if (!g_per_neighbour) {
    const int blocks = std::max(1, std::min((p->nB * dof + 255) / 256, 148));
    halo_push_all_kernel<<<blocks, 256, 0, ctx->stream>>>(p->nB, p->nnb, dof, p->d_brow, p->d_bptr, p->d_bent, p->d_rbuf + b * p->nnb, V,
                                                         p->d_rflag + b * p->nnb, seq, p->d_count);
    ctx->launches++;
    return halo_wait_add_all(ctx, p, dof, V, seq);
  }
  unsigned long long* myflags = reinterpret_cast<unsigned long long*>(p->base);
  for (size_t j = 0; j < ctx->neigh.size(); j++) {
    auto& nb = ctx->neigh[j];
    const int n = nb.n * dof;
    double* rbase = p->peer[nb.rank];
    double* rbuf = rbase + p->nb[j].peer_off + (size_t)b * 4 * std::max(nb.n, 1);
    unsigned long long* rflag = reinterpret_cast<unsigned long long*>(rbase) + (size_t)b * p->nranks + p->rank;
    const int blocks = std::max(1, std::min((n + 255) / 256, 148));
    halo_push_kernel<<<blocks, 256, 0, ctx->stream>>>(nb.n, dof, nb.d_ptr, V, rbuf, rflag, seq, p->nb[j].d_count);
    ctx->launches++;
  }
  for (size_t j = 0; j < ctx->neigh.size(); j++) {   // ascending rank order, like in_commu.cpp:128-135
    auto& nb = ctx->neigh[j];
    const int n = nb.n * dof;
    const double* buf = p->base + p->nb[j].my_off + (size_t)b * 4 * std::max(nb.n, 1);
    const int blocks = std::max(1, std::min((n + 255) / 256, 148));
    halo_wait_add_kernel<<<blocks, 256, 0, ctx->stream>>>(nb.n, dof, nb.d_ptr, buf, V, myflags + (size_t)b * p->nranks + nb.rank, seq,
                                                         myflags + p->off_err());
    ctx->launches++;
  }
  SVB_CUDA(cudaGetLastError());
  return SVB200_OK;
}

// K*U followed by the shared-node sum (every fsils_spar_mul_* ends in fsils_commuv, spar_mul.cpp:230), dof = 4, with the
// exchange hidden behind the interior rows: interface rows + push (one kernel), interior rows, wait-add.  Returns
// handled = false when this transport / layout cannot do it (the caller then runs SpMV + halo_sum).
int launch_spmv4_rows(svb200_ctx* ctx, int row0, int nrows, const double* Val, const double* U, double* KU);
int spmv4_halo_fused(svb200_ctx* ctx, const double* Val, const double* U, double* KU, bool* handled)
{
  *handled = false;
  P2P* p = p2p_of(ctx);
  if (ctx->nranks <= 1 || !p || !p->ready || !p->contiguous || g_per_neighbour || p->nB == 0) return SVB200_OK;
  static const bool no_overlap = getenv("SVB200_HALO_NO_OVERLAP") != nullptr;       // A/B knob
  if (no_overlap) return SVB200_OK;
  const unsigned long long seq = ++p->halo_seq;
  const int b = (int)(seq & 1ull);
  const long long threads = (long long)p->nB * 8;
  bsr_spmv4_bnd_push_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(
      p->nB, p->nnb, p->d_brow, p->d_bptr, p->d_bent, p->d_rbuf + b * p->nnb, p->d_rflag + b * p->nnb, seq, p->d_count, ctx->d_rowPtr,
      ctx->d_colPtr, Val, U, KU);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  int rc = launch_spmv4_rows(ctx, p->nLow, ctx->mynNo - p->nLow, Val, U, KU);
  if (rc) return rc;
  rc = halo_wait_add_all(ctx, p, 4, KU, seq);
  if (rc) return rc;
  *handled = true;
  return SVB200_OK;
}

// Second stage of multi_dot + cross-rank sum in one kernel (p2p transport); handled = false: the caller runs
// multi_dot_stage2 + allreduce_sum.
int dot_stage2_allreduce(svb200_ctx* ctx, int nblocks, int nvec, const double* d_part, double* d_out, bool* handled)
{
  *handled = false;
  P2P* p = p2p_of(ctx);
  static const bool off = getenv("SVB200_DOT_NO_FUSE") != nullptr;                  // A/B knob
  if (ctx->nranks <= 1 || !p || !p->ready || nvec > AR_MAX || off) return SVB200_OK;
  const unsigned long long seq = ++p->ar_seq;
  dot_stage2_allreduce_kernel<<<1, 1024, 0, ctx->stream>>>(nblocks, nvec, d_part, d_out, p->nranks, p->rank, p->d_peer, p->off_arflag(),
                                                          p->off_slots(), p->off_err(), seq);
  ctx->launches++;
  SVB_CUDA(cudaGetLastError());
  *handled = true;
  return SVB200_OK;
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
  if (p2p_of(ctx) && p2p_of(ctx)->ready) return halo_sum_p2p(ctx, dof, V);
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
  P2P* p = p2p_of(ctx);
  if (p && p->ready && n <= AR_MAX) {
    const unsigned long long seq = ++p->ar_seq;
    allreduce_p2p_kernel<<<1, 512, 0, ctx->stream>>>(n, p->nranks, p->rank, d_buf, p->d_peer, p->off_arflag(), p->off_slots(),
                                                    p->off_err(), seq);
    ctx->launches++;
    SVB_CUDA(cudaGetLastError());
    return SVB200_OK;
  }
  SVB_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return SVB200_OK;
}

}  // namespace svb
