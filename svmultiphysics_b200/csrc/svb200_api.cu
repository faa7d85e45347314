// svb200_api.cu — the C ABI of include/svb200.h: host-side orchestration only, every number is
// produced by the CUDA kernels of this library.  There is no CPU fallback anywhere in this file: when
// no usable CUDA device exists svb200_create fails and nothing else can be called.
#include <algorithm>
#include <cstring>
#include <numeric>
#include "svb200_internal.h"
#include <chrono>
#include "fsils_kernels.h"

namespace svb {

static thread_local std::string g_error;

void set_error(const std::string& msg) { g_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
  g_error = std::string("svb200: CUDA error '") + cudaGetErrorString(e) + "' in " + what + " (" + file + ":" +
            std::to_string(line) + ")";
  return SVB200_ERR_CUDA;
}

int nccl_unique_id(void* id128);
int nccl_init(svb200_ctx* ctx, int nranks, int rank, const void* id128);
void nccl_destroy(svb200_ctx* ctx);
int p2p_setup(svb200_ctx* ctx);
void p2p_destroy(svb200_ctx* ctx);
int p2p_check(svb200_ctx* ctx);
const char* comm_transport(svb200_ctx* ctx);
int add_bc_mul_device(svb200_ctx* ctx, int op, int dof, const double* X, double* Y, double* d_scal);

template <class T>
static int upload(svb200_ctx* ctx, T** d, const T* h, size_t n)
{
  if (*d) { cudaFree(*d); *d = nullptr; }
  if (n == 0) return SVB200_OK;
  SVB_CUDA(cudaMalloc(d, sizeof(T) * n));
  SVB_CUDA(cudaMemcpyAsync(*d, h, sizeof(T) * n, cudaMemcpyHostToDevice, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SVB200_OK;
}

static int ensure_stage(svb200_ctx* ctx, size_t bytes)
{
  if (bytes > ctx->stage_bytes) {
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    ctx->d_stage = nullptr;
    ctx->stage_bytes = 0;
    SVB_CUDA(cudaMalloc(&ctx->d_stage, bytes));
    ctx->stage_bytes = bytes;
  }
  return SVB200_OK;
}

// Host (rows, nNo) array in caller node order -> device array in internal order.
static int upload_nodal(svb200_ctx* ctx, int rows, const double* h, double** d)
{
  const size_t n = (size_t)rows * ctx->nNo;
  if (!*d) SVB_CUDA(cudaMalloc(d, sizeof(double) * std::max<size_t>(n, 1)));
  if (n == 0) return SVB200_OK;
  if (!ctx->has_map) {
    SVB_CUDA(cudaMemcpyAsync(*d, h, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  } else {
    int rc = ensure_stage(ctx, sizeof(double) * n);
    if (rc) return rc;
    SVB_CUDA(cudaMemcpyAsync(ctx->d_stage, h, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    rc = launch_permute_cols(ctx, rows, ctx->nNo, ctx->d_map, ctx->d_stage, *d, false);
    if (rc) return rc;
  }
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SVB200_OK;
}

static int download_nodal(svb200_ctx* ctx, int rows, const double* d, double* h)
{
  const size_t n = (size_t)rows * ctx->nNo;
  if (n == 0) return SVB200_OK;
  if (!ctx->has_map) {
    SVB_CUDA(cudaMemcpyAsync(h, d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  } else {
    int rc = ensure_stage(ctx, sizeof(double) * n);
    if (rc) return rc;
    rc = launch_permute_cols(ctx, rows, ctx->nNo, ctx->d_map, d, ctx->d_stage, true);
    if (rc) return rc;
    SVB_CUDA(cudaMemcpyAsync(h, ctx->d_stage, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SVB200_OK;
}

static void free_mesh(Mesh& m)
{
  cudaFree(m.d_IEN); cudaFree(m.d_eId); cudaFree(m.d_fN); cudaFree(m.d_slot); cudaFree(m.d_color_perm); cudaFree(m.d_gcolor_perm); cudaFree(m.d_gtab);
  cudaFree(m.d_uris_mask); cudaFree(m.d_uris_list); cudaFree(m.d_thtab);
  free_group_sched(m.schedK); free_group_sched(m.schedR);
  m = Mesh();
}

static void free_face(Face& f)
{
  cudaFree(f.d_glob); cudaFree(f.d_val); cudaFree(f.d_valM);
  cudaFree(f.d_cap_glob); cudaFree(f.d_cap_val); cudaFree(f.d_cap_valM);
  f = Face();
}

// Greedy colouring of items (elements, or 128-element groups) that conflict when they share a node: item k takes the
// smallest colour none of its nodes carries yet.  The per-node colour sets are bitsets that grow on demand, so meshes with
// high-valence nodes (more than 128 conflicting colours) are coloured as well instead of being rejected.
static int greedy_color(int nNo, int nItems, const std::vector<int>& ien, size_t nodes_per_item, size_t total_nodes,
                        std::vector<int>& color)
{
  int nw = 2;                                   // 64-bit words per node
  std::vector<uint64_t> mask((size_t)nNo * nw, 0);
  color.assign(nItems, 0);
  int ncol = 0;
  std::vector<uint64_t> used;
  for (int k = 0; k < nItems; k++) {
    const size_t b = (size_t)k * nodes_per_item, e = std::min(b + nodes_per_item, total_nodes);
    used.assign(nw, 0);
    for (size_t q = b; q < e; q++)
      for (int w = 0; w < nw; w++) used[w] |= mask[(size_t)ien[q] * nw + w];
    int c = -1;
    for (int w = 0; w < nw && c < 0; w++)
      if (~used[w]) c = 64 * w + __builtin_ctzll(~used[w]);
    if (c < 0) {                                // every colour of the current width is taken: widen the bitsets
      c = 64 * nw;
      const int nw2 = nw * 2;
      std::vector<uint64_t> wide((size_t)nNo * nw2, 0);
      for (int n = 0; n < nNo; n++)
        for (int w = 0; w < nw; w++) wide[(size_t)n * nw2 + w] = mask[(size_t)n * nw + w];
      mask.swap(wide);
      nw = nw2;
    }
    color[k] = c;
    ncol = std::max(ncol, c + 1);
    for (size_t q = b; q < e; q++) mask[(size_t)ien[q] * nw + (c >> 6)] |= (1ull << (c & 63));
  }
  return ncol;
}

static int upload_color_perm(svb200_ctx* ctx, int nItems, int ncol, const std::vector<int>& color, std::vector<int>& off, int** d_perm)
{
  off.assign(ncol + 1, 0);
  for (int k = 0; k < nItems; k++) off[color[k] + 1]++;
  for (int c = 0; c < ncol; c++) off[c + 1] += off[c];
  std::vector<int> pos(off.begin(), off.end() - 1), perm(nItems);
  for (int k = 0; k < nItems; k++) perm[pos[color[k]]++] = k;
  return upload(ctx, d_perm, perm.data(), perm.size());
}

// Element colouring (two elements of one colour never share a node: a colour can be scattered with plain
// read-modify-write, the deterministic mode) and, for TET4, the colouring of the 128-element GROUPS of the grouped scatter
// (group_sched.cu): groups of one colour share no node, hence no R row and no CSR block — launched colour by colour, the
// grouped kernel adds every value in a fixed order, so the deterministic mode keeps the pre-reduction and the locality of
// the default path.
static int build_coloring(svb200_ctx* ctx, Mesh& m, const std::vector<int>& ien)
{
  std::vector<int> color;
  int ncol = greedy_color(ctx->nNo, m.nEl, ien, (size_t)m.eNoN, ien.size(), color);
  { int rc = upload_color_perm(ctx, m.nEl, ncol, color, m.color_off, &m.d_color_perm); if (rc) return rc; }
  m.gcolor_off.clear();
  if (m.eNoN != 4) return SVB200_OK;
  const int nGrp = (m.nEl + ASM_GROUP - 1) / ASM_GROUP;
  ncol = greedy_color(ctx->nNo, nGrp, ien, (size_t)ASM_GROUP * 4, ien.size(), color);
  return upload_color_perm(ctx, nGrp, ncol, color, m.gcolor_off, &m.d_gcolor_perm);
}

}  // namespace svb

using namespace svb;

#define CTX_GUARD(ctx)                                        \
  do {                                                        \
    if (!(ctx)) { set_error("svb200: null context"); return SVB200_ERR_INVALID; } \
    cudaError_t e__ = cudaSetDevice((ctx)->device);           \
    if (e__ != cudaSuccess) return cuda_fail(e__, "cudaSetDevice", __FILE__, __LINE__); \
  } while (0)

#define TRY(call)                      \
  do {                                 \
    int rc__ = (call);                 \
    if (rc__ != SVB200_OK) return rc__; \
  } while (0)

extern "C" {

int svb200_abi_version(void) { return SVB200_ABI_VERSION; }

const char* svb200_last_error(void) { return g_error.c_str(); }

int svb200_create(svb200_ctx** out, int device)
{
  if (!out) { set_error("svb200_create: null output pointer"); return SVB200_ERR_INVALID; }
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error(std::string("svb200_create: no usable CUDA device (") + cudaGetErrorString(e) +
              "); this library has no CPU path");
    return SVB200_ERR_CUDA;
  }
  if (device < 0 || device >= count) { set_error("svb200_create: bad device index"); return SVB200_ERR_INVALID; }
  SVB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  SVB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error(std::string("svb200_create: device '") + prop.name + "' is not a Blackwell (sm_100a) GPU");
    return SVB200_ERR_UNSUPPORTED;
  }
  auto* ctx = new svb200_ctx();
  ctx->device = device;
  SVB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  SVB_CUDA(cudaEventCreate(&ctx->ev0));
  SVB_CUDA(cudaEventCreate(&ctx->ev1));
  SVB_CUDA(cudaEventCreate(&ctx->tm0));
  SVB_CUDA(cudaEventCreate(&ctx->tm1));
  SVB_CUDA(cudaStreamCreateWithFlags(&ctx->zstream, cudaStreamNonBlocking));
  SVB_CUDA(cudaStreamCreateWithFlags(&ctx->dstream, cudaStreamNonBlocking));
  for (auto& row : ctx->pev) for (auto& e : row) SVB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : ctx->zev) SVB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : ctx->hev) SVB_CUDA(cudaEventCreate(&e));
  SVB_CUDA(cudaMallocHost(&ctx->h_pinned, sizeof(double) * 1024));
  *out = ctx;
  return SVB200_OK;
}

int svb200_destroy(svb200_ctx* ctx)
{
  if (!ctx) return SVB200_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  p2p_destroy(ctx);
  nccl_destroy(ctx);
  for (auto& m : ctx->mesh) free_mesh(m);
  for (auto& f : ctx->face) free_face(f);
  for (auto& f : ctx->bface) { cudaFree(f.d_IENb); cudaFree(f.d_gE); }
  cudaFree(ctx->d_hg);
  for (auto& nb : ctx->neigh) { cudaFree(nb.d_ptr); cudaFree(nb.d_send); cudaFree(nb.d_recv); }
  cudaFree(ctx->d_map); cudaFree(ctx->d_rowPtr_in); cudaFree(ctx->d_rowPtr); cudaFree(ctx->d_colPtr); cudaFree(ctx->d_diagPtr);
  cudaFree(ctx->d_x); cudaFree(ctx->d_Ag); cudaFree(ctx->d_Yg); cudaFree(ctx->d_Dg); cudaFree(ctx->d_Bf); cudaFree(ctx->d_Do); cudaFree(ctx->d_Ya); cudaFree(ctx->d_uris); cudaFree(ctx->d_pS0); cudaFree(ctx->d_pSn);
  ris_build_plan(ctx, 0, nullptr, nullptr, nullptr);
  cudaFree(ctx->d_Ao); cudaFree(ctx->d_Yo); cudaFree(ctx->d_An); cudaFree(ctx->d_Yn); cudaFree(ctx->d_Dn); cudaFree(ctx->d_nodeflag);
  cudaFree(ctx->d_err); cudaFree(ctx->d_Kd); cudaFree(ctx->d_Ad); cudaFree(ctx->d_Rd);
  cudaFree(ctx->d_stage); cudaFree(ctx->d_R); cudaFree(ctx->d_Val); cudaFree(ctx->d_W);
  cudaFree(ctx->d_work); cudaFree(ctx->d_red); cudaFree(ctx->d_tslot);
  cudaFreeHost(ctx->h_pinned); cudaFreeHost(ctx->h_cg); cudaFree(ctx->d_cg);
  cudaFree(ctx->d_shared_rows); cudaFree(ctx->d_shared_off); cudaFree(ctx->d_shared_caller); cudaFreeHost(ctx->h_shared_buf);
  for (int k = 0; k < 2; k++) if (ctx->ev_cg[k]) cudaEventDestroy(ctx->ev_cg[k]);
  cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
  for (auto& e : ctx->zev) if (e) cudaEventDestroy(e);
  if (ctx->zstream) { cudaStreamSynchronize(ctx->zstream); cudaStreamDestroy(ctx->zstream); }
  if (ctx->dstream) { cudaStreamSynchronize(ctx->dstream); cudaStreamDestroy(ctx->dstream); }
  for (auto& row : ctx->pev) for (auto& e : row) if (e) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return SVB200_OK;
}

}  // extern "C"

namespace svb {
// Zeroing with a footprint of ONE warp per SM (148 CTAs x 32 threads, <= 32 registers): cudaMemsetAsync launches a grid
// that fills the machine, so a kernel on another stream cannot start until it has drained; this kernel leaves room
// (1024 registers are exactly what three resident CTAs of the 168-register assembly kernel leave free on an SM) for the
// element kernel to run beside it.  Stores are fire-and-forget, so one warp per SM saturates the HBM write bandwidth.
__global__ void __launch_bounds__(128, 1) zero_small_footprint_kernel(double2* __restrict__ p, long long n2)
{
  const double2 z = make_double2(0.0, 0.0);
  const int nt = blockDim.x;
  const long long stride = (long long)gridDim.x * nt * 4;
  for (long long k = (long long)blockIdx.x * nt * 4 + threadIdx.x; k < n2; k += stride) {
    p[k] = z;
    if (k + nt < n2) p[k + nt] = z;
    if (k + 2 * nt < n2) p[k + 2 * nt] = z;
    if (k + 3 * nt < n2) p[k + 3 * nt] = z;
  }
}

// svb200_alloc may defer the zeroing of Val (val_zero_pending): whoever touches Val next zeroes it first.
int flush_val_zero(svb200_ctx* ctx)
{
  if (!ctx->val_zero_pending) return SVB200_OK;
  ctx->val_zero_pending = false;
  const size_t nV = (size_t)ctx->dof * ctx->dof * ctx->nnz;
  if (nV) SVB_CUDA(cudaMemsetAsync(ctx->d_Val, 0, sizeof(double) * nV, ctx->stream));
  return SVB200_OK;
}
}  // namespace svb

extern "C" {

int svb200_comm_unique_id(void* id128)
{
  if (!id128) { set_error("svb200_comm_unique_id: null buffer"); return SVB200_ERR_INVALID; }
  return nccl_unique_id(id128);
}

int svb200_comm_init(svb200_ctx* ctx, int nranks, int rank, const void* id128)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks && id128, "svb200_comm_init: bad arguments");
  TRY(nccl_init(ctx, nranks, rank, id128));
  if (ctx->d_rowPtr) TRY(p2p_setup(ctx));     // graph already set: map the peers' mailboxes now
  return SVB200_OK;
}

const char* svb200_comm_transport(svb200_ctx* ctx)
{
  if (!ctx) return "none";
  return comm_transport(ctx);
}

int svb200_set_graph(svb200_ctx* ctx, int32_t nNo, int32_t nnz, const int32_t* rowPtr, const int32_t* colPtr,
                     int32_t mynNo, const int32_t* map, int32_t nReq, const int32_t* neigh_rank,
                     const int32_t* neigh_n, const int32_t* neigh_ptr)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(nNo >= 0 && nnz >= 0 && rowPtr && (colPtr || nnz == 0), "svb200_set_graph: bad arguments");
  SVB_REQUIRE(rowPtr[0] == 0 && rowPtr[nNo] == nnz, "svb200_set_graph: rowPtr does not span [0,nnz]");
  SVB_REQUIRE(mynNo >= 0 && mynNo <= nNo, "svb200_set_graph: mynNo out of range");
  ctx->nNo = nNo;
  ctx->nnz = nnz;
  ctx->mynNo = mynNo;
  ctx->has_map = (map != nullptr);
  ctx->h_rowPtr_in.assign(rowPtr, rowPtr + nNo + 1);
  ctx->h_map.resize(nNo);
  if (map) {
    std::vector<char> seen(nNo, 0);
    for (int a = 0; a < nNo; a++) {
      SVB_REQUIRE(map[a] >= 0 && map[a] < nNo && !seen[map[a]], "svb200_set_graph: map is not a permutation");
      seen[map[a]] = 1;
      ctx->h_map[a] = map[a];
    }
  } else {
    std::iota(ctx->h_map.begin(), ctx->h_map.end(), 0);
  }
  // internal CSR: rows in FSILS order, columns of a row kept in caller order
  std::vector<int> len(nNo);
  for (int a = 0; a < nNo; a++) {
    SVB_REQUIRE(rowPtr[a + 1] >= rowPtr[a], "svb200_set_graph: rowPtr is not monotone");
    len[ctx->h_map[a]] = rowPtr[a + 1] - rowPtr[a];
  }
  ctx->h_rowPtr.assign(nNo + 1, 0);
  for (int r = 0; r < nNo; r++) ctx->h_rowPtr[r + 1] = ctx->h_rowPtr[r] + len[r];
  std::vector<int> col(nnz);
  for (int a = 0; a < nNo; a++) {
    const int base = ctx->h_rowPtr[ctx->h_map[a]];
    for (int k = rowPtr[a]; k < rowPtr[a + 1]; k++) {
      SVB_REQUIRE(colPtr[k] >= 0 && colPtr[k] < nNo, "svb200_set_graph: column index out of range");
      col[base + (k - rowPtr[a])] = ctx->h_map[colPtr[k]];
    }
  }
  TRY(upload(ctx, &ctx->d_rowPtr, ctx->h_rowPtr.data(), (size_t)nNo + 1));
  TRY(upload(ctx, &ctx->d_colPtr, col.data(), (size_t)nnz));
  ctx->h_colPtr.swap(col);
  TRY(upload(ctx, &ctx->d_map, ctx->h_map.data(), (size_t)nNo));
  TRY(upload(ctx, &ctx->d_rowPtr_in, ctx->h_rowPtr_in.data(), (size_t)nNo + 1));
  if (ctx->d_diagPtr) { cudaFree(ctx->d_diagPtr); ctx->d_diagPtr = nullptr; }
  if (ctx->d_tslot) { cudaFree(ctx->d_tslot); ctx->d_tslot = nullptr; }
  SVB_CUDA(cudaMalloc(&ctx->d_diagPtr, sizeof(int) * std::max(nNo, 1)));
  if (nNo > 0) TRY(launch_find_diag(ctx));

  for (auto& nb : ctx->neigh) { cudaFree(nb.d_ptr); cudaFree(nb.d_send); cudaFree(nb.d_recv); }
  ctx->neigh.clear();
  size_t off = 0;
  for (int i = 0; i < nReq; i++) {
    SVB_REQUIRE(neigh_rank && neigh_n && neigh_ptr, "svb200_set_graph: neighbour lists missing");
    Neighbor nb;
    nb.rank = neigh_rank[i];
    nb.n = neigh_n[i];
    for (int k = 0; k < nb.n; k++)
      SVB_REQUIRE(neigh_ptr[off + k] >= 0 && neigh_ptr[off + k] < nNo, "svb200_set_graph: shared node id out of range");
    TRY(upload(ctx, &nb.d_ptr, neigh_ptr + off, (size_t)nb.n));
    nb.h_ptr.assign(neigh_ptr + off, neigh_ptr + off + nb.n);
    SVB_CUDA(cudaMalloc(&nb.d_send, sizeof(double) * 4 * std::max(nb.n, 1)));
    SVB_CUDA(cudaMalloc(&nb.d_recv, sizeof(double) * 4 * std::max(nb.n, 1)));
    off += nb.n;
    ctx->neigh.push_back(nb);
  }
  std::sort(ctx->neigh.begin(), ctx->neigh.end(), [](const Neighbor& a, const Neighbor& b) { return a.rank < b.rank; });
  ctx->shared_built = false;
  TRY(p2p_setup(ctx));      // collective: every rank calls svb200_set_graph
  // state arrays depend on nNo
  cudaFree(ctx->d_x); cudaFree(ctx->d_Ag); cudaFree(ctx->d_Yg); cudaFree(ctx->d_Dg); cudaFree(ctx->d_Bf); cudaFree(ctx->d_Do);
  cudaFree(ctx->d_Ya); ctx->d_Ya = nullptr; ctx->ya_sn_positive = false;
  cudaFree(ctx->d_uris); ctx->d_uris = nullptr; ctx->nUris = 0;
  ris_build_plan(ctx, 0, nullptr, nullptr, nullptr);
  cudaFree(ctx->d_pS0); cudaFree(ctx->d_pSn); ctx->d_pS0 = ctx->d_pSn = nullptr;
  cudaFree(ctx->d_Ao); cudaFree(ctx->d_Yo); cudaFree(ctx->d_An); cudaFree(ctx->d_Yn); cudaFree(ctx->d_Dn); cudaFree(ctx->d_nodeflag);
  ctx->d_Ao = ctx->d_Yo = ctx->d_An = ctx->d_Yn = ctx->d_Dn = nullptr; ctx->d_nodeflag = nullptr;
  ctx->d_x = ctx->d_Ag = ctx->d_Yg = ctx->d_Dg = ctx->d_Bf = ctx->d_Do = nullptr;
  ctx->bf_set = false;
  ctx->tDof = 0;
  return SVB200_OK;
}

int svb200_set_mesh(svb200_ctx* ctx, int32_t iM, int32_t eNoN, int32_t nEl, const int32_t* IEN, const int32_t* eId,
                    int32_t nFn, const double* fN, int32_t nG, const double* w, const double* N, const double* Nx)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ctx->d_rowPtr, "svb200_set_mesh: call svb200_set_graph first");
  SVB_REQUIRE(iM >= 0 && iM < 64, "svb200_set_mesh: bad mesh index");
  SVB_REQUIRE(eNoN >= 1 && eNoN <= MAX_ENON_ANY && nG >= 1 && nG <= MAX_NG_ANY, "svb200_set_mesh: unsupported eNoN / nG");
  SVB_REQUIRE(nEl >= 0 && (IEN || nEl == 0) && w && N && Nx, "svb200_set_mesh: bad arguments");
  if ((int)ctx->mesh.size() <= iM) ctx->mesh.resize(iM + 1);
  Mesh& m = ctx->mesh[iM];
  free_mesh(m);
  m.eNoN = eNoN; m.nEl = nEl; m.nG = nG; m.nFn = nFn;
  std::vector<int> ien((size_t)eNoN * nEl);
  for (size_t k = 0; k < ien.size(); k++) {
    SVB_REQUIRE(IEN[k] >= 0 && IEN[k] < ctx->nNo, "svb200_set_mesh: IEN entry out of range");
    ien[k] = ctx->h_map[IEN[k]];
  }
  TRY(upload(ctx, &m.d_IEN, ien.data(), ien.size()));
  if (eNoN == 4 && nEl > 0) {
    // caller-order node window of every 128-element group (svb200_assemble_host pipelines uploads / downloads on it)
    const int nGrp = (nEl + ASM_GROUP - 1) / ASM_GROUP;
    std::vector<int> gmax(nGrp, -1), gmin(nGrp, ctx->nNo);
    for (int e = 0; e < nEl; e++)
      for (int a = 0; a < 4; a++) {
        const int n = IEN[(size_t)e * 4 + a], g = e / ASM_GROUP;
        gmax[g] = std::max(gmax[g], n); gmin[g] = std::min(gmin[g], n);
      }
    m.grp_node_need.resize(nGrp); m.grp_node_done.resize(nGrp);
    int run = 0;
    for (int g = 0; g < nGrp; g++) { run = std::max(run, gmax[g] + 1); m.grp_node_need[g] = run; }
    run = ctx->nNo;
    for (int g = nGrp - 1; g >= 0; g--) { run = std::min(run, gmin[g]); m.grp_node_done[g] = run; }
  }
  if (eId) TRY(upload(ctx, &m.d_eId, eId, (size_t)nEl));
  if (fN && nFn > 0) TRY(upload(ctx, &m.d_fN, fN, (size_t)3 * nFn * nEl));
  m.w.assign(w, w + nG);
  m.N.assign(N, N + (size_t)eNoN * nG);
  m.Nx.assign(Nx, Nx + (size_t)3 * eNoN * nG);
  TRY(launch_build_slot_map(ctx, m));
  TRY(build_group_schedules(ctx, m));
  TRY(build_group_slot_need(ctx, m));
  TRY(build_coloring(ctx, m, ien));
  TRY(upload_fluid_gen_tables(ctx, m));
  m.set = true;
  return SVB200_OK;
}

int svb200_set_mesh_nxx(svb200_ctx* ctx, int32_t iM, const double* Nxx)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(iM >= 0 && iM < (int)ctx->mesh.size() && ctx->mesh[iM].set && Nxx, "svb200_set_mesh_nxx: mesh not set");
  Mesh& m = ctx->mesh[iM];
  m.Nxx.assign(Nxx, Nxx + (size_t)6 * m.eNoN * m.nG);
  return upload_fluid_gen_tables(ctx, m);
}

int svb200_set_mesh_thood(svb200_ctx* ctx, int32_t iM, int32_t eNoNq, int32_t nG2, int32_t lShpF_q, const double* Nq1, const double* Nqxi1,
                          const double* w2, const double* Nw2, const double* Nwxi2, const double* Nq2, const double* Nqxi2)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(iM >= 0 && iM < (int)ctx->mesh.size() && ctx->mesh[iM].set, "svb200_set_mesh_thood: mesh not set");
  Mesh& m = ctx->mesh[iM];
  if (eNoNq <= 0) {                                  // back to equal-order function spaces
    cudaFree(m.d_thtab); m.d_thtab = nullptr; m.th_eNoNq = m.th_nG2 = m.th_lShpFq = 0;
    return SVB200_OK;
  }
  SVB_REQUIRE(eNoNq < m.eNoN && nG2 >= 1 && Nq1 && Nqxi1 && w2 && Nw2 && Nwxi2 && Nq2 && Nqxi2, "svb200_set_mesh_thood: bad arguments");
  TRY(upload_thood_tables(ctx, m, eNoNq, nG2, Nq1, Nqxi1, w2, Nw2, Nwxi2, Nq2, Nqxi2));
  m.th_eNoNq = eNoNq; m.th_nG2 = nG2; m.th_lShpFq = lShpF_q ? 1 : 0;
  return SVB200_OK;
}

int svb200_thood_val_rc(svb200_ctx* ctx)
{
  CTX_GUARD(ctx);
  TRY(flush_val_zero(ctx));
  return run_thood_val_rc(ctx);
}

int svb200_set_coords(svb200_ctx* ctx, const double* x)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ctx->d_rowPtr && x, "svb200_set_coords: call svb200_set_graph first");
  for (auto& m : ctx->mesh) m.jac_checked = false;
  return upload_nodal(ctx, 3, x, &ctx->d_x);
}

int svb200_set_num_faces(svb200_ctx* ctx, int32_t nFaces)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(nFaces >= 0, "svb200_set_num_faces: negative count");
  for (auto& f : ctx->face) free_face(f);
  ctx->face.assign(nFaces, Face());
  return SVB200_OK;
}

int svb200_set_face(svb200_ctx* ctx, int32_t faIn, int32_t bGrp, int32_t face_dof, int32_t nNo, const int32_t* glob,
                    const double* val, int32_t sharedFlag)
{
  CTX_GUARD(ctx);
  // same messages as fsils_bc_create (linear_solver/bc.cpp:25-32)
  if (faIn >= (int)ctx->face.size()) {
    set_error("FSILS: faIn is exceeding lhs structure maximum number of faces (" + std::to_string(ctx->face.size()) +
              ") is less than " + std::to_string(faIn) + ".");
    return SVB200_ERR_INVALID;
  }
  if (faIn <= -1) { set_error("FSILS: faIn is smaller than zero"); return SVB200_ERR_INVALID; }
  SVB_REQUIRE(face_dof >= 1 && face_dof <= 4 && nNo >= 0 && (glob || nNo == 0), "svb200_set_face: bad arguments");
  Face& f = ctx->face[faIn];
  free_face(f);
  f.bGrp = bGrp; f.dof = face_dof; f.nNo = nNo; f.shared = sharedFlag != 0;
  std::vector<int> g(nNo);
  for (int a = 0; a < nNo; a++) {
    SVB_REQUIRE(glob[a] >= 0 && glob[a] < ctx->nNo, "svb200_set_face: node id out of range");
    g[a] = ctx->h_map[glob[a]];
  }
  std::vector<double> v((size_t)face_dof * nNo, 0.0);
  if (val) std::copy(val, val + v.size(), v.begin());
  TRY(upload(ctx, &f.d_glob, g.data(), g.size()));
  TRY(upload(ctx, &f.d_val, v.data(), v.size()));
  if (!v.empty()) {
    SVB_CUDA(cudaMalloc(&f.d_valM, sizeof(double) * v.size()));
    SVB_CUDA(cudaMemsetAsync(f.d_valM, 0, sizeof(double) * v.size(), ctx->stream));
  }
  if (sharedFlag == 1 && ctx->nranks > 1 && val) {
    // sharedFlag 2 = lhs.face[].val as the host's fsils_bc_create left it (already summed).
    // fsils_bc_create sums the face values of shared nodes across partitions (bc.cpp:70-102).
    const size_t n = (size_t)face_dof * ctx->nNo;
    TRY(ensure_stage(ctx, sizeof(double) * n));
    SVB_CUDA(cudaMemsetAsync(ctx->d_stage, 0, sizeof(double) * n, ctx->stream));
    std::vector<double> full(n, 0.0);
    for (int a = 0; a < nNo; a++)
      for (int i = 0; i < face_dof; i++) full[(size_t)g[a] * face_dof + i] = v[(size_t)a * face_dof + i];
    SVB_CUDA(cudaMemcpyAsync(ctx->d_stage, full.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    TRY(halo_sum(ctx, face_dof, ctx->d_stage));
    SVB_CUDA(cudaMemcpyAsync(full.data(), ctx->d_stage, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    SVB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int a = 0; a < nNo; a++)
      for (int i = 0; i < face_dof; i++) v[(size_t)a * face_dof + i] = full[(size_t)g[a] * face_dof + i];
    SVB_CUDA(cudaMemcpyAsync(f.d_val, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice, ctx->stream));
  }
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  f.set = true;
  return SVB200_OK;
}

int svb200_set_face_cap(svb200_ctx* ctx, int32_t faIn, int32_t cap_nNo, const int32_t* cap_glob, const double* cap_val)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(faIn >= 0 && faIn < (int)ctx->face.size() && ctx->face[faIn].set, "svb200_set_face_cap: call svb200_set_face first");
  SVB_REQUIRE(cap_nNo >= 0 && (cap_nNo == 0 || (cap_glob && cap_val)), "svb200_set_face_cap: bad arguments");
  Face& f = ctx->face[faIn];
  cudaFree(f.d_cap_glob); cudaFree(f.d_cap_val); cudaFree(f.d_cap_valM);
  f.d_cap_glob = nullptr; f.d_cap_val = f.d_cap_valM = nullptr; f.cap_n = 0;
  // cap nodes that are not on this partition carry a negative id (add_bc_mul.cpp:70) and are dropped here
  std::vector<int> g;
  std::vector<double> v;
  for (int a = 0; a < cap_nNo; a++) {
    if (cap_glob[a] < 0) continue;
    SVB_REQUIRE(cap_glob[a] < ctx->nNo, "svb200_set_face_cap: node id out of range");
    g.push_back(ctx->h_map[cap_glob[a]]);
    for (int i = 0; i < f.dof; i++) v.push_back(cap_val[(size_t)a * f.dof + i]);
  }
  f.cap_n = (int)g.size();
  f.has_cap = true;           // all ranks agree on the flag even when a rank holds no cap node (collective sums)
  if (f.cap_n == 0) return SVB200_OK;
  TRY(upload(ctx, &f.d_cap_glob, g.data(), g.size()));
  TRY(upload(ctx, &f.d_cap_val, v.data(), v.size()));
  SVB_CUDA(cudaMalloc(&f.d_cap_valM, sizeof(double) * v.size()));
  SVB_CUDA(cudaMemsetAsync(f.d_cap_valM, 0, sizeof(double) * v.size(), ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SVB200_OK;
}

int svb200_alloc(svb200_ctx* ctx, int32_t dof)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ctx->d_rowPtr, "svb200_alloc: call svb200_set_graph first");
  SVB_REQUIRE(dof >= 1 && dof <= 4, "svb200_alloc: dof must be in [1,4]");
  const size_t nR = (size_t)dof * ctx->nNo, nV = (size_t)dof * dof * ctx->nnz;
  if (nR > ctx->R_cap) {
    cudaFree(ctx->d_R); ctx->d_R = nullptr;
    SVB_CUDA(cudaMalloc(&ctx->d_R, sizeof(double) * std::max<size_t>(nR, 1)));
    ctx->R_cap = nR;
  }
  if (nV > ctx->Val_cap) {
    cudaFree(ctx->d_Val); ctx->d_Val = nullptr;
    SVB_CUDA(cudaMalloc(&ctx->d_Val, sizeof(double) * std::max<size_t>(nV, 1)));
    ctx->Val_cap = nV;
  }
  ctx->dof = dof;
  if (nR) SVB_CUDA(cudaMemsetAsync(ctx->d_R, 0, sizeof(double) * nR, ctx->stream));
  // Val (3.3 GB at 10 M tets) is zeroed here.  SVB200_LAZY_ZERO=1 defers it so that the TET4 fluid assembly overlaps the zeroing
  // with its first chunk of elements (run_assemble; every other consumer of Val then zeroes it in full first, flush_val_zero).
  // Measured on B200 (profiles/r2k_zero_overlap_ab.txt): the one-warp-per-SM zeroing kernel does reach 6.4 TB/s beside the element
  // kernel, but the assembly stage does not get shorter (4.87-4.94 ms against 4.86-4.93 ms): the chip sits at its power cap during
  // this FP64-heavy stage, so the stage is bound by the energy of the work, which overlapping does not change.  Kept as an option.
  static const bool lazy_zero = getenv("SVB200_LAZY_ZERO") != nullptr;
  ctx->val_zero_pending = false;
  if (nV) {
    if (!lazy_zero) SVB_CUDA(cudaMemsetAsync(ctx->d_Val, 0, sizeof(double) * nV, ctx->stream));
    else ctx->val_zero_pending = true;
  }
  // com_mod.Kd is zeroed with the linear system (solver/Integrator.cpp:106-109)
  if (ctx->d_Kd && ctx->nnz) SVB_CUDA(cudaMemsetAsync(ctx->d_Kd, 0, sizeof(double) * 12 * (size_t)ctx->nnz, ctx->stream));
  // the prestress accumulators pSn / pSa are zeroed once per Newton iteration by Integrator::initiator (Integrator.cpp:745-748)
  if (ctx->d_pSn && ctx->nNo) SVB_CUDA(cudaMemsetAsync(ctx->d_pSn, 0, sizeof(double) * 7 * (size_t)ctx->nNo, ctx->stream));
  if (ctx->d_Rd && ctx->nNo) SVB_CUDA(cudaMemsetAsync(ctx->d_Rd, 0, sizeof(double) * 3 * (size_t)ctx->nNo, ctx->stream));
  return SVB200_OK;
}

int svb200_set_state(svb200_ctx* ctx, int32_t tDof, const double* Ag, const double* Yg, const double* Dg,
                     const double* Bf)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ctx->d_rowPtr, "svb200_set_state: call svb200_set_graph first");
  SVB_REQUIRE(tDof >= 1 && tDof <= 16, "svb200_set_state: bad tDof");
  if (tDof != ctx->tDof) {
    cudaFree(ctx->d_Ag); cudaFree(ctx->d_Yg); cudaFree(ctx->d_Dg); cudaFree(ctx->d_Do);
    cudaFree(ctx->d_Ao); cudaFree(ctx->d_Yo); cudaFree(ctx->d_An); cudaFree(ctx->d_Yn); cudaFree(ctx->d_Dn);
    ctx->d_Ag = ctx->d_Yg = ctx->d_Dg = ctx->d_Do = ctx->d_Ao = ctx->d_Yo = ctx->d_An = ctx->d_Yn = ctx->d_Dn = nullptr;
    ctx->tDof = tDof;
  }
  const size_t n = (size_t)tDof * ctx->nNo;
  auto zero_if_new = [&](double** d, size_t cnt) -> int {
    if (!*d) {
      SVB_CUDA(cudaMalloc(d, sizeof(double) * std::max<size_t>(cnt, 1)));
      SVB_CUDA(cudaMemsetAsync(*d, 0, sizeof(double) * cnt, ctx->stream));
    }
    return SVB200_OK;
  };
  TRY(zero_if_new(&ctx->d_Ag, n));
  TRY(zero_if_new(&ctx->d_Yg, n));
  TRY(zero_if_new(&ctx->d_Dg, n));
  TRY(zero_if_new(&ctx->d_Bf, (size_t)3 * ctx->nNo));
  if (Ag) TRY(upload_nodal(ctx, tDof, Ag, &ctx->d_Ag));
  if (Yg) TRY(upload_nodal(ctx, tDof, Yg, &ctx->d_Yg));
  if (Dg) TRY(upload_nodal(ctx, tDof, Dg, &ctx->d_Dg));
  if (Bf) { TRY(upload_nodal(ctx, 3, Bf, &ctx->d_Bf)); ctx->bf_set = true; }
  return SVB200_OK;
}

static int fill_fluid_args(svb200_ctx* ctx, const Mesh& m, const svb200_eqparams* eq, const svb200_dmnparams* dmn,
                           int nDmn, FluidArgs& A)
{
  SVB_REQUIRE(nDmn >= 1 && nDmn <= MAX_DMN, "svb200_assemble: between 1 and 8 domains are supported");
  SVB_REQUIRE(eq->dof == 4 && ctx->dof == 4, "svb200_assemble: the fluid equation has dof = 4 (call svb200_alloc(4))");
  SVB_REQUIRE((eq->vmsStab == 1) == (m.th_eNoNq == 0),
              "svb200_assemble: vmsStab = 1 goes with equal-order function spaces, vmsStab = 0 with a Taylor-Hood mesh (svb200_set_mesh_thood)");
  SVB_REQUIRE(eq->tDof == ctx->tDof && ctx->d_Yg, "svb200_assemble: state not set (svb200_set_state) or tDof mismatch");
  SVB_REQUIRE(!eq->mvMsh || eq->tDof >= 7, "svb200_assemble: mvMsh needs the mesh velocity in state dofs 4..6");
  SVB_REQUIRE(ctx->d_x, "svb200_assemble: coordinates not set");
  memset(&A, 0, sizeof(A));
  A.IEN = m.d_IEN; A.eId = m.d_eId; A.slot = m.d_slot; A.perm = nullptr;
  A.kU_ptr = m.schedK.d_uptr; A.kU_ent = m.schedK.d_uent; A.kU_partner = m.schedK.d_upartner; A.kContrib = m.schedK.d_contrib;
  A.rU_ptr = m.schedR.d_uptr; A.rU_ent = m.schedR.d_uent; A.rContrib = m.schedR.d_contrib;
  A.x = ctx->d_x; A.Ag = ctx->d_Ag; A.Yg = ctx->d_Yg; A.Bf = ctx->d_Bf; A.Dg = ctx->d_Dg;
  // no body-force array was ever uploaded: com_mod.Bf is identically zero and the TET4 kernels skip the 12 loads per element
  // (a - 0.0 == a bit for bit); SVB200_BF_ALWAYS=1 keeps the gather (A/B)
  static const bool bf_always = getenv("SVB200_BF_ALWAYS") != nullptr;
  static const bool gather_scalar = getenv("SVB200_GATHER_SCALAR") != nullptr;     // A/B: skip Bf but keep the 8-byte loads
  A.bfZero = (!ctx->bf_set && !bf_always) ? (gather_scalar ? 1 : 2) : 0;
  A.R = ctx->d_R; A.Val = ctx->d_Val;
  if (!ctx->d_err) {
    SVB_CUDA(cudaMalloc(&ctx->d_err, sizeof(int)));
    SVB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
  }
  A.err = ctx->d_err;
  A.e0 = 0; A.e1 = m.nEl;
  A.tDof = eq->tDof; A.mvMsh = eq->mvMsh; A.nDmn = nDmn;
  A.uris = ctx->nUris ? ctx->d_uris : nullptr;
  A.nUris = ctx->nUris;
  for (int v = 0; v < ctx->nUris; v++) A.urisP[v] = ctx->urisP[v];
  A.ale = (eq->phys == SVB200_PHYS_FSI);
  SVB_REQUIRE(!A.ale || (eq->tDof >= 7 && ctx->d_Dg), "svb200_assemble: FSI needs tDof >= 7 and the displacement state");
  A.atomic = (eq->scatter == SVB200_SCATTER_ATOMIC);
  A.dt = eq->dt; A.af = eq->af; A.am = eq->am; A.gam = eq->gam;
  // (the general kernel reads the tables of the quadratic elements from m.d_gtab; the fixed-size copy serves TET4 / HEX8)
  if (m.nG <= MAX_NG && m.eNoN <= MAX_ENON)
    for (int g = 0; g < m.nG; g++) {
      A.w[g] = m.w[g];
      for (int a = 0; a < m.eNoN; a++) {
        A.N[g][a] = m.N[(size_t)g * m.eNoN + a];
        for (int k = 0; k < 3; k++) A.Nxi[g][a][k] = m.Nx[((size_t)g * m.eNoN + a) * 3 + k];
      }
    }
  bool needEId = false;
  for (int d = 0; d < nDmn; d++) {
    FluidDmn& o = A.dmn[d];
    o.rho = dmn[d].rho;
    for (int k = 0; k < 3; k++) o.f[k] = dmn[d].f[k];
    o.Kd = A.ale ? 0.0 : dmn[d].K_darcy;   // construct_fsi passes K_inverse_darcy_permeability = 0 (fsi.cpp:215,313)
    o.mu_i = dmn[d].mu_i; o.mu_o = dmn[d].mu_o; o.lam = dmn[d].lam; o.a = dmn[d].a; o.n = dmn[d].n;
    o.viscType = dmn[d].viscType;
    o.Id = dmn[d].Id;
    o.isFluid = (dmn[d].phys == SVB200_PHYS_FLUID);
    if (o.Id != -1) needEId = true;
    SVB_REQUIRE(o.Id >= -1 && o.Id < 31, "svb200_assemble: domain Id out of range");
    if (o.Id == -1) break;
  }
  // all_fun::domain throws "eId is not allocated" when no domain covers the whole mesh (all_fun.cpp:137-139)
  if (needEId && A.dmn[0].Id != -1) {
    bool whole = false;
    for (int d = 0; d < nDmn; d++) whole |= (dmn[d].Id == -1);
    if (!whole && !m.d_eId) { set_error("eId is not allocated"); return SVB200_ERR_INVALID; }
  }
  return SVB200_OK;
}

// construct_fluid / construct_fsi throw "Jacobian for element e is < 0." when utils::is_zero(Jac) (fluid.cpp:637-647,
// fsi.cpp:185-193); the kernels leave 1 + e in the device error word.
static int check_jacobian_word(svb200_ctx* ctx, bool fsi)
{
  int e = 0;
  SVB_CUDA(cudaMemcpyAsync(&e, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (e == 0) return SVB200_OK;
  SVB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
  set_error(std::string(fsi ? "[construct_fsi]" : "[construct_fluid]") + " Jacobian for element " + std::to_string(e - 1) + " is < 0.");
  return SVB200_ERR_NUMERIC;
}

static int run_assemble(svb200_ctx* ctx, const Mesh& m, FluidArgs& A, bool general = false)
{
  // linear tetrahedra have their own kernel (constant gradients, no second derivatives); everything else, or
  // SVB200_EQ_GENERAL_KERNEL, goes through the per-Gauss-point kernel of assemble_fluid_gen.cu
  // URIS valves on a TET4 mesh (atomic scatter): split launch — the band of elements around the valves through the per-Gauss-point
  // kernel, which has the URIS terms, everything else through the closed-form kernel below (exact there: zero valve factor)
  // Taylor-Hood function spaces (mshType::nFs = 2): construct_fluid's vmsStab = false branch has its own kernel
  if (m.th_eNoNq > 0) {
    TRY(flush_val_zero(ctx));
    TRY(run_assemble_fluid_thood(ctx, m, A));
    return check_jacobian_word(ctx, A.ale != 0);
  }
  static const bool uris_no_split = getenv("SVB200_URIS_NO_SPLIT") != nullptr;       // A/B: all elements through the general kernel
  const bool uris_split = A.nUris > 0 && m.eNoN == 4 && m.nG == 4 && !general && A.atomic && !uris_no_split;
  if (uris_split) {
    TRY(build_uris_element_mask(ctx, m));
    TRY(flush_val_zero(ctx));
    if (m.n_uris_el > 0) {
      FluidArgs B = A;
      B.emask = m.d_uris_mask; B.emask_val = 1;
      TRY(run_assemble_fluid_gen(ctx, m, B));
      TRY(check_jacobian_word(ctx, A.ale != 0));
    }
    A.emask = m.d_uris_mask; A.emask_val = 0;
    A.uris = nullptr; A.nUris = 0;
  } else if (m.eNoN != 4 || general || A.nUris > 0) {        // the URIS terms exist in the per-Gauss-point kernel only
    TRY(flush_val_zero(ctx));
    TRY(run_assemble_fluid_gen(ctx, m, A));
    return check_jacobian_word(ctx, A.ale != 0);
  }
  if (A.ale || !m.jac_checked) {
    TRY(launch_tet4_jacobian_check(ctx, m, A));
    TRY(check_jacobian_word(ctx, A.ale != 0));
    m.jac_checked = true;
  }
  if (A.atomic) {
    static const bool legacy = getenv("SVB200_ASM_LEGACY") != nullptr;
    const int nGrp = (m.nEl + ASM_GROUP - 1) / ASM_GROUP;
    if (ctx->val_zero_pending && !legacy && A.kU_ptr && (int)m.grp_need.size() == nGrp && nGrp >= 64) {
      // Overlapped zeroing: Val is zeroed on a second stream while the first chunk of element groups runs.  Chunk 0 = the
      // first ~1/7 of the groups (its kernel time covers the rest of the memset); it waits only for the part of Val its groups
      // add to (grp_need), chunk 1 for all of it.
      ctx->val_zero_pending = false;
      const long long nnz = ctx->nnz;
      const size_t blk = sizeof(double) * 16;
      static const int zfrac = getenv("SVB200_ZERO_CHUNK0_DIV") ? atoi(getenv("SVB200_ZERO_CHUNK0_DIV")) : 7;
      const int g1 = std::max(1, nGrp / std::max(zfrac, 2));
      const long long need0 = std::min(nnz, m.grp_need[g1 - 1]);
      SVB_CUDA(cudaEventRecord(ctx->zev[0], ctx->stream));               // Val's previous readers (the last solve) are done
      SVB_CUDA(cudaStreamWaitEvent(ctx->zstream, ctx->zev[0], 0));
      if (need0 > 0) SVB_CUDA(cudaMemsetAsync(ctx->d_Val, 0, blk * need0, ctx->zstream));
      SVB_CUDA(cudaEventRecord(ctx->zev[1], ctx->zstream));
      if (nnz > need0) {
        static const bool plain_memset = getenv("SVB200_ZERO_MEMSET") != nullptr;      // A/B knob
        if (plain_memset) {
          SVB_CUDA(cudaMemsetAsync(ctx->d_Val + 16 * need0, 0, blk * (nnz - need0), ctx->zstream));
        } else {
          int nsm = 148;
          cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
          static const int zthreads = getenv("SVB200_ZERO_THREADS") ? atoi(getenv("SVB200_ZERO_THREADS")) : 32;   // tuning knobs
          static const int zmult = getenv("SVB200_ZERO_CTAS_PER_SM") ? atoi(getenv("SVB200_ZERO_CTAS_PER_SM")) : 1;
          zero_small_footprint_kernel<<<nsm * zmult, zthreads, 0, ctx->zstream>>>(reinterpret_cast<double2*>(ctx->d_Val + 16 * need0),
                                                                                   8 * (nnz - need0));
          ctx->launches++;
        }
      }
      SVB_CUDA(cudaEventRecord(ctx->zev[2], ctx->zstream));
      SVB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->zev[1], 0));
      A.g0 = 0; A.nGrpLaunch = g1;
      TRY(launch_assemble_fluid(ctx, m, A));
      SVB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->zev[2], 0));
      A.g0 = g1; A.nGrpLaunch = nGrp - g1;
      if (A.nGrpLaunch > 0) TRY(launch_assemble_fluid(ctx, m, A));
      A.g0 = 0; A.nGrpLaunch = 0;
      return SVB200_OK;
    }
    TRY(flush_val_zero(ctx));
    return launch_assemble_fluid(ctx, m, A);
  }
  TRY(flush_val_zero(ctx));
  // deterministic mode: the grouped kernel, one launch per GROUP colour (SVB200_ASM_LEGACY=1: per-element colours, plain RMW)
  static const bool legacy_colored = getenv("SVB200_ASM_LEGACY") != nullptr;
  if (!legacy_colored && A.kU_ptr && m.d_gcolor_perm && !m.gcolor_off.empty()) {
    A.gperm = m.d_gcolor_perm;
    for (size_t c = 0; c + 1 < m.gcolor_off.size(); c++) {
      A.g0 = m.gcolor_off[c];
      A.nGrpLaunch = m.gcolor_off[c + 1] - m.gcolor_off[c];
      TRY(launch_assemble_fluid(ctx, m, A));
    }
    return SVB200_OK;
  }
  A.perm = m.d_color_perm;
  for (size_t c = 0; c + 1 < m.color_off.size(); c++) {
    A.e0 = m.color_off[c];
    A.e1 = m.color_off[c + 1];
    TRY(launch_assemble_fluid(ctx, m, A));
  }
  return SVB200_OK;
}

int svb200_set_prestress(svb200_ctx* ctx, const double* pS0)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ctx->d_rowPtr, "svb200_set_prestress: set the graph first");
  if (!pS0) { cudaFree(ctx->d_pS0); ctx->d_pS0 = nullptr; return SVB200_OK; }
  return upload_nodal(ctx, 6, pS0, &ctx->d_pS0);
}

int svb200_get_prestress(svb200_ctx* ctx, double* pSn, double* pSa)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(pSn && pSa, "svb200_get_prestress: null output");
  SVB_REQUIRE(ctx->d_pSn, "svb200_get_prestress: no prestress equation (SVB200_EQ_PRESTRESS) has been assembled");
  const size_t n = std::max<size_t>((size_t)ctx->nNo, 1);
  TRY(download_nodal(ctx, 6, ctx->d_pSn, pSn));
  return download_nodal(ctx, 1, ctx->d_pSn + 6 * n, pSa);
}

int svb200_set_active_tension(svb200_ctx* ctx, const double* Ya_f, const double* Ya_s, const double* Ya_n)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(Ya_f, "svb200_set_active_tension: Ya_f is null");
  SVB_REQUIRE(ctx->nNo > 0 || ctx->d_rowPtr, "svb200_set_active_tension: set the graph first");
  const size_t n = (size_t)ctx->nNo;
  std::vector<double> h(3 * n);
  bool sn = false;
  for (size_t a = 0; a < n; a++) {
    h[3 * a] = Ya_f[a];
    h[3 * a + 1] = Ya_s ? Ya_s[a] : 0.0;
    h[3 * a + 2] = Ya_n ? Ya_n[a] : 0.0;
    sn |= (h[3 * a + 1] > 0.0 || h[3 * a + 2] > 0.0);
  }
  ctx->ya_sn_positive = sn;
  return upload_nodal(ctx, 3, h.data(), &ctx->d_Ya);
}

int svb200_set_uris(svb200_ctx* ctx, int32_t nUris, const svb200_uris* valves, const double* sdf, const double* scaffold_udf,
                    const double* valve_vel)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(nUris >= 0 && nUris <= SVB200_MAX_URIS, "svb200_set_uris: between 0 and SVB200_MAX_URIS valves");
  if (nUris == 0) {
    cudaFree(ctx->d_uris); ctx->d_uris = nullptr; ctx->nUris = 0;
    ctx->uris_version++;
    return SVB200_OK;
  }
  SVB_REQUIRE(valves && sdf, "svb200_set_uris: null valve parameters or signed distance function");
  SVB_REQUIRE(ctx->d_rowPtr, "svb200_set_uris: set the graph first");
  const size_t n = (size_t)ctx->nNo;
  for (int v = 0; v < nUris; v++) {
    SVB_REQUIRE(!valves[v].scaffold || scaffold_udf, "svb200_set_uris: a valve has a scaffold but scaffold_udf is null");
    SVB_REQUIRE(!valves[v].include_velocity || valve_vel, "svb200_set_uris: a valve includes its velocity but valve_vel is null");
  }
  std::vector<double> h(5 * (size_t)nUris * n, 0.0);
  for (int v = 0; v < nUris; v++)
    for (size_t a = 0; a < n; a++) {
      double* r = h.data() + (a * nUris + v) * 5;
      r[0] = std::fabs(sdf[(size_t)v * n + a]);
      if (valves[v].scaffold) r[1] = std::fabs(scaffold_udf[(size_t)v * n + a]);
      if (valves[v].include_velocity)
        for (int i = 0; i < 3; i++) r[2 + i] = valve_vel[((size_t)v * n + a) * 3 + i];
    }
  if (ctx->d_uris && ctx->nUris != nUris) { cudaFree(ctx->d_uris); ctx->d_uris = nullptr; }
  TRY(upload_nodal(ctx, 5 * nUris, h.data(), &ctx->d_uris));
  ctx->nUris = nUris;
  ctx->uris_version++;
  for (int v = 0; v < nUris; v++) ctx->urisP[v] = valves[v];
  return SVB200_OK;
}

int svb200_set_ris(svb200_ctx* ctx, int32_t nProj, const int32_t* nMap, const int32_t* maps, const int32_t* closed)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(nProj >= 0 && (nProj == 0 || (nMap && maps && closed)), "svb200_set_ris: bad arguments");
  SVB_REQUIRE(nProj == 0 || !ctx->h_rowPtr.empty(), "svb200_set_ris: set the graph first");
  return ris_build_plan(ctx, nProj, nMap, maps, closed);
}

int svb200_set_old_disp(svb200_ctx* ctx, int32_t tDof, const double* Do)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ctx->d_rowPtr && Do, "svb200_set_old_disp: call svb200_set_graph first");
  SVB_REQUIRE(tDof == ctx->tDof, "svb200_set_old_disp: tDof differs from svb200_set_state");
  return upload_nodal(ctx, tDof, Do, &ctx->d_Do);
}

int svb200_set_bface(svb200_ctx* ctx, int32_t iFa, int32_t iM, int32_t eNoNb, int32_t nElb, const int32_t* IENb,
                     const int32_t* gE, int32_t nGb, const double* w, const double* N, const double* Nx)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(iFa >= 0 && iFa < 1024, "svb200_set_bface: bad face index");
  SVB_REQUIRE(iM >= 0 && iM < (int)ctx->mesh.size() && ctx->mesh[iM].set, "svb200_set_bface: parent mesh not set");
  SVB_REQUIRE((eNoNb == 3 || eNoNb == 4) && nGb >= 1 && nGb <= 4, "svb200_set_bface: TRI3 / QUD4 faces with at most 4 Gauss points");
  SVB_REQUIRE(nElb >= 0 && (nElb == 0 || (IENb && gE)) && w && N && Nx, "svb200_set_bface: bad arguments");
  if ((int)ctx->bface.size() <= iFa) ctx->bface.resize(iFa + 1);
  BFace& f = ctx->bface[iFa];
  cudaFree(f.d_IENb); cudaFree(f.d_gE);
  f = BFace();
  f.iM = iM; f.eNoNb = eNoNb; f.nElb = nElb; f.nGb = nGb;
  std::vector<int> ien((size_t)eNoNb * nElb);
  for (size_t k = 0; k < ien.size(); k++) {
    SVB_REQUIRE(IENb[k] >= 0 && IENb[k] < ctx->nNo, "svb200_set_bface: face node id out of range");
    ien[k] = ctx->h_map[IENb[k]];
  }
  for (int e = 0; e < nElb; e++) SVB_REQUIRE(gE[e] >= 0 && gE[e] < ctx->mesh[iM].nEl, "svb200_set_bface: parent element out of range");
  TRY(upload(ctx, &f.d_IENb, ien.data(), ien.size()));
  TRY(upload(ctx, &f.d_gE, gE, (size_t)nElb));
  f.w.assign(w, w + nGb);
  f.N.assign(N, N + (size_t)eNoNb * nGb);
  f.Nx.assign(Nx, Nx + (size_t)2 * eNoNb * nGb);
  f.set = true;
  return SVB200_OK;
}

int svb200_assemble_neu(svb200_ctx* ctx, int32_t iFa, const svb200_eqparams* eq, const svb200_dmnparams* dmn, int32_t nDmn,
                        const double* hg)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(eq && dmn && hg, "svb200_assemble_neu: null parameters");
  SVB_REQUIRE(iFa >= 0 && iFa < (int)ctx->bface.size() && ctx->bface[iFa].set, "svb200_assemble_neu: face not set");
  SVB_REQUIRE(ctx->d_R && ctx->d_Val, "svb200_assemble_neu: call svb200_alloc first");
  TRY(upload_nodal(ctx, 1, hg, &ctx->d_hg));
  TRY(flush_val_zero(ctx));
  return run_assemble_neu(ctx, ctx->bface[iFa], eq, dmn, nDmn, ctx->d_hg);
}

static double** sol_ptrs(svb200_ctx* ctx, int which, int k)
{
  double** tab[3][3] = {{&ctx->d_Ao, &ctx->d_Yo, &ctx->d_Do}, {&ctx->d_An, &ctx->d_Yn, &ctx->d_Dn}, {&ctx->d_Ag, &ctx->d_Yg, &ctx->d_Dg}};
  return tab[which][k];
}

static int ensure_solution_arrays(svb200_ctx* ctx)
{
  const size_t n = (size_t)ctx->tDof * ctx->nNo;
  for (int w = 0; w < 3; w++)
    for (int k = 0; k < 3; k++) {
      double** d = sol_ptrs(ctx, w, k);
      if (!*d) {
        SVB_CUDA(cudaMalloc(d, sizeof(double) * std::max<size_t>(n, 1)));
        SVB_CUDA(cudaMemsetAsync(*d, 0, sizeof(double) * n, ctx->stream));
      }
    }
  return SVB200_OK;
}

int svb200_set_solution(svb200_ctx* ctx, int32_t tDof, int32_t which, const double* A, const double* Y, const double* D)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ctx->d_rowPtr, "svb200_set_solution: call svb200_set_graph first");
  SVB_REQUIRE(which >= 0 && which <= 2, "svb200_set_solution: bad selector");
  if (ctx->tDof == 0) TRY(svb200_set_state(ctx, tDof, nullptr, nullptr, nullptr, nullptr));
  SVB_REQUIRE(tDof == ctx->tDof, "svb200_set_solution: tDof differs from svb200_set_state");
  TRY(ensure_solution_arrays(ctx));
  const double* h[3] = {A, Y, D};
  for (int k = 0; k < 3; k++)
    if (h[k]) TRY(upload_nodal(ctx, tDof, h[k], sol_ptrs(ctx, which, k)));
  return SVB200_OK;
}

int svb200_get_solution(svb200_ctx* ctx, int32_t which, double* A, double* Y, double* D)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(which >= 0 && which <= 2 && ctx->tDof > 0, "svb200_get_solution: bad selector or no state");
  TRY(ensure_solution_arrays(ctx));
  double* h[3] = {A, Y, D};
  for (int k = 0; k < 3; k++)
    if (h[k]) TRY(download_nodal(ctx, ctx->tDof, *sol_ptrs(ctx, which, k), h[k]));
  return SVB200_OK;
}

int svb200_predictor(svb200_ctx* ctx, int32_t nEq, const svb200_eqtime* eqs, double dt, int32_t dFlag)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ctx->tDof > 0, "svb200_predictor: no state (svb200_set_solution)");
  TRY(ensure_solution_arrays(ctx));
  for (int i = 0; i < nEq && eqs; i++) SVB_REQUIRE(eqs[i].e < ctx->tDof, "svb200_predictor: equation rows exceed tDof");
  return launch_predictor(ctx, nEq, eqs, dt, dFlag);
}

int svb200_initiator(svb200_ctx* ctx, int32_t nEq, const svb200_eqtime* eqs)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ctx->tDof > 0, "svb200_initiator: no state (svb200_set_solution)");
  TRY(ensure_solution_arrays(ctx));
  for (int i = 0; i < nEq && eqs; i++) SVB_REQUIRE(eqs[i].e < ctx->tDof, "svb200_initiator: equation rows exceed tDof");
  return launch_initiator(ctx, nEq, eqs);
}

int svb200_corrector(svb200_ctx* ctx, const svb200_eqtime* eq, double dt, int32_t mesh_s)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(eq && ctx->tDof > 0 && ctx->d_R, "svb200_corrector: no state or no linear system");
  SVB_REQUIRE(eq->s >= 0 && eq->e >= eq->s && eq->e < ctx->tDof && eq->e - eq->s + 1 == ctx->dof,
              "svb200_corrector: equation rows do not match the dof of the solved system");
  SVB_REQUIRE(mesh_s < 0 || (mesh_s + 3 <= ctx->tDof && ctx->d_nodeflag), "svb200_corrector: FSI copy needs svb200_set_node_flags");
  TRY(ensure_solution_arrays(ctx));
  if (eqtime_is_sst(*eq)) {
    SVB_REQUIRE(ctx->dof == 4 && ctx->d_Ad, "svb200_corrector: the ustruct update needs Ad (svb200_set_ad)");
    if (!ctx->d_Rd) {                       // ustruct_r was not called: Rd = 0 like Integrator.cpp:106-108
      SVB_CUDA(cudaMalloc(&ctx->d_Rd, sizeof(double) * 3 * std::max<size_t>((size_t)ctx->nNo, 1)));
      SVB_CUDA(cudaMemsetAsync(ctx->d_Rd, 0, sizeof(double) * 3 * (size_t)ctx->nNo, ctx->stream));
    }
  }
  return launch_corrector(ctx, eq, dt, mesh_s, ctx->d_nodeflag);
}

int svb200_set_node_flags(svb200_ctx* ctx, const int32_t* is_solid_node)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ctx->d_rowPtr && is_solid_node, "svb200_set_node_flags: call svb200_set_graph first");
  std::vector<int> f(ctx->nNo);
  for (int a = 0; a < ctx->nNo; a++) f[ctx->h_map[a]] = is_solid_node[a];
  return upload(ctx, &ctx->d_nodeflag, f.data(), f.size());
}

int svb200_set_dirichlet_rows(svb200_ctx* ctx, int32_t row0, int32_t nrow, int32_t n, const int32_t* nodes, const double* valA,
                              const double* valY, const double* valD)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ctx->tDof > 0 && row0 >= 0 && nrow >= 1 && row0 + nrow <= ctx->tDof && n >= 0 && (nodes || n == 0),
              "svb200_set_dirichlet_rows: bad arguments");
  TRY(ensure_solution_arrays(ctx));
  if (n == 0) return SVB200_OK;
  std::vector<int> g(n);
  for (int k = 0; k < n; k++) {
    SVB_REQUIRE(nodes[k] >= 0 && nodes[k] < ctx->nNo, "svb200_set_dirichlet_rows: node id out of range");
    g[k] = ctx->h_map[nodes[k]];
  }
  int* d_nodes = nullptr;
  TRY(upload(ctx, &d_nodes, g.data(), g.size()));
  const double* h[3] = {valA, valY, valD};
  int rc = SVB200_OK;
  for (int k = 0; k < 3 && rc == SVB200_OK; k++) {
    if (!h[k]) continue;
    double* d_val = nullptr;
    rc = upload(ctx, &d_val, h[k], (size_t)n * nrow);
    if (rc == SVB200_OK) rc = launch_set_rows(ctx, row0, nrow, n, d_nodes, d_val, *sol_ptrs(ctx, SVB200_SOL_CURRENT, k));
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_val);
  }
  cudaFree(d_nodes);
  return rc;
}

int svb200_dirichlet_ustruct(svb200_ctx* ctx, const svb200_eqtime* eq, double dt, int32_t n, const int32_t* nodes, int32_t dir_mask,
                             int32_t impD)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(eq && ctx->tDof > 0 && n >= 0 && (nodes || n == 0), "svb200_dirichlet_ustruct: bad arguments");
  SVB_REQUIRE(eq->s >= 0 && eq->s + 3 <= ctx->tDof, "svb200_dirichlet_ustruct: equation rows exceed tDof");
  SVB_REQUIRE(ctx->d_Ad, "svb200_dirichlet_ustruct: Ad was never set (svb200_set_ad)");
  TRY(ensure_solution_arrays(ctx));
  if (n == 0) return SVB200_OK;
  std::vector<int> g(n);
  for (int k = 0; k < n; k++) {
    SVB_REQUIRE(nodes[k] >= 0 && nodes[k] < ctx->nNo, "svb200_dirichlet_ustruct: node id out of range");
    g[k] = ctx->h_map[nodes[k]];
  }
  int* d_nodes = nullptr;
  TRY(upload(ctx, &d_nodes, g.data(), g.size()));
  int rc = launch_dirichlet_ustruct(ctx, eq, dt, n, d_nodes, dir_mask & 7, impD);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(d_nodes);
  return rc;
}

int svb200_advance_time_step(svb200_ctx* ctx)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ctx->tDof > 0, "svb200_advance_time_step: no state");
  TRY(ensure_solution_arrays(ctx));
  const size_t bytes = sizeof(double) * (size_t)ctx->tDof * ctx->nNo;
  for (int k = 0; k < 3; k++)
    SVB_CUDA(cudaMemcpyAsync(*sol_ptrs(ctx, SVB200_SOL_OLD, k), *sol_ptrs(ctx, SVB200_SOL_CURRENT, k), bytes,
                             cudaMemcpyDeviceToDevice, ctx->stream));
  return SVB200_OK;
}

int svb200_assemble(svb200_ctx* ctx, int32_t iM, const svb200_eqparams* eq, const svb200_dmnparams* dmn, int32_t nDmn)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(eq && dmn, "svb200_assemble: null parameters");
  SVB_REQUIRE(iM >= 0 && iM < (int)ctx->mesh.size() && ctx->mesh[iM].set, "svb200_assemble: mesh not set");
  SVB_REQUIRE(ctx->d_R && ctx->d_Val, "svb200_assemble: call svb200_alloc first");
  const Mesh& m = ctx->mesh[iM];
  SVB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  // a deferred zeroing of Val (svb200_alloc) is consumed by the TET4 fluid path; every other kernel needs Val zeroed first
  if (eq->phys != SVB200_PHYS_FLUID && eq->phys != SVB200_PHYS_FSI) TRY(flush_val_zero(ctx));
  // open fitted RIS surfaces (ris::doassem_ris, fluid.cpp:750-754, fsi.cpp:349-353): the rows of the mapped nodes are taken aside
  // around the element kernels of this mesh and what the mesh contributed to them is added to the twin rows afterwards (ris.cu)
  const bool ris = ris_active(ctx) && (eq->phys == SVB200_PHYS_FLUID || eq->phys == SVB200_PHYS_FSI);
  if (ris) {
    TRY(flush_val_zero(ctx));
    TRY(ris_begin(ctx));
  }
  switch (eq->phys) {
    case SVB200_PHYS_FLUID: {
      FluidArgs A;
      TRY(fill_fluid_args(ctx, m, eq, dmn, nDmn, A));
      TRY(run_assemble(ctx, m, A, (eq->reserved & SVB200_EQ_GENERAL_KERNEL) != 0));
    } break;
    case SVB200_PHYS_STRUCT:
      TRY(run_assemble_struct(ctx, m, eq, dmn, nDmn));
      break;
    case SVB200_PHYS_FSI: {
      // fsi::construct_fsi (fsi.cpp:24-362): per-element domain switch; fluid elements on the moved mesh
      // (ALE), solid elements through struct_3d writing the 3x3 part of the 4x4 blocks.
      bool anyFluid = false, anySolid = false, anyUstruct = false;
      for (int d = 0; d < nDmn; d++) {
        anyFluid |= (dmn[d].phys == SVB200_PHYS_FLUID);
        anySolid |= (dmn[d].phys == SVB200_PHYS_STRUCT);
        anyUstruct |= (dmn[d].phys == SVB200_PHYS_USTRUCT);
        // construct_fsi has branches for fluid, struct and ustruct domains (fsi.cpp:203-262) and throws for lElas (:236);
        // any other domain would be skipped silently there, here it is an error
        if (dmn[d].phys != SVB200_PHYS_FLUID && dmn[d].phys != SVB200_PHYS_STRUCT && dmn[d].phys != SVB200_PHYS_USTRUCT) {
          set_error("svb200_assemble: an FSI equation with a domain that is neither fluid, struct nor ustruct "
                    "([construct_fsi] LELAS3D not implemented)");
          return SVB200_ERR_UNSUPPORTED;
        }
        if (dmn[d].Id == -1) break;
      }
      if (anyFluid) {
        FluidArgs A;
        TRY(fill_fluid_args(ctx, m, eq, dmn, nDmn, A));
        TRY(run_assemble(ctx, m, A, (eq->reserved & SVB200_EQ_GENERAL_KERNEL) != 0));
      }
      TRY(flush_val_zero(ctx));
      if (anySolid) TRY(run_assemble_struct(ctx, m, eq, dmn, nDmn));
      // velocity-pressure solid inside FSI (fsi.cpp:243-262, 318-322, 343-346): ustruct_3d_m / ustruct_3d_c / ustruct_do_assem
      // on the reference configuration with the displacement rows eq.s..eq.s+2 of the tDof = 7 state; fills Kd as well
      if (anyUstruct) TRY(run_assemble_ustruct(ctx, m, eq, dmn, nDmn));
    } break;
    case SVB200_PHYS_MESH:
    case SVB200_PHYS_LELAS:
      TRY(run_assemble_mesh(ctx, m, eq, dmn, nDmn));
      break;
    case SVB200_PHYS_USTRUCT:
      TRY(run_assemble_ustruct(ctx, m, eq, dmn, nDmn));
      break;
    case SVB200_PHYS_HEATS:
    case SVB200_PHYS_HEATF:
      TRY(run_assemble_heat(ctx, m, eq, dmn, nDmn));
      break;
    default:
      set_error("svb200_assemble: this physics is not implemented in this build");
      return SVB200_ERR_UNSUPPORTED;
  }
  if (ris) TRY(ris_end(ctx));
  SVB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SVB_CUDA(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  SVB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  ctx->last_assemble_ms = ms;
  return SVB200_OK;
}

// The host-resident assembly stage in one pipelined call (see include/svb200.h).
// dst(:, caller[k]) = src(:, rows[k]): the completed interface rows written straight into the caller's page-locked residual
// (mapped host memory, a few 10^4 rows of dof doubles over PCIe).
__global__ void scatter_rows_to_host_kernel(int n, int dof, const int* __restrict__ rows, const int* __restrict__ caller,
                                            const double* __restrict__ src, double* __restrict__ dst)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * dof) return;
  const int k = t / dof, i = t - k * dof;
  dst[(size_t)caller[k] * dof + i] = src[(size_t)rows[k] * dof + i];
}

int svb200_assemble_host(svb200_ctx* ctx, int32_t iM, const svb200_eqparams* eq, const svb200_dmnparams* dmn, int32_t nDmn,
                         const double* Ag, const double* Yg, double* R_out)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(eq && dmn && Ag && Yg, "svb200_assemble_host: null parameters");
  SVB_REQUIRE(iM >= 0 && iM < (int)ctx->mesh.size() && ctx->mesh[iM].set, "svb200_assemble_host: mesh not set");
  const Mesh& m = ctx->mesh[iM];
  const int nGrp = (m.nEl + ASM_GROUP - 1) / ASM_GROUP;
  static const bool no_pipe = getenv("SVB200_HOST_NO_PIPELINE") != nullptr;       // A/B knob
  const bool fast = !no_pipe && eq->phys == SVB200_PHYS_FLUID && m.eNoN == 4 && eq->scatter == SVB200_SCATTER_ATOMIC &&
                    !(eq->reserved & SVB200_EQ_GENERAL_KERNEL) && m.schedK.d_uptr && (int)m.grp_node_need.size() == nGrp && nGrp >= 256 &&
                    m.jac_checked && ctx->tDof == eq->tDof && ctx->d_Ag && ctx->d_Yg && ctx->nUris == 0 && !ris_active(ctx) && getenv("SVB200_ASM_LEGACY") == nullptr;
  if (!fast) {
    // any other case: the plain sequence (also the first call on a mesh, which runs the Jacobian check)
    TRY(svb200_set_state(ctx, eq->tDof, Ag, Yg, nullptr, nullptr));
    TRY(svb200_alloc(ctx, eq->dof));
    TRY(svb200_assemble(ctx, iM, eq, dmn, nDmn));
    TRY(svb200_commu_R(ctx));
    if (R_out) TRY(svb200_download(ctx, SVB200_ARRAY_R, R_out));
    return SVB200_OK;
  }
  const int tDof = eq->tDof, dof = eq->dof;
  const size_t nV = (size_t)dof * dof * ctx->nnz, nR = (size_t)dof * ctx->nNo;
  SVB_REQUIRE(ctx->d_R && ctx->d_Val && ctx->dof == dof && nR <= ctx->R_cap && nV <= ctx->Val_cap,
              "svb200_assemble_host: call svb200_alloc(dof) once before the first pipelined assembly");
  auto wall_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = wall_ms();
  bool patched_on_device = false;
  SVB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  ctx->val_zero_pending = false;
  SVB_CUDA(cudaMemsetAsync(ctx->d_R, 0, sizeof(double) * nR, ctx->stream));
  SVB_CUDA(cudaMemsetAsync(ctx->d_Val, 0, sizeof(double) * nV, ctx->stream));
  if (ctx->d_Kd && ctx->nnz) SVB_CUDA(cudaMemsetAsync(ctx->d_Kd, 0, sizeof(double) * 12 * (size_t)ctx->nnz, ctx->stream));
  // the prestress accumulators pSn / pSa are zeroed once per Newton iteration by Integrator::initiator (Integrator.cpp:745-748)
  if (ctx->d_pSn && ctx->nNo) SVB_CUDA(cudaMemsetAsync(ctx->d_pSn, 0, sizeof(double) * 7 * (size_t)ctx->nNo, ctx->stream));
  FluidArgs A;
  TRY(fill_fluid_args(ctx, m, eq, dmn, nDmn, A));
  // the copy streams start after whatever the main stream was doing with the state arrays
  SVB_CUDA(cudaEventRecord(ctx->zev[0], ctx->stream));
  SVB_CUDA(cudaStreamWaitEvent(ctx->zstream, ctx->zev[0], 0));
  SVB_CUDA(cudaStreamWaitEvent(ctx->dstream, ctx->zev[0], 0));
  constexpr int K = 8;
  const bool direct = !ctx->has_map;
  // Residual rows come back while later chunks run.  On a partitioned mesh the interface rows are streamed with their partial
  // sums and fetched again after the shared-node sum (a few 10^4 rows).
  const bool stream_down = R_out != nullptr;
  const bool multi = ctx->nranks > 1 && !ctx->neigh.empty();
  if (stream_down && multi && !ctx->shared_built) {
    std::vector<int> inv(ctx->nNo);
    for (int a = 0; a < ctx->nNo; a++) inv[ctx->h_map[a]] = a;
    std::vector<int> rows;
    for (auto& nb : ctx->neigh) rows.insert(rows.end(), nb.h_ptr.begin(), nb.h_ptr.end());
    std::sort(rows.begin(), rows.end());
    rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
    ctx->h_shared_caller.resize(rows.size());
    std::vector<long long> off(rows.size() + 1);
    for (size_t k = 0; k < rows.size(); k++) { ctx->h_shared_caller[k] = inv[rows[k]]; off[k] = (long long)k; }
    off[rows.size()] = (long long)rows.size();
    TRY(upload(ctx, &ctx->d_shared_rows, rows.data(), rows.size()));
    TRY(upload(ctx, &ctx->d_shared_off, off.data(), off.size()));
    TRY(upload(ctx, &ctx->d_shared_caller, ctx->h_shared_caller.data(), ctx->h_shared_caller.size()));
    cudaFreeHost(ctx->h_shared_buf); ctx->h_shared_buf = nullptr;
    SVB_CUDA(cudaMallocHost(&ctx->h_shared_buf, sizeof(double) * 4 * std::max<size_t>(rows.size(), 1)));
    ctx->shared_built = true;
  }
  if (!direct) TRY(ensure_stage(ctx, sizeof(double) * (2 * (size_t)tDof + dof) * ctx->nNo));
  double* stageA = ctx->d_stage;
  double* stageY = ctx->d_stage + (size_t)tDof * ctx->nNo;
  double* stageR = ctx->d_stage + 2 * (size_t)tDof * ctx->nNo;
  int up = 0, down = 0;
  for (int c = 0; c < K; c++) {
    const int g0 = (int)((long long)nGrp * c / K), g1 = (int)((long long)nGrp * (c + 1) / K);
    if (g1 <= g0) continue;
    const int need = (c == K - 1) ? ctx->nNo : m.grp_node_need[g1 - 1];
    if (need > up) {
      const size_t off = (size_t)tDof * up, cnt = (size_t)tDof * (need - up);
      if (direct) {
        SVB_CUDA(cudaMemcpyAsync(ctx->d_Ag + off, Ag + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, ctx->zstream));
        SVB_CUDA(cudaMemcpyAsync(ctx->d_Yg + off, Yg + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, ctx->zstream));
      } else {
        SVB_CUDA(cudaMemcpyAsync(stageA + off, Ag + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, ctx->zstream));
        SVB_CUDA(cudaMemcpyAsync(stageY + off, Yg + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, ctx->zstream));
        TRY(launch_permute_cols(ctx, tDof, need - up, ctx->d_map + up, stageA + off, ctx->d_Ag, false, ctx->zstream));
        TRY(launch_permute_cols(ctx, tDof, need - up, ctx->d_map + up, stageY + off, ctx->d_Yg, false, ctx->zstream));
      }
      up = need;
    }
    SVB_CUDA(cudaEventRecord(ctx->pev[0][c], ctx->zstream));
    SVB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->pev[0][c], 0));
    A.g0 = g0; A.nGrpLaunch = g1 - g0; A.gperm = nullptr;
    TRY(launch_assemble_fluid(ctx, m, A));
    if (stream_down) {
      // residual rows no later group touches are final (up to the shared-node sum): bring them back while the next chunk runs
      const int fin = (c == K - 1) ? ctx->nNo : m.grp_node_done[g1];
      if (fin > down) {
        SVB_CUDA(cudaEventRecord(ctx->pev[1][c], ctx->stream));
        SVB_CUDA(cudaStreamWaitEvent(ctx->dstream, ctx->pev[1][c], 0));
        const size_t off = (size_t)dof * down, cnt = (size_t)dof * (fin - down);
        if (direct) {
          SVB_CUDA(cudaMemcpyAsync(R_out + off, ctx->d_R + off, sizeof(double) * cnt, cudaMemcpyDeviceToHost, ctx->dstream));
        } else {
          TRY(launch_permute_cols(ctx, dof, fin - down, ctx->d_map + down, ctx->d_R, stageR + off, true, ctx->dstream));
          SVB_CUDA(cudaMemcpyAsync(R_out + off, stageR + off, sizeof(double) * cnt, cudaMemcpyDeviceToHost, ctx->dstream));
        }
        down = fin;
      }
    }
  }
  SVB_CUDA(cudaEventRecord(ctx->hev[0], ctx->zstream));     // all uploads done
  SVB_CUDA(cudaEventRecord(ctx->hev[1], ctx->stream));      // all element groups done
  SVB_CUDA(cudaEventRecord(ctx->hev[3], ctx->dstream));     // streamed residual rows on the host
  if (multi) {
    TRY(halo_sum(ctx, dof, ctx->d_R));
    SVB_CUDA(cudaEventRecord(ctx->hev[2], ctx->stream));    // shared-node sum done
    if (stream_down && !ctx->h_shared_caller.empty()) {
      // the interface rows again, now complete
      const int ns = (int)ctx->h_shared_caller.size();
      static const bool no_zc = getenv("SVB200_HOST_NO_ZEROCOPY") != nullptr;      // A/B knob
      double* dR = nullptr;
      if (!no_zc && cudaHostGetDevicePointer(reinterpret_cast<void**>(&dR), R_out, 0) != cudaSuccess) { dR = nullptr; cudaGetLastError(); }
      if (dR) {
        // page-locked caller buffer: a kernel writes the rows in place (after the streamed copies of the same rows, which
        // carried partial sums: the main stream waits for the download stream)
        SVB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->hev[3], 0));
        scatter_rows_to_host_kernel<<<(ns * dof + 255) / 256, 256, 0, ctx->stream>>>(ns, dof, ctx->d_shared_rows, ctx->d_shared_caller,
                                                                                 ctx->d_R, dR);
        ctx->launches++;
        patched_on_device = true;
      } else {
        // pageable caller buffer: gather -> pinned scratch -> patched into R_out by the host
        SVB_CUDA(cudaStreamSynchronize(ctx->dstream));          // stageR is free, the streamed rows are in R_out
        TRY(launch_gather_row_blocks(ctx, ns, dof, ctx->d_shared_rows, false, ctx->d_shared_off, ctx->d_R, stageR));
        SVB_CUDA(cudaMemcpyAsync(ctx->h_shared_buf, stageR, sizeof(double) * dof * ns, cudaMemcpyDeviceToHost, ctx->stream));
      }
    }
  }
  const double t_enq = wall_ms();
  SVB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->dstream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  const double t_sync = wall_ms();
  if (multi && stream_down && !patched_on_device) {
    const int ns = (int)ctx->h_shared_caller.size();
    for (int k = 0; k < ns; k++)
      for (int i = 0; i < dof; i++) R_out[(size_t)dof * ctx->h_shared_caller[k] + i] = ctx->h_shared_buf[(size_t)dof * k + i];
  }
  SVB_CUDA(cudaStreamSynchronize(ctx->zstream));
  const double t_end = wall_ms();
  ctx->host_stage_ms[5] = t_enq - t_begin; ctx->host_stage_ms[6] = t_sync - t_begin; ctx->host_stage_ms[7] = t_end - t_begin;
  float ms = 0.f;
  SVB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  ctx->last_assemble_ms = ms;
  ctx->host_stage_ms[4] = ms;
  for (int k = 0; k < 4; k++) {
    float t = 0.f;
    if ((k == 2 && !multi) || cudaEventElapsedTime(&t, ctx->ev0, ctx->hev[k]) != cudaSuccess) { t = 0.f; cudaGetLastError(); }
    ctx->host_stage_ms[k] = t;
  }
  return SVB200_OK;
}

int svb200_last_host_stage(svb200_ctx* ctx, double* ms5)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ms5, "svb200_last_host_stage: null output (8 doubles)");
  for (int k = 0; k < 8; k++) ms5[k] = ctx->host_stage_ms[k];
  return SVB200_OK;
}

// dst(:, rows[k]) += add(:, k); the targets are distinct (duplicates were summed on the host), so plain adds suffice.
__global__ void add_rows_kernel(int n, int dof, const int* __restrict__ rows, const double* __restrict__ add,
                                double* __restrict__ dst)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * dof) return;
  dst[(size_t)rows[t / dof] * dof + t % dof] += add[t];
}

// Sum the d-double entries of `vals` that share a target, in the order they were staged (the order of the host's
// element loop, i.e. the reference's own do_assem order), and add the result to the device array: one deterministic add
// per distinct target instead of an atomic per staged entry.
static int add_reduced(svb200_ctx* ctx, const std::vector<int>& target, const double* vals, int d, double* d_dst)
{
  const size_t n = target.size();
  if (n == 0) return SVB200_OK;
  std::vector<size_t> order(n);
  std::iota(order.begin(), order.end(), (size_t)0);
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return target[a] < target[b]; });
  std::vector<int> uniq;
  std::vector<double> sum;
  for (size_t q = 0; q < n; q++) {
    const size_t k = order[q];
    if (q == 0 || target[k] != uniq.back()) {
      uniq.push_back(target[k]);
      sum.insert(sum.end(), vals + k * d, vals + (k + 1) * d);
    } else {
      double* s = sum.data() + sum.size() - d;
      for (int i = 0; i < d; i++) s[i] += vals[k * d + i];
    }
  }
  int* d_t = nullptr; double* d_a = nullptr;
  int rc = upload(ctx, &d_t, uniq.data(), uniq.size());
  if (!rc) rc = upload(ctx, &d_a, sum.data(), sum.size());
  if (!rc) {
    const int nu = (int)uniq.size();
    add_rows_kernel<<<(nu * d + 255) / 256, 256, 0, ctx->stream>>>(nu, d, d_t, d_a, d_dst);
    ctx->launches++;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "svb200_add_host_contrib", __FILE__, __LINE__);
  }
  cudaFree(d_t); cudaFree(d_a);
  return rc;
}

int svb200_add_host_contrib(svb200_ctx* ctx, int32_t dof, int32_t nR, const int32_t* rows, const double* R_add,
                            int32_t nK, const int32_t* krows, const int32_t* kcols, const double* K_add)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(dof == ctx->dof && ctx->d_R && ctx->d_Val, "svb200_add_host_contrib: call svb200_alloc(dof) first");
  TRY(flush_val_zero(ctx));
  if (nR > 0) {
    SVB_REQUIRE(rows && R_add, "svb200_add_host_contrib: null residual arrays");
    std::vector<int> r(nR);
    for (int k = 0; k < nR; k++) {
      SVB_REQUIRE(rows[k] >= 0 && rows[k] < ctx->nNo, "svb200_add_host_contrib: row out of range");
      r[k] = ctx->h_map[rows[k]];
    }
    TRY(add_reduced(ctx, r, R_add, dof, ctx->d_R));
  }
  if (nK > 0) {
    SVB_REQUIRE(krows && kcols && K_add, "svb200_add_host_contrib: null tangent arrays");
    const std::vector<int>& col = ctx->h_colPtr;
    std::vector<int> s(nK);
    for (int k = 0; k < nK; k++) {
      SVB_REQUIRE(krows[k] >= 0 && krows[k] < ctx->nNo && kcols[k] >= 0 && kcols[k] < ctx->nNo,
                  "svb200_add_host_contrib: index out of range");
      const int r = ctx->h_map[krows[k]], c = ctx->h_map[kcols[k]];
      int slot = -1;
      for (int q = ctx->h_rowPtr[r]; q < ctx->h_rowPtr[r + 1]; q++)
        if (col[q] == c) { slot = q; break; }
      SVB_REQUIRE(slot >= 0, "svb200_add_host_contrib: (row,col) pair is not in the CSR graph");
      s[k] = slot;
    }
    TRY(add_reduced(ctx, s, K_add, dof * dof, ctx->d_Val));
  }
  return SVB200_OK;
}

int svb200_commu_R(svb200_ctx* ctx)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ctx->d_R, "svb200_commu_R: call svb200_alloc first");
  TRY(halo_sum(ctx, ctx->dof, ctx->d_R));
  return SVB200_OK;
}

int svb200_ustruct_r(svb200_ctx* ctx, const svb200_eqparams* eq, int32_t itr, const double* Ad)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(eq, "svb200_ustruct_r: null parameters");
  SVB_REQUIRE(Ad || ctx->d_Ad, "svb200_ustruct_r: no Ad (pass it, or svb200_set_ad once for the device-resident loop)");
  SVB_REQUIRE(eq->phys == SVB200_PHYS_USTRUCT || eq->phys == SVB200_PHYS_FSI, "svb200_ustruct_r: the equation is neither ustruct nor FSI");
  // FSI (sstEq): only nodes of a ustruct domain take part (all_fun::is_domain, ustruct.cpp:1776-1779, 1791-1793)
  SVB_REQUIRE(eq->phys != SVB200_PHYS_FSI || ctx->d_nodeflag, "svb200_ustruct_r: an FSI equation needs the solid-node flags (svb200_set_node_flags)");
  SVB_REQUIRE(ctx->dof == 4 && ctx->d_R && ctx->d_Kd && ctx->d_Yg, "svb200_ustruct_r: assemble the ustruct equation first");
  SVB_REQUIRE(eq->tDof == ctx->tDof && eq->s >= 0 && eq->s + 4 <= eq->tDof, "svb200_ustruct_r: tDof / eq.s mismatch");
  if (Ad) TRY(upload_nodal(ctx, 3, Ad, &ctx->d_Ad));
  TRY(run_ustruct_r(ctx, eq, itr, ctx->d_Ad));
  return SVB200_OK;
}

int svb200_set_ad(svb200_ctx* ctx, const double* Ad)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ctx->d_rowPtr && Ad, "svb200_set_ad: call svb200_set_graph first");
  return upload_nodal(ctx, 3, Ad, &ctx->d_Ad);
}

int svb200_get_ad(svb200_ctx* ctx, double* Ad)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ctx->d_Ad && Ad, "svb200_get_ad: Ad was never set");
  return download_nodal(ctx, 3, ctx->d_Ad, Ad);
}

int svb200_solve(svb200_ctx* ctx, int32_t dof, int32_t ls_type, int32_t prec, const svb200_lsparams* ls,
                 int32_t nFaces, const int32_t* incL, const double* res, double* R_out, svb200_lsresult* result)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ls, "svb200_solve: null solver parameters");
  SVB_REQUIRE(dof == ctx->dof && ctx->d_R && ctx->d_Val, "svb200_solve: call svb200_alloc(dof) and assemble first");
  SVB_REQUIRE(prec == SVB200_PREC_FSILS || prec == SVB200_PREC_RCS, "svb200_solve: preconditioner must be SVB200_PREC_FSILS or SVB200_PREC_RCS");
  SVB_REQUIRE(nFaces <= (int)ctx->face.size(), "svb200_solve: nFaces exceeds svb200_set_num_faces");
  TRY(flush_val_zero(ctx));
  SVB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  TRY(fsils_solve_device(ctx, dof, ls_type, prec, ls, nFaces, incL, res, result));
  SVB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SVB_CUDA(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  SVB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  ctx->last_solve_ms = ms;
  TRY(p2p_check(ctx));
  if (R_out) TRY(download_nodal(ctx, dof, ctx->d_R, R_out));
  return SVB200_OK;
}

// Val is returned in the caller's CSR slot order: caller row a occupies internal row map[a].
// Per-CSR-entry blocks of d2 doubles (Val: dof*dof, Kd: 12) between the caller's and the internal row order.
static int copy_blocks(svb200_ctx* ctx, double* d_blocks, size_t d2, double* host, bool to_host)
{
  if (!ctx->has_map) {
    if (to_host)
      SVB_CUDA(cudaMemcpyAsync(host, d_blocks, sizeof(double) * d2 * ctx->nnz, cudaMemcpyDeviceToHost, ctx->stream));
    else
      SVB_CUDA(cudaMemcpyAsync(d_blocks, host, sizeof(double) * d2 * ctx->nnz, cudaMemcpyHostToDevice, ctx->stream));
    SVB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SVB200_OK;
  }
  // permuted rows: a device kernel moves whole chunks of caller rows through the staging buffer (one memcpy per chunk)
  const size_t chunk_doubles = (size_t)32 << 20;      // 256 MB of staging
  int a0 = 0;
  while (a0 < ctx->nNo) {
    int a1 = a0;
    size_t cnt = 0;
    while (a1 < ctx->nNo) {
      const size_t len = (size_t)(ctx->h_rowPtr_in[a1 + 1] - ctx->h_rowPtr_in[a1]) * d2;
      if (cnt + len > chunk_doubles && a1 > a0) break;
      cnt += len;
      a1++;
    }
    TRY(ensure_stage(ctx, sizeof(double) * std::max<size_t>(cnt, 1)));
    double* h = host + (size_t)ctx->h_rowPtr_in[a0] * d2;
    if (to_host) {
      TRY(launch_permute_row_blocks(ctx, a0, a1, (int)d2, ctx->d_rowPtr_in, d_blocks, ctx->d_stage, true));
      SVB_CUDA(cudaMemcpyAsync(h, ctx->d_stage, sizeof(double) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
      SVB_CUDA(cudaMemcpyAsync(ctx->d_stage, h, sizeof(double) * cnt, cudaMemcpyHostToDevice, ctx->stream));
      TRY(launch_permute_row_blocks(ctx, a0, a1, (int)d2, ctx->d_rowPtr_in, d_blocks, ctx->d_stage, false));
    }
    SVB_CUDA(cudaStreamSynchronize(ctx->stream));     // the staging buffer is reused by the next chunk
    a0 = a1;
  }
  return SVB200_OK;
}

static int copy_val(svb200_ctx* ctx, int dof, double* host, bool to_host)
{
  return copy_blocks(ctx, ctx->d_Val, (size_t)dof * dof, host, to_host);
}

int svb200_download(svb200_ctx* ctx, int32_t what, double* dst)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(dst, "svb200_download: null destination");
  switch (what) {
    case SVB200_ARRAY_R:
      SVB_REQUIRE(ctx->d_R, "svb200_download: R not allocated");
      return download_nodal(ctx, ctx->dof, ctx->d_R, dst);
    case SVB200_ARRAY_VAL:
      SVB_REQUIRE(ctx->d_Val, "svb200_download: Val not allocated");
      TRY(flush_val_zero(ctx));
      return copy_val(ctx, ctx->dof, dst, true);
    case SVB200_ARRAY_W:
      SVB_REQUIRE(ctx->d_W, "svb200_download: W not computed yet");
      return download_nodal(ctx, ctx->dof, ctx->d_W, dst);
    case SVB200_ARRAY_KD:
      SVB_REQUIRE(ctx->d_Kd, "svb200_download: Kd exists only after a ustruct assembly");
      return copy_blocks(ctx, ctx->d_Kd, 12, dst, true);
    case SVB200_ARRAY_RD:
      SVB_REQUIRE(ctx->d_Rd, "svb200_download: Rd exists only after svb200_ustruct_r");
      return download_nodal(ctx, 3, ctx->d_Rd, dst);
  }
  set_error("svb200_download: unknown array id");
  return SVB200_ERR_INVALID;
}

int svb200_download_rows(svb200_ctx* ctx, int32_t what, int32_t n, const int32_t* nodes, double* dst)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(n >= 0 && (n == 0 || (nodes && dst)), "svb200_download_rows: bad arguments");
  if (n == 0) return SVB200_OK;
  const double* src = nullptr;
  int d2 = 0;
  bool blocks = false;
  switch (what) {
    case SVB200_ARRAY_R: src = ctx->d_R; d2 = ctx->dof; break;
    case SVB200_ARRAY_W: src = ctx->d_W; d2 = ctx->dof; break;
    case SVB200_ARRAY_VAL: TRY(flush_val_zero(ctx)); src = ctx->d_Val; d2 = ctx->dof * ctx->dof; blocks = true; break;
    case SVB200_ARRAY_KD: src = ctx->d_Kd; d2 = 12; blocks = true; break;
    default: set_error("svb200_download_rows: unknown array id"); return SVB200_ERR_INVALID;
  }
  SVB_REQUIRE(src && d2 > 0, "svb200_download_rows: the array does not exist yet");
  std::vector<int> rows(n);
  std::vector<long long> off(n + 1, 0);
  for (int k = 0; k < n; k++) {
    SVB_REQUIRE(nodes[k] >= 0 && nodes[k] < ctx->nNo, "svb200_download_rows: node id out of range");
    rows[k] = ctx->h_map[nodes[k]];
    off[k + 1] = off[k] + (blocks ? ctx->h_rowPtr[rows[k] + 1] - ctx->h_rowPtr[rows[k]] : 1);
  }
  int* d_rows = nullptr; long long* d_off = nullptr;
  const size_t total = (size_t)off[n] * d2;
  TRY(ensure_stage(ctx, sizeof(double) * std::max<size_t>(total, 1)));
  int rc = upload(ctx, &d_rows, rows.data(), rows.size());
  if (!rc) rc = upload(ctx, &d_off, off.data(), off.size());
  if (!rc) rc = launch_gather_row_blocks(ctx, n, d2, d_rows, blocks, d_off, src, ctx->d_stage);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(dst, ctx->d_stage, sizeof(double) * total, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "svb200_download_rows copy", __FILE__, __LINE__);
  }
  cudaFree(d_rows); cudaFree(d_off);
  return rc;
}

int svb200_upload(svb200_ctx* ctx, int32_t what, int32_t dof, const double* src)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(src, "svb200_upload: null source");
  SVB_REQUIRE(dof == ctx->dof && ctx->d_R && ctx->d_Val, "svb200_upload: call svb200_alloc(dof) first");
  switch (what) {
    case SVB200_ARRAY_R: return upload_nodal(ctx, dof, src, &ctx->d_R);
    case SVB200_ARRAY_VAL: ctx->val_zero_pending = false; return copy_val(ctx, dof, const_cast<double*>(src), false);
    case SVB200_ARRAY_RD: return upload_nodal(ctx, 3, src, &ctx->d_Rd);
  }
  set_error("svb200_upload: unknown array id");
  return SVB200_ERR_INVALID;
}

int svb200_spmv(svb200_ctx* ctx, int32_t dof, const double* U, double* KU)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(U && KU, "svb200_spmv: null vectors");
  SVB_REQUIRE(dof == ctx->dof && ctx->d_Val, "svb200_spmv: call svb200_alloc(dof) first");
  TRY(flush_val_zero(ctx));
  const size_t n = (size_t)dof * ctx->nNo;
  double* d_u = nullptr; double* d_ku = nullptr;
  SVB_CUDA(cudaMalloc(&d_u, sizeof(double) * std::max<size_t>(n, 1)));
  SVB_CUDA(cudaMalloc(&d_ku, sizeof(double) * std::max<size_t>(n, 1)));
  int rc = upload_nodal(ctx, dof, U, &d_u);
  if (!rc) rc = launch_spmv(ctx, dof, ctx->d_Val, d_u, d_ku);
  if (!rc) rc = halo_sum(ctx, dof, d_ku);
  if (!rc) rc = download_nodal(ctx, dof, d_ku, KU);
  cudaFree(d_u); cudaFree(d_ku);
  return rc;
}

int svb200_last_timing(svb200_ctx* ctx, double* assemble_ms, double* solve_ms)
{
  if (!ctx) { set_error("svb200: null context"); return SVB200_ERR_INVALID; }
  if (assemble_ms) *assemble_ms = ctx->last_assemble_ms;
  if (solve_ms) *solve_ms = ctx->last_solve_ms;
  return SVB200_OK;
}

int svb200_bench_assemble(svb200_ctx* ctx, int32_t iM, const svb200_eqparams* eq, const svb200_dmnparams* dmn,
                          int32_t nDmn, int32_t reps, double* ms_per_launch)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(eq && dmn && ms_per_launch && reps >= 1, "svb200_bench_assemble: bad arguments");
  SVB_REQUIRE(iM >= 0 && iM < (int)ctx->mesh.size() && ctx->mesh[iM].set, "svb200_bench_assemble: mesh not set");
  SVB_REQUIRE(ctx->d_R && ctx->d_Val, "svb200_bench_assemble: call svb200_alloc first");
  const Mesh& m = ctx->mesh[iM];
  FluidArgs A;
  TRY(fill_fluid_args(ctx, m, eq, dmn, nDmn, A));
  SVB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  for (int r = 0; r < reps; r++) {
    FluidArgs B = A;
    TRY(run_assemble(ctx, m, B));
  }
  SVB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SVB_CUDA(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  SVB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *ms_per_launch = ms / reps;
  return SVB200_OK;
}

int svb200_bench_spmv(svb200_ctx* ctx, int32_t dof, int32_t reps, double* ms_per_launch)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ms_per_launch && reps >= 1, "svb200_bench_spmv: bad arguments");
  SVB_REQUIRE(dof == ctx->dof && ctx->d_Val, "svb200_bench_spmv: call svb200_alloc(dof) first");
  TRY(flush_val_zero(ctx));
  const size_t n = (size_t)dof * ctx->nNo;
  double* d_u = nullptr; double* d_ku = nullptr;
  SVB_CUDA(cudaMalloc(&d_u, sizeof(double) * std::max<size_t>(n, 1)));
  SVB_CUDA(cudaMalloc(&d_ku, sizeof(double) * std::max<size_t>(n, 1)));
  SVB_CUDA(cudaMemsetAsync(d_u, 0, sizeof(double) * n, ctx->stream));
  int rc = launch_spmv(ctx, dof, ctx->d_Val, d_u, d_ku);   // warm-up
  SVB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  for (int r = 0; r < reps && !rc; r++) rc = launch_spmv(ctx, dof, ctx->d_Val, d_u, d_ku);
  SVB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SVB_CUDA(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  SVB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *ms_per_launch = ms / reps;
  cudaFree(d_u); cudaFree(d_ku);
  return rc;
}

// ---- rectangular-block SpMV / Schur operator with caller-supplied matrices (tests, A/B measurements) ---------------------------
namespace {
struct DevBuf {
  double* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t n, bool zero, cudaStream_t st)
  {
    SVB_CUDA(cudaMalloc(&p, sizeof(double) * std::max<size_t>(n, 2) + 16));
    if (zero) SVB_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * std::max<size_t>(n, 2), st));
    return SVB200_OK;
  }
  int put(const double* h, size_t n, cudaStream_t st)
  {
    if (n) SVB_CUDA(cudaMemcpyAsync(p, h, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    return SVB200_OK;
  }
};
}  // namespace

int svb200_spmv_rc_variants(int32_t R, int32_t C) { return spmv_rc_num_variants(R, C); }
int svb200_schur_sp_variants(void) { return schur_sp4_num_variants(); }

int svb200_spmv_rc(svb200_ctx* ctx, int32_t R, int32_t C, int32_t variant, const double* K, const double* U, double* KU)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(K && U && KU, "svb200_spmv_rc: null arrays");
  SVB_REQUIRE(ctx->d_rowPtr && !ctx->has_map && ctx->nranks == 1, "svb200_spmv_rc: needs a graph on a single-partition context");
  const size_t nK = (size_t)R * C * ctx->nnz, nU = (size_t)C * ctx->nNo, nKU = (size_t)R * ctx->nNo;
  DevBuf dK, dU, dKU;
  TRY(dK.alloc(nK, false, ctx->stream)); TRY(dU.alloc(nU, false, ctx->stream)); TRY(dKU.alloc(nKU, true, ctx->stream));
  TRY(dK.put(K, nK, ctx->stream)); TRY(dU.put(U, nU, ctx->stream));
  TRY(variant < 0 ? spmv_rc(ctx, R, C, dK.p, dU.p, dKU.p) : spmv_rc_variant(ctx, R, C, variant, dK.p, dU.p, dKU.p));
  if (nKU) SVB_CUDA(cudaMemcpyAsync(KU, dKU.p, sizeof(double) * nKU, cudaMemcpyDeviceToHost, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SVB200_OK;
}

int svb200_bench_spmv_rc(svb200_ctx* ctx, int32_t R, int32_t C, int32_t variant, int32_t reps, double* ms_per_launch)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ms_per_launch && reps >= 1 && ctx->d_rowPtr, "svb200_bench_spmv_rc: bad arguments");
  const size_t nK = (size_t)R * C * ctx->nnz, nU = (size_t)C * ctx->nNo, nKU = (size_t)R * ctx->nNo;
  DevBuf dK, dU, dKU;
  TRY(dK.alloc(nK, true, ctx->stream)); TRY(dU.alloc(nU, true, ctx->stream)); TRY(dKU.alloc(nKU, true, ctx->stream));
  auto run = [&]() { return variant < 0 ? spmv_rc(ctx, R, C, dK.p, dU.p, dKU.p) : spmv_rc_variant(ctx, R, C, variant, dK.p, dU.p, dKU.p); };
  TRY(run());   // warm-up
  SVB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  for (int r = 0; r < reps; r++) TRY(run());
  SVB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SVB_CUDA(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  SVB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *ms_per_launch = ms / reps;
  return SVB200_OK;
}

static int schur_run(svb200_ctx* ctx, int variant, const double* dL, const double* dGt, const double* dDL, const double* dP,
                     const double* dGP, double* dSP, double* dPart, int* nparts)
{
  if (variant == -2) {
    if (nparts) *nparts = 0;
    return schur_sp(ctx, 3, dL, dGt, dP, dGP, dSP);
  }
  return schur_sp4(ctx, variant, dDL, dP, dGP, dSP, dPart, nparts);
}

int svb200_schur_sp(svb200_ctx* ctx, int32_t variant, const double* L, const double* Gt, const double* P, const double* GP,
                    double* SP, double* p_dot_sp)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(L && Gt && P && GP && SP, "svb200_schur_sp: null arrays");
  SVB_REQUIRE(ctx->d_rowPtr && !ctx->has_map && ctx->nranks == 1, "svb200_schur_sp: needs a graph on a single-partition context");
  const size_t nnz = ctx->nnz, nNo = ctx->nNo;
  std::vector<double> hDL(4 * nnz);
  for (size_t k = 0; k < nnz; k++) {
    hDL[4 * k] = Gt[3 * k]; hDL[4 * k + 1] = Gt[3 * k + 1]; hDL[4 * k + 2] = Gt[3 * k + 2]; hDL[4 * k + 3] = L[k];
  }
  DevBuf dL, dGt, dDL, dP, dGP, dSP, dPart;
  TRY(dL.alloc(nnz, false, ctx->stream)); TRY(dGt.alloc(3 * nnz, false, ctx->stream)); TRY(dDL.alloc(4 * nnz, false, ctx->stream));
  TRY(dP.alloc(nNo, false, ctx->stream)); TRY(dGP.alloc(3 * nNo, false, ctx->stream)); TRY(dSP.alloc(nNo, true, ctx->stream));
  TRY(dPart.alloc(148 * 8 + 16, true, ctx->stream));
  TRY(dL.put(L, nnz, ctx->stream)); TRY(dGt.put(Gt, 3 * nnz, ctx->stream)); TRY(dDL.put(hDL.data(), 4 * nnz, ctx->stream));
  TRY(dP.put(P, nNo, ctx->stream)); TRY(dGP.put(GP, 3 * nNo, ctx->stream));
  int nparts = 0;
  TRY(schur_run(ctx, variant, dL.p, dGt.p, dDL.p, dP.p, dGP.p, dSP.p, dPart.p, &nparts));
  if (nNo) SVB_CUDA(cudaMemcpyAsync(SP, dSP.p, sizeof(double) * nNo, cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<double> hp(std::max(nparts, 1), 0.0);
  if (nparts) SVB_CUDA(cudaMemcpyAsync(hp.data(), dPart.p, sizeof(double) * nparts, cudaMemcpyDeviceToHost, ctx->stream));
  SVB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (p_dot_sp) {
    double s = 0.0;
    if (nparts) for (int b = 0; b < nparts; b++) s += hp[b];
    else for (int a = 0; a < ctx->mynNo; a++) s += P[a] * SP[a];
    *p_dot_sp = s;
  }
  return SVB200_OK;
}

int svb200_bench_schur_sp(svb200_ctx* ctx, int32_t variant, int32_t reps, double* ms_per_launch)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ms_per_launch && reps >= 1 && ctx->d_rowPtr, "svb200_bench_schur_sp: bad arguments");
  const size_t nnz = ctx->nnz, nNo = ctx->nNo;
  DevBuf dL, dGt, dDL, dP, dGP, dSP, dPart;
  TRY(dL.alloc(nnz, true, ctx->stream)); TRY(dGt.alloc(3 * nnz, true, ctx->stream)); TRY(dDL.alloc(4 * nnz, true, ctx->stream));
  TRY(dP.alloc(nNo, true, ctx->stream)); TRY(dGP.alloc(3 * nNo, true, ctx->stream)); TRY(dSP.alloc(nNo, true, ctx->stream));
  TRY(dPart.alloc(148 * 8 + 16, true, ctx->stream));
  TRY(schur_run(ctx, variant, dL.p, dGt.p, dDL.p, dP.p, dGP.p, dSP.p, dPart.p, nullptr));
  SVB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  for (int r = 0; r < reps; r++) TRY(schur_run(ctx, variant, dL.p, dGt.p, dDL.p, dP.p, dGP.p, dSP.p, dPart.p, nullptr));
  SVB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SVB_CUDA(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  SVB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *ms_per_launch = ms / reps;
  return SVB200_OK;
}

int svb200_measure_fp64_peak(svb200_ctx* ctx, double* tflops)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(tflops, "svb200_measure_fp64_peak: null output");
  return fp64_peak(ctx, tflops);
}

int64_t svb200_launch_count(svb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

int svb200_host_register(svb200_ctx* ctx, void* ptr, size_t bytes)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ptr && bytes > 0, "svb200_host_register: bad arguments");
  SVB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return SVB200_OK;
}

int svb200_host_unregister(svb200_ctx* ctx, void* ptr)
{
  CTX_GUARD(ctx);
  SVB_CUDA(cudaHostUnregister(ptr));
  return SVB200_OK;
}

int svb200_timer_mark(svb200_ctx* ctx, int32_t which)
{
  CTX_GUARD(ctx);
  SVB_CUDA(cudaEventRecord(which == 0 ? ctx->tm0 : ctx->tm1, ctx->stream));
  return SVB200_OK;
}

int svb200_timer_elapsed(svb200_ctx* ctx, double* ms)
{
  CTX_GUARD(ctx);
  SVB_REQUIRE(ms, "svb200_timer_elapsed: null output");
  SVB_CUDA(cudaEventSynchronize(ctx->tm1));
  float f = 0.f;
  SVB_CUDA(cudaEventElapsedTime(&f, ctx->tm0, ctx->tm1));
  *ms = f;
  return SVB200_OK;
}

}  // extern "C"
