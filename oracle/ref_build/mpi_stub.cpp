// MPI subset the reference hot path links against.  Test infrastructure only (see mpi.h in this directory).
//
// Default: a single rank (collectives are copies).  With SVREF_MPI_SIZE=N (2..64), SVREF_MPI_RANK=r and SVREF_MPI_SHM=/name in
// the environment of N cooperating PROCESSES, the calls FSILS makes (Comm_rank/size, Allreduce, Allgather(v), Bcast, Reduce,
// Send/Recv, Isend/Irecv/Wait, Barrier) run over one POSIX shared-memory segment: per-rank collective slots read in rank
// order between two barriers, and a one-message mailbox per ordered rank pair.  This is the multi-rank reference of
// SURVEY.md 8(c) (ranks as processes instead of threads: no shared C++ state to worry about) — enough for
// fsils_lhs_create, fsils_commuv/commus and the Krylov solvers; ParMETIS-scale MPI is out of its reach.
#include "mpi.h"
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <chrono>
#include <atomic>
#include <vector>
#include <sched.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>

namespace {
constexpr int MAXR = 64;      // the segment is sparse: only the slots and mailboxes a run touches are ever backed by memory
constexpr size_t COLL_BYTES = 8u << 20, BOX_BYTES = 1u << 20;
struct Box { std::atomic<int> full; int bytes; char data[BOX_BYTES]; };
struct Shm {
  std::atomic<int> bar_count, bar_gen;
  char coll[MAXR][COLL_BYTES];
  Box box[MAXR][MAXR];          // box[src][dst]
};
int g_size = -1, g_rank = 0;
Shm* g_shm = nullptr;
struct PendingRecv { void* buf; size_t bytes; int src; bool done; };
std::vector<PendingRecv> g_recv;

void mp_init()
{
  if (g_size >= 0) return;
  const char* sz = std::getenv("SVREF_MPI_SIZE");
  g_size = sz ? std::atoi(sz) : 1;
  if (g_size <= 1) { g_size = 1; return; }
  const char* rk = std::getenv("SVREF_MPI_RANK");
  const char* nm = std::getenv("SVREF_MPI_SHM");
  if (!rk || !nm || g_size > MAXR) { std::fprintf(stderr, "[mpi_stub] bad SVREF_MPI_* environment\n"); std::abort(); }
  g_rank = std::atoi(rk);
  // "/name" = POSIX shared memory; a name with a directory part ("/tmp/x") = an ordinary file mapped MAP_SHARED
  const bool plain_file = std::strchr(nm + 1, '/') != nullptr;
  int fd = plain_file ? open(nm, O_CREAT | O_RDWR, 0600) : shm_open(nm, O_CREAT | O_RDWR, 0600);
  if (fd < 0 || ftruncate(fd, sizeof(Shm)) != 0) { std::perror("[mpi_stub] shm_open/ftruncate"); std::abort(); }
  void* p = mmap(nullptr, sizeof(Shm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) { std::perror("[mpi_stub] mmap"); std::abort(); }
  g_shm = static_cast<Shm*>(p);      // a fresh segment is zero-filled: counters and mailbox flags start at 0
}

inline bool multi() { mp_init(); return g_size > 1; }

void mp_barrier()
{
  const int gen = g_shm->bar_gen.load(std::memory_order_acquire);
  if (g_shm->bar_count.fetch_add(1, std::memory_order_acq_rel) + 1 == g_size) {
    g_shm->bar_count.store(0, std::memory_order_relaxed);
    g_shm->bar_gen.fetch_add(1, std::memory_order_release);
  } else {
    while (g_shm->bar_gen.load(std::memory_order_acquire) == gen) sched_yield();
  }
}

void mp_send(const void* buf, size_t bytes, int dst)
{
  if (bytes > BOX_BYTES) { std::fprintf(stderr, "[mpi_stub] message of %zu bytes exceeds the mailbox\n", bytes); std::abort(); }
  Box& b = g_shm->box[g_rank][dst];
  while (b.full.load(std::memory_order_acquire) != 0) sched_yield();
  std::memcpy(b.data, buf, bytes);
  b.bytes = (int)bytes;
  b.full.store(1, std::memory_order_release);
}

void mp_recv(void* buf, size_t bytes, int src)
{
  Box& b = g_shm->box[src][g_rank];
  while (b.full.load(std::memory_order_acquire) != 1) sched_yield();
  std::memcpy(buf, b.data, bytes < (size_t)b.bytes ? bytes : (size_t)b.bytes);
  b.full.store(0, std::memory_order_release);
}

template <class T> void reduce_into(T* acc, const T* x, int n, int op)
{
  for (int i = 0; i < n; i++) {
    switch (op) {
      case MPI_SUM: acc[i] += x[i]; break;
      case MPI_MAX: if (x[i] > acc[i]) acc[i] = x[i]; break;
      case MPI_MIN: if (x[i] < acc[i]) acc[i] = x[i]; break;
      case MPI_LOR: acc[i] = (acc[i] || x[i]); break;
      case MPI_LAND: acc[i] = (acc[i] && x[i]); break;
      default: std::fprintf(stderr, "[mpi_stub] unknown reduction %d\n", op); std::abort();
    }
  }
}
}  // namespace

static size_t dt_size(MPI_Datatype t)
{
  switch (t) {
    case MPI_CHARACTER: return 1;
    case MPI_INTEGER: case MPI_INT: case MPI_UNSIGNED: case MPI_LOGICAL: case MPI_FLOAT: return 4;
    case MPI_CXX_BOOL: return sizeof(bool);
    case MPI_DOUBLE: case MPI_DOUBLE_PRECISION: case MPI_LONG: return 8;
  }
  std::fprintf(stderr, "[mpi_stub] unknown datatype %d\n", t);
  std::abort();
}

static int copy_if(const void* s, void* r, size_t bytes)
{
  if (s != MPI_IN_PLACE && s != r && bytes) std::memcpy(r, s, bytes);
  return MPI_SUCCESS;
}

// Every rank deposits `bytes` in its slot; between the two barriers all slots are readable.
static void mp_exchange(const void* mine, size_t bytes)
{
  if (bytes > COLL_BYTES) { std::fprintf(stderr, "[mpi_stub] collective of %zu bytes exceeds the slot\n", bytes); std::abort(); }
  std::memcpy(g_shm->coll[g_rank], mine, bytes);
  mp_barrier();
}

static int mp_allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root)
{
  const size_t bytes = (size_t)n * dt_size(t);
  mp_exchange(s == MPI_IN_PLACE ? r : s, bytes);
  if (root < 0 || root == g_rank) {
    std::vector<char> acc(g_shm->coll[0], g_shm->coll[0] + bytes);
    for (int q = 1; q < g_size; q++) {
      if (t == MPI_DOUBLE || t == MPI_DOUBLE_PRECISION) reduce_into((double*)acc.data(), (const double*)g_shm->coll[q], n, op);
      else if (t == MPI_INTEGER || t == MPI_INT || t == MPI_LOGICAL) reduce_into((int*)acc.data(), (const int*)g_shm->coll[q], n, op);
      else if (t == MPI_LONG) reduce_into((long*)acc.data(), (const long*)g_shm->coll[q], n, op);
      else if (t == MPI_CXX_BOOL) reduce_into((bool*)acc.data(), (const bool*)g_shm->coll[q], n, op);
      else { std::fprintf(stderr, "[mpi_stub] reduction on datatype %d\n", t); std::abort(); }
    }
    std::memcpy(r, acc.data(), bytes);
  }
  mp_barrier();
  return MPI_SUCCESS;
}

static int unreachable(const char* what)
{
  std::fprintf(stderr, "[mpi_stub] %s called with a single rank\n", what);
  std::abort();
  return 1;
}

extern "C" {
int MPI_Init(int*, char***) { return MPI_SUCCESS; }
int MPI_Initialized(int* f) { *f = 1; return MPI_SUCCESS; }
int MPI_Finalized(int* f) { *f = 0; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm, int c) { std::exit(c); }
int MPI_Barrier(MPI_Comm) { if (multi()) mp_barrier(); return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm, int* r) { *r = multi() ? g_rank : 0; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int* s) { *s = multi() ? g_size : 1; return MPI_SUCCESS; }
double MPI_Wtime(void)
{
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}
int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm)
{
  if (!multi()) return MPI_SUCCESS;
  const size_t bytes = (size_t)n * dt_size(t);
  if (bytes > COLL_BYTES) { std::fprintf(stderr, "[mpi_stub] broadcast too large\n"); std::abort(); }
  if (g_rank == root) std::memcpy(g_shm->coll[root], b, bytes);
  mp_barrier();
  if (g_rank != root) std::memcpy(b, g_shm->coll[root], bytes);
  mp_barrier();
  return MPI_SUCCESS;
}
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm)
{ return multi() ? mp_allreduce(s, r, n, t, op, -1) : copy_if(s, r, n*dt_size(t)); }
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm)
{ return multi() ? mp_allreduce(s, r, n, t, op, root) : copy_if(s, r, n*dt_size(t)); }
int MPI_Scan(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) { return copy_if(s, r, n*dt_size(t)); }
int MPI_Allgather(const void* s, int n, MPI_Datatype t, void* r, int, MPI_Datatype, MPI_Comm)
{
  if (!multi()) return copy_if(s, r, n*dt_size(t));
  const size_t bytes = (size_t)n * dt_size(t);
  mp_exchange(s, bytes);
  for (int q = 0; q < g_size; q++) std::memcpy((char*)r + q * bytes, g_shm->coll[q], bytes);
  mp_barrier();
  return MPI_SUCCESS;
}
int MPI_Allgatherv(const void* s, int n, MPI_Datatype t, void* r, const int* cnt, const int* d, MPI_Datatype, MPI_Comm)
{
  if (!multi()) return copy_if(s, (char*)r + (d ? d[0] : 0)*dt_size(t), n*dt_size(t));
  const size_t sz = dt_size(t);
  mp_exchange(s, (size_t)n * sz);
  for (int q = 0; q < g_size; q++) std::memcpy((char*)r + (size_t)d[q] * sz, g_shm->coll[q], (size_t)cnt[q] * sz);
  mp_barrier();
  return MPI_SUCCESS;
}
int MPI_Gather(const void* s, int n, MPI_Datatype t, void* r, int, MPI_Datatype, int, MPI_Comm) { return copy_if(s, r, n*dt_size(t)); }
int MPI_Gatherv(const void* s, int n, MPI_Datatype t, void* r, const int*, const int* d, MPI_Datatype, int, MPI_Comm)
{ return copy_if(s, (char*)r + (d ? d[0] : 0)*dt_size(t), n*dt_size(t)); }
int MPI_Scatter(const void* s, int, MPI_Datatype, void* r, int n, MPI_Datatype t, int, MPI_Comm) { return copy_if(s, r, n*dt_size(t)); }
int MPI_Scatterv(const void* s, const int*, const int* d, MPI_Datatype, void* r, int n, MPI_Datatype t, int, MPI_Comm)
{ return copy_if((const char*)s + (d ? d[0] : 0)*dt_size(t), r, n*dt_size(t)); }
int MPI_Send(const void* b, int n, MPI_Datatype t, int dst, int, MPI_Comm)
{ if (!multi()) return unreachable("MPI_Send"); mp_send(b, (size_t)n * dt_size(t), dst); return MPI_SUCCESS; }
int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int, MPI_Comm, MPI_Status*)
{ if (!multi()) return unreachable("MPI_Recv"); mp_recv(b, (size_t)n * dt_size(t), src); return MPI_SUCCESS; }
int MPI_Isend(const void* b, int n, MPI_Datatype t, int dst, int, MPI_Comm, MPI_Request* rq)
{
  if (!multi()) return unreachable("MPI_Isend");
  mp_send(b, (size_t)n * dt_size(t), dst);       // eager: the data is in the mailbox when this returns
  *rq = 0;
  return MPI_SUCCESS;
}
int MPI_Irecv(void* b, int n, MPI_Datatype t, int src, int, MPI_Comm, MPI_Request* rq)
{
  if (!multi()) return unreachable("MPI_Irecv");
  g_recv.push_back({b, (size_t)n * dt_size(t), src, false});
  *rq = (int)g_recv.size();                       // 1-based id of the pending receive
  return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request* rq, MPI_Status*)
{
  if (!multi() || *rq <= 0) return MPI_SUCCESS;
  PendingRecv& p = g_recv[*rq - 1];
  if (!p.done) { mp_recv(p.buf, p.bytes, p.src); p.done = true; }
  bool all = true;
  for (auto& q : g_recv) all = all && q.done;
  if (all) g_recv.clear();
  *rq = 0;
  return MPI_SUCCESS;
}
int MPI_File_open(MPI_Comm, const char*, int, MPI_Info, MPI_File*) { return unreachable("MPI_File_open"); }
int MPI_File_close(MPI_File*) { return unreachable("MPI_File_close"); }
int MPI_File_set_view(MPI_File, MPI_Offset, MPI_Datatype, MPI_Datatype, const char*, MPI_Info) { return unreachable("MPI_File_set_view"); }
int MPI_File_read(MPI_File, void*, int, MPI_Datatype, MPI_Status*) { return unreachable("MPI_File_read"); }
int MPI_File_write(MPI_File, const void*, int, MPI_Datatype, MPI_Status*) { return unreachable("MPI_File_write"); }
}
