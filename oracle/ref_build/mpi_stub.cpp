// Single-rank implementation of the MPI subset the reference hot path links against.
// Test infrastructure only (see mpi.h in this directory).
#include "mpi.h"
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <chrono>

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
int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int* s) { *s = 1; return MPI_SUCCESS; }
double MPI_Wtime(void)
{
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}
int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) { return copy_if(s, r, n*dt_size(t)); }
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) { return copy_if(s, r, n*dt_size(t)); }
int MPI_Scan(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) { return copy_if(s, r, n*dt_size(t)); }
int MPI_Allgather(const void* s, int n, MPI_Datatype t, void* r, int, MPI_Datatype, MPI_Comm) { return copy_if(s, r, n*dt_size(t)); }
int MPI_Allgatherv(const void* s, int n, MPI_Datatype t, void* r, const int*, const int* d, MPI_Datatype, MPI_Comm)
{ return copy_if(s, (char*)r + (d ? d[0] : 0)*dt_size(t), n*dt_size(t)); }
int MPI_Gather(const void* s, int n, MPI_Datatype t, void* r, int, MPI_Datatype, int, MPI_Comm) { return copy_if(s, r, n*dt_size(t)); }
int MPI_Gatherv(const void* s, int n, MPI_Datatype t, void* r, const int*, const int* d, MPI_Datatype, int, MPI_Comm)
{ return copy_if(s, (char*)r + (d ? d[0] : 0)*dt_size(t), n*dt_size(t)); }
int MPI_Scatter(const void* s, int, MPI_Datatype, void* r, int n, MPI_Datatype t, int, MPI_Comm) { return copy_if(s, r, n*dt_size(t)); }
int MPI_Scatterv(const void* s, const int*, const int* d, MPI_Datatype, void* r, int n, MPI_Datatype t, int, MPI_Comm)
{ return copy_if((const char*)s + (d ? d[0] : 0)*dt_size(t), r, n*dt_size(t)); }
int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm) { return unreachable("MPI_Send"); }
int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*) { return unreachable("MPI_Recv"); }
int MPI_Isend(const void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*) { return unreachable("MPI_Isend"); }
int MPI_Irecv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*) { return unreachable("MPI_Irecv"); }
int MPI_Wait(MPI_Request*, MPI_Status*) { return MPI_SUCCESS; }
int MPI_File_open(MPI_Comm, const char*, int, MPI_Info, MPI_File*) { return unreachable("MPI_File_open"); }
int MPI_File_close(MPI_File*) { return unreachable("MPI_File_close"); }
int MPI_File_set_view(MPI_File, MPI_Offset, MPI_Datatype, MPI_Datatype, const char*, MPI_Info) { return unreachable("MPI_File_set_view"); }
int MPI_File_read(MPI_File, void*, int, MPI_Datatype, MPI_Status*) { return unreachable("MPI_File_read"); }
int MPI_File_write(MPI_File, const void*, int, MPI_Datatype, MPI_Status*) { return unreachable("MPI_File_write"); }
}
