/* Single-rank MPI stand-in used ONLY to compile the unmodified reference hot-path
 * sources into oracle/_ref/libsvref.so (test infrastructure; never linked into the
 * product library).  Every collective degenerates to a local copy; point-to-point
 * calls are never reached with one rank (the reference guards them with
 * `nTasks == 1` early returns, e.g. linear_solver/in_commu.cpp:86-88).
 */
#ifndef SVB200_ORACLE_MPI_STUB_H
#define SVB200_ORACLE_MPI_STUB_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Info;
typedef int MPI_File;
typedef long long MPI_Offset;
typedef long MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_SUCCESS 0
#define MPI_STATUS_SIZE 3
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_INFO_NULL 0
#define MPI_IN_PLACE ((void*)1)

#define MPI_CHARACTER 1
#define MPI_CHAR 1
#define MPI_INTEGER 4
#define MPI_INT 5
#define MPI_UNSIGNED 6
#define MPI_LOGICAL 7
#define MPI_CXX_BOOL 8
#define MPI_DOUBLE 9
#define MPI_DOUBLE_PRECISION 10
#define MPI_LONG 11
#define MPI_FLOAT 12

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_LOR 4
#define MPI_LAND 5

#define MPI_MODE_RDONLY 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_CREATE 4

int MPI_Init(int*, char***);
int MPI_Initialized(int*);
int MPI_Finalized(int*);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm, int);
int MPI_Barrier(MPI_Comm);
int MPI_Comm_rank(MPI_Comm, int*);
int MPI_Comm_size(MPI_Comm, int*);
double MPI_Wtime(void);
int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Allreduce(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Reduce(const void*, void*, int, MPI_Datatype, MPI_Op, int, MPI_Comm);
int MPI_Scan(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Allgather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm);
int MPI_Allgatherv(const void*, int, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, MPI_Comm);
int MPI_Gather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Gatherv(const void*, int, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, int, MPI_Comm);
int MPI_Scatter(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Scatterv(const void*, const int*, const int*, MPI_Datatype, void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*);
int MPI_Isend(const void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*);
int MPI_Irecv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*);
int MPI_Wait(MPI_Request*, MPI_Status*);
int MPI_File_open(MPI_Comm, const char*, int, MPI_Info, MPI_File*);
int MPI_File_close(MPI_File*);
int MPI_File_set_view(MPI_File, MPI_Offset, MPI_Datatype, MPI_Datatype, const char*, MPI_Info);
int MPI_File_read(MPI_File, void*, int, MPI_Datatype, MPI_Status*);
int MPI_File_write(MPI_File, const void*, int, MPI_Datatype, MPI_Status*);

#ifdef __cplusplus
}
#endif
#endif
