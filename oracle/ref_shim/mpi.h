/* Single-rank MPI shim, written for this repo (TEST INFRASTRUCTURE, not product code).
 * The reference (/root/reference, C99+MPI) is compiled against this header so that its own CPU
 * implementation can serve as the parity oracle and the CPU baseline.  Only the symbols the reference
 * uses are provided, with one-rank semantics (self send/recv matched by tag through memcpy). */
#ifndef DDA_MPI_SHIM_H
#define DDA_MPI_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int MPI_Comm;
typedef int MPI_Group;
typedef int MPI_Datatype;   /* value = size in bytes */
typedef int MPI_Op;
typedef int MPI_Info;
typedef struct { int src, tag; } MPI_Status;
typedef struct { void *buf; int bytes, tag, kind, done; } *MPI_Request_ptr;
typedef struct mpi_shim_req { void *buf; int bytes, tag, kind, active; } MPI_Request;

#define MPI_COMM_WORLD 0
#define MPI_SUM 1
#define MPI_CHAR 1
#define MPI_INT 4
#define MPI_FLOAT 4
#define MPI_DOUBLE 8
#define MPI_COMPLEX 8
#define MPI_DOUBLE_COMPLEX 16
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_INFO_NULL 0
#define MPI_SUCCESS 0

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm c, int code);
double MPI_Wtime(void);
int MPI_Comm_rank(MPI_Comm c, int *rank);
int MPI_Comm_size(MPI_Comm c, int *size);
int MPI_Cart_create(MPI_Comm c, int ndims, const int *dims, const int *periods, int reorder, MPI_Comm *out);
int MPI_Cart_coords(MPI_Comm c, int rank, int maxdims, int *coords);
int MPI_Cart_rank(MPI_Comm c, const int *coords, int *rank);
int MPI_Comm_group(MPI_Comm c, MPI_Group *g);
int MPI_Group_incl(MPI_Group g, int n, const int *ranks, MPI_Group *out);
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm *out);
int MPI_Group_free(MPI_Group *g);
int MPI_Comm_free(MPI_Comm *c);
int MPI_Allreduce(const void *s, void *r, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c);
int MPI_Iallreduce(const void *s, void *r, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c, MPI_Request *req);
int MPI_Bcast(void *b, int count, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Isend(const void *b, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request *req);
int MPI_Irecv(void *b, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *req);
int MPI_Send(const void *b, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c);
int MPI_Recv(void *b, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st);
int MPI_Wait(MPI_Request *req, MPI_Status *st);
int MPI_Info_create(MPI_Info *i);
int MPI_Info_set(MPI_Info i, const char *k, const char *v);
#ifdef __cplusplus
}
#endif
#endif
