/* Single-rank MPI shim implementation (TEST INFRASTRUCTURE; see mpi.h). */
#include "mpi.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* pending self-messages: sends are buffered, receives match by tag (FIFO per tag) */
typedef struct msg { int tag, bytes; void *data; struct msg *next; } msg_t;
static msg_t *sendq = NULL;
typedef struct rcv { int tag, bytes; void *buf; struct rcv *next; } rcv_t;
static rcv_t *recvq = NULL;

static void deliver(void) {
  rcv_t **rp = &recvq;
  while (*rp) {
    rcv_t *r = *rp; msg_t **mp = &sendq; int hit = 0;
    while (*mp) {
      msg_t *m = *mp;
      if (m->tag == r->tag) {
        memcpy(r->buf, m->data, (size_t)(m->bytes < r->bytes ? m->bytes : r->bytes));
        *mp = m->next; free(m->data); free(m); hit = 1; break;
      }
      mp = &m->next;
    }
    if (hit) { *rp = r->next; free(r); } else rp = &r->next;
  }
}
static void push_send(const void *b, int bytes, int tag) {
  msg_t *m = (msg_t*)malloc(sizeof(msg_t)), **p = &sendq;
  m->tag = tag; m->bytes = bytes; m->data = malloc((size_t)bytes > 0 ? (size_t)bytes : 1); m->next = NULL;
  memcpy(m->data, b, (size_t)bytes);
  while (*p) p = &(*p)->next;
  *p = m; deliver();
}
static void push_recv(void *b, int bytes, int tag) {
  rcv_t *r = (rcv_t*)malloc(sizeof(rcv_t)), **p = &recvq;
  r->tag = tag; r->bytes = bytes; r->buf = b; r->next = NULL;
  while (*p) p = &(*p)->next;
  *p = r; deliver();
}

int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Abort(MPI_Comm c, int code) { (void)c; fprintf(stderr, "MPI_Abort(%d)\n", code); fflush(NULL); abort(); return 0; }
double MPI_Wtime(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9*ts.tv_nsec; }
int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = 0; return 0; }
int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = 1; return 0; }
int MPI_Cart_create(MPI_Comm c, int nd, const int *dims, const int *per, int re, MPI_Comm *out) {
  (void)c; (void)per; (void)re;
  for (int i = 0; i < nd; i++) if (dims[i] != 1) { fprintf(stderr, "mpi shim: only 1 rank supported\n"); abort(); }
  *out = 1; return 0; }
int MPI_Cart_coords(MPI_Comm c, int rank, int maxd, int *coords) { (void)c; (void)rank; for (int i = 0; i < maxd; i++) coords[i] = 0; return 0; }
int MPI_Cart_rank(MPI_Comm c, const int *coords, int *rank) { (void)c; (void)coords; *rank = 0; return 0; }
int MPI_Comm_group(MPI_Comm c, MPI_Group *g) { (void)c; *g = 1; return 0; }
int MPI_Group_incl(MPI_Group g, int n, const int *ranks, MPI_Group *out) { (void)g; (void)n; (void)ranks; *out = 1; return 0; }
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm *out) { (void)c; (void)g; *out = 2; return 0; }
int MPI_Group_free(MPI_Group *g) { (void)g; return 0; }
int MPI_Comm_free(MPI_Comm *c) { (void)c; return 0; }
int MPI_Allreduce(const void *s, void *r, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  (void)op; (void)c; if (s != r) memmove(r, s, (size_t)count*(size_t)t); return 0; }
int MPI_Iallreduce(const void *s, void *r, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c, MPI_Request *req) {
  req->active = 0; return MPI_Allreduce(s, r, count, t, op, c); }
int MPI_Bcast(void *b, int count, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)count; (void)t; (void)root; (void)c; return 0; }
int MPI_Isend(const void *b, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request *req) {
  (void)dest; (void)c; push_send(b, count*t, tag); req->active = 0; return 0; }
int MPI_Irecv(void *b, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *req) {
  (void)src; (void)c; push_recv(b, count*t, tag); req->active = 0; return 0; }
int MPI_Send(const void *b, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c) {
  (void)dest; (void)c; push_send(b, count*t, tag); return 0; }
int MPI_Recv(void *b, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st) {
  (void)src; (void)c; (void)st; push_recv(b, count*t, tag); return 0; }
int MPI_Wait(MPI_Request *req, MPI_Status *st) { (void)req; (void)st; deliver(); return 0; }
int MPI_Info_create(MPI_Info *i) { *i = 0; return 0; }
int MPI_Info_set(MPI_Info i, const char *k, const char *v) { (void)i; (void)k; (void)v; return 0; }
