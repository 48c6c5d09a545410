/* mini-MPI implementation — TEST INFRASTRUCTURE ONLY (part of oracle/; see mpi.h).
 *
 * Ranks are separate processes started by oracle/mpi_shim/mpirun.c (fork+exec) that share one POSIX shared-memory
 * segment: a P x P matrix of single-producer/single-consumer byte rings.  Every message is a framed byte stream
 * (header + payload) on ring[src][dst]; matching follows MPI rules (source, communicator context, tag, in order)
 * with an unexpected-message queue.  Collectives are built from point-to-point in deadlock-free orders.
 * One-sided MPI_Put/MPI_Win_fence (used by the reference's split-dimensional Cannon, spcannon.cxx:64-71,139-152)
 * are implemented as "queue at Put, exchange at fence".
 * Without the launcher (MINIMPI_SIZE unset) MPI_Init creates a private 1-rank world.
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#define RING_BYTES (4u << 20) /* 4 MiB per directed pair */
#define MAX_COMMS 256
#define MAX_REQS 1024
#define MAX_WINS 32

typedef struct {
  volatile uint64_t head; /* bytes produced (written by sender)   */
  char pad0[56];
  volatile uint64_t tail; /* bytes consumed (written by receiver) */
  char pad1[56];
  unsigned char data[RING_BYTES];
} ring_t;

typedef struct {
  volatile int abort_flag;
  char pad[60];
} shm_hdr_t;

typedef struct {
  int ctx, tag;
  int64_t nbytes;
} msg_hdr_t;

typedef struct {
  int used;
  int np, rank; /* my rank in this communicator */
  int ctx;      /* context id agreed by all members */
  int* members; /* world ranks */
} comm_t;

enum { REQ_FREE = 0, REQ_SEND, REQ_RECV };
typedef struct {
  int kind, done;
  int peer; /* world rank */
  int ctx, tag;
  char* buf;
  int64_t nbytes;
  int64_t progress;   /* payload bytes moved so far */
  int hdr_done;       /* send: header written */
  uint64_t seq;       /* posting order */
} req_t;

typedef struct umsg {
  int src, ctx, tag;
  int64_t nbytes;
  char* data;
  struct umsg* next;
} umsg_t;

/* incoming stream state per source */
typedef struct {
  int active;
  msg_hdr_t hdr;
  int64_t got;
  int req;      /* bound posted recv or -1 */
  umsg_t* um;   /* or unexpected buffer */
} instream_t;

typedef struct {
  int used;
  char* base;
  int64_t size;
  int disp_unit;
  int comm;
  /* queued puts */
  int nput, capput;
  struct put_rec {
    int target; /* comm rank */
    int64_t disp_bytes, nbytes;
    const char* src;
  } * puts;
} win_t;

static int g_np = 1, g_rank = 0, g_inited = 0;
static shm_hdr_t* g_hdr = NULL;
static ring_t* g_rings = NULL; /* [src * np + dst] */
static comm_t g_comms[MAX_COMMS];
static req_t g_reqs[MAX_REQS];
static win_t g_wins[MAX_WINS];
static instream_t* g_in = NULL;
static umsg_t *g_uq_head = NULL, *g_uq_tail = NULL;
static uint64_t g_seq = 0;
static int g_next_ctx = 1;
static int* g_send_order = NULL; /* per destination: FIFO of send request ids is implied by seq */

static void die(const char* msg) {
  fprintf(stderr, "[mini-mpi rank %d] fatal: %s\n", g_rank, msg);
  if (g_hdr) g_hdr->abort_flag = 1;
  _exit(86);
}

static inline ring_t* ring(int src, int dst) { return &g_rings[(size_t)src * g_np + dst]; }
static inline int dt_size(MPI_Datatype t) { return t & 0xff; }

static void check_abort(void) {
  if (g_hdr && g_hdr->abort_flag) _exit(87);
}

/* ---- ring primitives ---- */
static int64_t ring_write(ring_t* r, const char* src, int64_t n) {
  uint64_t head = r->head;
  uint64_t tail = __atomic_load_n(&r->tail, __ATOMIC_ACQUIRE);
  uint64_t space = RING_BYTES - (head - tail);
  if ((uint64_t)n > space) n = (int64_t)space;
  if (n <= 0) return 0;
  uint64_t off = head % RING_BYTES;
  uint64_t first = RING_BYTES - off;
  if (first > (uint64_t)n) first = (uint64_t)n;
  memcpy(r->data + off, src, first);
  if ((uint64_t)n > first) memcpy(r->data, src + first, (size_t)n - first);
  __atomic_store_n(&r->head, head + (uint64_t)n, __ATOMIC_RELEASE);
  return n;
}
static int64_t ring_avail(ring_t* r) {
  uint64_t head = __atomic_load_n(&r->head, __ATOMIC_ACQUIRE);
  return (int64_t)(head - r->tail);
}
static void ring_read(ring_t* r, char* dst, int64_t n) { /* caller checked availability */
  uint64_t tail = r->tail;
  uint64_t off = tail % RING_BYTES;
  uint64_t first = RING_BYTES - off;
  if (first > (uint64_t)n) first = (uint64_t)n;
  memcpy(dst, r->data + off, first);
  if ((uint64_t)n > first) memcpy(dst + first, r->data, (size_t)n - first);
  __atomic_store_n(&r->tail, tail + (uint64_t)n, __ATOMIC_RELEASE);
}

/* ---- progress engine ---- */
static int find_posted_recv(int src, int ctx, int tag) {
  int best = -1;
  for (int i = 0; i < MAX_REQS; ++i) {
    req_t* q = &g_reqs[i];
    if (q->kind == REQ_RECV && !q->done && q->progress < 0 /* unbound */ && q->peer == src && q->ctx == ctx &&
        (q->tag == tag || q->tag == MPI_ANY_TAG)) {
      if (best < 0 || q->seq < g_reqs[best].seq) best = i;
    }
  }
  return best;
}

static void complete_unexpected(umsg_t* um) {
  int r = find_posted_recv(um->src, um->ctx, um->tag);
  if (r >= 0) {
    req_t* q = &g_reqs[r];
    if (um->nbytes > q->nbytes) die("message longer than posted receive buffer");
    memcpy(q->buf, um->data, (size_t)um->nbytes);
    q->done = 1;
    free(um->data);
    free(um);
    return;
  }
  um->next = NULL;
  if (g_uq_tail) g_uq_tail->next = um; else g_uq_head = um;
  g_uq_tail = um;
}

static int progress_recv_from(int src) {
  int moved = 0;
  instream_t* in = &g_in[src];
  ring_t* r = ring(src, g_rank);
  for (;;) {
    if (!in->active) {
      if (ring_avail(r) < (int64_t)sizeof(msg_hdr_t)) break;
      ring_read(r, (char*)&in->hdr, sizeof(msg_hdr_t));
      in->active = 1;
      in->got = 0;
      in->um = NULL;
      in->req = find_posted_recv(src, in->hdr.ctx, in->hdr.tag);
      if (in->req >= 0) {
        if (in->hdr.nbytes > g_reqs[in->req].nbytes) die("message longer than posted receive buffer");
        g_reqs[in->req].progress = 0; /* bound */
      } else {
        in->um = (umsg_t*)malloc(sizeof(umsg_t));
        in->um->src = src;
        in->um->ctx = in->hdr.ctx;
        in->um->tag = in->hdr.tag;
        in->um->nbytes = in->hdr.nbytes;
        in->um->data = (char*)malloc(in->hdr.nbytes > 0 ? (size_t)in->hdr.nbytes : 1);
      }
      moved = 1;
    }
    int64_t want = in->hdr.nbytes - in->got;
    if (want > 0) {
      int64_t av = ring_avail(r);
      if (av <= 0) break;
      if (av > want) av = want;
      char* dst = (in->req >= 0) ? g_reqs[in->req].buf + in->got : in->um->data + in->got;
      ring_read(r, dst, av);
      in->got += av;
      moved = 1;
    }
    if (in->got == in->hdr.nbytes) {
      if (in->req >= 0) g_reqs[in->req].done = 1; else complete_unexpected(in->um);
      in->active = 0;
    } else {
      break;
    }
  }
  return moved;
}

static int progress_sends(void) {
  int moved = 0;
  /* per destination, only the oldest unfinished send may write (messages are framed back to back) */
  for (int dst = 0; dst < g_np; ++dst) {
    for (;;) {
      int cur = -1;
      for (int i = 0; i < MAX_REQS; ++i) {
        req_t* q = &g_reqs[i];
        if (q->kind == REQ_SEND && !q->done && q->peer == dst && (cur < 0 || q->seq < g_reqs[cur].seq)) cur = i;
      }
      if (cur < 0) break;
      req_t* q = &g_reqs[cur];
      ring_t* r = ring(g_rank, dst);
      if (!q->hdr_done) {
        uint64_t space = RING_BYTES - (r->head - __atomic_load_n(&r->tail, __ATOMIC_ACQUIRE));
        if (space < sizeof(msg_hdr_t)) break;
        msg_hdr_t h;
        h.ctx = q->ctx;
        h.tag = q->tag;
        h.nbytes = q->nbytes;
        ring_write(r, (const char*)&h, sizeof(h));
        q->hdr_done = 1;
        moved = 1;
      }
      if (q->progress < q->nbytes) {
        int64_t w = ring_write(r, q->buf + q->progress, q->nbytes - q->progress);
        q->progress += w;
        if (w > 0) moved = 1;
      }
      if (q->progress == q->nbytes) {
        q->done = 1;
        continue; /* next send to the same destination */
      }
      break;
    }
  }
  return moved;
}

static void progress(void) {
  int moved = progress_sends();
  for (int s = 0; s < g_np; ++s) moved |= progress_recv_from(s);
  if (!moved) {
    check_abort();
    sched_yield();
  }
}

static int alloc_req(void) {
  for (int i = 0; i < MAX_REQS; ++i)
    if (g_reqs[i].kind == REQ_FREE) return i;
  die("out of request slots");
  return -1;
}

static comm_t* get_comm(MPI_Comm c) {
  if (c < 0 || c >= MAX_COMMS || !g_comms[c].used) die("invalid communicator");
  return &g_comms[c];
}

static int isend_bytes(const void* buf, int64_t nbytes, int dest, int tag, MPI_Comm comm) {
  comm_t* c = get_comm(comm);
  if (dest < 0 || dest >= c->np) die("send: bad destination rank");
  int i = alloc_req();
  req_t* q = &g_reqs[i];
  memset(q, 0, sizeof(*q));
  q->kind = REQ_SEND;
  q->peer = c->members[dest];
  q->ctx = c->ctx;
  q->tag = tag;
  q->buf = (char*)buf;
  q->nbytes = nbytes;
  q->seq = g_seq++;
  return i;
}

static int irecv_bytes(void* buf, int64_t nbytes, int source, int tag, MPI_Comm comm) {
  comm_t* c = get_comm(comm);
  if (source < 0 || source >= c->np) die("recv: bad source rank");
  int i = alloc_req();
  req_t* q = &g_reqs[i];
  memset(q, 0, sizeof(*q));
  q->kind = REQ_RECV;
  q->peer = c->members[source];
  q->ctx = c->ctx;
  q->tag = tag;
  q->buf = (char*)buf;
  q->nbytes = nbytes;
  q->progress = -1; /* unbound */
  q->seq = g_seq++;
  /* already arrived? */
  umsg_t *prev = NULL, *u = g_uq_head;
  while (u) {
    if (u->src == q->peer && u->ctx == q->ctx && (tag == MPI_ANY_TAG || u->tag == tag)) {
      if (u->nbytes > nbytes) die("message longer than posted receive buffer");
      memcpy(buf, u->data, (size_t)u->nbytes);
      if (prev) prev->next = u->next; else g_uq_head = u->next;
      if (g_uq_tail == u) g_uq_tail = prev;
      free(u->data);
      free(u);
      q->done = 1;
      break;
    }
    prev = u;
    u = u->next;
  }
  return i;
}

static void wait_req(int i) {
  if (i < 0) return;
  while (!g_reqs[i].done) progress();
  g_reqs[i].kind = REQ_FREE;
}

static void send_bytes(const void* buf, int64_t n, int dest, int tag, MPI_Comm comm) {
  wait_req(isend_bytes(buf, n, dest, tag, comm));
}
static void recv_bytes(void* buf, int64_t n, int src, int tag, MPI_Comm comm) {
  wait_req(irecv_bytes(buf, n, src, tag, comm));
}

/* ---- init / finalize ---- */
int MPI_Init(int* argc, char*** argv) {
  (void)argc;
  (void)argv;
  if (g_inited) return MPI_SUCCESS;
  const char* ssize = getenv("MINIMPI_SIZE");
  const char* srank = getenv("MINIMPI_RANK");
  const char* sshm = getenv("MINIMPI_SHM");
  size_t bytes;
  void* mem;
  if (ssize && srank && sshm) {
    g_np = atoi(ssize);
    g_rank = atoi(srank);
    bytes = sizeof(shm_hdr_t) + sizeof(ring_t) * (size_t)g_np * g_np;
    int fd = shm_open(sshm, O_RDWR, 0600);
    if (fd < 0) die("shm_open failed");
    mem = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
  } else {
    g_np = 1;
    g_rank = 0;
    bytes = sizeof(shm_hdr_t) + sizeof(ring_t);
    mem = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  }
  if (mem == MAP_FAILED) die("mmap failed");
  g_hdr = (shm_hdr_t*)mem;
  g_rings = (ring_t*)((char*)mem + sizeof(shm_hdr_t));
  g_in = (instream_t*)calloc((size_t)g_np, sizeof(instream_t));
  memset(g_comms, 0, sizeof(g_comms));
  memset(g_reqs, 0, sizeof(g_reqs));
  memset(g_wins, 0, sizeof(g_wins));
  comm_t* w = &g_comms[0];
  w->used = 1;
  w->np = g_np;
  w->rank = g_rank;
  w->ctx = 0;
  w->members = (int*)malloc(sizeof(int) * (size_t)g_np);
  for (int i = 0; i < g_np; ++i) w->members[i] = i;
  g_inited = 1;
  (void)g_send_order;
  return MPI_SUCCESS;
}

int MPI_Finalize(void) {
  if (g_inited) MPI_Barrier(MPI_COMM_WORLD);
  return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int code) {
  (void)comm;
  fflush(stdout);
  if (g_hdr) g_hdr->abort_flag = 1;
  _exit(code ? (code & 0xff) | 1 : 1);
}

double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int MPI_Comm_size(MPI_Comm comm, int* size) { *size = get_comm(comm)->np; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int* rank) { *rank = get_comm(comm)->rank; return MPI_SUCCESS; }

/* ---- collectives (tag space above user tags) ---- */
#define TAG_BARRIER 0x40000001
#define TAG_BCAST 0x40000002
#define TAG_REDUCE 0x40000003
#define TAG_SPLIT 0x40000004
#define TAG_PUT 0x40000005
#define TAG_PUTCNT 0x40000006

int MPI_Barrier(MPI_Comm comm) {
  comm_t* c = get_comm(comm);
  char t = 0;
  if (c->np == 1) return MPI_SUCCESS;
  if (c->rank == 0) {
    for (int i = 1; i < c->np; ++i) recv_bytes(&t, 1, i, TAG_BARRIER, comm);
    for (int i = 1; i < c->np; ++i) send_bytes(&t, 1, i, TAG_BARRIER, comm);
  } else {
    send_bytes(&t, 1, 0, TAG_BARRIER, comm);
    recv_bytes(&t, 1, 0, TAG_BARRIER, comm);
  }
  return MPI_SUCCESS;
}

static void bcast_bytes(void* buf, int64_t n, int root, MPI_Comm comm) {
  comm_t* c = get_comm(comm);
  if (c->np == 1) return;
  /* binomial tree rooted at `root` */
  int vr = (c->rank - root + c->np) % c->np;
  int mask = 1;
  while (mask < c->np) {
    if (vr & mask) {
      recv_bytes(buf, n, (vr - mask + root) % c->np, TAG_BCAST, comm);
      break;
    }
    mask <<= 1;
  }
  mask >>= 1;
  int reqs[32], nr = 0;
  while (mask > 0) {
    if (vr + mask < c->np) reqs[nr++] = isend_bytes(buf, n, (vr + mask + root) % c->np, TAG_BCAST, comm);
    mask >>= 1;
  }
  for (int i = 0; i < nr; ++i) wait_req(reqs[i]);
}

int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm) {
  bcast_bytes(buf, (int64_t)count * dt_size(type), root, comm);
  return MPI_SUCCESS;
}

static void apply_op(void* inout, const void* in, int64_t count, MPI_Datatype type, MPI_Op op) {
#define LOOP(T, EXPR)                                  \
  do {                                                 \
    T* a = (T*)inout;                                  \
    const T* b = (const T*)in;                         \
    for (int64_t i = 0; i < count; ++i) a[i] = (EXPR); \
  } while (0)
#define ARITH(T)                                                       \
  switch (op) {                                                        \
    case MPI_SUM: LOOP(T, a[i] + b[i]); break;                         \
    case MPI_PROD: LOOP(T, a[i] * b[i]); break;                        \
    case MPI_MAX: LOOP(T, a[i] > b[i] ? a[i] : b[i]); break;           \
    case MPI_MIN: LOOP(T, a[i] < b[i] ? a[i] : b[i]); break;           \
    default: die("unsupported reduction op for this datatype");        \
  }
#define INTEG(T)                                                       \
  switch (op) {                                                        \
    case MPI_SUM: LOOP(T, a[i] + b[i]); break;                         \
    case MPI_PROD: LOOP(T, a[i] * b[i]); break;                        \
    case MPI_MAX: LOOP(T, a[i] > b[i] ? a[i] : b[i]); break;           \
    case MPI_MIN: LOOP(T, a[i] < b[i] ? a[i] : b[i]); break;           \
    case MPI_BAND: LOOP(T, a[i] & b[i]); break;                        \
    case MPI_BOR: LOOP(T, a[i] | b[i]); break;                         \
    case MPI_LAND: LOOP(T, a[i] && b[i]); break;                       \
    case MPI_LOR: LOOP(T, a[i] || b[i]); break;                        \
    default: die("unsupported reduction op");                          \
  }
  switch (type) {
    case MPI_DOUBLE: ARITH(double); break;
    case MPI_FLOAT: ARITH(float); break;
    case MPI_INT: INTEG(int); break;
    case MPI_UNSIGNED: INTEG(unsigned); break;
    case MPI_INT64_T: INTEG(int64_t); break;
    case MPI_LONG: INTEG(long); break;
    case MPI_LONG_LONG: INTEG(long long); break;
    case MPI_CHAR: INTEG(char); break;
    case MPI_BYTE: INTEG(unsigned char); break;
    default: die("unsupported reduction datatype");
  }
}

/* binomial-tree reduce to `root`; non-roots use a scratch accumulator */
static void reduce_impl(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, int root,
                        MPI_Comm comm) {
  comm_t* c = get_comm(comm);
  const int64_t n = (int64_t)count * dt_size(type);
  const int is_root = (c->rank == root);
  char* acc;
  if (is_root) {
    acc = (char*)recvbuf;
    if (sendbuf != MPI_IN_PLACE && sendbuf != recvbuf) memcpy(acc, sendbuf, (size_t)n);
  } else {
    acc = (char*)malloc(n > 0 ? (size_t)n : 1);
    memcpy(acc, sendbuf == MPI_IN_PLACE ? recvbuf : sendbuf, (size_t)n);
  }
  if (c->np > 1) {
    char* tmp = (char*)malloc(n > 0 ? (size_t)n : 1);
    int vr = (c->rank - root + c->np) % c->np;
    int mask = 1;
    while (mask < c->np) {
      if (vr & mask) {
        send_bytes(acc, n, (vr - mask + root) % c->np, TAG_REDUCE, comm);
        break;
      }
      if (vr + mask < c->np) {
        recv_bytes(tmp, n, (vr + mask + root) % c->np, TAG_REDUCE, comm);
        apply_op(acc, tmp, count, type, op);
      }
      mask <<= 1;
    }
    free(tmp);
  }
  if (!is_root) free(acc);
}

int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm) {
  reduce_impl(sendbuf, recvbuf, count, type, op, root, comm);
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm) {
  comm_t* c = get_comm(comm);
  const int64_t n = (int64_t)count * dt_size(type);
  if (c->rank == 0) {
    reduce_impl(sendbuf, recvbuf, count, type, op, 0, comm);
  } else {
    if (sendbuf != MPI_IN_PLACE && sendbuf != recvbuf) {
      reduce_impl(sendbuf, NULL, count, type, op, 0, comm);
    } else {
      reduce_impl(MPI_IN_PLACE, recvbuf, count, type, op, 0, comm);
    }
  }
  bcast_bytes(recvbuf, n, 0, comm);
  return MPI_SUCCESS;
}

int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm* newcomm) {
  comm_t* c = get_comm(comm);
  int np = c->np;
  int* all = (int*)malloc(sizeof(int) * 3 * (size_t)np);
  int mine[3] = {color, key, g_next_ctx};
  /* allgather (color, key, next_ctx) through rank 0 */
  if (c->rank == 0) {
    memcpy(all, mine, sizeof(mine));
    for (int i = 1; i < np; ++i) recv_bytes(all + 3 * i, sizeof(mine), i, TAG_SPLIT, comm);
  } else {
    send_bytes(mine, sizeof(mine), 0, TAG_SPLIT, comm);
  }
  bcast_bytes(all, (int64_t)sizeof(int) * 3 * np, 0, comm);
  int ctx = 0;
  for (int i = 0; i < np; ++i)
    if (all[3 * i + 2] > ctx) ctx = all[3 * i + 2];
  g_next_ctx = ctx + 1;
  if (color == MPI_UNDEFINED) {
    *newcomm = MPI_COMM_NULL;
    free(all);
    return MPI_SUCCESS;
  }
  int slot = -1;
  for (int i = 1; i < MAX_COMMS; ++i)
    if (!g_comms[i].used) { slot = i; break; }
  if (slot < 0) die("out of communicator slots");
  comm_t* nc = &g_comms[slot];
  nc->used = 1;
  nc->ctx = ctx;
  nc->members = (int*)malloc(sizeof(int) * (size_t)np);
  nc->np = 0;
  /* stable selection sort by (key, parent rank) over members with my color */
  char* taken = (char*)calloc((size_t)np, 1);
  for (;;) {
    int best = -1;
    for (int i = 0; i < np; ++i) {
      if (taken[i] || all[3 * i] != color) continue;
      if (best < 0 || all[3 * i + 1] < all[3 * best + 1]) best = i;
    }
    if (best < 0) break;
    taken[best] = 1;
    if (best == c->rank) nc->rank = nc->np;
    nc->members[nc->np++] = c->members[best];
  }
  free(taken);
  free(all);
  *newcomm = slot;
  return MPI_SUCCESS;
}

int MPI_Comm_free(MPI_Comm* comm) {
  if (*comm > 0 && *comm < MAX_COMMS && g_comms[*comm].used) {
    free(g_comms[*comm].members);
    g_comms[*comm].used = 0;
  }
  *comm = MPI_COMM_NULL;
  return MPI_SUCCESS;
}

/* ---- point to point ---- */
int MPI_Send(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm) {
  send_bytes(buf, (int64_t)count * dt_size(type), dest, tag, comm);
  return MPI_SUCCESS;
}
int MPI_Recv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status* status) {
  recv_bytes(buf, (int64_t)count * dt_size(type), source, tag, comm);
  if (status) { status->MPI_SOURCE = source; status->MPI_TAG = tag; status->MPI_ERROR = 0; }
  return MPI_SUCCESS;
}
int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request* req) {
  *req = isend_bytes(buf, (int64_t)count * dt_size(type), dest, tag, comm);
  progress();
  return MPI_SUCCESS;
}
int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request* req) {
  *req = irecv_bytes(buf, (int64_t)count * dt_size(type), source, tag, comm);
  return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request* req, MPI_Status* status) {
  if (*req != MPI_REQUEST_NULL) wait_req(*req);
  *req = MPI_REQUEST_NULL;
  if (status) status->MPI_ERROR = 0;
  return MPI_SUCCESS;
}
int MPI_Waitall(int n, MPI_Request* reqs, MPI_Status* statuses) {
  (void)statuses;
  for (int i = 0; i < n; ++i) MPI_Wait(&reqs[i], MPI_STATUS_IGNORE);
  return MPI_SUCCESS;
}
int MPI_Sendrecv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, int dest, int sendtag, void* recvbuf,
                 int recvcount, MPI_Datatype recvtype, int source, int recvtag, MPI_Comm comm, MPI_Status* status) {
  int r = irecv_bytes(recvbuf, (int64_t)recvcount * dt_size(recvtype), source, recvtag, comm);
  int s = isend_bytes(sendbuf, (int64_t)sendcount * dt_size(sendtype), dest, sendtag, comm);
  wait_req(s);
  wait_req(r);
  if (status) { status->MPI_SOURCE = source; status->MPI_TAG = recvtag; status->MPI_ERROR = 0; }
  return MPI_SUCCESS;
}

int MPI_Sendrecv_replace(void* buf, int count, MPI_Datatype type, int dest, int sendtag, int source, int recvtag,
                         MPI_Comm comm, MPI_Status* status) {
  const int64_t n = (int64_t)count * dt_size(type);
  char* tmp = (char*)malloc(n > 0 ? (size_t)n : 1);
  memcpy(tmp, buf, (size_t)n);
  int r = irecv_bytes(buf, n, source, recvtag, comm);
  int s = isend_bytes(tmp, n, dest, sendtag, comm);
  wait_req(s);
  wait_req(r);
  free(tmp);
  if (status) { status->MPI_SOURCE = source; status->MPI_TAG = recvtag; status->MPI_ERROR = 0; }
  return MPI_SUCCESS;
}

#define TAG_GATHER 0x40000007

int MPI_Gatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int* recvcounts,
                const int* displs, MPI_Datatype recvtype, int root, MPI_Comm comm) {
  comm_t* c = get_comm(comm);
  const int64_t sn = (int64_t)sendcount * dt_size(sendtype);
  if (c->rank == root) {
    const int rs = dt_size(recvtype);
    for (int r = 0; r < c->np; ++r) {
      char* dst = (char*)recvbuf + (int64_t)displs[r] * rs;
      if (r == root) {
        if (sendbuf != MPI_IN_PLACE) memcpy(dst, sendbuf, (size_t)sn);
      } else {
        recv_bytes(dst, (int64_t)recvcounts[r] * rs, r, TAG_GATHER, comm);
      }
    }
  } else {
    send_bytes(sendbuf, sn, root, TAG_GATHER, comm);
  }
  return MPI_SUCCESS;
}

int MPI_Gather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
               MPI_Datatype recvtype, int root, MPI_Comm comm) {
  comm_t* c = get_comm(comm);
  int* counts = (int*)malloc(sizeof(int) * (size_t)c->np);
  int* displs = (int*)malloc(sizeof(int) * (size_t)c->np);
  for (int r = 0; r < c->np; ++r) { counts[r] = recvcount; displs[r] = r * recvcount; }
  MPI_Gatherv(sendbuf, sendcount, sendtype, recvbuf, counts, displs, recvtype, root, comm);
  free(counts);
  free(displs);
  return MPI_SUCCESS;
}

int MPI_Allgather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
                  MPI_Datatype recvtype, MPI_Comm comm) {
  comm_t* c = get_comm(comm);
  if (sendbuf == MPI_IN_PLACE) {  /* my contribution already sits in its slot */
    char* mine = (char*)recvbuf + (int64_t)c->rank * recvcount * dt_size(recvtype);
    char* tmp = (char*)malloc((size_t)recvcount * dt_size(recvtype) + 1);
    memcpy(tmp, mine, (size_t)recvcount * dt_size(recvtype));
    MPI_Gather(tmp, recvcount, recvtype, recvbuf, recvcount, recvtype, 0, comm);
    free(tmp);
  } else {
    MPI_Gather(sendbuf, sendcount, sendtype, recvbuf, recvcount, recvtype, 0, comm);
  }
  bcast_bytes(recvbuf, (int64_t)c->np * recvcount * dt_size(recvtype), 0, comm);
  return MPI_SUCCESS;
}

/* ---- one-sided: Put is queued, data moves at the closing fence ---- */
int MPI_Win_create(void* base, MPI_Aint size, int disp_unit, MPI_Info info, MPI_Comm comm, MPI_Win* win) {
  (void)info;
  int slot = -1;
  for (int i = 0; i < MAX_WINS; ++i)
    if (!g_wins[i].used) { slot = i; break; }
  if (slot < 0) {
    /* the reference never frees its windows (spcannon.cxx:270-273): recycle the oldest pair */
    static int recycle = 0;
    slot = recycle;
    recycle = (recycle + 1) % MAX_WINS;
    free(g_wins[slot].puts);
  }
  win_t* w = &g_wins[slot];
  memset(w, 0, sizeof(*w));
  w->used = 1;
  w->base = (char*)base;
  w->size = size;
  w->disp_unit = disp_unit;
  w->comm = comm;
  *win = slot;
  MPI_Barrier(comm);
  return MPI_SUCCESS;
}

int MPI_Win_free(MPI_Win* win) {
  if (*win >= 0 && *win < MAX_WINS) {
    free(g_wins[*win].puts);
    g_wins[*win].used = 0;
  }
  return MPI_SUCCESS;
}

int MPI_Put(const void* origin, int origin_count, MPI_Datatype origin_type, int target_rank, MPI_Aint target_disp,
            int target_count, MPI_Datatype target_type, MPI_Win win) {
  (void)target_count;
  (void)target_type;
  win_t* w = &g_wins[win];
  if (w->nput == w->capput) {
    w->capput = w->capput ? 2 * w->capput : 16;
    w->puts = (struct put_rec*)realloc(w->puts, sizeof(struct put_rec) * (size_t)w->capput);
  }
  struct put_rec* p = &w->puts[w->nput++];
  p->target = target_rank;
  p->disp_bytes = (int64_t)target_disp * w->disp_unit;
  p->nbytes = (int64_t)origin_count * dt_size(origin_type);
  p->src = (const char*)origin;
  return MPI_SUCCESS;
}

int MPI_Win_fence(int assert_, MPI_Win win) {
  (void)assert_;
  win_t* w = &g_wins[win];
  comm_t* c = get_comm(w->comm);
  const int np = c->np;
  /* 1. tell every rank how many puts are coming from me */
  int* outcnt = (int*)calloc((size_t)np, sizeof(int));
  int* incnt = (int*)calloc((size_t)np, sizeof(int));
  for (int i = 0; i < w->nput; ++i) outcnt[w->puts[i].target]++;
  int* rq = (int*)malloc(sizeof(int) * 2 * (size_t)np);
  for (int r = 0; r < np; ++r) {
    rq[2 * r] = irecv_bytes(&incnt[r], sizeof(int), r, TAG_PUTCNT, w->comm);
    rq[2 * r + 1] = isend_bytes(&outcnt[r], sizeof(int), r, TAG_PUTCNT, w->comm);
  }
  for (int r = 0; r < 2 * np; ++r) wait_req(rq[r]);
  /* 2. ship (disp,len) + payload for each queued put; local puts are plain copies */
  int64_t(*meta)[2] = (int64_t(*)[2])malloc(sizeof(int64_t) * 2 * (size_t)(w->nput ? w->nput : 1));
  int* sreq = (int*)malloc(sizeof(int) * 2 * (size_t)(w->nput ? w->nput : 1));
  int ns = 0;
  for (int i = 0; i < w->nput; ++i) {
    struct put_rec* p = &w->puts[i];
    meta[i][0] = p->disp_bytes;
    meta[i][1] = p->nbytes;
    sreq[ns++] = isend_bytes(meta[i], sizeof(int64_t) * 2, p->target, TAG_PUT, w->comm);
    sreq[ns++] = isend_bytes(p->src, p->nbytes, p->target, TAG_PUT, w->comm);
  }
  for (int r = 0; r < np; ++r) {
    for (int k = 0; k < incnt[r]; ++k) {
      int64_t m[2];
      recv_bytes(m, sizeof(m), r, TAG_PUT, w->comm);
      if (m[0] < 0 || m[0] + m[1] > w->size) die("MPI_Put outside the target window");
      recv_bytes(w->base + m[0], m[1], r, TAG_PUT, w->comm);
    }
  }
  for (int i = 0; i < ns; ++i) wait_req(sreq[i]);
  w->nput = 0;
  free(meta);
  free(sreq);
  free(rq);
  free(outcnt);
  free(incnt);
  MPI_Barrier(w->comm);
  return MPI_SUCCESS;
}

/* ---- PMPI aliases ---- */
int PMPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { return MPI_Allreduce(s, r, n, t, op, c); }
int PMPI_Send(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) { return MPI_Send(b, n, t, d, tag, c); }
int PMPI_Recv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status* st) { return MPI_Recv(b, n, t, s, tag, c, st); }
int PMPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) { return MPI_Bcast(b, n, t, root, c); }
int PMPI_Barrier(MPI_Comm c) { return MPI_Barrier(c); }
int PMPI_Comm_rank(MPI_Comm c, int* r) { return MPI_Comm_rank(c, r); }
int PMPI_Comm_size(MPI_Comm c, int* s) { return MPI_Comm_size(c, s); }
double PMPI_Wtime(void) { return MPI_Wtime(); }
