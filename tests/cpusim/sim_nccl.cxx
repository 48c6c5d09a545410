// tests/cpusim — the NCCL calls the product makes, re-implemented between PROCESSES over a named POSIX shared-memory
// segment (TEST INFRASTRUCTURE ONLY, see sim_device.h).  Declarations come from the real <nccl.h>.
//
// Model: every ordered pair of world ranks owns a single-producer / single-consumer byte ring.  An operation is executed
// when it is enqueued (streams are synchronous in the simulator) — or at ncclGroupEnd for grouped calls — by a progress
// loop over nonblocking sends and receives; collectives are built from those.  Messages carry (communicator id,
// operation kind, byte count): a receive whose communicator differs from the message at the head of the ring parks that
// message in an "unexpected" queue, so operations on DIFFERENT communicators may be issued in different orders by
// different ranks, as NCCL allows for independent streams.  On the SAME communicator the order must match, and it is
// checked: a kind or size mismatch between the two ends aborts with both descriptions, and so does a progress loop that
// is stuck for CPUSIM_TIMEOUT seconds (default 60) — a deadlock in a host schedule becomes a readable test failure.
// Reductions sum in rank order, so results are deterministic.
#include <cuda_runtime.h>
#include <errno.h>
#include <fcntl.h>
#include <nccl.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <deque>
#include <functional>
#include <vector>

#include "sim_internal.h"

namespace {

constexpr int kMaxRanks = 16;
constexpr size_t kRing = 1u << 18;

struct Channel {
  std::atomic<uint64_t> head;  // bytes ever written (producer)
  char pad0[56];
  std::atomic<uint64_t> tail;  // bytes ever read (consumer)
  char pad1[56];
  char data[kRing];
};

struct Segment {
  std::atomic<int> attached;
  std::atomic<int> detached;
  Channel ch[kMaxRanks * kMaxRanks];  // ch[src * kMaxRanks + dst]
};

enum Kind : uint32_t { K_P2P = 1, K_BCAST, K_ALLREDUCE, K_REDUCESCATTER, K_ALLGATHER, K_SPLIT, K_INIT };
const char* kind_name(uint32_t k) {
  static const char* n[] = {"?", "send/recv", "broadcast", "all-reduce", "reduce-scatter", "all-gather", "comm-split", "init"};
  return k < 8 ? n[k] : "?";
}

struct Header {
  uint64_t tag;
  uint64_t bytes;
  uint32_t kind;
  uint32_t seq;
};

struct Unexpected {
  Header h;
  std::vector<char> data;
  bool complete = false;
};

struct Op {
  bool send;
  int peer;  // world rank
  char* buf;
  size_t bytes;
  Header h;
  size_t done = 0;
  size_t hdr_done = 0;
  bool complete = false;
  bool direct = false;           // (recv) the message at the head of the ring is streaming into buf
  Unexpected* waiting = nullptr; // (recv) an earlier unexpected message with my tag is still arriving
};

struct RxState {
  bool in_msg = false;
  Header h;
  size_t hdr_got = 0, got = 0;
  Op* op = nullptr;
  Unexpected* ux = nullptr;
  std::deque<Unexpected*> unexpected;
};

struct World {
  Segment* seg = nullptr;
  int nranks = 0, rank = 0, refs = 0;
  char name[128];
  RxState rx[kMaxRanks];
};

}  // namespace

struct ncclComm {
  World* w;
  uint64_t id;
  std::vector<int> members;  // world ranks, indexed by communicator rank
  int rank;
  uint32_t splits = 0;
  uint32_t seq = 0;
};

namespace {

struct Request {
  std::vector<Op> ops;
  std::function<void()> start;    // local copies that belong to the operation (root's send -> recv, own all-gather slot)
  std::function<void()> finish;   // reductions, once every contribution has arrived
  World* w;
  cudaStream_t stream = nullptr;
};

}  // namespace

struct cpusim::NcclBatch {
  std::vector<Request*> reqs;
  bool started = false;
};

namespace {

int g_group_depth = 0;
std::vector<Request*> g_pending;
char g_errbuf[256] = "";

double now_s() {
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + t.tv_nsec * 1e-9;
}

size_t type_size(ncclDataType_t t) {
  switch (t) {
    case ncclChar:
    case ncclUint8:
      return 1;
    case ncclInt32:
    case ncclUint32:
    case ncclFloat32:
      return 4;
    case ncclInt64:
    case ncclUint64:
    case ncclFloat64:
      return 8;
    default:
      fprintf(stderr, "cpusim nccl: unsupported datatype %d\n", (int)t);
      abort();
  }
}

void describe(const Op& o, int me) {
  fprintf(stderr, "  [rank %d] %s %s world-rank %d: %zu bytes, comm %016llx, %s #%u, %zu done%s\n", me,
          o.send ? "send to" : "recv from", o.complete ? "(complete)" : "(PENDING)", o.peer, o.bytes,
          (unsigned long long)o.h.tag, kind_name(o.h.kind), o.h.seq, o.done, o.waiting ? " (behind an unexpected message)" : "");
}

// ---- ring primitives ---------------------------------------------------------------------------------------------------
size_t ring_write(Channel& c, const char* src, size_t n) {
  const uint64_t head = c.head.load(std::memory_order_relaxed), tail = c.tail.load(std::memory_order_acquire);
  size_t space = kRing - (size_t)(head - tail);
  if (n > space) n = space;
  if (!n) return 0;
  size_t off = head % kRing, first = std::min(n, kRing - off);
  memcpy(c.data + off, src, first);
  memcpy(c.data, src + first, n - first);
  c.head.store(head + n, std::memory_order_release);
  return n;
}
size_t ring_read(Channel& c, char* dst, size_t n) {
  const uint64_t tail = c.tail.load(std::memory_order_relaxed), head = c.head.load(std::memory_order_acquire);
  size_t avail = (size_t)(head - tail);
  if (n > avail) n = avail;
  if (!n) return 0;
  size_t off = tail % kRing, first = std::min(n, kRing - off);
  memcpy(dst, c.data + off, first);
  memcpy(dst + first, c.data, n - first);
  c.tail.store(tail + n, std::memory_order_release);
  return n;
}

[[noreturn]] void mismatch(World* w, const Op& o, const Header& h, int src) {
  fprintf(stderr,
          "cpusim nccl: MISMATCHED OPERATIONS on communicator %016llx between world ranks %d -> %d:\n"
          "  sender issued   %s #%u of %llu bytes\n  receiver issued %s #%u of %zu bytes\n",
          (unsigned long long)h.tag, src, w->rank, kind_name(h.kind), h.seq, (unsigned long long)h.bytes, kind_name(o.h.kind),
          o.h.seq, o.bytes);
  abort();
}

// one pass over the sends of `ops` (per destination: in list order)
bool progress_sends(World* w, std::vector<Op*>& ops) {
  bool any = false;
  bool busy[kMaxRanks] = {false};
  for (Op* o : ops) {
    if (!o->send || o->complete) continue;
    if (busy[o->peer]) continue;
    busy[o->peer] = true;  // later sends to this peer wait for this one
    Channel& c = w->seg->ch[w->rank * kMaxRanks + o->peer];
    if (o->hdr_done < sizeof(Header)) {
      size_t n = ring_write(c, reinterpret_cast<const char*>(&o->h) + o->hdr_done, sizeof(Header) - o->hdr_done);
      o->hdr_done += n;
      any |= n > 0;
      if (o->hdr_done < sizeof(Header)) continue;
    }
    if (o->done < o->bytes) {
      size_t n = ring_write(c, o->buf + o->done, o->bytes - o->done);
      o->done += n;
      any |= n > 0;
    }
    if (o->done == o->bytes) {
      o->complete = true;
      any = true;
    }
  }
  return any;
}

bool progress_recvs(World* w, std::vector<Op*>& ops) {
  bool any = false;
  // 1. receives that an already-parked message satisfies (oldest matching message first)
  for (Op* o : ops) {
    if (o->send || o->complete || o->direct) continue;
    RxState& rx = w->rx[o->peer];
    o->waiting = nullptr;
    for (auto it = rx.unexpected.begin(); it != rx.unexpected.end(); ++it) {
      Unexpected* u = *it;
      if (u->h.tag != o->h.tag) continue;
      bool claimed = false;  // an earlier pending receive of this call on the same communicator comes first
      for (Op* p : ops) {
        if (p == o) break;
        if (!p->send && !p->complete && p->peer == o->peer && p->h.tag == o->h.tag && p->waiting == u) claimed = true;
      }
      if (claimed) continue;
      if (u->h.kind != o->h.kind || u->h.bytes != o->bytes) mismatch(w, *o, u->h, o->peer);
      if (u->complete) {
        memcpy(o->buf, u->data.data(), o->bytes);
        o->done = o->bytes;
        o->complete = true;
        rx.unexpected.erase(it);
        delete u;
        any = true;
      } else {
        o->waiting = u;
      }
      break;
    }
  }
  // 2. pull from the rings
  for (int src = 0; src < w->nranks; ++src) {
    if (src == w->rank) continue;
    RxState& rx = w->rx[src];
    Channel& c = w->seg->ch[src * kMaxRanks + w->rank];
    for (;;) {
      if (!rx.in_msg) {
        size_t n = ring_read(c, reinterpret_cast<char*>(&rx.h) + rx.hdr_got, sizeof(Header) - rx.hdr_got);
        rx.hdr_got += n;
        any |= n > 0;
        if (rx.hdr_got < sizeof(Header)) break;
        rx.hdr_got = 0;
        rx.in_msg = true;
        rx.got = 0;
        rx.op = nullptr;
        rx.ux = nullptr;
        for (Op* o : ops) {
          if (o->send || o->complete || o->direct || o->waiting || o->peer != src || o->h.tag != rx.h.tag) continue;
          if (o->h.kind != rx.h.kind || o->bytes != rx.h.bytes) mismatch(w, *o, rx.h, src);
          rx.op = o;
          o->direct = true;
          break;
        }
        if (!rx.op) {
          rx.ux = new Unexpected();
          rx.ux->h = rx.h;
          rx.ux->data.resize(rx.h.bytes);
          rx.unexpected.push_back(rx.ux);
        }
      }
      char* dst = rx.op ? rx.op->buf : rx.ux->data.data();
      size_t n = ring_read(c, dst + rx.got, rx.h.bytes - rx.got);
      rx.got += n;
      any |= n > 0;
      if (rx.got < rx.h.bytes) break;
      if (rx.op) {
        rx.op->done = rx.op->bytes;
        rx.op->complete = true;
        rx.op->direct = false;
      } else {
        rx.ux->complete = true;
      }
      rx.in_msg = false;
      any = true;
    }
  }
  return any;
}

// group the requests of one ncclGroup (or one ungrouped call) by stream and hand them to the stream scheduler
void flush_pending() {
  std::vector<Request*> reqs;
  reqs.swap(g_pending);
  std::vector<cudaStream_t> streams;
  for (Request* r : reqs)
    if (std::find(streams.begin(), streams.end(), r->stream) == streams.end()) streams.push_back(r->stream);
  for (cudaStream_t st : streams) {
    cpusim::NcclBatch* b = new cpusim::NcclBatch();
    for (Request* r : reqs)
      if (r->stream == st) b->reqs.push_back(r);
    cpusim::stream_submit_nccl(st, b);
  }
}

void submit(Request* r, cudaStream_t st) {
  r->stream = st;
  g_pending.push_back(r);
  if (g_group_depth == 0) flush_pending();
}

}  // namespace

namespace cpusim {

bool nccl_progress(std::vector<NcclBatch*>& active) {
  std::vector<World*> worlds;
  std::vector<std::vector<Op*>> ops;
  for (NcclBatch* b : active) {
    if (!b->started) {
      b->started = true;
      for (Request* r : b->reqs)
        if (r->start) r->start();
    }
    for (Request* r : b->reqs) {
      size_t wi = std::find(worlds.begin(), worlds.end(), r->w) - worlds.begin();
      if (wi == worlds.size()) {
        worlds.push_back(r->w);
        ops.emplace_back();
      }
      for (Op& o : r->ops) ops[wi].push_back(&o);
    }
  }
  bool any = false;
  for (size_t wi = 0; wi < worlds.size(); ++wi) {
    any |= progress_sends(worlds[wi], ops[wi]);
    any |= progress_recvs(worlds[wi], ops[wi]);
  }
  return any;
}

bool nccl_done(const NcclBatch* b) {
  if (!b->started) return false;
  for (const Request* r : b->reqs)
    for (const Op& o : r->ops)
      if (!o.complete) return false;
  return true;
}

void nccl_finish(NcclBatch* b) {
  for (Request* r : b->reqs) {
    if (r->finish) r->finish();
    delete r;
  }
  delete b;
}

void nccl_describe(const NcclBatch* b) {
  for (const Request* r : b->reqs)
    for (const Op& o : r->ops) describe(o, r->w->rank);
}

void run_batch_blocking(NcclBatch* b) {
  static double timeout = getenv("CPUSIM_TIMEOUT") ? atof(getenv("CPUSIM_TIMEOUT")) : 60.0;
  std::vector<NcclBatch*> active{b};
  double last = now_s();
  long idle = 0;
  for (;;) {
    const bool any = nccl_progress(active);
    if (nccl_done(b)) break;
    if (any) {
      idle = 0;
      last = now_s();
      continue;
    }
    if (++idle < 200) {
      sched_yield();
    } else {
      timespec ts = {0, 50000};
      nanosleep(&ts, nullptr);
      if ((idle & 1023) == 0 && now_s() - last > timeout) {
        fprintf(stderr, "cpusim nccl: NO PROGRESS for %.0f s — deadlock in the communication schedule?  Operations of this call:\n",
                timeout);
        nccl_describe(b);
        abort();
      }
    }
  }
  nccl_finish(b);
}

}  // namespace cpusim

namespace {

Request* new_request(ncclComm* c) {
  Request* r = new Request();
  r->w = c->w;
  return r;
}

void add_op(Request* r, ncclComm* c, bool send, int peer_in_comm, const void* buf, size_t bytes, uint32_t kind, uint32_t seq) {
  Op o;
  o.send = send;
  o.peer = c->members[peer_in_comm];
  o.buf = const_cast<char*>(static_cast<const char*>(buf));
  o.bytes = bytes;
  o.h.tag = c->id;
  o.h.bytes = bytes;
  o.h.kind = kind;
  o.h.seq = seq;
  r->ops.push_back(o);
}

uint64_t mix(uint64_t a, uint64_t b) {
  uint64_t x = a * 0x9E3779B97F4A7C15ull ^ (b + 0xD1B54A32D192ED03ull + (a << 6) + (a >> 2));
  x ^= x >> 31;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 29;
  return x ? x : 1;
}

void allgather_now(ncclComm* c, const void* mine, void* all, size_t bytes, uint32_t kind) {
  Request* r = new_request(c);
  const uint32_t seq = c->seq++;
  const int P = (int)c->members.size();
  for (int p = 0; p < P; ++p) {
    if (p == c->rank) continue;
    add_op(r, c, true, p, mine, bytes, kind, seq);
    add_op(r, c, false, p, static_cast<char*>(all) + (size_t)p * bytes, bytes, kind, seq);
  }
  memmove(static_cast<char*>(all) + (size_t)c->rank * bytes, mine, bytes);
  cpusim::NcclBatch* b = new cpusim::NcclBatch();
  b->reqs.push_back(r);
  cpusim::run_batch_blocking(b);
}

void world_release(World* w) {
  if (--w->refs > 0) return;
  munmap(w->seg, sizeof(Segment));
  for (int s = 0; s < kMaxRanks; ++s)
    for (Unexpected* u : w->rx[s].unexpected) delete u;
  delete w;
}

}  // namespace

extern "C" {

const char* ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : (g_errbuf[0] ? g_errbuf : "cpusim nccl error"); }
const char* ncclGetLastError(ncclComm_t) { return g_errbuf; }

ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
  static int counter = 0;
  memset(id, 0, sizeof(*id));
  timespec t;
  clock_gettime(CLOCK_REALTIME, &t);
  snprintf(id->internal, sizeof(id->internal), "/cpusim.%d.%d.%lx", (int)getpid(), counter++, (unsigned long)t.tv_nsec);
  int fd = shm_open(id->internal, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0 || ftruncate(fd, sizeof(Segment)) != 0) {
    snprintf(g_errbuf, sizeof(g_errbuf), "cpusim: shm_open(%s): %s", id->internal, strerror(errno));
    if (fd >= 0) close(fd);
    return ncclSystemError;
  }
  close(fd);  // fresh tmpfs pages are zero: heads, tails and counters start at 0
  return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t* out, int nranks, ncclUniqueId id, int rank) {
  if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) {
    snprintf(g_errbuf, sizeof(g_errbuf), "cpusim: %d ranks unsupported (max %d)", nranks, kMaxRanks);
    return ncclInvalidArgument;
  }
  int fd = shm_open(id.internal, O_RDWR, 0600);
  if (fd < 0) {
    snprintf(g_errbuf, sizeof(g_errbuf), "cpusim: shm_open(%s): %s", id.internal, strerror(errno));
    return ncclSystemError;
  }
  void* m = mmap(nullptr, sizeof(Segment), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (m == MAP_FAILED) return ncclSystemError;
  World* w = new World();
  w->seg = static_cast<Segment*>(m);
  w->nranks = nranks;
  w->rank = rank;
  w->refs = 1;
  snprintf(w->name, sizeof(w->name), "%s", id.internal);
  ncclComm* c = new ncclComm();
  c->w = w;
  c->id = mix(0x51D, 1);
  c->rank = rank;
  for (int r = 0; r < nranks; ++r) c->members.push_back(r);
  // everyone has mapped the segment once this exchange completes: the name can go away (no leak if a test dies later)
  std::vector<char> all(nranks);
  char mine = 1;
  allgather_now(c, &mine, all.data(), 1, K_INIT);
  if (w->seg->attached.fetch_add(1) + 1 == nranks) shm_unlink(w->name);
  *out = c;
  return ncclSuccess;
}

ncclResult_t ncclCommSplit(ncclComm_t parent, int color, int key, ncclComm_t* out, ncclConfig_t*) {
  if (g_group_depth) {
    snprintf(g_errbuf, sizeof(g_errbuf), "cpusim: ncclCommSplit inside a group");
    return ncclInvalidUsage;
  }
  const int P = (int)parent->members.size();
  struct CK {
    int color, key;
  } mine = {color, key};
  std::vector<CK> all(P);
  allgather_now(parent, &mine, all.data(), sizeof(CK), K_SPLIT);
  const uint32_t nth = parent->splits++;
  if (color < 0) {  // NCCL_SPLIT_NOCOLOR
    *out = nullptr;
    return ncclSuccess;
  }
  std::vector<int> idx;
  for (int p = 0; p < P; ++p)
    if (all[p].color == color) idx.push_back(p);
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return all[a].key < all[b].key; });
  ncclComm* c = new ncclComm();
  c->w = parent->w;
  c->w->refs++;
  c->id = mix(mix(parent->id, nth), (uint64_t)color + 0x1000);
  c->rank = -1;
  for (size_t i = 0; i < idx.size(); ++i) {
    c->members.push_back(parent->members[idx[i]]);
    if (idx[i] == parent->rank) c->rank = (int)i;
  }
  *out = c;
  return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t c) {
  if (!c) return ncclSuccess;
  world_release(c->w);
  delete c;
  return ncclSuccess;
}
ncclResult_t ncclCommCount(const ncclComm_t c, int* n) {
  *n = (int)c->members.size();
  return ncclSuccess;
}
ncclResult_t ncclCommUserRank(const ncclComm_t c, int* r) {
  *r = c->rank;
  return ncclSuccess;
}

ncclResult_t ncclGroupStart(void) {
  ++g_group_depth;
  return ncclSuccess;
}
ncclResult_t ncclGroupEnd(void) {
  if (g_group_depth <= 0) return ncclInvalidUsage;
  if (--g_group_depth == 0) flush_pending();
  return ncclSuccess;
}

ncclResult_t ncclSend(const void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t st) {
  if (peer < 0 || peer >= (int)c->members.size() || peer == c->rank) {
    snprintf(g_errbuf, sizeof(g_errbuf), "cpusim: ncclSend to invalid peer %d (comm size %zu, my rank %d)", peer, c->members.size(),
             c->rank);
    return ncclInvalidArgument;
  }
  Request* r = new_request(c);
  add_op(r, c, true, peer, buf, count * type_size(t), K_P2P, 0);
  submit(r, st);
  return ncclSuccess;
}
ncclResult_t ncclRecv(void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t st) {
  if (peer < 0 || peer >= (int)c->members.size() || peer == c->rank) {
    snprintf(g_errbuf, sizeof(g_errbuf), "cpusim: ncclRecv from invalid peer %d (comm size %zu, my rank %d)", peer,
             c->members.size(), c->rank);
    return ncclInvalidArgument;
  }
  Request* r = new_request(c);
  add_op(r, c, false, peer, buf, count * type_size(t), K_P2P, 0);
  submit(r, st);
  return ncclSuccess;
}

ncclResult_t ncclBroadcast(const void* send, void* recv, size_t count, ncclDataType_t t, int root, ncclComm_t c, cudaStream_t st) {
  const int P = (int)c->members.size();
  if (root < 0 || root >= P) return ncclInvalidArgument;
  const size_t bytes = count * type_size(t);
  Request* r = new_request(c);
  const uint32_t seq = c->seq++;
  if (c->rank == root) {
    if (send != recv) r->start = [=]() { memmove(recv, send, bytes); };
    for (int p = 0; p < P; ++p)
      if (p != root) add_op(r, c, true, p, recv, bytes, K_BCAST, seq);
  } else {
    add_op(r, c, false, root, recv, bytes, K_BCAST, seq);
  }
  submit(r, st);
  return ncclSuccess;
}

static ncclResult_t reduce_common(const void* send, void* recv, size_t count, ncclDataType_t t, ncclRedOp_t op, ncclComm_t c,
                                  bool scatter, cudaStream_t st) {
  if (t != ncclFloat64 || op != ncclSum) {
    snprintf(g_errbuf, sizeof(g_errbuf), "cpusim: only float64 sums are implemented");
    return ncclInvalidArgument;
  }
  const int P = (int)c->members.size(), me = c->rank;
  const size_t bytes = count * sizeof(double);  // per received contribution
  Request* r = new_request(c);
  const uint32_t seq = c->seq++;
  const uint32_t kind = scatter ? K_REDUCESCATTER : K_ALLREDUCE;
  std::vector<double>* tmp = new std::vector<double>((size_t)P * count);
  const char* s = static_cast<const char*>(send);
  for (int p = 0; p < P; ++p) {
    if (p == me) continue;
    add_op(r, c, true, p, scatter ? s + (size_t)p * bytes : s, bytes, kind, seq);
    add_op(r, c, false, p, tmp->data() + (size_t)p * count, bytes, kind, seq);
  }
  r->finish = [=]() {
    memcpy(tmp->data() + (size_t)me * count, scatter ? s + (size_t)me * bytes : s, bytes);
    double* out = static_cast<double*>(recv);
    for (size_t i = 0; i < count; ++i) {
      double acc = 0.0;
      for (int p = 0; p < P; ++p) acc += (*tmp)[(size_t)p * count + i];
      out[i] = acc;
    }
    delete tmp;
  };
  submit(r, st);
  return ncclSuccess;
}

ncclResult_t ncclAllReduce(const void* send, void* recv, size_t count, ncclDataType_t t, ncclRedOp_t op, ncclComm_t c,
                           cudaStream_t st) {
  return reduce_common(send, recv, count, t, op, c, false, st);
}
ncclResult_t ncclReduceScatter(const void* send, void* recv, size_t recvcount, ncclDataType_t t, ncclRedOp_t op, ncclComm_t c,
                               cudaStream_t st) {
  return reduce_common(send, recv, recvcount, t, op, c, true, st);
}

ncclResult_t ncclAllGather(const void* send, void* recv, size_t sendcount, ncclDataType_t t, ncclComm_t c, cudaStream_t st) {
  const int P = (int)c->members.size(), me = c->rank;
  const size_t bytes = sendcount * type_size(t);
  Request* r = new_request(c);
  const uint32_t seq = c->seq++;
  char* out = static_cast<char*>(recv);
  if (out + (size_t)me * bytes != send) r->start = [=]() { memmove(out + (size_t)me * bytes, send, bytes); };
  for (int p = 0; p < P; ++p) {
    if (p == me) continue;
    add_op(r, c, true, p, out + (size_t)me * bytes, bytes, K_ALLGATHER, seq);
    add_op(r, c, false, p, out + (size_t)p * bytes, bytes, K_ALLGATHER, seq);
  }
  submit(r, st);
  return ncclSuccess;
}

}  // extern "C"
