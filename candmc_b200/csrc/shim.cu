// candmc_b200 — C++ host layer with the reference's own entry-point signatures (include/CANDMC.h and
// include/candmc/*.h), forwarding to the C ABI, plus the MPI subset the reference's CANMM drivers need
// (include/candmc/mpi.h) on top of the NCCL grid communicators.  Error behaviour mirrors the reference: a violated
// precondition or a runtime failure prints a message and aborts the process (ASSERT / ABORT, alg/shared/util.h:127-138).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../../include/CANDMC.h"
#include "comm.h"
#include "runtime.h"
#include "staging.h"

using namespace candmc;

extern "C" void candmc_shim_abort(const char* file, int line) {
  fflush(stdout);
  fprintf(stderr, "candmc_b200: ABORT at %s:%d\n", file, line);
  fflush(stderr);
  _exit(134);
}

extern "C" void candmc_shim_check(int status, const char* what) {
  if (status == CANDMC_OK) return;
  fflush(stdout);
  fprintf(stderr, "candmc_b200: %s failed (%d): %s\n", what, status, candmc_last_error());
  fflush(stderr);
  _exit(134);
}

// ---- reference C++ entry points -------------------------------------------------------------------------------------
namespace {
candmc_ctb_args_t to_c(ctb_args_t const* a) {
  candmc_ctb_args_t c;
  c.trans_A = a->trans_A;
  c.trans_B = a->trans_B;
  c.n = a->n;
  c.lda_A = a->lda_A;
  c.lda_B = a->lda_B;
  c.lda_C = a->lda_C;
  c.buffer_size = a->buffer_size;
  c.ovp = a->ovp;
  return c;
}
}  // namespace

void summa(ctb_args_t const* args, double const* mat_A, double const* mat_B, double* mat_C, double* buffer,
           CommData_t cdt_row, CommData_t cdt_col) {
  candmc_ctb_args_t c = to_c(args);
  candmc_shim_check(candmc_summa(&c, mat_A, mat_B, mat_C, buffer, cdt_row.cm, cdt_col.cm, 0), "summa");
}

void d25_summa(ctb_args_t const* args, double* mat_A, double* mat_B, double* mat_C, double* buffer, CommData_t cdt_row,
               CommData_t cdt_col, CommData_t cdt_kdir) {
  candmc_ctb_args_t c = to_c(args);
  candmc_shim_check(candmc_d25_summa(&c, mat_A, mat_B, mat_C, buffer, cdt_row.cm, cdt_col.cm, cdt_kdir.cm, 0, 0),
                    "d25_summa");
}

void d25_summa_ovp(ctb_args_t const* args, double* mat_A, double* mat_B, double* mat_C, double* buffer,
                   CommData_t cdt_row, CommData_t cdt_col, CommData_t cdt_kdir) {
  candmc_ctb_args_t c = to_c(args);
  candmc_shim_check(candmc_d25_summa(&c, mat_A, mat_B, mat_C, buffer, cdt_row.cm, cdt_col.cm, cdt_kdir.cm, 1, 0),
                    "d25_summa_ovp");
}

void bcast_cannon_4d(ctb_args_t const* args, double* mat_A, double* mat_B, double* mat_C, double* buffer,
                     CommData_t cdt_x1, CommData_t cdt_y1, CommData_t cdt_x2, CommData_t cdt_y2) {
  candmc_ctb_args_t c = to_c(args);
  candmc_shim_check(
      candmc_bcast_cannon_4d(&c, mat_A, mat_B, mat_C, buffer, cdt_x1.cm, cdt_y1.cm, cdt_x2.cm, cdt_y2.cm, 0),
      "bcast_cannon_4d");
}

void kput_cannon(int const rank, int const kary, int const ndim, MPI_Comm const comm, int const n, int const m,
                 int const k, char const transp_A, double const alpha, double* A, char const transp_B,
                 double const beta, double* B, double* C) {
  (void)comm;  // the reference ignores it too and works on MPI_COMM_WORLD (spcannon.cxx:270-273)
  candmc_shim_check(candmc_spcannon(1, rank, kary, ndim, candmc_mpi_comm_world(), n, m, k, transp_A, alpha, A, transp_B,
                                    beta, B, C, 0),
                    "kput_cannon");
}

void kuni_cannon(int const rank, int const kary, int const ndim, MPI_Comm const comm, int const n, int const m,
                 int const k, char const transp_A, double const alpha, double* A, char const transp_B,
                 double const beta, double* B, double* C) {
  (void)comm;
  candmc_shim_check(candmc_spcannon(0, rank, kary, ndim, candmc_mpi_comm_world(), n, m, k, transp_A, alpha, A, transp_B,
                                    beta, B, C, 0),
                    "kuni_cannon");
}

void cdgemm(char transa, char transb, int m, int n, int k, double a, const double* A, int lda, const double* B, int ldb,
            double b, double* C, int ldc) {
  candmc_shim_check(candmc_dgemm(transa, transb, m, n, k, a, A, lda, B, ldb, b, C, ldc, 0), "cdgemm");
}

void csgemm(char transa, char transb, int m, int n, int k, float a, const float* A, int lda, const float* B, int ldb, float b,
            float* C, int ldc) {
  candmc_shim_check(candmc_sgemm(transa, transb, m, n, k, a, A, lda, B, ldb, b, C, ldc, 0), "csgemm");
}

void print_matrix(double const* M, int n, int m) { print_matrix(M, n, m, n); }
void print_matrix(double const* M, int n, int m, int lda) {  // alg/shared/util.cxx print_matrix
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < m; j++) printf("%+.4lf ", M[i + (size_t)j * lda]);
    printf("\n");
  }
}

// ---- the MPI subset (include/candmc/mpi.h) --------------------------------------------------------------------------
namespace {
candmc_comm_t* g_world = nullptr;
int dt_size(MPI_Datatype t) { return t & 0xff; }

void die(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fflush(stdout);
  fprintf(stderr, "candmc_b200 (mpi subset): ");
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
  _exit(134);
}

int env_int(const char* a, const char* b, int dflt) {
  const char* v = getenv(a);
  if (!v && b) v = getenv(b);
  return v ? atoi(v) : dflt;
}

void cuda_ok(cudaError_t e, const char* what) {
  if (e != cudaSuccess) die("%s: %s", what, cudaGetErrorString(e));
}
void nccl_ok(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) die("%s: %s", what, ncclGetErrorString(r));
}

// gather `bytes` from every rank of `c` into a host vector (rank-major)
std::vector<unsigned char> allgather_host(candmc_comm_t* c, const void* mine, size_t bytes) {
  std::vector<unsigned char> out(bytes * c->size);
  if (c->size == 1) {
    memcpy(out.data(), mine, bytes);
    return out;
  }
  cudaStream_t st = runtime().comm_stream;
  unsigned char *ds = nullptr, *dr = nullptr;
  cuda_ok(cudaMalloc(&ds, bytes), "cudaMalloc");
  cuda_ok(cudaMalloc(&dr, bytes * c->size), "cudaMalloc");
  cuda_ok(cudaMemcpyAsync(ds, mine, bytes, cudaMemcpyHostToDevice, st), "H2D");
  nccl_ok(ncclAllGather(ds, dr, bytes, ncclChar, c->nccl, st), "ncclAllGather");
  cuda_ok(cudaMemcpyAsync(out.data(), dr, bytes * c->size, cudaMemcpyDeviceToHost, st), "D2H");
  cuda_ok(cudaStreamSynchronize(st), "sync");
  cudaFree(ds);
  cudaFree(dr);
  return out;
}

template <typename T>
void reduce_host(T* acc, const T* in, int count, MPI_Op op, bool integral) {
  for (int i = 0; i < count; ++i) {
    switch (op) {
      case MPI_SUM: acc[i] = acc[i] + in[i]; break;
      case MPI_MAX: acc[i] = acc[i] > in[i] ? acc[i] : in[i]; break;
      case MPI_MIN: acc[i] = acc[i] < in[i] ? acc[i] : in[i]; break;
      case MPI_BAND:
      case MPI_BOR: {
        if (!integral) die("bitwise reduction on a floating type");
        long long a = (long long)acc[i], b = (long long)in[i];
        acc[i] = (T)(op == MPI_BAND ? (a & b) : (a | b));
        break;
      }
      default: die("unsupported MPI_Op %d", op);
    }
  }
}

void reduce_any(void* acc, const void* in, int count, MPI_Datatype t, MPI_Op op) {
  switch (t) {
    case MPI_DOUBLE: reduce_host((double*)acc, (const double*)in, count, op, false); break;
    case MPI_FLOAT: reduce_host((float*)acc, (const float*)in, count, op, false); break;
    case MPI_INT: reduce_host((int*)acc, (const int*)in, count, op, true); break;
    case MPI_INT64_T:
    case MPI_LONG: reduce_host((int64_t*)acc, (const int64_t*)in, count, op, true); break;
    case MPI_CHAR:
    case MPI_BYTE: reduce_host((char*)acc, (const char*)in, count, op, true); break;
    default: die("unsupported MPI_Datatype 0x%x", t);
  }
}
}  // namespace

extern "C" {

MPI_Comm candmc_mpi_comm_world(void) {
  if (!g_world) die("MPI_COMM_WORLD used before MPI_Init");
  return g_world;
}

int MPI_Init(int* argc, char*** argv) {
  (void)argc;
  (void)argv;
  if (g_world) return MPI_SUCCESS;
  const int rank = env_int("RANK", "CANDMC_RANK", 0);
  const int size = env_int("WORLD_SIZE", "CANDMC_WORLD_SIZE", 1);
  const int local = env_int("LOCAL_RANK", "CANDMC_LOCAL_RANK", rank);
  candmc_shim_check(candmc_init(local), "candmc_init");
  unsigned char id[CANDMC_UNIQUE_ID_BYTES];
  memset(id, 0, sizeof(id));
  if (size > 1) {
    const char* dir = getenv("CANDMC_RENDEZVOUS");
    if (!dir) die("WORLD_SIZE=%d but CANDMC_RENDEZVOUS is not set (launch with tools/candmc_run)", size);
    const std::string path = std::string(dir) + "/nccl_id";
    if (rank == 0) {
      candmc_shim_check(candmc_get_unique_id(id), "candmc_get_unique_id");
      const std::string tmp = path + ".tmp";
      FILE* f = fopen(tmp.c_str(), "wb");
      if (!f || fwrite(id, 1, sizeof(id), f) != sizeof(id)) die("cannot write %s", tmp.c_str());
      fclose(f);
      if (rename(tmp.c_str(), path.c_str()) != 0) die("cannot publish %s", path.c_str());
    } else {
      const time_t t0 = time(nullptr);
      for (;;) {
        FILE* f = fopen(path.c_str(), "rb");
        if (f) {
          const size_t got = fread(id, 1, sizeof(id), f);
          fclose(f);
          if (got == sizeof(id)) break;
        }
        if (time(nullptr) - t0 > 120) die("timed out waiting for %s", path.c_str());
        usleep(2000);
      }
    }
  } else {
    candmc_shim_check(candmc_get_unique_id(id), "candmc_get_unique_id");
  }
  candmc_shim_check(candmc_comm_init_rank(id, size, rank, &g_world), "candmc_comm_init_rank");
  return MPI_SUCCESS;
}

int MPI_Finalize(void) {
  if (!g_world) return MPI_SUCCESS;
  candmc_comm_barrier(g_world);
  candmc_comm_free(g_world);
  g_world = nullptr;
  return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int code) {
  (void)comm;
  fflush(stdout);
  fflush(stderr);
  _exit(code ? (code & 0xff) | 1 : 1);
}

double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int MPI_Comm_size(MPI_Comm comm, int* size) {
  candmc_shim_check(candmc_comm_size(comm, size), "MPI_Comm_size");
  return MPI_SUCCESS;
}
int MPI_Comm_rank(MPI_Comm comm, int* rank) {
  candmc_shim_check(candmc_comm_rank(comm, rank), "MPI_Comm_rank");
  return MPI_SUCCESS;
}
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm* newcomm) {
  candmc_shim_check(candmc_comm_split(comm, color, key, newcomm), "MPI_Comm_split");
  return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm* comm) {
  if (*comm && *comm != g_world) candmc_comm_free(*comm);
  *comm = MPI_COMM_NULL;
  return MPI_SUCCESS;
}
int MPI_Barrier(MPI_Comm comm) {
  candmc_shim_check(candmc_comm_barrier(comm), "MPI_Barrier");
  return MPI_SUCCESS;
}

int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm) {
  const size_t bytes = (size_t)count * dt_size(type);
  if (bytes == 0 || comm->size == 1) return MPI_SUCCESS;
  cudaStream_t st = runtime().comm_stream;
  if (is_device_ptr(buf)) {
    // MPI semantics: the buffer is complete when the call is made.  A device buffer may still be being written by an earlier
    // asynchronous call on some other stream (this library's entry points are asynchronous for device operands), and the
    // broadcast runs on the library's own stream: drain the device first.
    cuda_ok(cudaDeviceSynchronize(), "sync before a broadcast of a device buffer");
    nccl_ok(ncclBroadcast(buf, buf, bytes, ncclChar, root, comm->nccl, st), "ncclBroadcast");
    cuda_ok(cudaStreamSynchronize(st), "sync");
    return MPI_SUCCESS;
  }
  void* d = nullptr;
  cuda_ok(cudaMalloc(&d, bytes), "cudaMalloc");
  if (comm->rank == root) cuda_ok(cudaMemcpyAsync(d, buf, bytes, cudaMemcpyHostToDevice, st), "H2D");
  nccl_ok(ncclBroadcast(d, d, bytes, ncclChar, root, comm->nccl, st), "ncclBroadcast");
  if (comm->rank != root) cuda_ok(cudaMemcpyAsync(buf, d, bytes, cudaMemcpyDeviceToHost, st), "D2H");
  cuda_ok(cudaStreamSynchronize(st), "sync");
  cudaFree(d);
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm) {
  const void* mine = (sendbuf == MPI_IN_PLACE) ? recvbuf : sendbuf;
  if (type == MPI_DOUBLE && op == MPI_SUM) {
    candmc_shim_check(candmc_comm_allreduce_sum(comm, (const double*)mine, (double*)recvbuf, count,
                                                runtime().comm_stream),
                      "MPI_Allreduce");
    cuda_ok(cudaStreamSynchronize(runtime().comm_stream), "sync");
    return MPI_SUCCESS;
  }
  const size_t bytes = (size_t)count * dt_size(type);
  std::vector<unsigned char> all = allgather_host(comm, mine, bytes);
  std::vector<unsigned char> acc(all.begin(), all.begin() + bytes);
  for (int r = 1; r < comm->size; ++r) reduce_any(acc.data(), all.data() + r * bytes, count, type, op);
  memcpy(recvbuf, acc.data(), bytes);
  return MPI_SUCCESS;
}

int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm) {
  const size_t bytes = (size_t)count * dt_size(type);
  const void* mine = (sendbuf == MPI_IN_PLACE) ? recvbuf : sendbuf;
  std::vector<unsigned char> all = allgather_host(comm, mine, bytes);
  if (comm->rank == root) {
    std::vector<unsigned char> acc(all.begin(), all.begin() + bytes);
    for (int r = 1; r < comm->size; ++r) reduce_any(acc.data(), all.data() + r * bytes, count, type, op);
    memcpy(recvbuf, acc.data(), bytes);
  }
  return MPI_SUCCESS;
}

}  // extern "C"
