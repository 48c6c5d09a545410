/* mini-MPI launcher — TEST INFRASTRUCTURE ONLY (part of oracle/).
 *   mpirun [-np P] [-timeout SECONDS] [-threads T] prog [args...]
 * Creates the shared-memory segment of rings, then fork+execs P ranks with MINIMPI_{SIZE,RANK,SHM} set.
 * OPENBLAS_NUM_THREADS (and OMP_NUM_THREADS) default to max(1, ncores / P) per rank unless -threads is given or the
 * variable is already set.  Exit status: 0 iff every rank exited 0; on the first failure or on timeout the
 * remaining ranks are killed (MPI_Abort semantics).
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#define RING_BYTES (4u << 20)
#define RING_T_BYTES (128 + RING_BYTES) /* must match ring_t in mpi_shim.c */
#define HDR_BYTES 64

int main(int argc, char** argv) {
  int np = 1, timeout = 0, threads = 0, i = 1;
  while (i < argc && argv[i][0] == '-') {
    if ((!strcmp(argv[i], "-np") || !strcmp(argv[i], "-n")) && i + 1 < argc) {
      np = atoi(argv[i + 1]);
      i += 2;
    } else if (!strcmp(argv[i], "-timeout") && i + 1 < argc) {
      timeout = atoi(argv[i + 1]);
      i += 2;
    } else if (!strcmp(argv[i], "-threads") && i + 1 < argc) {
      threads = atoi(argv[i + 1]);
      i += 2;
    } else {
      break;
    }
  }
  if (i >= argc || np < 1 || np > 64) {
    fprintf(stderr, "usage: mpirun [-np P] [-timeout S] [-threads T] prog [args...]\n");
    return 2;
  }
  char name[64];
  snprintf(name, sizeof(name), "/minimpi_%d_%ld", (int)getpid(), (long)time(NULL));
  int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0) {
    perror("shm_open");
    return 2;
  }
  size_t bytes = HDR_BYTES + (size_t)RING_T_BYTES * np * np;
  if (ftruncate(fd, (off_t)bytes) != 0) {
    perror("ftruncate");
    shm_unlink(name);
    return 2;
  }
  close(fd);

  long ncores = sysconf(_SC_NPROCESSORS_ONLN);
  if (threads <= 0) threads = (int)(ncores / np > 0 ? ncores / np : 1);
  char sthreads[16], ssize[16];
  snprintf(sthreads, sizeof(sthreads), "%d", threads);
  snprintf(ssize, sizeof(ssize), "%d", np);

  pid_t* pids = (pid_t*)calloc((size_t)np, sizeof(pid_t));
  for (int r = 0; r < np; ++r) {
    pid_t p = fork();
    if (p < 0) {
      perror("fork");
      for (int k = 0; k < r; ++k) kill(pids[k], SIGKILL);
      shm_unlink(name);
      return 2;
    }
    if (p == 0) {
      char srank[16];
      snprintf(srank, sizeof(srank), "%d", r);
      setenv("MINIMPI_SIZE", ssize, 1);
      setenv("MINIMPI_RANK", srank, 1);
      setenv("MINIMPI_SHM", name, 1);
      setenv("OPENBLAS_NUM_THREADS", sthreads, 0);
      setenv("OMP_NUM_THREADS", sthreads, 0);
      setenv("LOCAL_RANK", srank, 0); /* programs that drive an accelerator pick device (LOCAL_RANK mod #devices) */
      execvp(argv[i], &argv[i]);
      perror("execvp");
      _exit(127);
    }
    pids[r] = p;
  }

  int alive = np, status_all = 0;
  time_t t0 = time(NULL);
  while (alive > 0) {
    int st;
    pid_t p = waitpid(-1, &st, WNOHANG);
    if (p > 0) {
      alive--;
      int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + WTERMSIG(st);
      for (int r = 0; r < np; ++r)
        if (pids[r] == p) pids[r] = 0;
      if (code != 0 && status_all == 0) {
        status_all = code;
        for (int r = 0; r < np; ++r)
          if (pids[r] > 0) kill(pids[r], SIGKILL);
      }
    } else {
      if (timeout > 0 && time(NULL) - t0 > timeout) {
        fprintf(stderr, "mpirun: timeout after %d s, killing ranks\n", timeout);
        for (int r = 0; r < np; ++r)
          if (pids[r] > 0) kill(pids[r], SIGKILL);
        status_all = 124;
        timeout = 0;
      }
      struct timespec ts = {0, 2000000};
      nanosleep(&ts, NULL);
    }
  }
  shm_unlink(name);
  free(pids);
  return status_all;
}
