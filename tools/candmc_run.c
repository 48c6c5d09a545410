/* candmc_run — native launcher for programs built on libcandmc_b200.so's MPI subset (include/candmc/mpi.h):
 *     candmc_run -np P [-timeout S] prog [args...]
 * fork+execs P ranks, one per GPU, with RANK / WORLD_SIZE / LOCAL_RANK and a fresh CANDMC_RENDEZVOUS directory.
 * Plays the role of `mpirun -np P` in the reference's scripts (scripts/test_all.sh:10-13).  Exit status 0 iff all
 * ranks exit 0; the first failure (or the timeout) kills the remaining ranks. */
#define _GNU_SOURCE
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

int main(int argc, char** argv) {
  int np = 1, timeout = 0, i = 1;
  while (i < argc && argv[i][0] == '-') {
    if ((!strcmp(argv[i], "-np") || !strcmp(argv[i], "-n")) && i + 1 < argc) { np = atoi(argv[i + 1]); i += 2; }
    else if (!strcmp(argv[i], "-timeout") && i + 1 < argc) { timeout = atoi(argv[i + 1]); i += 2; }
    else break;
  }
  if (i >= argc || np < 1 || np > 64) {
    fprintf(stderr, "usage: candmc_run -np P [-timeout S] prog [args...]\n");
    return 2;
  }
  char dir[] = "/tmp/candmc_rdv_XXXXXX";
  if (!mkdtemp(dir)) { perror("mkdtemp"); return 2; }
  char ssize[16];
  snprintf(ssize, sizeof(ssize), "%d", np);
  pid_t* pids = (pid_t*)calloc((size_t)np, sizeof(pid_t));
  for (int r = 0; r < np; ++r) {
    pid_t p = fork();
    if (p < 0) { perror("fork"); return 2; }
    if (p == 0) {
      char srank[16];
      snprintf(srank, sizeof(srank), "%d", r);
      setenv("RANK", srank, 1);
      setenv("LOCAL_RANK", srank, 1);
      setenv("WORLD_SIZE", ssize, 1);
      setenv("CANDMC_RENDEZVOUS", dir, 1);
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
      for (int r = 0; r < np; ++r) if (pids[r] == p) pids[r] = 0;
      if (code != 0 && status_all == 0) {
        status_all = code;
        for (int r = 0; r < np; ++r) if (pids[r] > 0) kill(pids[r], SIGKILL);
      }
    } else {
      if (timeout > 0 && time(NULL) - t0 > timeout) {
        fprintf(stderr, "candmc_run: timeout after %d s\n", timeout);
        for (int r = 0; r < np; ++r) if (pids[r] > 0) kill(pids[r], SIGKILL);
        status_all = 124;
        timeout = 0;
      }
      struct timespec ts = {0, 2000000};
      nanosleep(&ts, NULL);
    }
  }
  char cmd[128];
  snprintf(cmd, sizeof(cmd), "rm -rf %s", dir);
  if (system(cmd) != 0) { /* best effort */ }
  free(pids);
  return status_all;
}
