/* ab2run -- starts the N ranks of a drop-in MPI build on one node, one process per GPU (mpi.h):
 *     ab2run -np N ./athena -i athinput ... domain1/NGrid_x3=N
 * creates the job's shared-memory segment, forks N copies of the program with AB2_MPI_RANK / AB2_MPI_SIZE / AB2_MPI_SHM
 * (and CUDA device = rank, which athena_shim.c reads from AB2_MPI_RANK), waits for them, and returns the first non-zero
 * exit status (the other ranks stop at their next collective, see die_if_aborted). */
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

unsigned long ab2_mpi_shm_bytes(void);

int main(int argc, char **argv)
{
  int np = 1, a = 1, r, fd, status, rc = 0;
  char name[64], buf[32];
  pid_t pid[64];
  if (argc >= 3 && strcmp(argv[1], "-np") == 0) { np = atoi(argv[2]); a = 3; }
  if (np < 1 || np > 64 || a >= argc) { fprintf(stderr, "usage: ab2run -np N program [args...]\n"); return 2; }
  snprintf(name, sizeof name, "/ab2mpi_%d", (int)getpid());
  fd = shm_open(name, O_CREAT|O_RDWR|O_TRUNC, 0600);
  if (fd < 0 || ftruncate(fd, (off_t)ab2_mpi_shm_bytes()) != 0) { perror("ab2run: shm"); return 2; }
  close(fd);                                           /* a fresh segment is zero-filled: all barriers and mailboxes idle */
  setenv("AB2_MPI_SHM", name, 1);
  snprintf(buf, sizeof buf, "%d", np); setenv("AB2_MPI_SIZE", buf, 1);
  for (r=0; r<np; r++) {
    pid[r] = fork();
    if (pid[r] == 0) {
      snprintf(buf, sizeof buf, "%d", r); setenv("AB2_MPI_RANK", buf, 1);
      execvp(argv[a], argv + a);
      perror("ab2run: exec"); _exit(127);
    }
  }
  for (r=0; r<np; r++) {
    pid_t p = wait(&status);
    int code = WIFEXITED(status) ? WEXITSTATUS(status) : 128 + WTERMSIG(status);
    if (p > 0 && code != 0 && rc == 0) {               /* first failure: record it and stop the others */
      int q;
      rc = code;
      for (q=0; q<np; q++) if (pid[q] != p) kill(pid[q], SIGTERM);
    }
  }
  shm_unlink(name);
  return rc;
}
