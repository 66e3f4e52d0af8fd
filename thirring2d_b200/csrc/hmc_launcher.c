/* hmc_launcher.c — runs the reference's UNMODIFIED hmc.c driver (built as libhmcref_<NT>x<NX>_*.so with
 * -Dmain=hmc_main) on top of libthirring_hmc.so.  Usage:
 *     hmc_b200 <libhmcref.so> <NT> <NX> <compat|adjoint> [device] [coarse]  < parameter
 * With "coarse" (or THIRRING_COARSE=1) libthirring_hmc_coarse.so is loaded in front of the driver, so update_gauge
 * itself is one device-resident trajectory.
 * The launcher links libthirring_hmc.so, so its fm_mul/fm_conjugate_mul/fmdm_invert_cg/... sit earlier in
 * the global symbol scope than the driver's own copies and every PLT call inside hmc.c lands on the GPU. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "../../include/thirring_b200.h"
#include "../../include/thirring_hmc_abi.h"

int main(int argc, char **argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s <libhmcref.so> <NT> <NX> <compat|adjoint> [device]\n", argv[0]);
    return 2;
  }
  int nt = atoi(argv[2]), nx = atoi(argv[3]);
  int mode = strcmp(argv[4], "adjoint") == 0 ? TB_MODE_ADJOINT : TB_MODE_REF_COMPAT;
  int dev = argc > 5 ? atoi(argv[5]) : 0;
  int coarse = (argc > 6 && strcmp(argv[6], "coarse") == 0) || (getenv("THIRRING_COARSE") && atoi(getenv("THIRRING_COARSE")));
  tb_hmc_configure(nt, nx, mode, dev);
  if (coarse) {
    /* next to this executable */
    char path[4096];
    ssize_t n = readlink("/proc/self/exe", path, sizeof(path) - 64);
    if (n <= 0) { fprintf(stderr, "hmc_b200: cannot locate myself\n"); return 2; }
    path[n] = 0;
    char *slash = strrchr(path, '/');
    strcpy(slash ? slash + 1 : path, "libthirring_hmc_coarse.so");
    if (!dlopen(path, RTLD_NOW | RTLD_GLOBAL)) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  }
  void *h = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  int (*hmc_main)(void) = (int (*)(void))dlsym(h, "hmc_main");
  if (!hmc_main) { fprintf(stderr, "hmc_main not found: %s\n", dlerror()); return 2; }
  struct timespec t0, t1;   /* wall clock of the driver alone: CUDA start-up happened in tb_hmc_configure */
  clock_gettime(CLOCK_MONOTONIC, &t0);
  int rc = hmc_main();
  clock_gettime(CLOCK_MONOTONIC, &t1);
  fflush(stdout);
  fprintf(stderr, "hmc_main_seconds=%.6f\n", (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec));
  fprintf(stderr, "hmc_b200: %ld CG solves and %ld Dirac applies served by the GPU library, %ld whole trajectories\n",
          tb_hmc_cg_calls(), tb_hmc_apply_calls(), tb_hmc_trajectory_calls());
  tb_hmc_shutdown();
  return rc;
}
