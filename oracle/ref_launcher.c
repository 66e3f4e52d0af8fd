/* ref_launcher.c — runs the reference's own driver (hmc_main inside oracle/_ref/libhmcref_*.so) on the CPU,
 * with nothing interposed.  TEST INFRASTRUCTURE ONLY (CPU side of the driver-parity test and CPU baseline).
 *   ref_hmc <libhmcref.so> < parameter */
#include <dlfcn.h>
#include <stdio.h>
#include <time.h>
int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s <libhmcref.so>\n", argv[0]); return 2; }
  void *h = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  int (*hmc_main)(void) = (int (*)(void))dlsym(h, "hmc_main");
  if (!hmc_main) { fprintf(stderr, "hmc_main missing\n"); return 2; }
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  int rc = hmc_main();
  clock_gettime(CLOCK_MONOTONIC, &t1);
  fflush(stdout);
  fprintf(stderr, "hmc_main_seconds=%.6f\n", (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec));
  return rc;
}
