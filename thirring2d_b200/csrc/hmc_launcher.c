/* hmc_launcher.c — runs the reference's UNMODIFIED hmc.c driver (built as libhmcref_<NT>x<NX>_*.so with
 * -Dmain=hmc_main) on top of libthirring_hmc.so.  Usage:
 *     hmc_b200 <libhmcref.so> <NT> <NX> <compat|adjoint> [device]  < parameter
 * The launcher links libthirring_hmc.so, so its fm_mul/fm_conjugate_mul/fmdm_invert_cg/... sit earlier in
 * the global symbol scope than the driver's own copies and every PLT call inside hmc.c lands on the GPU. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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
  tb_hmc_configure(nt, nx, mode, dev);
  void *h = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  int (*hmc_main)(void) = (int (*)(void))dlsym(h, "hmc_main");
  if (!hmc_main) { fprintf(stderr, "hmc_main not found: %s\n", dlerror()); return 2; }
  int rc = hmc_main();
  fflush(stdout);
  fprintf(stderr, "hmc_b200: %ld CG solves and %ld Dirac applies served by the GPU library\n",
          tb_hmc_cg_calls(), tb_hmc_apply_calls());
  tb_hmc_shutdown();
  return rc;
}
