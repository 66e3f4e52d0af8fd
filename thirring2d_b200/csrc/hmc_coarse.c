/* hmc_coarse.c — libthirring_hmc_coarse.so: the optional coarse override of SURVEY 8(b).  Binds the reference's
 * update_gauge (hmc.c:671-746) to the device-resident trajectory of libthirring_hmc.so, so the unmodified driver
 * crosses PCIe once per trajectory instead of once per solve.  Load it (RTLD_GLOBAL / LD_PRELOAD / link line) before
 * the reference's own definition; without it every solve and apply is still served one by one. */
#include "../../include/thirring_hmc_abi.h"

void update_gauge(double ***A) { tb_hmc_update_gauge(A); }
