/* hmc_interpose.c — libthirring_hmc.so: the reference's family-A symbols (hmc.c:105-414) implemented on the
 * B200 through the handle C-ABI.  Plain C host code; see include/thirring_hmc_abi.h for the contract. */
#define _GNU_SOURCE
#include <complex.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/thirring_b200.h"
#include "../../include/thirring_hmc_abi.h"

static struct {
  tb_ctx *ctx;
  int nt, nx, mode, device;
  double *A_flat;      /* last uploaded angles [t][x][2] */
  double *A_tmp;
  double *vin, *vout;  /* contiguous staging [t][x] complex */
  int have_A;
  int mu_frozen;
  double mu_at_first_call;
  int params_set;       /* what the context holds: uploaded again only when the driver's m moved */
  double m_set, mu_set;
  double *p_m, *p_mu;  /* the driver's globals, hmc.c:38,40 */
  long cg_calls, apply_calls, traj_calls;
  /* coarse override (update_gauge as one device-resident trajectory) */
  double *p_g;          /* the driver's global g, hmc.c:39 */
  int *p_mers_i;        /* the driver's Mersenne state, mersenne.h:8-14 */
  double *p_mers_array;
  double (*p_mers_generate)(void);
  double *xi, *mom, *st;
  unsigned int traj_index;
} S;

static void die(const char *what) {
  fprintf(stderr, "libthirring_hmc: %s: %s\n", what, tb_last_error());
  abort(); /* a void ABI has no channel to report device errors */
}

int tb_hmc_configure(int nt, int nx, int mode, int device) {
  if (S.ctx) tb_hmc_shutdown();
  memset(&S, 0, sizeof(S));
  S.nt = nt; S.nx = nx; S.mode = mode; S.device = device;
  if (tb_create(&S.ctx, nt, nx, 1, mode, device) != TB_OK) die("tb_create");
  size_t v = (size_t)nt * nx;
  S.A_flat = malloc(v * 2 * sizeof(double));
  S.A_tmp = malloc(v * 2 * sizeof(double));
  S.vin = malloc(v * 2 * sizeof(double));
  S.vout = malloc(v * 2 * sizeof(double));
  return 0;
}

void tb_hmc_shutdown(void) {
  if (S.ctx) tb_destroy(S.ctx);
  free(S.A_flat); free(S.A_tmp); free(S.vin); free(S.vout); free(S.xi); free(S.mom); free(S.st);
  memset(&S, 0, sizeof(S));
}

long tb_hmc_cg_calls(void) { return S.cg_calls; }
long tb_hmc_apply_calls(void) { return S.apply_calls; }
long tb_hmc_trajectory_calls(void) { return S.traj_calls; }

static void lazy_init(void) {
  if (S.ctx) return;
  const char *nt = getenv("THIRRING_NT"), *nx = getenv("THIRRING_NX");
  const char *mode = getenv("THIRRING_MODE"), *dev = getenv("THIRRING_DEVICE");
  if (!nt || !nx) {
    fprintf(stderr, "libthirring_hmc: lattice size unknown: call tb_hmc_configure() or set THIRRING_NT/THIRRING_NX\n");
    abort();
  }
  tb_hmc_configure(atoi(nt), atoi(nx), (mode && !strcmp(mode, "adjoint")) ? TB_MODE_ADJOINT : TB_MODE_REF_COMPAT,
                   dev ? atoi(dev) : 0);
}

/* m is read at every call, exp(+-mu) is frozen at the first one (hmc.c:124-130) */
static void sync_params(void) {
  if (!S.p_m) {
    S.p_m = (double *)dlsym(RTLD_DEFAULT, "m");
    S.p_mu = (double *)dlsym(RTLD_DEFAULT, "mu");
    if (!S.p_m || !S.p_mu) {
      fprintf(stderr, "libthirring_hmc: the driver's globals `m` and `mu` (hmc.c:38,40) are not visible\n");
      abort();
    }
  }
  if (!S.mu_frozen) { S.mu_at_first_call = *S.p_mu; S.mu_frozen = 1; }
  double m = *S.p_m, mu = S.mu_at_first_call;
  if (S.params_set && m == S.m_set && mu == S.mu_set) return;   /* read at every call, uploaded when it moved */
  if (tb_set_params(S.ctx, &m, &mu, 1) != TB_OK) die("tb_set_params");
  S.params_set = 1; S.m_set = m; S.mu_set = mu;
}

static void sync_gauge(double ***A) {
  double *d = S.A_tmp;
  for (int t = 0; t < S.nt; t++) for (int x = 0; x < S.nx; x++) {
    *d++ = A[t][x][0];
    *d++ = A[t][x][1];
  }
  size_t bytes = (size_t)S.nt * S.nx * 2 * sizeof(double);
  if (S.have_A && memcmp(S.A_tmp, S.A_flat, bytes) == 0) return; /* links already on the device */
  memcpy(S.A_flat, S.A_tmp, bytes);
  if (tb_set_gauge(S.ctx, S.A_flat) != TB_OK) die("tb_set_gauge");
  S.have_A = 1;
}

static const double *gather(_Complex double **v) {
  /* rows handed out by our alloc_vector are contiguous: no copy */
  if (v[S.nt - 1] == v[0] + (size_t)(S.nt - 1) * S.nx) return (const double *)v[0];
  for (int t = 0; t < S.nt; t++) memcpy(S.vin + (size_t)t * S.nx * 2, v[t], (size_t)S.nx * 2 * sizeof(double));
  return S.vin;
}

static double *out_buffer(_Complex double **v) {
  if (v[S.nt - 1] == v[0] + (size_t)(S.nt - 1) * S.nx) return (double *)v[0];
  return S.vout;
}

static void scatter(_Complex double **v, const double *buf) {
  if ((const double *)v[0] == buf) return;
  for (int t = 0; t < S.nt; t++) memcpy(v[t], buf + (size_t)t * S.nx * 2, (size_t)S.nx * 2 * sizeof(double));
}

static void apply(int op, _Complex double **v_in, _Complex double **v_out, double ***A) {
  lazy_init();
  sync_params();
  sync_gauge(A);
  const double *in = gather(v_in);
  double *out = out_buffer(v_out);
  if (tb_apply(S.ctx, op, in, out) != TB_OK) die("tb_apply");
  scatter(v_out, out);
  S.apply_calls++;
}

_Complex double **alloc_vector(void) {
  lazy_init();
  size_t table = ((size_t)S.nt * sizeof(_Complex double *) + 63) & ~(size_t)63;
  char *blk = malloc(table + (size_t)S.nt * S.nx * sizeof(_Complex double));
  _Complex double **v = (_Complex double **)blk;
  for (int t = 0; t < S.nt; t++) v[t] = (_Complex double *)(blk + table) + (size_t)t * S.nx;
  return v;
}

void free_vector(_Complex double **v) { free(v); }

void fm_mul(_Complex double **v_in, _Complex double **v_out, double ***A) { apply(TB_OP_M, v_in, v_out, A); }

void fm_conjugate_mul(_Complex double **v_in, _Complex double **v_out, double ***A) {
  apply(TB_OP_MCONJ, v_in, v_out, A);
}

void fmdm_mul(_Complex double **v_in, _Complex double **v_out, double ***A) { apply(TB_OP_MDM, v_in, v_out, A); }

static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void solve(int with_conj, _Complex double **v_in, _Complex double **v_out, double ***A) {
  lazy_init();
  const int trace = getenv("TB_HMC_TRACE") != NULL;
  const double t0 = trace ? now_ms() : 0.0;
  sync_params();
  const double t1 = trace ? now_ms() : 0.0;
  sync_gauge(A);
  const double t2 = trace ? now_ms() : 0.0;
  const double *in = gather(v_in);
  double *out = out_buffer(v_out);
  int status = 0, iters = 0;
  double rr = 0;
  int rc = with_conj ? tb_invert(S.ctx, in, out, &status, &iters, &rr) : tb_cg(S.ctx, in, out, &status, &iters, &rr);
  if (rc != TB_OK) die("tb_cg");
  S.cg_calls++;
  if (trace)
    fprintf(stderr, "libthirring_hmc: solve %ld: params %.3f ms, gauge %.3f ms, cg %.3f ms (%d iterations)\n", S.cg_calls,
            t1 - t0, t2 - t1, now_ms() - t2, iters);
  if (status == TB_CG_DIVERGED) { /* hmc.c:383-388 */
    printf("Cannot invert fermion matrix\n");
    exit(1);
  }
  scatter(v_out, out);
}

void fmdm_invert_cg(_Complex double **v_in, _Complex double **v_out, double ***A) { solve(0, v_in, v_out, A); }

void fm_invert_cg(_Complex double **v_in, _Complex double **v_out, double ***A) { solve(1, v_in, v_out, A); }

void test_conjugate(double ***A) {
  /* hmc.c:759-789; the random vector must come from the driver's own stochastic_vector so the
   * Mersenne stream stays in step with an un-interposed run */
  static void (*stoch)(_Complex double **) = NULL;
  lazy_init();
  if (!stoch) stoch = (void (*)(_Complex double **))dlsym(RTLD_DEFAULT, "stochastic_vector");
  if (!stoch) { fprintf(stderr, "libthirring_hmc: stochastic_vector (hmc.c:439) not visible\n"); abort(); }
  _Complex double **c = alloc_vector(), **mc = alloc_vector(), **mdc = alloc_vector();
  stoch(c);
  fm_mul(c, mc, A);
  fm_conjugate_mul(c, mdc, A);
  _Complex double cmdc = 0, cmc = 0;
  for (int t = 0; t < S.nt; t++) for (int x = 0; x < S.nx; x++) {
    cmdc += conj(c[t][x]) * mdc[t][x];
    cmc += conj(mc[t][x]) * c[t][x];
  }
  _Complex double cdiff = (S.mode == TB_MODE_ADJOINT) ? cmdc - cmc : cmdc - conj(cmc);
  double diff = creal(cdiff) * creal(cdiff) + cimag(cdiff) * cimag(cdiff);
  if (diff > 0.001) {
    printf("ERROR fm_conjugate_mul is not the conjugate of fm_mul\n");
    printf("Difference = %g\n", diff);
    exit(1);
  }
  free_vector(c); free_vector(mc); free_vector(mdc);
}

/* ---- coarse override: update_gauge (hmc.c:671-746) as ONE device-resident trajectory ------------------------------
 * Exported under a tb_ name here; libthirring_hmc_coarse.so (hmc_coarse.c) binds the reference's symbol to it, so the
 * fine-grained path above stays the default.  The random numbers are the DRIVER's: the Box-Muller fields of
 * random_pseudofermion (hmc.c:418-430), random_momentum (hmc.c:483-499) and stochastic_vector (hmc.c:439-447) and the
 * Metropolis uniform (hmc.c:738) are drawn here from the driver's Mersenne state in the reference's order (nothing
 * else draws in between), so the stream stays in step with an un-interposed run.  The two stdout lines and the
 * verdict are printed in the reference's format (hmc.c:701,735,739,743). */
static double drv_mersenne(void) {   /* the mersenne() macro, mersenne.h:11 */
  return *S.p_mers_i > 0 ? S.p_mers_array[--*S.p_mers_i] : S.p_mers_generate();
}

static void box_muller_pairs(double *out, size_t npairs) {
  for (size_t i = 0; i < npairs; i++) {
    double x1 = drv_mersenne();
    double x2 = drv_mersenne();
    out[2 * i] = sqrt(-2 * log(x1)) * cos(2 * M_PI * x2);
    out[2 * i + 1] = sqrt(-2 * log(x1)) * sin(2 * M_PI * x2);
  }
}

void tb_hmc_update_gauge(double ***A) {
  lazy_init();
  sync_params();
  const size_t v = (size_t)S.nt * S.nx;
  if (!S.p_g) {
    S.p_g = (double *)dlsym(RTLD_DEFAULT, "g");
    S.p_mers_i = (int *)dlsym(RTLD_DEFAULT, "mersenne_i");
    S.p_mers_array = (double *)dlsym(RTLD_DEFAULT, "mersenne_array");
    S.p_mers_generate = (double (*)(void))dlsym(RTLD_DEFAULT, "mersenne_generate");
    if (!S.p_g || !S.p_mers_i || !S.p_mers_array || !S.p_mers_generate) {
      fprintf(stderr, "libthirring_hmc: the driver's `g` (hmc.c:39) or its Mersenne generator is not visible\n");
      abort();
    }
    S.xi = malloc(v * 2 * sizeof(double));
    S.mom = malloc(v * 2 * sizeof(double));
    S.st = malloc(v * 2 * sizeof(double));
  }
  const char *ns = getenv("THIRRING_NSTEPS");   /* hmc.c:708 hard-codes 10 */
  const int nsteps = ns ? atoi(ns) : 10;
  double g = *S.p_g;
  if (tb_hmc_set_coupling(S.ctx, &g, 1) != TB_OK) die("tb_hmc_set_coupling");
  sync_gauge(A);
  box_muller_pairs(S.xi, v);    /* random_pseudofermion: re, im per site */
  box_muller_pairs(S.mom, v);   /* random_momentum: mom[t][x][0], mom[t][x][1] */
  box_muller_pairs(S.st, v);    /* stochastic_vector */
  double obs[10];
  int accepted = 0;
  /* the trajectory needs the uniform before it starts; the reference draws it after the leapfrog, with nothing drawn
   * in between, so the position in the stream is the same */
  double u = drv_mersenne();
  if (tb_hmc_trajectory(S.ctx, nsteps, 1.0, 0, S.traj_index++, S.xi, S.mom, S.st, &u, obs, &accepted, NULL) != TB_OK)
    die("tb_hmc_trajectory");
  S.traj_calls++;
  printf("Start HMC: Sg %g, Smdm %g, Smd %g, Smom %g\n", obs[0], obs[1], obs[2], obs[3]);
  int failed = 0;
  if (tb_hmc_cg_failures(S.ctx, &failed) != TB_OK) die("tb_hmc_cg_failures");
  if (failed & (1 << TB_CG_DIVERGED)) {   /* hmc.c:383-388 */
    printf("Cannot invert fermion matrix\n");
    exit(1);
  }
  printf("HMC End, dS %g, Sg %g, Smdm %g, Smd %g, Sm %g\n", obs[8], obs[4], obs[5], obs[6], obs[7]);
  if (accepted) {
    printf("HMC ACCEPTED\n");
    if (tb_get_gauge(S.ctx, S.A_flat) != TB_OK) die("tb_get_gauge");
    const double *d = S.A_flat;
    for (int t = 0; t < S.nt; t++) for (int x = 0; x < S.nx; x++) {
      A[t][x][0] = *d++;
      A[t][x][1] = *d++;
    }
    S.have_A = 1;   /* the device already holds this field and its links */
  } else {
    printf("HMC REJECTED\n");
  }
}
