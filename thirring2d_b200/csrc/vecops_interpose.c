/* vecops_interpose.c — libthirring_vecops.so: the reference's family-B symbols (vec_ops.c) on the B200.
 * Plain C host code over the handle C-ABI; contract in include/thirring_vecops_abi.h. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/thirring_b200.h"
#include "../../include/thirring_vecops_abi.h"

#define VEC_CG_MAX_ITER 10000 /* Thirring.h:43 */

static struct {
  tb_ctx *ctx;
  int nt, nx, device, bc;
  int *field_flat, *field_last;
  double *cin, *cout;   /* complex staging of the flat-array family (imaginary parts stay zero) */
  double *rin, *rout;   /* real staging: one flat vector each */
  int have_field, mu_frozen;
  double mu0, m_last;
  int ***p_field_dummy;
  int ***p_field;       /* &field  (int **field, Thirring.h:76) */
  double *p_m, *p_mu;
  long gpu_calls;
  /* flat-array family (vec_ops.c:345-461): massless operator, see the section at the end of the file */
  tb_ctx *ctx0;
  int *field0_last;
  int have_field0;
  double mu_flat;
  double *inv0;
} S;

static void die(const char *what) {
  fprintf(stderr, "libthirring_vecops: %s: %s\n", what, tb_last_error());
  abort();
}

int tb_vecops_configure(int nt, int nx, int device) {
  if (S.ctx) tb_vecops_shutdown();
  memset(&S, 0, sizeof(S));
  S.nt = nt; S.nx = nx; S.device = device;
  /* the boundary variant is a #define in the reference (Thirring.h:27-29): THIRRING_BC = antisymmetric | symmetric | openx */
  const char *bcs = getenv("THIRRING_BC");
  S.bc = (bcs && !strcmp(bcs, "symmetric")) ? TB_BC_SYMMETRIC : ((bcs && !strcmp(bcs, "openx")) ? TB_BC_OPENX : TB_BC_ANTISYMMETRIC);
  if (tb_create(&S.ctx, nt, nx, 1, TB_MODE_ADJOINT, device) != TB_OK) die("tb_create");
  if (tb_set_cg(S.ctx, 1e-30, VEC_CG_MAX_ITER) != TB_OK) die("tb_set_cg");  /* Thirring.h:42-43 */
  size_t v = (size_t)nt * nx;
  S.field_flat = malloc(v * sizeof(int));
  S.field_last = malloc(v * sizeof(int));
  S.cin = calloc(2 * v, sizeof(double));
  S.cout = calloc(2 * v, sizeof(double));
  S.rin = calloc(v, sizeof(double));
  S.rout = calloc(v, sizeof(double));
  return 0;
}

void tb_vecops_shutdown(void) {
  if (S.ctx) tb_destroy(S.ctx);
  if (S.ctx0) tb_destroy(S.ctx0);
  free(S.field_flat); free(S.field_last); free(S.cin); free(S.cout); free(S.rin); free(S.rout); free(S.field0_last); free(S.inv0);
  memset(&S, 0, sizeof(S));
}

long tb_vecops_gpu_calls(void) { return S.gpu_calls; }

static void lazy_init(void) {
  if (S.ctx) return;
  const char *nt = getenv("THIRRING_NT"), *nx = getenv("THIRRING_NX"), *dev = getenv("THIRRING_DEVICE");
  if (!nt || !nx) {
    fprintf(stderr, "libthirring_vecops: call tb_vecops_configure() or set THIRRING_NT/THIRRING_NX\n");
    abort();
  }
  tb_vecops_configure(atoi(nt), atoi(nx), dev ? atoi(dev) : 0);
}

/* mass, mu and the occupation field are the driver's globals; re-upload only what changed */
static void sync_state(void) {
  lazy_init();
  if (!S.p_m) {
    S.p_m = (double *)dlsym(RTLD_DEFAULT, "m");
    S.p_mu = (double *)dlsym(RTLD_DEFAULT, "mu");
    S.p_field = (int ***)dlsym(RTLD_DEFAULT, "field");
    if (!S.p_m || !S.p_mu || !S.p_field) {
      fprintf(stderr, "libthirring_vecops: the driver's globals m, mu, field (Thirring.h:63-76) are not visible\n");
      abort();
    }
  }
  if (!S.mu_frozen) { S.mu0 = *S.p_mu; S.mu_frozen = 1; }   /* vec_ops.c:98-104 */
  int **field = *S.p_field;
  int *d = S.field_flat;
  for (int t = 0; t < S.nt; t++) for (int x = 0; x < S.nx; x++) *d++ = field[t][x];
  size_t bytes = (size_t)S.nt * S.nx * sizeof(int);
  int changed = !S.have_field || memcmp(S.field_flat, S.field_last, bytes) != 0 || S.m_last != *S.p_m;
  if (changed) {
    double m = *S.p_m, mu = S.mu0;
    if (tb_set_params(S.ctx, &m, &mu, 1) != TB_OK) die("tb_set_params");
    if (tb_set_occupancy_bc(S.ctx, S.field_flat, S.bc, 0) != TB_OK) die("tb_set_occupancy");
    memcpy(S.field_last, S.field_flat, bytes);
    S.m_last = m;
    S.have_field = 1;
  }
}

/* ---- element-wise helpers: host rows, as in the reference ------------------------------------------------ */
#define FORALL for (int t = 0; t < S.nt; t++) for (int x = 0; x < S.nx; x++)
double **alloc_vector(void) {
  lazy_init();
  size_t table = ((size_t)S.nt * sizeof(double *) + 63) & ~(size_t)63;
  char *blk = malloc(table + (size_t)S.nt * S.nx * sizeof(double));
  double **a = (double **)blk;
  for (int t = 0; t < S.nt; t++) a[t] = (double *)(blk + table) + (size_t)t * S.nx;
  return a;
}
void free_vector(double **a) { free(a); }
void vec_neg(double **a) { lazy_init(); FORALL a[t][x] = -a[t][x]; }
void vec_zero(double **a) { lazy_init(); FORALL a[t][x] = 0; }
void vec_one(double **a) { lazy_init(); FORALL a[t][x] = 1; }
void vec_set(double **a, double d) { lazy_init(); FORALL a[t][x] = d; }
void vec_d_mul(double **a, double d) { lazy_init(); FORALL a[t][x] = a[t][x] * d; }
void vec_assign(double **a, double **b) { lazy_init(); FORALL a[t][x] = b[t][x]; }
void vec_add(double **a, double **b) { lazy_init(); FORALL a[t][x] += b[t][x]; }
void vec_dmul_add(double **a, double **b, double **d, double e) { lazy_init(); FORALL a[t][x] = b[t][x] + e * d[t][x]; }
double vec_dot(double **a, double **b) { lazy_init(); double s = 0; FORALL s += a[t][x] * b[t][x]; return s; }
void vec_zero_occupied(double **a) {
  sync_state();
  int **field = *S.p_field;
  FORALL if (field[t][x] > 0) a[t][x] = 0;
}
void vec_print_lat(double **a) {
  lazy_init();
  for (int t = 0; t < S.nt; t++) { for (int x = 0; x < S.nx; x++) printf(" %8.2f ", a[t][x]); printf(" \n"); }
  printf(" \n");
}

/* ---- the hot path: GPU ------------------------------------------------------------------------------------ */
/* real vectors straight through (tb_real.cu): the rows are gathered into / scattered from one flat double[NT*NX] */
static void to_flat(double **v) {
  for (int t = 0; t < S.nt; t++) memcpy(S.rin + (size_t)t * S.nx, v[t], (size_t)S.nx * sizeof(double));
}
static void from_flat(double **v) {
  for (int t = 0; t < S.nt; t++) memcpy(v[t], S.rout + (size_t)t * S.nx, (size_t)S.nx * sizeof(double));
}

static void apply(int op, double **chi, double **psi) {
  sync_state();
  to_flat(psi);
  if (tb_apply_real(S.ctx, op, S.rin, S.rout) != TB_OK) die("tb_apply_real");
  from_flat(chi);
  S.gpu_calls++;
}

void fM(double **chi, double **psi) { apply(TB_OP_M, chi, psi); }
void fM_transpose(double **chi, double **psi) { apply(TB_OP_MDAG, chi, psi); }

static void solve(int propagator, double **inv, double **source) {
  sync_state();
  to_flat(source);
  int status = 0, iters = 0;
  double rr = 0;
  if (tb_cg_real(S.ctx, S.rin, S.rout, propagator, &status, &iters, &rr) != TB_OK) die("tb_cg_real");
  S.gpu_calls++;
  from_flat(inv);   /* a diverged solve comes back filled with 1e50 (vec_ops.c:292-296) */
}

void cg_MdM(double **inv, double **source) { solve(0, inv, source); }
void cg_propagator(double **propagator, double **source) { solve(1, propagator, source); }

/* ---- the flat-array family of vec_ops.c:345-461 (declared Thirring.h:102-106, no caller in the reference) -----------
 * F = fM_occupied is fM_transpose without the mass term (vec_ops.c:345-380): F = T (+) 1 with T the hop matrix between
 * free sites and 1 on the occupied ones.  The GPU context ctx0 holds M' = fM at mass 0 and chemical potential -mu,
 * which is (-T) (+) 1 (transposing the hop matrix flips its sign and exchanges exp(mu) <-> exp(-mu)), so
 *   F v = -M' v on free sites, v on occupied sites;  F F = M' M'  (what TB_OP_MDM and tb_cg apply in REF_COMPAT mode).
 * Vectors are flat double[VOLUME], index t*NX+x. */
static void sync_state0(void) {
  sync_state();   /* globals, frozen mu, S.field_flat = the driver's current field */
  size_t v = (size_t)S.nt * S.nx, bytes = v * sizeof(int);
  if (!S.ctx0) {
    if (tb_create(&S.ctx0, S.nt, S.nx, 1, TB_MODE_REF_COMPAT, S.device) != TB_OK) die("tb_create (flat-array family)");
    if (tb_set_cg(S.ctx0, 1e-30, VEC_CG_MAX_ITER) != TB_OK) die("tb_set_cg");
    S.field0_last = malloc(bytes);
    S.inv0 = calloc(2 * v, sizeof(double));
    S.have_field0 = 0;
  }
  /* fM_occupied never clears its `init` flag (vec_ops.c:347-352): exp(+-mu) follows the driver's current mu */
  if (!S.have_field0 || S.mu_flat != *S.p_mu) {
    double m0 = 0, mu0 = -*S.p_mu;
    if (tb_set_params(S.ctx0, &m0, &mu0, 1) != TB_OK) die("tb_set_params");
    S.mu_flat = *S.p_mu;
    S.have_field0 = 0;
  }
  if (!S.have_field0 || memcmp(S.field_flat, S.field0_last, bytes) != 0) {
    if (tb_set_occupancy(S.ctx0, S.field_flat) != TB_OK) die("tb_set_occupancy");
    memcpy(S.field0_last, S.field_flat, bytes);
    S.have_field0 = 1;
  }
}

/* chi = F psi from M' psi (S.cout holds M' of S.cin) */
static void f_from_mprime(double *chi) {
  size_t v = (size_t)S.nt * S.nx;
  for (size_t i = 0; i < v; i++) chi[i] = S.field_flat[i] == 0 ? -S.cout[2 * i] : S.cin[2 * i];
}

double *alloc_field(void) {   /* vec_ops.c:252-255 sizes it with sizeof(double *): the same 8 bytes per site */
  lazy_init();
  return malloc((size_t)S.nt * S.nx * sizeof(double *));
}

void fM_occupied(double *chi, double *psi) {
  sync_state0();
  size_t v = (size_t)S.nt * S.nx;
  for (size_t i = 0; i < v; i++) S.cin[2 * i] = psi[i];
  if (tb_apply(S.ctx0, TB_OP_M, S.cin, S.cout) != TB_OK) die("tb_apply");
  f_from_mprime(chi);
  S.gpu_calls++;
}

void fM_occupied_sq(double *chi, double *psi) {   /* vec_ops.c:384-390 */
  sync_state0();
  size_t v = (size_t)S.nt * S.nx;
  for (size_t i = 0; i < v; i++) S.cin[2 * i] = psi[i];
  if (tb_apply(S.ctx0, TB_OP_MDM, S.cin, S.cout) != TB_OK) die("tb_apply");
  for (size_t i = 0; i < v; i++) chi[i] = S.cout[2 * i];
  S.gpu_calls++;
}

double action(double *psi) {   /* vec_ops.c:392-397 */
  lazy_init();
  double s = 0;
  size_t v = (size_t)S.nt * S.nx;
  for (size_t i = 0; i < v; i++) s += psi[i] * psi[i];
  return 0.5 * s;
}

/* vec_ops.c:399-409.  mersenne() is a macro over the generator's state (mersenne.h:8-14), which lives in the driver's
 * mersenne_inline.o: the same state is used here, so the driver's random stream stays in step. */
void vec_gaussian(double *a) {
  lazy_init();
  static int *p_i;
  static double *p_arr;
  static double (*p_gen)(void);
  if (!p_gen) {
    p_i = (int *)dlsym(RTLD_DEFAULT, "mersenne_i");
    p_arr = (double *)dlsym(RTLD_DEFAULT, "mersenne_array");
    p_gen = (double (*)(void))dlsym(RTLD_DEFAULT, "mersenne_generate");
    if (!p_i || !p_arr || !p_gen) {
      fprintf(stderr, "libthirring_vecops: the driver's Mersenne generator (mersenne.h:8-14) is not visible\n");
      abort();
    }
  }
#define MERSENNE() (*p_i > 0 ? p_arr[--*p_i] : p_gen())
  int v = S.nt * S.nx;
  for (int t = 0; t < v; t++) {
    double x1 = MERSENNE();
    double x2 = MERSENNE();
    a[t] = sqrt(-2 * log(x1)) * cos(2 * M_PI * x2);
    if (t < v - 1) a[++t] = sqrt(-2 * log(x1)) * sin(2 * M_PI * x2);
  }
#undef MERSENNE
}

/* vec_ops.c:413-461: CG on F F from x0 = 0, then psi = F inv.  Returns 0 when ||r||^2 < CG_ACCURACY, 1 on
 * divergence, NaN or CG_MAX_ITER (psi is left untouched then, as in the reference). */
int cg_MdM_occupied(double *psi, double *source) {
  sync_state0();
  size_t v = (size_t)S.nt * S.nx;
  for (size_t i = 0; i < v; i++) S.cin[2 * i] = source[i];
  int status = 0, iters = 0;
  double rr = 0;
  if (tb_cg(S.ctx0, S.cin, S.inv0, &status, &iters, &rr) != TB_OK) die("tb_cg");
  S.gpu_calls++;
  if (status == TB_CG_ZERO_SOURCE) {
    /* no zero-source exit in vec_ops.c:413-461: rr_old = 0 gives a = 0/0 and the NaN branch returns 1; a source
     * that is only tiny is solved by x = 0 to the requested accuracy */
    double s = 0;
    for (size_t i = 0; i < v; i++) s += source[i] * source[i];
    if (s == 0) return 1;
  } else if (status != TB_CG_CONVERGED) {
    return 1;
  }
  memcpy(S.cin, S.inv0, 2 * v * sizeof(double));
  if (tb_apply(S.ctx0, TB_OP_M, S.cin, S.cout) != TB_OK) die("tb_apply");
  f_from_mprime(psi);
  return 0;
}
