/* thirring_oracle.c — CPU restatement of the Thirring2D HMC fermion hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (thirring2d_b200/, include/, the C-ABI
 * libraries) may include, link or execute this file; it is the checker for tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity status: PINNED.  tests/test_oracle_pinned.py checks every function below
 *   (i)  bit-for-bit against the reference itself (/root/reference/hmc.c compiled unmodified into
 *        oracle/_ref/libhmcref_*.so by oracle/build_ref.sh) whenever oracle/_ref is present, and
 *   (ii) against the committed fixtures tests/golden/ (npz) that tests/golden/make_golden.py generated
 *        from those reference builds (the reference ships no golden vectors of its own).
 *
 * Conventions (all from /root/reference/hmc.c):
 *   vectors   complex FP64, flat [t][x], interleaved (re,im)             (hmc.c:105-112 row-pointer arrays)
 *   gauge     real FP64 angles, flat [t][x][dir], dir 0 = t, 1 = x      (hmc.c:47, 889-897)
 *   eta       eta[t][x][0] = +1 for even x, -1 for odd x; eta[..][1]=1  (hmc.c:914-921)
 *   boundary  antiperiodic in BOTH directions: the wrap-around hop enters with the opposite sign
 *             ("if (t2 > t) += else -=", hmc.c:143-148,154-158,165-170,175-180)
 *   mode 0    REF_COMPAT: fm_conjugate_mul is a verbatim copy of fm_mul (hmc.c:188-249)
 *   mode 1    ADJOINT   : fm_conjugate_mul applies the true M^dagger (all hop signs flipped,
 *                         exp(mu) <-> exp(-mu)); equals the one-hunk corrected reference build.
 *
 * Every floating-point expression keeps the reference's evaluation order; compile with
 * -ffp-contract=off (as gcc does for the reference under -std=c99) and results are bit-identical.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_CG_ACCURACY 1e-30      /* hmc.c:34 */
#define ORC_CG_MAX_ITER 100000     /* hmc.c:35 */

enum { ORC_MODE_REF_COMPAT = 0, ORC_MODE_ADJOINT = 1 };
enum { ORC_CG_CONVERGED = 0, ORC_CG_MAXITER = 1, ORC_CG_DIVERGED = 2, ORC_CG_ZERO_SOURCE = 3 };

typedef struct { double re, im; } cplx;

static inline int eta0(int x) { return (x % 2 == 0) ? 1 : -1; }   /* hmc.c:917-921 */

/* one hop term:  +-( 0.5*(cA + s*sA*I) * eta * ex * w ), evaluated left to right as the reference does
 * (hmc.c:144).  conj_link selects (cA - sA*I). */
static inline void hop(cplx *v, double cA, double sA, int conj_link, int eta, double ex, int use_ex,
                       cplx w, int add)
{
  double lr = 0.5 * cA, li = 0.5 * (conj_link ? -sA : sA);
  lr = lr * (double)eta;  li = li * (double)eta;
  if (use_ex) { lr = lr * ex; li = li * ex; }
  double pr = lr * w.re - li * w.im;
  double pi = lr * w.im + li * w.re;
  if (add) { v->re += pr; v->im += pi; } else { v->re -= pr; v->im -= pi; }
}

/* Shared body of fm_mul (hmc.c:132-183) and of the corrected fm_conjugate_mul.
 * dagger = 0: M.   dagger = 1: M^dagger = every "+=" <-> "-=" on the hops and expmu <-> expmmu. */
static void apply(int nt, int nx, double m, double mu, int dagger,
                  const cplx *in, cplx *out, const double *A)
{
  const double expmu = exp(mu), expmmu = exp(-mu);                 /* hmc.c:127-128 */
  const double e_fwd = dagger ? expmmu : expmu;                    /* factor on the +t hop */
  const double e_bwd = dagger ? expmu : expmmu;                    /* factor on the -t hop */
  for (int t = 0; t < nt; t++) for (int x = 0; x < nx; x++) {
    cplx v;
    const cplx c = in[t * nx + x];
    v.re = m * c.re;  v.im = m * c.im;                             /* hmc.c:137 */

    /* positive time direction, hmc.c:140-148 */
    int t2 = (t + 1) % nt;
    double a = A[(t * nx + x) * 2 + 0];
    int add = (t2 > t);
    hop(&v, cos(a), sin(a), 0, eta0(x), e_fwd, 1, in[t2 * nx + x], dagger ? !add : add);

    /* negative time direction, hmc.c:151-159: link and eta taken at (t2,x) */
    t2 = (t - 1 + nt) % nt;
    a = A[(t2 * nx + x) * 2 + 0];
    add = (t2 > t);
    hop(&v, cos(a), sin(a), 1, eta0(x), e_bwd, 1, in[t2 * nx + x], dagger ? !add : add);

    /* positive x direction, hmc.c:162-170 */
    int x2 = (x + 1) % nx;
    a = A[(t * nx + x) * 2 + 1];
    add = (x2 > x);
    hop(&v, cos(a), sin(a), 0, 1, 1.0, 0, in[t * nx + x2], dagger ? !add : add);

    /* negative x direction, hmc.c:172-180: link taken at (t,x2) */
    x2 = (x - 1 + nx) % nx;
    a = A[(t * nx + x2) * 2 + 1];
    add = (x2 > x);
    hop(&v, cos(a), sin(a), 1, 1, 1.0, 0, in[t * nx + x2], dagger ? !add : add);

    out[t * nx + x] = v;                                           /* hmc.c:182 */
  }
}

/* fm_mul, hmc.c:123-184 */
void orc_fm_mul(int nt, int nx, double m, double mu, const double *in, double *out, const double *A)
{ apply(nt, nx, m, mu, 0, (const cplx *)in, (cplx *)out, A); }

/* true adjoint M^dagger (what the corrected fm_conjugate_mul applies) */
void orc_fm_dagger_mul(int nt, int nx, double m, double mu, const double *in, double *out, const double *A)
{ apply(nt, nx, m, mu, 1, (const cplx *)in, (cplx *)out, A); }

/* fm_conjugate_mul, hmc.c:188-249: identical to fm_mul as shipped (mode 0), M^dagger in mode 1 */
void orc_fm_conjugate_mul(int nt, int nx, double m, double mu, int mode,
                          const double *in, double *out, const double *A)
{ apply(nt, nx, m, mu, mode == ORC_MODE_ADJOINT, (const cplx *)in, (cplx *)out, A); }

/* fmdm_invert_cg, hmc.c:341-404.  Returns a status code and, through the optional pointers, the
 * number of loop passes executed (k at exit; 0 for the zero-source early return) and the last ||r||^2.
 * The reference prints "Cannot invert fermion matrix" and exit(1)s on divergence (hmc.c:383-388); the
 * oracle returns ORC_CG_DIVERGED instead so a test can observe it. */
/* Sum of n terms: sequentially in index order as the reference does (tree = 0), or as a balanced pairwise tree
 * (tree = 1).  The tree is NOT the reference's arithmetic: it exists so that a test can attribute the 1-3 iterations
 * by which a parallel reduction moves the ||r||^2 < 1e-30 crossing of a 700+ iteration solve to the summation order
 * alone (same recursion, same per-site arithmetic, only the two dot products of the loop summed differently). */
static double sum_terms(double *t, int n, int tree)
{
  if (!tree) { double s = 0; for (int i = 0; i < n; i++) s += t[i]; return s; }
  for (; n > 1; n = (n + 1) / 2) for (int i = 0; i < n / 2; i++) t[i] = t[i] + t[i + (n + 1) / 2];
  return t[0];
}

static int cg_impl(int nt, int nx, double m, double mu, int mode, const double *b_, double *x_,
                   const double *A, int max_iter, int *iters, double *rr_final, int tree);

int orc_fmdm_invert_cg(int nt, int nx, double m, double mu, int mode, const double *b_, double *x_,
                       const double *A, int max_iter, int *iters, double *rr_final)
{ return cg_impl(nt, nx, m, mu, mode, b_, x_, A, max_iter, iters, rr_final, 0); }

/* the same recursion with the loop's two dot products summed as a pairwise tree (see sum_terms) */
int orc_fmdm_invert_cg_treesum(int nt, int nx, double m, double mu, int mode, const double *b_, double *x_,
                               const double *A, int max_iter, int *iters, double *rr_final)
{ return cg_impl(nt, nx, m, mu, mode, b_, x_, A, max_iter, iters, rr_final, 1); }

static int cg_impl(int nt, int nx, double m, double mu, int mode, const double *b_, double *x_,
                   const double *A, int max_iter, int *iters, double *rr_final, int tree)
{
  const int V = nt * nx;
  const cplx *b = (const cplx *)b_;
  cplx *xo = (cplx *)x_;
  cplx *r = malloc(sizeof(cplx) * V), *p = malloc(sizeof(cplx) * V);
  cplx *Mp = malloc(sizeof(cplx) * V), *MMp = malloc(sizeof(cplx) * V);
  double *term = malloc(sizeof(double) * V);
  double rr = 0, rr_old = 0, rr_init, pMp, a;
  int status = ORC_CG_MAXITER, k = 0;
  if (max_iter <= 0) max_iter = ORC_CG_MAX_ITER;

  for (int i = 0; i < V; i++) {                                    /* hmc.c:350-356 */
    xo[i].re = 0; xo[i].im = 0;
    r[i] = b[i];
    p[i] = r[i];
    rr_old += r[i].re * r[i].re + r[i].im * r[i].im;
  }
  rr_init = rr_old;
  rr = rr_old;
  if (rr_old < ORC_CG_ACCURACY) { status = ORC_CG_ZERO_SOURCE; goto done; }   /* hmc.c:359-361 */

  for (k = 1; k < max_iter; k++) {                                 /* hmc.c:364 */
    apply(nt, nx, m, mu, 0, p, Mp, A);                             /* hmc.c:366 */
    apply(nt, nx, m, mu, mode == ORC_MODE_ADJOINT, Mp, MMp, A);    /* hmc.c:367 */
    for (int i = 0; i < V; i++) term[i] = p[i].re * MMp[i].re + p[i].im * MMp[i].im;   /* hmc.c:369-370 */
    pMp = sum_terms(term, V, tree);
    a = rr_old / pMp;
    for (int i = 0; i < V; i++) { xo[i].re += a * p[i].re; xo[i].im += a * p[i].im; }     /* :372-373 */
    for (int i = 0; i < V; i++) { r[i].re -= a * MMp[i].re; r[i].im -= a * MMp[i].im; }   /* :374-375 */
    for (int i = 0; i < V; i++) term[i] = r[i].re * r[i].re + r[i].im * r[i].im;          /* :377-379 */
    rr = sum_terms(term, V, tree);
    if (rr < ORC_CG_ACCURACY) { status = ORC_CG_CONVERGED; break; }                       /* :381 */
    if (rr / rr_init > 1e10) { status = ORC_CG_DIVERGED; break; }                         /* :383 */
    double beta = rr / rr_old;                                                            /* :390 */
    for (int i = 0; i < V; i++) { p[i].re = r[i].re + beta * p[i].re; p[i].im = r[i].im + beta * p[i].im; }
    rr_old = rr;
  }
done:
  if (iters) *iters = (status == ORC_CG_MAXITER) ? k - 1 : k;   /* loop passes executed */
  if (rr_final) *rr_final = rr;
  free(r); free(p); free(Mp); free(MMp); free(term);
  return status;
}

/* fm_invert_cg, hmc.c:408-414:  x = (M~ M)^-1 M~ v */
int orc_fm_invert_cg(int nt, int nx, double m, double mu, int mode, const double *v, double *x,
                     const double *A, int max_iter, int *iters, double *rr_final)
{
  double *tmp = malloc(sizeof(double) * 2 * nt * nx);
  orc_fm_conjugate_mul(nt, nx, m, mu, mode, v, tmp, A);
  int st = orc_fmdm_invert_cg(nt, nx, m, mu, mode, tmp, x, A, max_iter, iters, rr_final);
  free(tmp);
  return st;
}

/* fermion_matrix, hmc.c:269-310: dense V x V, column-major M[row + col*V], interleaved complex */
void orc_fermion_matrix(int nt, int nx, double m, double mu, const double *A, double *M_)
{
  const int V = nt * nx;
  cplx *M = (cplx *)M_;
  const double expmu = exp(mu), expmmu = exp(-mu);
  memset(M, 0, sizeof(cplx) * (size_t)V * V);
  for (int t = 0; t < nt; t++) for (int x = 0; x < nx; x++) {
    const int row = nx * t + x;
    M[row + (size_t)row * V].re = m;
    int t2 = (t + 1) % nt;
    double a = A[(t * nx + x) * 2], s = (t2 > t) ? 0.5 : -0.5;
    cplx *e = &M[row + (size_t)(nx * t2 + x) * V];
    e->re = s * cos(a) * eta0(x) * expmu;  e->im = s * sin(a) * eta0(x) * expmu;
    t2 = (t - 1 + nt) % nt;
    a = A[(t2 * nx + x) * 2];  s = (t2 > t) ? 0.5 : -0.5;
    e = &M[row + (size_t)(nx * t2 + x) * V];
    e->re = s * cos(a) * eta0(x) * expmmu;  e->im = -(s * sin(a)) * eta0(x) * expmmu;
    int x2 = (x + 1) % nx;
    a = A[(t * nx + x) * 2 + 1];  s = (x2 > x) ? 0.5 : -0.5;
    e = &M[row + (size_t)(nx * t + x2) * V];
    e->re = s * cos(a);  e->im = s * sin(a);
    x2 = (x - 1 + nx) % nx;
    a = A[(t * nx + x2) * 2 + 1];  s = (x2 > x) ? 0.5 : -0.5;
    e = &M[row + (size_t)(nx * t + x2) * V];
    e->re = s * cos(a);  e->im = -(s * sin(a));
  }
}

/* Re<a,b> summed sequentially in (t,x) order, as every action in hmc.c does (e.g. :456-459) */
double orc_re_dot(int n, const double *a, const double *b)
{
  double s = 0;
  for (int i = 0; i < n; i++) s += a[2 * i] * b[2 * i] + a[2 * i + 1] * b[2 * i + 1];
  return s;
}

/* ====================================================================================================
 * Family B: the real, occupation-masked operator of vec_ops.c behind Thirring.h.  Vectors are real FP64 flat
 * [t][x]; field[t][x] != 0 marks an occupied site (identity row, hops into it dropped).  Boundary variants
 * (Thirring.h:27-29): ORC_BC_ANTISYMMETRIC (the default; vec_ops.c:95-172), ORC_BC_SYMMETRIC (periodic in x,
 * vec_ops.c:175-249), ORC_BC_OPENX (the x hops across the boundary are dropped: the driver points xup[NX-1] and xdn[0]
 * at a phantom column whose field is EMPTY, fermionbag.c:713-717,761-765).  t is antiperiodic in all three.
 * Pinned bit-for-bit against oracle/_ref/libvecopsref_*.so (tests/test_oracle_pinned.py).
 * ==================================================================================================== */
#define ORC_B_CG_MAX_ITER 10000   /* Thirring.h:43 */

/* fM (transpose = 0, vec_ops.c:96-133) and fM_transpose (transpose = 1, vec_ops.c:135-172) */
enum { ORC_BC_ANTISYMMETRIC = 0, ORC_BC_SYMMETRIC = 1, ORC_BC_OPENX = 2 };
static int orc_bc = ORC_BC_ANTISYMMETRIC;   /* like the reference's compile-time choice: set once per run */
void orc_set_boundary(int bc) { orc_bc = bc; }

static void apply_b(int nt, int nx, double m, double mu, int transpose, const int *field,
                    const double *psi, double *chi)
{
  const int bc = orc_bc;
  const double expmu = exp(mu), expmmu = exp(-mu);                /* vec_ops.c:101-102 */
  const double e_up = transpose ? expmmu : expmu, e_dn = transpose ? expmu : expmmu;
  for (int t = 0; t < nt; t++) for (int x = 0; x < nx; x++) {
    const int k = t * nx + x;
    double c = 0;
    if (field[k] == 0) {
      c = m * psi[k];
      int t2 = (t + 1) % nt;
      if (field[t2 * nx + x] == 0) {
        double h = 0.5 * eta0(x) * e_up * psi[t2 * nx + x];
        if ((t2 > t) != transpose) c += h; else c -= h;
      }
      t2 = (t - 1 + nt) % nt;
      if (field[t2 * nx + x] == 0) {
        double h = 0.5 * eta0(x) * e_dn * psi[t2 * nx + x];
        if ((t2 > t) != transpose) c += h; else c -= h;
      }
      int x2 = (x + 1) % nx;
      if (field[t * nx + x2] == 0 && !(bc == ORC_BC_OPENX && x2 < x)) {
        double h = 0.5 * 1 * psi[t * nx + x2];
        const int plus = bc == ORC_BC_SYMMETRIC ? 1 : (x2 > x);       /* vec_ops.c:203 vs :123-124 */
        if (plus != transpose) c += h; else c -= h;
      }
      x2 = (x - 1 + nx) % nx;
      if (field[t * nx + x2] == 0 && !(bc == ORC_BC_OPENX && x2 > x)) {
        double h = 0.5 * 1 * psi[t * nx + x2];
        const int plus = bc == ORC_BC_SYMMETRIC ? 0 : (x2 > x);       /* vec_ops.c:207 vs :128-129 */
        if (plus != transpose) c += h; else c -= h;
      }
    } else {
      c = psi[k];                                                   /* vec_ops.c:130 */
    }
    chi[k] = c;
  }
}

void orc_fM(int nt, int nx, double m, double mu, const int *field, const double *psi, double *chi)
{ apply_b(nt, nx, m, mu, 0, field, psi, chi); }

void orc_fM_transpose(int nt, int nx, double m, double mu, const int *field, const double *psi, double *chi)
{ apply_b(nt, nx, m, mu, 1, field, psi, chi); }

static double dot_b(int n, const double *a, const double *b)      /* vec_dot, vec_ops.c:56-62 */
{ double s = 0; for (int i = 0; i < n; i++) s += a[i] * b[i]; return s; }

/* cg_MdM, vec_ops.c:261-307.  Divergence fills the solution with 1e50 (vec_ops.c:292-296). */
int orc_cg_MdM(int nt, int nx, double m, double mu, const int *field, const double *source, double *inv,
               int *iters, double *rr_final)
{
  const int V = nt * nx;
  double *r = malloc(sizeof(double) * V), *p = malloc(sizeof(double) * V);
  double *Mp = malloc(sizeof(double) * V), *MMp = malloc(sizeof(double) * V);
  int status = ORC_CG_MAXITER, k = 0;
  for (int i = 0; i < V; i++) { inv[i] = 0; r[i] = source[i]; p[i] = r[i]; }
  double rr_old = dot_b(V, r, r), rr_init = rr_old, rr = rr_old;
  if (rr_old < ORC_CG_ACCURACY) { status = ORC_CG_ZERO_SOURCE; goto done; }
  for (k = 1; k < ORC_B_CG_MAX_ITER; k++) {
    apply_b(nt, nx, m, mu, 0, field, p, Mp);
    apply_b(nt, nx, m, mu, 1, field, Mp, MMp);
    double pMp = dot_b(V, p, MMp);
    double a = rr_old / pMp;
    for (int i = 0; i < V; i++) inv[i] = inv[i] + a * p[i];          /* vec_dmul_add, vec_ops.c:286 */
    for (int i = 0; i < V; i++) r[i] = r[i] + (-a) * MMp[i];         /* vec_ops.c:287 */
    rr = dot_b(V, r, r);
    if (rr < ORC_CG_ACCURACY) { status = ORC_CG_CONVERGED; break; }
    if (rr / rr_init > 1e10) { for (int i = 0; i < V; i++) inv[i] = 1e50; status = ORC_CG_DIVERGED; break; }
    double b = rr / rr_old;
    for (int i = 0; i < V; i++) p[i] = r[i] + b * p[i];
    rr_old = rr;
  }
done:
  if (iters) *iters = (status == ORC_CG_MAXITER) ? k - 1 : k;
  if (rr_final) *rr_final = rr;
  free(r); free(p); free(Mp); free(MMp);
  return status;
}

/* ---- flat-array family, vec_ops.c:345-461 (declared Thirring.h:102-106; no caller in the reference) ------------- */

/* fM_occupied, vec_ops.c:345-380: the hop terms of fM_transpose without the mass term (chi starts from 0 and every
 * hop is 0.5*exp*eta*psi; the factors 0.5 and eta = +-1 are exact, so the product order of vec_ops.c:355 and
 * vec_ops.c:150 gives the same double) */
void orc_fM_occupied(int nt, int nx, double mu, const int *field, const double *psi, double *chi)
{ apply_b(nt, nx, 0.0, mu, 1, field, psi, chi); }

/* fM_occupied_sq, vec_ops.c:384-390 */
void orc_fM_occupied_sq(int nt, int nx, double mu, const int *field, const double *psi, double *chi)
{
  double *tmp = malloc(sizeof(double) * nt * nx);
  orc_fM_occupied(nt, nx, mu, field, psi, tmp);
  orc_fM_occupied(nt, nx, mu, field, tmp, chi);
  free(tmp);
}

/* action, vec_ops.c:392-397 */
double orc_action(int n, const double *psi)
{ double s = 0; for (int i = 0; i < n; i++) s += psi[i] * psi[i]; return 0.5 * s; }

/* cg_MdM_occupied, vec_ops.c:413-461: CG on fM_occupied_sq from inv = 0, then psi = fM_occupied(inv).
 * Returns 0 (converged; psi written) or 1 (divergence, NaN, or CG_MAX_ITER passes; psi untouched). */
int orc_cg_MdM_occupied(int nt, int nx, double mu, const int *field, const double *source, double *psi, int *iters)
{
  const int V = nt * nx;
  double *r = malloc(sizeof(double) * V), *p = malloc(sizeof(double) * V);
  double *MMp = malloc(sizeof(double) * V), *inv = malloc(sizeof(double) * V);
  int ret = 1, k;
  for (int i = 0; i < V; i++) inv[i] = 0;
  for (int i = 0; i < V; i++) p[i] = r[i] = source[i];
  double rr_old = 0;
  for (int i = 0; i < V; i++) rr_old += r[i] * r[i];
  const double rr_init = rr_old;
  for (k = 1; k < ORC_B_CG_MAX_ITER; k++) {
    orc_fM_occupied_sq(nt, nx, mu, field, p, MMp);
    double pMp = 0;
    for (int i = 0; i < V; i++) pMp += p[i] * MMp[i];
    const double a = rr_old / pMp;
    for (int i = 0; i < V; i++) { inv[i] = inv[i] + a * p[i]; r[i] = r[i] - a * MMp[i]; }
    double rr = 0;
    for (int i = 0; i < V; i++) rr += r[i] * r[i];
    if (rr < ORC_CG_ACCURACY) { orc_fM_occupied(nt, nx, mu, field, inv, psi); ret = 0; break; }
    if (rr / rr_init > 1e10 || isnan(rr)) break;                       /* vec_ops.c:448-451 */
    const double b = rr / rr_old;
    for (int i = 0; i < V; i++) p[i] = r[i] + b * p[i];
    rr_old = rr;
  }
  if (iters) *iters = k < ORC_B_CG_MAX_ITER ? k : k - 1;
  free(r); free(p); free(MMp); free(inv);
  return ret;
}

/* cg_propagator, vec_ops.c:311-321:  M^-1 source = (M^T M)^-1 M^T source */
int orc_cg_propagator(int nt, int nx, double m, double mu, const int *field, const double *source, double *prop,
                      int *iters, double *rr_final)
{
  double *tmp = malloc(sizeof(double) * nt * nx);
  apply_b(nt, nx, m, mu, 1, field, source, tmp);
  int st = orc_cg_MdM(nt, nx, m, mu, field, tmp, prop, iters, rr_final);
  free(tmp);
  return st;
}
