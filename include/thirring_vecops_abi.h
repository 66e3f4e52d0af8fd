/* thirring_vecops_abi.h — the reference's "family B" interface (vec_ops.c, prototypes in Thirring.h:85-95,
 * 102-106) exported by libthirring_vecops.so as a replacement for vec_ops.o (Makefile:18-19).
 *
 * Real FP64 vectors as row-pointer arrays double **v, v[t][x] (vec_ops.c:79-90); ARGUMENT ORDER IS (out, in).
 * The occupation field `int **field`, the mass `m` and chemical potential `mu` are the driver's globals
 * (Thirring.h:63-76), looked up with dlsym at every call; exp(+-mu) is frozen at the first call like the
 * function-local statics of vec_ops.c:98-104.  The Dirac applies and the CG run on the GPU through
 * include/thirring_b200.h (tb_set_occupancy + tb_apply / tb_cg / tb_invert); the element-wise vec_* helpers
 * operate on the caller's host rows in place, exactly as the reference does (they are not on the hot path).
 *
 * Cannot be linked together with libthirring_hmc.so: both families define alloc_vector / free_vector with
 * different types, as in the reference (hmc.c:105 vs vec_ops.c:79).
 */
#ifndef THIRRING_VECOPS_ABI_H
#define THIRRING_VECOPS_ABI_H
#ifdef __cplusplus
extern "C" {
#endif

/* lattice size (compile-time NT/NX of Thirring.h:14-15) and device; or env THIRRING_NT / THIRRING_NX / THIRRING_DEVICE */
int tb_vecops_configure(int nt, int nx, int device);
void tb_vecops_shutdown(void);
long tb_vecops_gpu_calls(void);

#ifndef THIRRING_VECOPS_ABI_NO_PROTOTYPES
double **alloc_vector(void);                                   /* vec_ops.c:79-84 */
void free_vector(double **a);                                  /* vec_ops.c:86-90 */
void vec_neg(double **a);                                      /* vec_ops.c:16 */
void vec_zero(double **a);                                     /* vec_ops.c:21 */
void vec_one(double **a);                                      /* vec_ops.c:26 */
void vec_set(double **a, double d);                            /* vec_ops.c:31 */
void vec_d_mul(double **a, double d);                          /* vec_ops.c:36 */
void vec_assign(double **a, double **b);                       /* vec_ops.c:41 */
void vec_add(double **a, double **b);                          /* vec_ops.c:46 */
void vec_dmul_add(double **a, double **b, double **d, double e); /* vec_ops.c:51: a = b + e d */
double vec_dot(double **a, double **b);                        /* vec_ops.c:56 */
void vec_zero_occupied(double **a);                            /* vec_ops.c:64 */
void vec_print_lat(double **a);                                /* vec_ops.c:69 */
void fM(double **chi, double **psi);                           /* vec_ops.c:96-133  chi = M psi */
void fM_transpose(double **chi, double **psi);                 /* vec_ops.c:135-172 chi = M^T psi */
void cg_MdM(double **inv, double **source);                    /* vec_ops.c:261-307; divergence -> 1e50 fill */
void cg_propagator(double **propagator, double **source);      /* vec_ops.c:311-321 */
/* flat-array family, double[NT*NX] indexed t*NX+x (Thirring.h:102-106; no caller in the reference).
 * fM_occupied is fM_transpose without the mass term; on the GPU through a second context at mass 0. */
double *alloc_field(void);                                     /* vec_ops.c:252-255 */
void fM_occupied(double *chi, double *psi);                    /* vec_ops.c:345-380 */
void fM_occupied_sq(double *chi, double *psi);                 /* vec_ops.c:384-390: fM_occupied twice */
double action(double *psi);                                    /* vec_ops.c:392-397: 0.5 sum psi^2 */
void vec_gaussian(double *a);                                  /* vec_ops.c:399-409: Box-Muller on the driver's mersenne() */
int cg_MdM_occupied(double *psi, double *source);              /* vec_ops.c:413-461: 0 = converged, 1 = not */
#endif

#ifdef __cplusplus
}
#endif
#endif
