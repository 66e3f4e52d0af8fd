/* thirring_hmc_abi.h — the reference's own function signatures for the HMC fermion solve ("family A"),
 * exported by libthirring_hmc.so and backed by the sm_100a kernels through include/thirring_b200.h.
 *
 * rantahar/Thirring2D has no plugin or FFI layer: hmc.c is one self-contained translation unit that
 * defines and calls these functions directly.  The drop-in seam is ELF symbol interposition: the
 * unmodified hmc.c is built as a shared object (-fPIC -shared -Dmain=hmc_main), its
 * internal calls go through the PLT, and a library that is earlier in the global lookup scope and
 * defines the same names replaces them.  Names, argument order ((in, out, A)), void returns and error
 * behaviour are the reference's.
 *
 * Vectors are the reference's row-pointer arrays: v[t] points to NX _Complex double (hmc.c:105-112).
 * The gauge field is double ***A with A[t][x][dir] (hmc.c:47,889-897).  The mass m and chemical
 * potential mu are read from the driver's globals `m` and `mu` (hmc.c:38,40), found with dlsym; exp(+-mu)
 * is frozen at the first apply, as the function-local statics of hmc.c:124-130 do.
 */
#ifndef THIRRING_HMC_ABI_H
#define THIRRING_HMC_ABI_H

#ifdef __cplusplus
extern "C" {
#endif

/* Must be called once before the first use (the reference fixes NT/NX at compile time, hmc.c:14-15).
 * mode is TB_MODE_REF_COMPAT or TB_MODE_ADJOINT.  Without it the environment variables
 * THIRRING_NT, THIRRING_NX, THIRRING_MODE (compat|adjoint) and THIRRING_DEVICE are consulted. */
int tb_hmc_configure(int nt, int nx, int mode, int device);
void tb_hmc_shutdown(void);
/* number of CG solves / Dirac applies served since configure (for tests) */
long tb_hmc_cg_calls(void);
long tb_hmc_apply_calls(void);
long tb_hmc_trajectory_calls(void);

/* update_gauge (hmc.c:671-746) as ONE device-resident trajectory of the context (tb_hmc_trajectory), drawing the
 * random fields and the Metropolis uniform from the driver's own Mersenne state in the reference's order and printing
 * the reference's stdout lines (hmc.c:701,735,739,743).  The leapfrog has 10 steps like hmc.c:708 unless
 * THIRRING_NSTEPS is set.  libthirring_hmc_coarse.so exports the reference's symbol `update_gauge` bound to this
 * function: the optional coarse override of SURVEY 8(b); without that library the driver's own update_gauge runs and
 * only its solves and applies are interposed. */
void tb_hmc_update_gauge(double ***A);

#ifndef THIRRING_HMC_ABI_NO_PROTOTYPES
/* replaces hmc.c:105-112 — one block holds the row table and the rows, so the stray free(tmp) of
 * hmc.c:434 releases everything */
_Complex double **alloc_vector(void);
/* replaces hmc.c:115-120 */
void free_vector(_Complex double **v);
/* replaces hmc.c:123-184: v_out = M v_in */
void fm_mul(_Complex double **v_in, _Complex double **v_out, double ***A);
/* replaces hmc.c:188-249: v_out = M~ v_in (M in REF_COMPAT, M^dagger in ADJOINT) */
void fm_conjugate_mul(_Complex double **v_in, _Complex double **v_out, double ***A);
/* replaces hmc.c:259-264 with what it intends: v_out = M~ (M v_in) */
void fmdm_mul(_Complex double **v_in, _Complex double **v_out, double ***A);
/* replaces hmc.c:341-404: CG for (M~ M) v_out = v_in; on divergence prints "Cannot invert fermion matrix"
 * and exit(1)s exactly like hmc.c:383-388 */
void fmdm_invert_cg(_Complex double **v_in, _Complex double **v_out, double ***A);
/* replaces hmc.c:408-414: v_out = (M~ M)^-1 M~ v_in */
void fm_invert_cg(_Complex double **v_in, _Complex double **v_out, double ***A);
/* replaces hmc.c:759-789.  REF_COMPAT: the check as coded (cmdc - conj(cmc)).  ADJOINT: the correct
 * identity <c, M^dagger c> = <M c, c>; the as-coded one would abort the run. */
void test_conjugate(double ***A);
#endif

#ifdef __cplusplus
}
#endif
#endif
