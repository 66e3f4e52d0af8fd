/* thirring_b200.h — C-ABI of the B200-native Thirring2D fermion hot path.
 *
 * Scope: the staggered Dirac apply with U(1) auxiliary-field links (M, M^dagger, M~M), the CG inverter on
 * M~M and the BLAS-1 reductions it needs, FP64, batched over independent Markov chains / sources.
 * Each entry point cites the reference interface it replaces (file:line under rantahar/Thirring2D).
 * The reference-signature symbols themselves (fm_mul, fm_conjugate_mul, fmdm_invert_cg, ...) live in
 * thirring_hmc_abi.h and are implemented on top of this handle API.
 *
 * Plain C: pointers and sizes only.  Every function returns TB_OK (0) or a negative TB_E* code;
 * tb_last_error() gives the message.  There is NO CPU fallback: without a CUDA device tb_create fails.
 *
 * Host ("canonical") layouts — what the reference's row-pointer arrays flatten to:
 *   vector  double[nchains][NT][NX][2]   (re,im) of the reference's _Complex double v[t][x]  (hmc.c:105-112)
 *   gauge   double[nchains][NT][NX][2]   A[t][x][dir], dir 0 = t-link, 1 = x-link            (hmc.c:47,889-897)
 * Device layout (tb_*_dev entry points): site-major, chain-minor structure of arrays,
 *   vector  double2[NT][NX][nchains]     so neighbouring lanes are neighbouring chains of the same site.
 */
#ifndef THIRRING_B200_H
#define THIRRING_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tb_ctx tb_ctx;

/* operator modes: what "fm_conjugate_mul" means */
#define TB_MODE_REF_COMPAT 0 /* M~ == M, as shipped (hmc.c:188-249 is a copy of hmc.c:123-184) */
#define TB_MODE_ADJOINT    1 /* M~ == M^dagger, the mathematically intended operator            */

/* operators for tb_apply* */
#define TB_OP_M     0 /* fm_mul            hmc.c:123-184 */
#define TB_OP_MDAG  1 /* true adjoint of fm_mul (all hop signs flipped, e^{mu} <-> e^{-mu}) */
#define TB_OP_MCONJ 2 /* fm_conjugate_mul  hmc.c:188-249: M in REF_COMPAT, M^dagger in ADJOINT */
#define TB_OP_MDM   3 /* M~ (M v): what fmdm_mul (hmc.c:259-264) intends; the CG matrix */

/* return codes */
#define TB_OK          0
#define TB_EINVAL     -1
#define TB_ECUDA      -2
#define TB_ENODEVICE  -3
#define TB_ENOMEM     -4

/* per-chain CG status (hmc.c:341-404) */
#define TB_CG_CONVERGED   0 /* rr < accuracy                      hmc.c:381 */
#define TB_CG_MAXITER     1 /* loop ran to CG_MAX_ITER-1 passes   hmc.c:364 (silent in the reference) */
#define TB_CG_DIVERGED    2 /* rr/rr_init > 1e10                  hmc.c:383-388 (reference: print + exit(1)) */
#define TB_CG_ZERO_SOURCE 3 /* ||b||^2 < accuracy, x = 0 returned hmc.c:359-361 */

const char *tb_last_error(void);
int tb_device_count(void);

/* Create a context for nchains independent NT x NX lattices on CUDA device `device`.
 * Replaces the compile-time NT/NX (hmc.c:14-15) and the file-scope globals (hmc.c:38-50). */
int tb_create(tb_ctx **out, int nt, int nx, int nchains, int mode, int device);
int tb_destroy(tb_ctx *ctx);

/* Use an existing CUDA stream (cudaStream_t passed as void*), e.g. torch's current stream. */
int tb_set_stream(tb_ctx *ctx, void *cuda_stream);
int tb_synchronize(tb_ctx *ctx);

/* Mass m and chemical potential mu (globals hmc.c:38,40).  n == 1 broadcasts one value to every chain,
 * n == nchains sets them per chain.  exp(+-mu) is evaluated on the host with libm, as hmc.c:127-128 does. */
int tb_set_params(tb_ctx *ctx, const double *m, const double *mu, int n);

/* CG stopping rule: absolute ||r||^2 < accuracy (CG_ACCURACY 1e-30, hmc.c:34) and at most max_iter-1
 * passes (CG_MAX_ITER 100000, hmc.c:35).  Defaults are the reference's. */
int tb_set_cg(tb_ctx *ctx, double accuracy, int max_iter);

/* Tuning knobs (0 = automatic).  rows_per_thread: t-rows marched per thread by the streaming stencil
 * (1,2,4,8,16); for the resident solver the same field selects the site tile per thread (44 = 4x4,
 * 18 = 1x8, 28 = 2x8, 24 = 2x4).  iters_per_launch: CG iterations per CUDA-graph launch (streaming).
 * solver: 0 auto (on-chip when the lattice has a resident or cluster shape, see tb_solver_info), 1 streaming
 * multi-kernel (3 fused kernels per iteration when M~ = M^dagger, else 4), 2 on-chip or fail, 3 streaming with
 * the 4-kernel iteration always, 5 the strict solver: fmdm_invert_cg in the reference's own floating-point evaluation
 * order (no FMA contraction, every lattice sum accumulated sequentially in (t, x) order as hmc.c:354-379 does) -- slow,
 * for parity work: it lands on the reference's iteration count where the fast solvers' tree sums move the
 * ||r||^2 < 1e-30 crossing of a 700+ iteration solve by 1-3 iterations. */
int tb_set_tuning(tb_ctx *ctx, int rows_per_thread, int iters_per_launch, int solver);

/* Which CG solver the context will use with the current tuning: kind 0 = streaming kernels, 1 = on-chip, one CTA
 * per chain (16^2/32^2/64^2), 2 = on-chip, one thread-block cluster per chain (lattices of CS * 4096/NX rows by
 * NX in {64,128,256} sites, 2 <= CS <= 16: 128^2, 256^2, ...).  chains_in_flight: chains the on-chip solver
 * runs concurrently on this device (0 for streaming).  Either pointer may be NULL. */
int tb_solver_info(tb_ctx *ctx, int *kind, int *chains_in_flight);

/* Which streaming kernels an apply / a CG iteration of the streaming solver runs on a family-A field with the current
 * tuning and environment: kernels 0 = register-marching, 1 = TMA-staged, every chain of the batch in a tile (bulk
 * copies, batches of up to 16 chains), 2 = TMA-staged, tiles of 16 chains x 16 sites moved as 2-D boxes through tensor
 * maps (batches of a multiple of 16 chains).  tile_chains x tile_sites threads per block, rows_per_block rows of the
 * lattice per block.  Any pointer may be NULL. */
int tb_streaming_info(tb_ctx *ctx, int *kernels, int *tile_chains, int *tile_sites, int *rows_per_block);

/* Diagnostics, host only (no GPU involved): the schedule of a planned launch of the on-chip solvers.  A batch that
 * fills its last wave of SMs (or of co-resident clusters) badly is cut into one equal share of CG iterations per
 * machine; est[c] / status[c] are the iteration count and the TB_CG_* status of chain c in the previous solve.
 * Output: segs4[4*s .. 4*s+3] = (chain, first iteration, end iteration, 0) for up to nchains + machines - 1
 * segments; machine b runs segments [seg_lo[b], seg_hi[b]) in order.  first iteration > 1: the tail of a split
 * chain (it waits for the head, which is the FIRST segment of a machine with a lower index); end iteration
 * 0x7fffffff: until the chain ends.  The same code runs inside the library's plan_kernel. */
int tb_plan_schedule(const int *est, const int *status, int nchains, int machines, int *segs4, int *seg_lo,
                     int *seg_hi);

/* ---- host-buffer entry points (copies inside; this is what the reference-facing shim calls) ---------- */

/* Upload angles A and build the link fields W_mu = s * 1/2 * eta_mu * exp(iA_mu) once (replaces the
 * sin/cos re-evaluated inside every apply, hmc.c:140-141,152-153,163-164,173-174). */
int tb_set_gauge(tb_ctx *ctx, const double *A_host);

/* One gauge field for EVERY chain of the context: the chains are then the right-hand sides of a multi-RHS solve on that
 * field -- the N_src random sources of fermion_phase (hmc.c:794-815), the point sources of a propagator.  A_host:
 * double[NT][NX][2].  Every kernel works as after tb_set_gauge with nchains copies of the field; the TMA-staged streaming
 * kernels read the links once per site instead of once per site and chain (32 + 32 / N_src bytes per site
 * and apply; a CG iteration moves 180 instead of 240 bytes per site and source).  Single-GPU contexts. */
int tb_set_gauge_shared(tb_ctx *ctx, const double *A_host);
int tb_set_gauge_shared_dev(tb_ctx *ctx, const double *d_A_one_field);

/* Links from cos / sin evaluated by the caller's libm: trig_t_host / trig_x_host = double[nchains][NT][NX][2] holding
 * (cos A_t, sin A_t) and (cos A_x, sin A_x).  The links are then bit for bit what hmc.c:140-174 computes on the host
 * (the device's sincos may differ from glibc's in the last bit), which together with the strict solver
 * (tb_set_tuning solver = 5) reproduces fmdm_invert_cg's recursion exactly.  The stored angles are not updated. */
int tb_set_links_trig(tb_ctx *ctx, const double *trig_t_host, const double *trig_x_host);

/* Family B (vec_ops.c behind Thirring.h): occupation field int[nchains][NT][NX], 0 = free site (vec_ops.c:107).
 * Replaces the gauge field: the links become the real constants s*1/2*eta masked by the field (a hop into or out
 * of an occupied site is dropped, vec_ops.c:110-128) and occupied sites become identity rows (vec_ops.c:130).
 * With it, TB_OP_M is fM (vec_ops.c:96), TB_OP_MDAG is fM_transpose (vec_ops.c:135), tb_cg is cg_MdM
 * (vec_ops.c:261) and tb_invert is cg_propagator (vec_ops.c:311) on vectors whose imaginary parts are zero.
 * The masses of tb_set_params are baked into the per-site mass field; a later tb_set_params re-bakes it from the
 * stored occupation field.  tb_set_gauge* switches the context back to family A. */
int tb_set_occupancy(tb_ctx *ctx, const int *field_host);

/* Boundary variants of family B, a compile-time choice in the reference (Thirring.h:27-29): t is antiperiodic in all
 * three; in x ANTISYMMETRIC (the default, vec_ops.c:95-172), SYMMETRIC (periodic, vec_ops.c:175-249) or OPENX (no hop
 * across the x boundary: the driver's phantom column, fermionbag.c:713-717,761-765). */
#define TB_BC_ANTISYMMETRIC 0
#define TB_BC_SYMMETRIC     1
#define TB_BC_OPENX         2

/* tb_set_occupancy with a boundary variant; shared != 0: field_host is ONE field int[NT][NX] used by every source of
 * the batch (multi-RHS: the 2 NX point sources of measure_propagator, fermionbag.c:389-435, the V/2 of calc_Dinv_cg,
 * fluctuation_determinant.c:973-1003). */
int tb_set_occupancy_bc(tb_ctx *ctx, const int *field_host, int bc, int shared);

/* Family B on REAL vectors, host layout double[nchains][NT][NX] (what the reference's double** rows flatten to):
 * tb_apply_real: TB_OP_M = fM (vec_ops.c:96), TB_OP_MDAG = fM_transpose (vec_ops.c:135).
 * tb_cg_real: cg_MdM (propagator = 0, vec_ops.c:261) or cg_propagator (1, vec_ops.c:311) for every source; a source
 * whose solve diverges is returned filled with 1e50 (vec_ops.c:292-296).  8 bytes per site cross PCIe, and on
 * lattices of whole 8-row tiles with at most 4096 sites (64 x 64, the size Thirring.h compiles in) the solve runs as
 * one real-arithmetic on-chip kernel per sub-batch of sources (tb_real.cu). */
int tb_apply_real(tb_ctx *ctx, int op, const double *in_host, double *out_host);
int tb_cg_real(tb_ctx *ctx, const double *b_host, double *x_host, int propagator, int *status, int *iters, double *rr);

/* Batched BLAS-1 of vec_ops.c on device-resident real vectors double[nchains][NT*NX]: vec_dot (vec_ops.c:56-62), one
 * value per vector to the host, summed in a fixed order; vec_dmul_add (vec_ops.c:51-55), a = b + e[chain] * d. */
int tb_vec_dot_real_dev(tb_ctx *ctx, const double *d_a, const double *d_b, double *out_host);
int tb_vec_dmul_add_real_dev(tb_ctx *ctx, double *d_a, const double *d_b, const double *d_d, const double *e_host);

/* out = Op in, Op one of TB_OP_* (fm_mul hmc.c:123, fm_conjugate_mul hmc.c:188, fmdm_mul hmc.c:259). */
int tb_apply(tb_ctx *ctx, int op, const double *in_host, double *out_host);

/* fmdm_invert_cg (hmc.c:341-404), batched: solve (M~ M) x = b for every chain from x0 = 0.
 * status / iters / rr may be NULL, else arrays of nchains. */
int tb_cg(tb_ctx *ctx, const double *b_host, double *x_host, int *status, int *iters, double *rr);

/* tb_set_gauge + tb_cg in one call (what momentum_step does at every leapfrog step: new angles, then the solve,
 * hmc.c:504-516).  Links and sources travel per sub-batch of chains, so the first solves start while the rest
 * of the input is still crossing PCIe. */
int tb_cg_gauge(tb_ctx *ctx, const double *A_host, const double *b_host, double *x_host, int *status, int *iters,
                double *rr);

/* fm_invert_cg (hmc.c:408-414), batched: x = (M~ M)^-1 M~ v. */
int tb_invert(tb_ctx *ctx, const double *v_host, double *x_host, int *status, int *iters, double *rr);

/* ---- device-resident entry points (no PCIe traffic; device layout unless stated) -------------------- */

size_t tb_vec_doubles(const tb_ctx *ctx); /* doubles in one device vector = 2*NT*NX*nchains */

/* canonical [chain][t][x] <-> device [t][x][chain] re-layout, both buffers on the device */
int tb_pack_dev(tb_ctx *ctx, const double *d_canonical, double *d_vec);
int tb_unpack_dev(tb_ctx *ctx, const double *d_vec, double *d_canonical);

/* links from angles already on the device, canonical gauge layout */
int tb_set_gauge_dev(tb_ctx *ctx, const double *d_A_canonical);

int tb_apply_dev(tb_ctx *ctx, int op, const double *d_in, double *d_out);

/* Batched CG on device vectors; results of the last solve are read with tb_cg_result. */
int tb_cg_dev(tb_ctx *ctx, const double *d_b, double *d_x);
int tb_invert_dev(tb_ctx *ctx, const double *d_v, double *d_x);
int tb_cg_result(tb_ctx *ctx, int *status, int *iters, double *rr);

/* Re<a,b> per chain, summed in a fixed (run-to-run deterministic) tree; out = nchains doubles on the HOST.
 * Replaces the sequential action sums, e.g. hmc.c:456-459, 472-475. */
int tb_re_dot_dev(tb_ctx *ctx, const double *d_a, const double *d_b, double *out_host);

/* ---- slab decomposition of ONE large lattice over the GPUs of a box --------------
 * One process per GPU; rank r owns global rows [r*NT/nranks, (r+1)*NT/nranks) and passes only its slab
 * (host layout double[nchains][NT/nranks][NX][2]) to every entry point above.  The stencil reads the
 * neighbour ranks' boundary rows straight out of their HBM (CUDA-IPC peer mapping over NVLink, epoch flags),
 * CG dot products are all-reduced by one-shot peer stores; no NCCL on the data path.  Every apply / solve /
 * gauge update is COLLECTIVE: all ranks call it in the same order.  After tb_set_gauge* the caller must
 * tb_synchronize and barrier across ranks (the neighbour reads the boundary row of the t-links).
 * Set-up: every rank tb_create_slab -> tb_slab_export -> all-gather the handles -> tb_slab_connect -> barrier. */
int tb_create_slab(tb_ctx **out, int nt_global, int nx, int nchains, int mode, int device, int rank, int nranks);
int tb_slab_handle_bytes(void);
int tb_slab_export(tb_ctx *ctx, void *handle_out);
int tb_slab_connect(tb_ctx *ctx, const void *all_handles);

/* ---- device-resident batched HMC trajectory: the caller of the hot path -----------
 * update_gauge (hmc.c:671-746) for every chain at once, as coded (leapfrog of hmc.c:708-716 with runtime nsteps
 * and trajectory length instead of the hard-coded 10 and 1, forces of hmc.c:504-661, Metropolis test of
 * hmc.c:738).  Nothing crosses PCIe during a trajectory except the per-chain observables at the end. */

/* coupling g (global hmc.c:39): Nf/g with Nf = 2 (hmc.c:28); n == 1 broadcasts, n == nchains per chain */
int tb_hmc_set_coupling(tb_ctx *ctx, const double *g, int n);

/* Global index of this context's first chain.  It enters the key of the device random-number stream (Philox keyed
 * by seed and GLOBAL chain index), so an ensemble sharded over several GPUs draws the same numbers per chain whatever
 * the number of shards.  Default 0. */
int tb_hmc_set_chain_offset(tb_ctx *ctx, unsigned int first_chain);

/* `sweeps` quenched heat-bath sweeps of every link (update_puregauge_hb, hmc.c:82-93; main() does 100 from
 * A = 0, hmc.c:927-929).  Random numbers: device Philox keyed by (seed, chain). */
int tb_hmc_heatbath(tb_ctx *ctx, int sweeps, unsigned long long seed);

/* One trajectory per chain.  The four random inputs are drawn on the device (Philox keyed by seed, chain,
 * traj_index) when the pointer is NULL, or taken from the host for parity tests: xi_host / st_host complex
 * [chain][t][x] (random_pseudofermion hmc.c:418, stochastic_vector hmc.c:439), mom_host real [chain][t][x][2]
 * (random_momentum hmc.c:483), u_host [chain] (the Metropolis uniform, hmc.c:738).
 * obs_host (may be NULL): 10 doubles per chain = Sg, Smdm, Smd, Smom at the start (hmc.c:701), the same four
 * at the end (hmc.c:735), dS, accepted.  accepted_host[chain], cg_iters_host (sum over chains and solves)
 * may be NULL.  A chain whose CG diverges or hits max-iter is rejected. */
int tb_hmc_trajectory(tb_ctx *ctx, int nsteps, double traj_length, unsigned long long seed,
                      unsigned int traj_index, const double *xi_host, const double *mom_host,
                      const double *st_host, const double *u_host, double *obs_host, int *accepted_host,
                      long long *cg_iters_host);

/* dS/dA for every link of every chain on the current field: what one momentum_step subtracts, times eps, from the
 * momenta (hmc.c:504-661): gauge force (Nf/g) sin A, pseudofermion force of Re<psi, (M~M)^-1 psi>, and the force of
 * the stochastic Re<st, M~ st> term as coded (st_host NULL: no such term).  psi_host complex [chain][t][x], force_host
 * real [chain][t][x][2].  The quantity the reference's disabled CHECK_FORCE block (hmc.c:502,535-559) compares with a
 * finite difference of pseudofermion_action. */
int tb_hmc_force(tb_ctx *ctx, const double *psi_host, const double *st_host, double *force_host);

/* Per chain, how the CG solves of the last tb_hmc_trajectory ended: bit (1 << TB_CG_MAXITER) and / or
 * (1 << TB_CG_DIVERGED) set, 0 when every solve converged.  The reference exit(1)s on divergence (hmc.c:383-388)
 * and is silent at max-iter; a batched trajectory rejects such a chain and reports it here. */
int tb_hmc_cg_failures(tb_ctx *ctx, int *mask_host);

/* measure() (hmc.c:823-842) per chain: Magnetisation = sum A / V and Phase = (1/nsrc) sum Im<c, M~ c> over nsrc
 * stochastic vectors (fermion_phase hmc.c:794-815 uses 20).  sources_host (optional): complex
 * [nsrc][chain][t][x]. */
int tb_hmc_measure(tb_ctx *ctx, int nsrc, unsigned long long seed, unsigned int meas_index,
                   const double *sources_host, double *magnetisation_host, double *phase_host);

/* Chiral condensate per chain (the reference's measure() lacks it):
 * (1/V) Tr M^-1 estimated as (1/(2 V nsrc)) sum_i Re<eta_i, M^-1 eta_i> with M^-1 eta = fm_invert_cg(eta)
 * (hmc.c:408-414) over nsrc stochastic vectors (stochastic_vector, hmc.c:439-447; E|eta|^2 = 2 per site).
 * sources_host (optional): complex [nsrc][chain][t][x]; cg_iters_host (optional): iterations summed over
 * chains and sources.  A chain whose solve does not converge gets NaN. */
int tb_hmc_condensate(tb_ctx *ctx, int nsrc, unsigned long long seed, unsigned int meas_index,
                      const double *sources_host, double *condensate_host, long long *cg_iters_host);

/* current gauge angles to the host, double[nchains][NT][NX][2] */
int tb_get_gauge(tb_ctx *ctx, double *A_host);

/* Gauge-field checkpoint (hmc.c has none; mirrors the raw dump of fermionbag.c:125-161): 64-byte header
 * ("THIRRING2D-A-V1", NT, NX, nchains, mode, index of the next trajectory) + raw FP64 A[chain][t][x][dir]. */
int tb_checkpoint_write(tb_ctx *ctx, const char *path);
int tb_checkpoint_read(tb_ctx *ctx, const char *path);

/* The trajectory counter that travels in the checkpoint header.  The device random stream is keyed by (seed, chain,
 * trajectory index): a resumed run that restarted the index at 1 would replay the momenta, the pseudofermion noise,
 * the Metropolis uniforms and the measurement sources of its first leg.  Set it before tb_checkpoint_write (the index
 * the NEXT trajectory will use); tb_checkpoint_read restores it (0: the file does not record it). */
int tb_checkpoint_set_next_trajectory(tb_ctx *ctx, unsigned int next_traj);
int tb_checkpoint_next_trajectory(const tb_ctx *ctx, unsigned int *next_traj);

/* Counters for bench.py: kernels launched by this context since creation / since the last reset. */
long long tb_launch_count(const tb_ctx *ctx);
int tb_reset_launch_count(tb_ctx *ctx);

/* FP64 FMA issue rate of the context's device in TFLOP/s (a kernel of independent DFMA chains, best of `repeats`
 * launches, CUDA events): the measured denominator for the on-chip solvers, which are bound by FP64 issue. */
int tb_measure_fp64_peak(tb_ctx *ctx, int repeats, double *tflops_out);
/* The same measurement with a choice of instruction mix: kind 0 = tb_measure_fp64_peak (two of the three source operands
 * are shared by all FMAs), kind 1 = every source operand of every FMA in its own register, which is what the FMAs of a
 * stencil look like to the register file. */
int tb_measure_fp64_rate(tb_ctx *ctx, int kind, int repeats, double *tflops_out);

/* Milliseconds the device spent in the last tb_cg_dev/tb_invert_dev call (CUDA events on the context stream). */
double tb_last_solve_ms(const tb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* THIRRING_B200_H */
