// tb_common.cuh — context, geometry and error plumbing shared by the sm_100a translation units.
#pragma once

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/thirring_b200.h"

#define TB_NUM_SMS_B200 148
#define TB_MAX_BLOCK 256
#define TB_MAX_SUB 8
#define TB_DIVERGENCE_RATIO 1e10 /* hmc.c:383 */

void tb_set_error(const char *fmt, ...);

#define TB_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      tb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
      return TB_ECUDA;                                                                       \
    }                                                                                        \
  } while (0)

#define TB_CHECK(call)            \
  do {                            \
    int r__ = (call);             \
    if (r__ != TB_OK) return r__; \
  } while (0)

// Thread-block geometry of the streaming kernels.  A block covers BC chains x BX x-sites and marches over
// TT consecutive t rows; lanes of a warp are chains of the same site (or neighbouring x when C < 32), so
// every neighbour access is a contiguous run of BC double2.
struct TbGeom {
  int nt, nx, C;      // lattice and number of chains
  int R;              // row length in sites*chains = nx*C
  int bc, bx;         // block tile: chains x x-sites (powers of two), bc*bx threads
  int bc_shift;       // log2(bc)
  int ta_shift;       // log2 of the chain-tile width that TbCgState::tile_active counts in (the marching kernels' bc)
  int nctiles, nxtiles, nttiles;
  int tt;             // rows marched per thread
  int nslots;         // partial sums per chain = nxtiles*nttiles
  int Cpad;           // nctiles*bc
};

// Per-chain scalars of the batched CG (device arrays of length Cpad unless noted).
struct TbCgState {
  double *rr_old, *rr_init, *rr, *pq, *alpha, *beta, *dot;
  int *active, *status, *iters;
  int *tile_active;  // [nctiles] number of active chains per chain tile
  int *n_active;     // [1]
  double *partial;   // [nslots][Cpad]
  unsigned int *ticket;  // [nctiles]
  double accuracy;
  int max_iter;
};

// Slab decomposition of ONE large lattice over P GPUs (1-D in t, rank r owns global rows [r*nt, (r+1)*nt)).
// Nothing is copied for the halos: the stencil reads the neighbour rank's boundary row straight out of the
// neighbour's HBM through an IPC-mapped peer pointer (NVLink), guarded by epoch flags that the producing
// kernel's last block writes into the consumer's memory.  CG dot products are all-reduced by every rank
// storing its partial into every peer's slot table (one-shot, P <= 8), summed in rank order => bitwise
// identical scalars, hence identical convergence decisions, on all ranks.
#define TB_SLAB_MAX_RANKS 8
enum { TB_FLAG_PREADY = 0, TB_FLAG_MPREADY = 1, TB_FLAG_PDONE = 2, TB_FLAG_MPDONE = 3, TB_NFLAGS = 4 };
enum { TB_RED_PQ = 0, TB_RED_RR = 1, TB_RED_INIT = 2, TB_NRED = 3 };
struct TbSlab {
  int P, rank;
  // peer views of the two exchange vectors and of the t-links
  const double2 *p_prev, *p_next, *mp_prev, *mp_next, *W0_prev;
  // one-launch solve: the neighbours' residual and second direction buffer (p is double-buffered there)
  const double2 *r_prev, *r_next, *p1_prev, *p1_next;
  // my flags: flags[kind*2 + side], side 0 = written by the previous rank, 1 = by the next rank
  volatile int *flags;
  int *sig_prev;  // previous rank's flags + 1 (its "from next" side):  sig_prev[kind*2]
  int *sig_next;  // next rank's flags + 0 (its "from prev" side):      sig_next[kind*2]
  // all-reduce slots in MY memory: red[(kind*P + q)*Cpad + c], red_flag[(kind*P + q)*nctiles + ctile]
  double *red;
  volatile int *red_flag;
  double *peer_red[TB_SLAB_MAX_RANKS];
  int *peer_red_flag[TB_SLAB_MAX_RANKS];
  int *seq;                    // device epoch counter = generation of the exchange vector p
  unsigned int *done_ticket;   // [TB_NFLAGS] completion counters of the signalling kernels
  int *err;                    // set when a flag wait timed out
  // persistent solve (tb_stream.cu: slab_cg_persistent_kernel): grid barrier counter and block 0's broadcast flag
  unsigned long long *gbar;
  int *go;
  // one-launch solve: all-reduce slots {partial, tag} written by one 16-byte store, [(kind*P + q)*Cpad + c] in MY memory
  double2 *red2;
  double2 *peer_red2[TB_SLAB_MAX_RANKS];
  // one-launch solve, every block polls (slab_cg_onelaunch_kernel): the same {partial, tag} slots, replicated NREP times
  // on separate L2 lines so that the pollers of one GPU spread over the replicas:
  //   red3[rep * rep_stride + (kind*P + q)*Cpad + c]
  double2 *red3;
  double2 *peer_red3[TB_SLAB_MAX_RANKS];
  int nrep, rep_stride;
  // the totals {total, tag} the last-arriving block of THIS GPU publishes to its other blocks: bcast[rep * bcast_stride + kind*Cpad + c]
  double2 *bcast;
  int bcast_stride;
  unsigned long long *timeline;   // optional: globaltimer stamps of the first iterations (TB_SLAB_TIMELINE), else nullptr
};
#define TB_SLAB_NREP_MAX 8
#define TB_SLAB_TL_ITERS 64   /* iterations the timeline records */
#define TB_SLAB_TL_WORDS 8    /* stamps per iteration: phase A first start, last end, stores issued, last total seen; same for B */

// buffers of the device-resident HMC trajectory (tb_hmc.cu), allocated at first use
struct TbHmc {
  double2 *mom, *newA, *psi, *st, *chi, *phi, *gauss;
  double *sums, *obs, *u, *nf_over_g;
  int *accept, *failed;
  unsigned long long *iter_sum;   // CG iterations of the trajectory, summed over chains and solves on the device
};

#define TB_TMAP_CACHE 12

struct tb_ctx {
  int nt, nx, C, mode, device;
  size_t V;          // nt*nx
  size_t nsite;      // V*C
  cudaStream_t stream;
  bool own_stream;
  TbGeom g;
  TbGeom gp;        // geometry of the TMA-staged streaming kernels (tb_stream.cu: *_pipe_kernel): same tiles, taller blocks
  bool pipe_ok;     // the lattice / batch has a TMA-staged shape (whole tiles, chain runs of whole 16-byte multiples)
  bool pipe_tiled;  // the staged tiles hold 16 of MORE chains: their rows are 2-D boxes, copied through tensor maps
  // tensor maps of the vectors the tiled staged kernels read (128-byte CUtensorMap objects, kept opaque here)
  unsigned char tmap_store[TB_TMAP_CACHE][128];
  const void *tmap_ptr[TB_TMAP_CACHE];
  int tmap_next;
  int xp_parity;    // two-launch staged iteration: which of the two direction buffers (p, q) holds the current direction
  int tune_tt, tune_chunk, tune_solver;
  int resident_x_tmem;  // resident solver: keep x in tensor memory (1, default) or in an L2 workspace (0)
  int cluster_capacity; // cluster solver: co-resident clusters of this lattice's shape (-1 = not queried yet)
  int cg_variant;  // 0 auto (fused 3-kernel iteration when M~ = M^dagger), 4 = always the 4-kernel form
  // parameters
  double *d_mass, *d_emu, *d_emmu;  // [Cpad]
  double *h_mass, *h_mu;
  bool has_mu;  // some chain has mu != 0
  // links, device layout [t][x][c]
  double2 *W0, *W1;
  bool have_gauge;
  // one gauge field shared by every chain of the batch (multi-RHS solves, tb_set_gauge_shared): W0 / W1 above hold it
  // replicated per chain for every kernel that indexes links by chain; Ws holds it once, W0 [V] then W1 [V], for the
  // staged streaming kernels, which then read 2 x 16 B of links per SITE instead of per site and chain
  double2 *Ws;
  bool gauge_shared;
  int gp_tt0;       // rows per staged block as tb_choose_geom chose them (a shared gauge field runs shorter blocks)
  // family B (vec_ops.c): per-site mass (occupied site = identity row); msite == nullptr for family A
  double *msite, *msite_buf;
  int *occ_dev, *occ_stage;
  int occ_bc;   // family B boundary variant (TB_BC_*), Thirring.h:27-29
  // work vectors (device layout)
  double2 *r, *p, *Mp, *q, *xw, *tmp, *vin, *vout;
  double2 *p1;    // slab mode: second direction buffer of the one-launch solve (in the exchange block, like p, Mp, r)
  double2 *Adev;  // angles (A0,A1) in device layout
  double *stage;  // canonical-layout device staging buffer (2*nsite doubles)
  double *h_pinned;  // pinned host staging (2*nsite doubles)
  TbCgState cg;
  int *h_flag;  // pinned: n_active readback (ring of 2)
  int *h_status, *h_iters;
  double *h_rr;
  cudaGraphExec_t cg_graph;
  int cg_graph_chunk;
  cudaEvent_t ev0, ev1, ev_flag[2];
  // host-buffer pipeline: sub-batches of chains on their own streams (H2D -> pack -> solve -> unpack -> D2H)
  cudaStream_t sub_stream[TB_MAX_SUB];
  cudaEvent_t sub_done[TB_MAX_SUB], fork_ev;
  int nsub;
  int sub_c0[TB_MAX_SUB + 1];   // sub-batch s holds chains [sub_c0[s], sub_c0[s + 1])
  bool sub_pending;  // work queued on the sub-streams that the context stream has not joined yet
  double2 *stage_x;  // second canonical staging buffer (results)
  double last_solve_ms;
  long long launches;
  int plan_m;     // machines the plan buffer was sized for
  int *plan_buf;  // planned launches of the on-chip solver (tb_resident.cu: TbPlan): segments, ranges, hand-over flags
  // slab mode (nranks > 1): ctx->nt is the LOCAL number of rows
  int nranks, rank, nt_global, t_off;
  void *slab_block;                 // the IPC-exported allocation (p, Mp, W0, slots, flags)
  size_t slab_bytes;
  void *peer_block[TB_SLAB_MAX_RANKS];
  bool slab_connected;
  TbSlab slab;
  cudaGraphExec_t slab_graph;
  int slab_graph_chunk;
  TbHmc hmc;
  unsigned int hmc_chain_offset;  // global index of chain 0 (RNG key), tb_hmc_set_chain_offset
  unsigned int ckpt_next_traj;    // index of the next trajectory, stored in / restored from the checkpoint header
};

int tb_choose_geom(tb_ctx *ctx);
int tb_stream_kernels(const tb_ctx *ctx, int *tile_chains, int *tile_sites, int *rows_per_block);

// kernels / launch wrappers (tb_dirac.cu, tb_cg.cu)
int tb_launch_links(tb_ctx *ctx, const double *d_A_dev_layout);
int tb_launch_pack(tb_ctx *ctx, const double *d_canonical, double2 *d_vec);
int tb_launch_unpack(tb_ctx *ctx, const double2 *d_vec, double *d_canonical);
int tb_launch_dslash(tb_ctx *ctx, bool dagger, const double2 *in, double2 *out, bool masked);
int tb_run_cg_stream(tb_ctx *ctx, const double2 *b, double2 *x);
int tb_run_cg_resident(tb_ctx *ctx, const double2 *b, double2 *x);
int tb_run_cg_resident_slice(tb_ctx *ctx, const double2 *b, double2 *x, int c0, int n, cudaStream_t st);
int tb_launch_links_slice(tb_ctx *ctx, const double2 *d_A_dev_layout, int c0, int n, cudaStream_t st);
int tb_launch_pack_slice(tb_ctx *ctx, const double2 *d_canonical_slice, double2 *d_vec, int c0, int n, cudaStream_t st);
int tb_launch_unpack_slice(tb_ctx *ctx, const double2 *d_vec, double2 *d_canonical_slice, int c0, int n, cudaStream_t st);
bool tb_resident_supported(const tb_ctx *ctx);
bool tb_resident_canon_supported(const tb_ctx *ctx);
int tb_run_cg_resident_canon(tb_ctx *ctx, const double2 *b_canon, double2 *x_canon, const double2 *A_canon, int c0, int n,
                             cudaStream_t st);
bool tb_cluster_supported(tb_ctx *ctx);
int tb_cluster_capacity(tb_ctx *ctx);
int tb_run_cg_cluster_slice(tb_ctx *ctx, const double2 *b, double2 *x, int c0, int n, cudaStream_t st);
int tb_launch_dot(tb_ctx *ctx, const double2 *a, const double2 *b, double *d_out);
int tb_launch_occupancy(tb_ctx *ctx, const int *d_field_canonical);
bool tb_real_cg_supported(const tb_ctx *ctx);
int tb_launch_real_apply(tb_ctx *ctx, bool transpose, const double *d_in, double *d_out);
int tb_run_cg_real(tb_ctx *ctx, const double *d_b, double *d_x, bool propagator, int c0, int n, cudaStream_t st);
int tb_slab_apply(tb_ctx *ctx, int op, const double2 *in, double2 *out);
int tb_slab_layout(tb_ctx *ctx);
int tb_run_cg_any(tb_ctx *ctx, const double2 *b, double2 *x);
int tb_run_cg_strict(tb_ctx *ctx, const double2 *b, double2 *x);   // tb_strict.cu: reference evaluation order
int tb_launch_links_from_trig(tb_ctx *ctx, const double2 *T0, const double2 *T1);
int tb_launch_links_shared(tb_ctx *ctx, const double2 *d_A_one_field);
void tb_gauge_sharing(tb_ctx *ctx, bool shared);   // sets gauge_shared and the staged geometry that goes with it   // [t][x] angles -> Ws, W0 / W1 / Adev replicated
void tb_hmc_release(tb_ctx *ctx);
int tb_create_common(tb_ctx **out, int nt_local, int nx, int nchains, int mode, int device, int rank, int nranks,
                     int nt_global);

static inline bool tb_conj_is_dagger(const tb_ctx *ctx) { return ctx->mode == TB_MODE_ADJOINT; }
