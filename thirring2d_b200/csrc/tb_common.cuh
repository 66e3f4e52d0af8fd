// tb_common.cuh — context, geometry and error plumbing shared by the sm_100a translation units.
#pragma once

#include <cuda_runtime.h>

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

struct tb_ctx {
  int nt, nx, C, mode, device;
  size_t V;          // nt*nx
  size_t nsite;      // V*C
  cudaStream_t stream;
  bool own_stream;
  TbGeom g;
  int tune_tt, tune_chunk, tune_solver;
  // parameters
  double *d_mass, *d_emu, *d_emmu;  // [Cpad]
  double *h_mass, *h_mu;
  bool has_mu;  // some chain has mu != 0
  // links, device layout [t][x][c]
  double2 *W0, *W1;
  bool have_gauge;
  // work vectors (device layout)
  double2 *r, *p, *Mp, *q, *xw, *tmp, *vin, *vout;
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
  bool sub_pending;  // work queued on the sub-streams that the context stream has not joined yet
  double2 *stage_x;  // second canonical staging buffer (results)
  double last_solve_ms;
  long long launches;
};

int tb_choose_geom(tb_ctx *ctx);

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
int tb_launch_dot(tb_ctx *ctx, const double2 *a, const double2 *b, double *d_out);

static inline bool tb_conj_is_dagger(const tb_ctx *ctx) { return ctx->mode == TB_MODE_ADJOINT; }
