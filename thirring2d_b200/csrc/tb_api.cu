// tb_api.cu — the C-ABI of include/thirring_b200.h: context life cycle, host-buffer entry points and the
// dispatch of the batched CG.  No CPU fallback: every entry point runs CUDA kernels or fails.
#include <cmath>

#include "tb_common.cuh"

static thread_local char g_err[512] = "";

void tb_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char *tb_last_error(void) { return g_err; }

extern "C" int tb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

template <typename T>
static int dev_alloc(T **p, size_t n) {
  cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
  if (e != cudaSuccess) {
    tb_set_error("cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
    return TB_ENOMEM;
  }
  return TB_OK;
}

static void invalidate_graph(tb_ctx *ctx) {
  if (ctx->cg_graph) {
    cudaGraphExecDestroy(ctx->cg_graph);
    ctx->cg_graph = nullptr;
  }
}

static int alloc_cg_state(tb_ctx *ctx) {
  // sized for the finest geometry any tuning can choose: tt = 1
  const TbGeom &g = ctx->g;
  const size_t cp = g.Cpad;
  const size_t max_slots = (size_t)g.nxtiles * ctx->nt;
  TbCgState &s = ctx->cg;
  double *dbl = nullptr;
  TB_CHECK(dev_alloc(&dbl, 7 * cp));
  s.rr_old = dbl;
  s.rr_init = dbl + cp;
  s.rr = dbl + 2 * cp;
  s.pq = dbl + 3 * cp;
  s.alpha = dbl + 4 * cp;
  s.beta = dbl + 5 * cp;
  s.dot = dbl + 6 * cp;
  int *ints = nullptr;
  TB_CHECK(dev_alloc(&ints, 3 * cp + g.nctiles + 1));
  s.active = ints;
  s.status = ints + cp;
  s.iters = ints + 2 * cp;
  s.tile_active = ints + 3 * cp;
  s.n_active = ints + 3 * cp + g.nctiles;
  TB_CHECK(dev_alloc(&s.partial, max_slots * cp));
  TB_CHECK(dev_alloc(&s.ticket, cp));   // one per chain tile of ANY geometry (the staged kernels' tiles can be narrower)
  TB_CUDA(cudaMemset(dbl, 0, 7 * cp * sizeof(double)));
  TB_CUDA(cudaMemset(ints, 0, (3 * cp + g.nctiles + 1) * sizeof(int)));
  TB_CUDA(cudaMemset(s.ticket, 0, cp * sizeof(unsigned int)));
  s.accuracy = 1e-30;   // CG_ACCURACY, hmc.c:34
  s.max_iter = 100000;  // CG_MAX_ITER, hmc.c:35
  return TB_OK;
}

void tb_slab_release(tb_ctx *ctx);

extern "C" int tb_create(tb_ctx **out, int nt, int nx, int nchains, int mode, int device) {
  return tb_create_common(out, nt, nx, nchains, mode, device, 0, 1, nt);
}

int tb_create_common(tb_ctx **out, int nt, int nx, int nchains, int mode, int device, int rank, int nranks,
                     int nt_global) {
  if (!out || nt < 2 || nx < 2 || nchains < 1 || (mode != TB_MODE_REF_COMPAT && mode != TB_MODE_ADJOINT)) {
    tb_set_error("tb_create: invalid arguments (nt=%d nx=%d nchains=%d mode=%d)", nt, nx, nchains, mode);
    return TB_EINVAL;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    tb_set_error("tb_create: no CUDA device available (this library has no CPU path)");
    return TB_ENODEVICE;
  }
  if (device < 0 || device >= ndev) {
    tb_set_error("tb_create: device %d out of range (have %d)", device, ndev);
    return TB_EINVAL;
  }
  TB_CUDA(cudaSetDevice(device));
  tb_ctx *ctx = (tb_ctx *)calloc(1, sizeof(tb_ctx));
  if (!ctx) return TB_ENOMEM;
  ctx->nt = nt;
  ctx->nx = nx;
  ctx->C = nchains;
  ctx->mode = mode;
  ctx->device = device;
  ctx->V = (size_t)nt * nx;
  ctx->nsite = ctx->V * nchains;
  ctx->rank = rank;
  ctx->nranks = nranks;
  ctx->nt_global = nt_global;
  ctx->t_off = rank * nt;
  const char *e;
  ctx->tune_tt = (e = getenv("TB_ROWS_PER_THREAD")) ? atoi(e) : 0;
  ctx->tune_chunk = (e = getenv("TB_ITERS_PER_LAUNCH")) ? atoi(e) : 0;
  ctx->tune_solver = (e = getenv("TB_SOLVER")) ? atoi(e) : 0;
  ctx->cg_variant = (e = getenv("TB_CG_VARIANT")) ? atoi(e) : 0;
  ctx->resident_x_tmem = (e = getenv("TB_RESIDENT_X_TMEM")) ? atoi(e) : 1;
  ctx->cluster_capacity = -1;
  tb_choose_geom(ctx);
  int rc = TB_OK;
  cudaError_t ce = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (ce != cudaSuccess) {
    tb_set_error("cudaStreamCreate: %s", cudaGetErrorString(ce));
    free(ctx);
    return TB_ECUDA;
  }
  ctx->own_stream = true;
  const size_t n = ctx->nsite;
  const size_t cp = ctx->g.Cpad;
#define A_(ptr, cnt)                              \
  if (rc == TB_OK) rc = dev_alloc(&(ptr), (cnt));
  A_(ctx->d_mass, cp) A_(ctx->d_emu, cp) A_(ctx->d_emmu, cp)
  if (nranks > 1) {  // p, Mp and W0 live in the IPC-exported exchange block
    if (rc == TB_OK) rc = tb_slab_layout(ctx);
  } else {
    A_(ctx->W0, n) A_(ctx->p, n) A_(ctx->Mp, n) A_(ctx->r, n)
  }
  A_(ctx->W1, n) A_(ctx->Adev, n)
  A_(ctx->q, n) A_(ctx->xw, n) A_(ctx->tmp, n)
  A_(ctx->vin, n) A_(ctx->vout, n)
  A_(ctx->stage, 2 * n)
  A_(ctx->stage_x, n)
#undef A_
  if (rc == TB_OK) rc = alloc_cg_state(ctx);
  if (rc == TB_OK) {
    ctx->h_mass = (double *)malloc(cp * sizeof(double));
    ctx->h_mu = (double *)malloc(cp * sizeof(double));
    if (cudaHostAlloc((void **)&ctx->h_flag, 4 * sizeof(int), cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void **)&ctx->h_status, cp * sizeof(int), cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void **)&ctx->h_iters, cp * sizeof(int), cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void **)&ctx->h_rr, cp * sizeof(double), cudaHostAllocDefault) != cudaSuccess) {
      tb_set_error("cudaHostAlloc failed");
      rc = TB_ENOMEM;
    }
  }
  if (rc == TB_OK) {
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    cudaEventCreateWithFlags(&ctx->ev_flag[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_flag[1], cudaEventDisableTiming);
    // host-buffer pipeline: chains are processed in nsub sub-batches on their own streams.  Where one CTA per chain
    // serves the context (64^2) a batch larger than the SM count runs in waves of nsm chains: every wave is cut into
    // pieces of its own, so that the first wave's CTAs start as their inputs arrive (256 chains on 148 SMs: 4 x 37, then
    // 4 x 27) instead of 4 x 64, of which at most 128 fit the first wave.
    const char *es = getenv("TB_SUBBATCHES");
    int nsm = TB_NUM_SMS_B200;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    const bool cta_per_chain = nranks == 1 && nt == 64 && nx == 64;
    int nb = 0;
    if (!es && cta_per_chain && ctx->C > nsm) {
      const int waves = (ctx->C + nsm - 1) / nsm;
      const int pieces = waves >= TB_MAX_SUB ? 1 : TB_MAX_SUB / waves;
      for (int w = 0; w < waves && nb < TB_MAX_SUB; w++) {
        const int lo = w * nsm, hi = (w + 1) * nsm < ctx->C ? (w + 1) * nsm : ctx->C;
        const bool last_slot = w == waves - 1 || nb + pieces >= TB_MAX_SUB;   // the last pieces take whatever is left
        const int end = last_slot && w < waves - 1 ? ctx->C : hi;
        for (int q = 0; q < pieces && nb < TB_MAX_SUB; q++) ctx->sub_c0[nb++] = lo + (int)((long long)(end - lo) * q / pieces);
        if (end == ctx->C) break;
      }
    } else {
      int k = es ? atoi(es) : (ctx->C >= 128 ? 4 : (ctx->C >= 32 ? 2 : 1));
      if (nranks > 1) k = 1;
      if (k < 1) k = 1;
      if (k > TB_MAX_SUB) k = TB_MAX_SUB;
      if (k > ctx->C) k = ctx->C;
      const int per = (ctx->C + k - 1) / k;
      for (int q = 0; q < k && q * per < ctx->C; q++) ctx->sub_c0[nb++] = q * per;
    }
    ctx->nsub = nb;
    ctx->sub_c0[nb] = ctx->C;
    cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming);
    for (int s = 0; s < ctx->nsub; s++) {
      cudaStreamCreateWithFlags(&ctx->sub_stream[s], cudaStreamNonBlocking);
      cudaEventCreateWithFlags(&ctx->sub_done[s], cudaEventDisableTiming);
    }
    const double one = 1.0, zero = 0.0;
    rc = tb_set_params(ctx, &one, &zero, 1);
  }
  if (rc != TB_OK) {
    tb_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return TB_OK;
}

extern "C" int tb_destroy(tb_ctx *ctx) {
  if (!ctx) return TB_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (int s = 0; s < ctx->nsub; s++) {
    if (ctx->sub_stream[s]) { cudaStreamSynchronize(ctx->sub_stream[s]); cudaStreamDestroy(ctx->sub_stream[s]); }
    if (ctx->sub_done[s]) cudaEventDestroy(ctx->sub_done[s]);
  }
  if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
  invalidate_graph(ctx);
  const bool slab = ctx->nranks > 1;
  tb_slab_release(ctx);
  tb_hmc_release(ctx);
  void *dev[] = {ctx->d_mass, ctx->d_emu, ctx->d_emmu, slab ? nullptr : (void *)ctx->W0, ctx->W1, ctx->Adev, slab ? nullptr : (void *)ctx->r,
                 slab ? nullptr : (void *)ctx->p, slab ? nullptr : (void *)ctx->Mp, ctx->q, ctx->xw, ctx->tmp, ctx->vin, ctx->vout, ctx->stage, ctx->stage_x, ctx->cg.rr_old, ctx->cg.active,
                 ctx->cg.partial, ctx->cg.ticket};
  for (void *p : dev)
    if (p) cudaFree(p);
  if (ctx->plan_buf) cudaFree(ctx->plan_buf);
  if (ctx->Ws) cudaFree(ctx->Ws);
  if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
  if (ctx->h_status) cudaFreeHost(ctx->h_status);
  if (ctx->h_iters) cudaFreeHost(ctx->h_iters);
  if (ctx->h_rr) cudaFreeHost(ctx->h_rr);
  free(ctx->h_mass);
  free(ctx->h_mu);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev_flag[0]) cudaEventDestroy(ctx->ev_flag[0]);
  if (ctx->ev_flag[1]) cudaEventDestroy(ctx->ev_flag[1]);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  free(ctx);
  return TB_OK;
}

// ---- sub-batch streams (host-buffer pipeline) -----------------------------------------------------------
static void sub_range(const tb_ctx *ctx, int s, int *c0, int *n) {
  *c0 = ctx->sub_c0[s];
  *n = ctx->sub_c0[s + 1] - ctx->sub_c0[s];
}

// sub-streams start after everything already queued on the context stream
static int fork_subs(tb_ctx *ctx) {
  TB_CUDA(cudaEventRecord(ctx->fork_ev, ctx->stream));
  for (int s = 0; s < ctx->nsub; s++) TB_CUDA(cudaStreamWaitEvent(ctx->sub_stream[s], ctx->fork_ev, 0));
  ctx->sub_pending = true;
  return TB_OK;
}

// the context stream continues after everything queued on the sub-streams (no-op when nothing is pending)
static int join_subs(tb_ctx *ctx) {
  if (!ctx->sub_pending) return TB_OK;
  for (int s = 0; s < ctx->nsub; s++) {
    TB_CUDA(cudaEventRecord(ctx->sub_done[s], ctx->sub_stream[s]));
    TB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->sub_done[s], 0));
  }
  ctx->sub_pending = false;
  return TB_OK;
}

static int sync_all(tb_ctx *ctx) {
  TB_CHECK(join_subs(ctx));
  TB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->nranks > 1) {  // a peer-flag wait that timed out is an error, never a hang
    int err = 0;
    TB_CUDA(cudaMemcpy(&err, ctx->slab.err, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) {
      tb_set_error("slab mode: wait on a neighbour's flag timed out (code %d): a rank died or the ranks issued "
                   "collective calls in different orders", err);
      return TB_ECUDA;
    }
  }
  return TB_OK;
}

extern "C" int tb_set_stream(tb_ctx *ctx, void *cuda_stream) {
  if (!ctx) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(sync_all(ctx));
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return TB_OK;
}

extern "C" int tb_synchronize(tb_ctx *ctx) {
  if (!ctx) return TB_EINVAL;
  return sync_all(ctx);
}

extern "C" int tb_set_params(tb_ctx *ctx, const double *m, const double *mu, int n) {
  if (!ctx || !m || !mu || (n != 1 && n != ctx->C)) {
    tb_set_error("tb_set_params: n must be 1 or nchains");
    return TB_EINVAL;
  }
  TB_CUDA(cudaSetDevice(ctx->device));
  const int cp = ctx->g.Cpad;
  double *buf = (double *)malloc(3 * (size_t)cp * sizeof(double));
  ctx->has_mu = false;
  for (int c = 0; c < cp; c++) {
    const int k = (n == 1 || c >= ctx->C) ? 0 : c;
    ctx->h_mass[c] = m[k];
    ctx->h_mu[c] = mu[k];
    if (mu[k] != 0.0) ctx->has_mu = true;
    buf[c] = m[k];
    buf[cp + c] = exp(mu[k]);       // hmc.c:127
    buf[2 * cp + c] = exp(-mu[k]);  // hmc.c:128
  }
  join_subs(ctx);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpy(ctx->d_mass, buf, cp * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(ctx->d_emu, buf + cp, cp * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(ctx->d_emmu, buf + 2 * cp, cp * sizeof(double), cudaMemcpyHostToDevice);
  free(buf);
  TB_CUDA(e);
  // family B: the site masses (m on free sites, 1 on occupied ones) were built from the old masses
  if (ctx->msite) TB_CHECK(tb_launch_occupancy(ctx, ctx->occ_stage));
  return TB_OK;
}

extern "C" int tb_set_cg(tb_ctx *ctx, double accuracy, int max_iter) {
  if (!ctx || !(accuracy >= 0) || max_iter < 2) {
    tb_set_error("tb_set_cg: invalid arguments");
    return TB_EINVAL;
  }
  ctx->cg.accuracy = accuracy;
  ctx->cg.max_iter = max_iter;
  invalidate_graph(ctx);
  return TB_OK;
}

extern "C" int tb_set_tuning(tb_ctx *ctx, int rows_per_thread, int iters_per_launch, int solver) {
  if (!ctx) return TB_EINVAL;
  TB_CHECK(sync_all(ctx));
  ctx->tune_tt = rows_per_thread;
  ctx->tune_chunk = iters_per_launch;
  ctx->tune_solver = solver == 3 ? 1 : solver;   // 3 = streaming solver, 4-kernel iteration
  ctx->cg_variant = solver == 3 ? 4 : 0;
  tb_choose_geom(ctx);
  invalidate_graph(ctx);
  return TB_OK;
}

extern "C" size_t tb_vec_doubles(const tb_ctx *ctx) { return ctx ? 2 * ctx->nsite : 0; }

extern "C" long long tb_launch_count(const tb_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int tb_reset_launch_count(tb_ctx *ctx) {
  if (!ctx) return TB_EINVAL;
  ctx->launches = 0;
  return TB_OK;
}
extern "C" double tb_last_solve_ms(const tb_ctx *ctx) { return ctx ? ctx->last_solve_ms : 0.0; }

// ---- device-resident entry points -----------------------------------------------------------------------

extern "C" int tb_pack_dev(tb_ctx *ctx, const double *d_canonical, double *d_vec) {
  if (!ctx || !d_canonical || !d_vec) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(join_subs(ctx));
  return tb_launch_pack(ctx, d_canonical, (double2 *)d_vec);
}

extern "C" int tb_unpack_dev(tb_ctx *ctx, const double *d_vec, double *d_canonical) {
  if (!ctx || !d_canonical || !d_vec) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(join_subs(ctx));
  return tb_launch_unpack(ctx, (const double2 *)d_vec, d_canonical);
}

extern "C" int tb_set_gauge_dev(tb_ctx *ctx, const double *d_A_canonical) {
  if (!ctx || !d_A_canonical) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(join_subs(ctx));
  TB_CHECK(tb_launch_pack(ctx, d_A_canonical, ctx->Adev));
  TB_CHECK(tb_launch_links(ctx, (const double *)ctx->Adev));
  ctx->msite = nullptr;
  ctx->have_gauge = true;
  return TB_OK;
}

// One gauge field for every chain of the context (the chains are then the right-hand sides of a multi-RHS solve on that
// field: fermion_phase's sources, hmc.c:794-815).  A: [NT][NX][2] angles, device or host.
extern "C" int tb_set_gauge_shared_dev(tb_ctx *ctx, const double *d_A_one_field) {
  if (!ctx || !d_A_one_field) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(join_subs(ctx));
  TB_CHECK(tb_launch_links_shared(ctx, (const double2 *)d_A_one_field));
  ctx->have_gauge = true;
  return TB_OK;
}

extern "C" int tb_set_gauge_shared(tb_ctx *ctx, const double *A_host) {
  if (!ctx || !A_host) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(join_subs(ctx));
  TB_CUDA(cudaMemcpyAsync(ctx->stage, A_host, ctx->V * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
  TB_CHECK(tb_launch_links_shared(ctx, (const double2 *)ctx->stage));
  TB_CUDA(cudaStreamSynchronize(ctx->stream));   // the staging buffer is free again
  ctx->have_gauge = true;
  return TB_OK;
}

static int need_gauge(tb_ctx *ctx) {
  if (ctx->nranks > 1 && !ctx->slab_connected) {
    tb_set_error("slab context is not connected (call tb_slab_connect on every rank first)");
    return TB_EINVAL;
  }
  if (!ctx->have_gauge) {
    tb_set_error("no gauge field set (call tb_set_gauge / tb_set_gauge_dev first)");
    return TB_EINVAL;
  }
  return TB_OK;
}

extern "C" int tb_apply_dev(tb_ctx *ctx, int op, const double *d_in, double *d_out) {
  if (!ctx || !d_in || !d_out || d_in == d_out) {
    tb_set_error("tb_apply_dev: invalid arguments (in-place apply is not supported)");
    return TB_EINVAL;
  }
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(join_subs(ctx));
  TB_CHECK(need_gauge(ctx));
  const double2 *in = (const double2 *)d_in;
  double2 *out = (double2 *)d_out;
  if (ctx->nranks > 1) {
    if (op < TB_OP_M || op > TB_OP_MDM) { tb_set_error("tb_apply_dev: unknown op %d", op); return TB_EINVAL; }
    return tb_slab_apply(ctx, op, in, out);
  }
  switch (op) {
    case TB_OP_M: return tb_launch_dslash(ctx, false, in, out, false);
    case TB_OP_MDAG: return tb_launch_dslash(ctx, true, in, out, false);
    case TB_OP_MCONJ: return tb_launch_dslash(ctx, tb_conj_is_dagger(ctx), in, out, false);
    case TB_OP_MDM:
      TB_CHECK(tb_launch_dslash(ctx, false, in, ctx->tmp, false));
      return tb_launch_dslash(ctx, tb_conj_is_dagger(ctx), ctx->tmp, out, false);
    default: tb_set_error("tb_apply_dev: unknown op %d", op); return TB_EINVAL;
  }
}

// 0 = streaming, 1 = one CTA per chain (tb_resident.cu), 2 = one thread-block cluster per chain (tb_cluster.cu)
static int onchip_solver(tb_ctx *ctx) {
  if (ctx->tune_solver == 1 || ctx->tune_solver == 5) return 0;
  if (tb_resident_supported(ctx)) return 1;
  if (tb_cluster_supported(ctx)) return 2;
  return 0;
}

static int run_onchip_slice(tb_ctx *ctx, int kind, const double2 *b, double2 *x, int c0, int n, cudaStream_t st) {
  return kind == 1 ? tb_run_cg_resident_slice(ctx, b, x, c0, n, st) : tb_run_cg_cluster_slice(ctx, b, x, c0, n, st);
}

int tb_run_cg_any(tb_ctx *ctx, const double2 *b, double2 *x) {
  TB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  // solver selection: 0 auto (on-chip when the lattice has a resident or cluster shape), 1 streaming, 2 on-chip
  const int onchip = onchip_solver(ctx);
  if (ctx->tune_solver == 2 && !onchip) {
    tb_set_error("on-chip solver requested but %dx%d is not supported", ctx->nt, ctx->nx);
    return TB_EINVAL;
  }
  if (onchip) TB_CHECK(run_onchip_slice(ctx, onchip, b, x, 0, ctx->C, ctx->stream));
  else if (ctx->tune_solver == 5) TB_CHECK(tb_run_cg_strict(ctx, b, x));   // the reference's own evaluation order
  else TB_CHECK(tb_run_cg_stream(ctx, b, x));
  TB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  TB_CUDA(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  TB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  ctx->last_solve_ms = ms;
  return TB_OK;
}

// The same solve without the event pair and the host synchronisation when an on-chip solver serves the context (one
// or two launches, nothing for the host to poll): callers that queue more work behind the solve (the device-resident
// trajectory) keep the GPU busy.  The streaming solver needs its host loop and stays synchronous.
int tb_run_cg_async(tb_ctx *ctx, const double2 *b, double2 *x) {
  const int onchip = onchip_solver(ctx);
  if (!onchip || ctx->tune_solver == 2) return tb_run_cg_any(ctx, b, x);
  return run_onchip_slice(ctx, onchip, b, x, 0, ctx->C, ctx->stream);
}

extern "C" int tb_solver_info(tb_ctx *ctx, int *kind, int *chains_in_flight) {
  if (!ctx) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  const int k = onchip_solver(ctx);
  if (kind) *kind = k;
  if (chains_in_flight) {
    int n = 0;
    if (k == 1) {
      cudaDeviceProp prop;
      TB_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
      n = prop.multiProcessorCount;   // one CTA per SM at 64^2; smaller lattices pack more
    } else if (k == 2) {
      n = tb_cluster_capacity(ctx);
    }
    *chains_in_flight = n;
  }
  return TB_OK;
}

extern "C" int tb_streaming_info(tb_ctx *ctx, int *kernels, int *tile_chains, int *tile_sites, int *rows_per_block) {
  if (!ctx) return TB_EINVAL;
  const int k = tb_stream_kernels(ctx, tile_chains, tile_sites, rows_per_block);
  if (kernels) *kernels = k;
  return TB_OK;
}

extern "C" int tb_cg_dev(tb_ctx *ctx, const double *d_b, double *d_x) {
  if (!ctx || !d_b || !d_x) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(join_subs(ctx));
  TB_CHECK(need_gauge(ctx));
  return tb_run_cg_any(ctx, (const double2 *)d_b, (double2 *)d_x);
}

extern "C" int tb_invert_dev(tb_ctx *ctx, const double *d_v, double *d_x) {
  if (!ctx || !d_v || !d_x) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(join_subs(ctx));
  TB_CHECK(need_gauge(ctx));
  // fm_invert_cg, hmc.c:408-414
  if (ctx->nranks > 1) TB_CHECK(tb_slab_apply(ctx, TB_OP_MCONJ, (const double2 *)d_v, ctx->tmp));
  else TB_CHECK(tb_launch_dslash(ctx, tb_conj_is_dagger(ctx), (const double2 *)d_v, ctx->tmp, false));
  return tb_run_cg_any(ctx, ctx->tmp, (double2 *)d_x);
}

extern "C" int tb_cg_result(tb_ctx *ctx, int *status, int *iters, double *rr) {
  if (!ctx) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(join_subs(ctx));
  const size_t c = ctx->C;
  cudaStream_t st = ctx->stream;
  TB_CUDA(cudaMemcpyAsync(ctx->h_status, ctx->cg.status, c * sizeof(int), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaMemcpyAsync(ctx->h_iters, ctx->cg.iters, c * sizeof(int), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaMemcpyAsync(ctx->h_rr, ctx->cg.rr, c * sizeof(double), cudaMemcpyDeviceToHost, st));
  TB_CHECK(sync_all(ctx));
  if (status) memcpy(status, ctx->h_status, c * sizeof(int));
  if (iters) memcpy(iters, ctx->h_iters, c * sizeof(int));
  if (rr) memcpy(rr, ctx->h_rr, c * sizeof(double));
  return TB_OK;
}

extern "C" int tb_re_dot_dev(tb_ctx *ctx, const double *d_a, const double *d_b, double *out_host) {
  if (!ctx || !d_a || !d_b || !out_host) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(join_subs(ctx));
  TB_CHECK(tb_launch_dot(ctx, (const double2 *)d_a, (const double2 *)d_b, ctx->cg.dot));
  TB_CUDA(cudaMemcpyAsync(ctx->h_rr, ctx->cg.dot, ctx->C * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  TB_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(out_host, ctx->h_rr, ctx->C * sizeof(double));
  return TB_OK;
}

// ---- host-buffer entry points ---------------------------------------------------------------------------
// Chains are processed in ctx->nsub sub-batches, each on its own stream: H2D -> re-layout -> kernels ->
// re-layout -> D2H of one sub-batch overlap the copies and kernels of the others (pinned host buffers make the
// copies asynchronous; pageable ones still work, without overlap).

// H2D + re-layout of the chains of sub-batch s on its stream
static int upload_slice(tb_ctx *ctx, int s, const double *host, double2 *d_vec) {
  int c0, n;
  sub_range(ctx, s, &c0, &n);
  if (n == 0) return TB_OK;
  const size_t off = (size_t)c0 * ctx->V;  // double2 elements
  double2 *stg = (double2 *)ctx->stage + off;
  double2 *dst = ctx->C == 1 ? d_vec : stg;
  TB_CUDA(cudaMemcpyAsync(dst, (const double2 *)host + off, (size_t)n * ctx->V * sizeof(double2),
                          cudaMemcpyHostToDevice, ctx->sub_stream[s]));
  if (ctx->C > 1) TB_CHECK(tb_launch_pack_slice(ctx, stg, d_vec, c0, n, ctx->sub_stream[s]));
  return TB_OK;
}

static int upload_vec(tb_ctx *ctx, const double *host, double2 *d_vec) {
  TB_CHECK(fork_subs(ctx));
  for (int s = 0; s < ctx->nsub; s++) TB_CHECK(upload_slice(ctx, s, host, d_vec));
  return TB_OK;
}

// queue re-layout + D2H of every sub-batch behind whatever is queued on its stream, then wait for all
static int download_vec(tb_ctx *ctx, const double2 *d_vec, double *host) {
  if (!ctx->sub_pending) TB_CHECK(fork_subs(ctx));
  for (int s = 0; s < ctx->nsub; s++) {
    int c0, n;
    sub_range(ctx, s, &c0, &n);
    if (n == 0) continue;
    const size_t off = (size_t)c0 * ctx->V;
    const double2 *src = d_vec;
    if (ctx->C > 1) {
      TB_CHECK(tb_launch_unpack_slice(ctx, d_vec, ctx->stage_x + off, c0, n, ctx->sub_stream[s]));
      src = ctx->stage_x + off;
    }
    TB_CUDA(cudaMemcpyAsync((double2 *)host + off, src, (size_t)n * ctx->V * sizeof(double2),
                            cudaMemcpyDeviceToHost, ctx->sub_stream[s]));
  }
  return sync_all(ctx);
}

// Small contexts (one sub-batch: fewer than 32 chains, e.g. the single chain of the interposed reference driver) run
// a host-buffer call on the context's stream alone -- copy in, kernels, copy out, ONE synchronisation -- instead of
// forking and joining sub-batch streams around every stage: an interposed fm_mul or fmdm_invert_cg is a handful of
// driver calls instead of some twenty-five.
static bool single_stream_path(const tb_ctx *ctx) { return ctx->nsub == 1 && ctx->nranks == 1 && !getenv("TB_NO_FASTPATH"); }

static int h2d_vec(tb_ctx *ctx, const double *host, double2 *d_vec) {
  if (ctx->C == 1) return cudaMemcpyAsync(d_vec, host, ctx->nsite * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess ? TB_OK : TB_ECUDA;
  TB_CUDA(cudaMemcpyAsync(ctx->stage, host, ctx->nsite * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
  return tb_launch_pack(ctx, ctx->stage, d_vec);
}

static int d2h_vec(tb_ctx *ctx, const double2 *d_vec, double *host) {
  const double2 *src = d_vec;
  if (ctx->C > 1) {
    TB_CHECK(tb_launch_unpack(ctx, d_vec, (double *)ctx->stage_x));
    src = ctx->stage_x;
  }
  TB_CUDA(cudaMemcpyAsync(host, src, ctx->nsite * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
  return TB_OK;
}

extern "C" int tb_set_gauge(tb_ctx *ctx, const double *A_host) {
  if (!ctx || !A_host) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  if (single_stream_path(ctx)) {
    TB_CHECK(join_subs(ctx));
    TB_CHECK(h2d_vec(ctx, A_host, ctx->Adev));
    TB_CHECK(tb_launch_links(ctx, (const double *)ctx->Adev));
    ctx->msite = nullptr;
    ctx->have_gauge = true;
    return TB_OK;
  }
  TB_CHECK(upload_vec(ctx, A_host, ctx->Adev));
  for (int s = 0; s < ctx->nsub; s++) {
    int c0, n;
    sub_range(ctx, s, &c0, &n);
    if (n) TB_CHECK(tb_launch_links_slice(ctx, ctx->Adev, c0, n, ctx->sub_stream[s]));
  }
  ctx->msite = nullptr;
  ctx->have_gauge = true;
  return TB_OK;
}

// Links from cos / sin of the angles computed by the CALLER (its libm), canonical layout double[nchains][NT][NX][2] =
// (cos A_mu, sin A_mu) per direction: the links are then bit for bit what hmc.c:140-174 evaluates on the host, which
// with the strict solver (tb_set_tuning solver = 5) makes the whole CG recursion the reference's.  The stored angles
// are not touched: use it for solves and applies, not in front of the device-resident trajectory.
extern "C" int tb_set_links_trig(tb_ctx *ctx, const double *trig_t_host, const double *trig_x_host) {
  if (!ctx || !trig_t_host || !trig_x_host) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(upload_vec(ctx, trig_t_host, ctx->vin));
  TB_CHECK(join_subs(ctx));
  TB_CUDA(cudaStreamSynchronize(ctx->stream));   // the staging buffer is reused by the second upload
  TB_CHECK(upload_vec(ctx, trig_x_host, ctx->vout));
  TB_CHECK(join_subs(ctx));
  TB_CHECK(tb_launch_links_from_trig(ctx, ctx->vin, ctx->vout));
  ctx->msite = nullptr;
  ctx->have_gauge = true;
  return TB_OK;
}

// Family B: occupation field `field` (ints, 0 = free site; vec_ops.c:107) in the canonical layout
// int[nchains][NT][NX].  Replaces the gauge field: links become the real masked constants of fM / fM_transpose
// and occupied sites identity rows.  The current tb_set_params masses are baked into the site masses.
extern "C" int tb_set_occupancy(tb_ctx *ctx, const int *field_host) {
  return tb_set_occupancy_bc(ctx, field_host, TB_BC_ANTISYMMETRIC, 0);
}

// The same with the boundary variant of Thirring.h:27-29 and, for shared != 0, ONE field int[NT][NX] for every source
// of the batch (the multi-RHS case: measure_propagator's 2 NX point sources on one configuration).
extern "C" int tb_set_occupancy_bc(tb_ctx *ctx, const int *field_host, int bc, int shared) {
  if (!ctx || !field_host || bc < TB_BC_ANTISYMMETRIC || bc > TB_BC_OPENX) return TB_EINVAL;
  if (ctx->nranks > 1) { tb_set_error("tb_set_occupancy: slab contexts are not supported"); return TB_EINVAL; }
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(sync_all(ctx));
  if (!ctx->msite_buf) {
    TB_CHECK(dev_alloc(&ctx->msite_buf, ctx->nsite));
    TB_CHECK(dev_alloc(&ctx->occ_dev, ctx->nsite));
    TB_CHECK(dev_alloc(&ctx->occ_stage, ctx->nsite));
  }
  ctx->occ_bc = bc;
  if (shared) {
    for (int c = 0; c < ctx->C; c++)
      TB_CUDA(cudaMemcpyAsync(ctx->occ_stage + (size_t)c * ctx->V, field_host, ctx->V * sizeof(int),
                              cudaMemcpyHostToDevice, ctx->stream));
  } else {
    TB_CUDA(cudaMemcpyAsync(ctx->occ_stage, field_host, ctx->nsite * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  }
  TB_CHECK(tb_launch_occupancy(ctx, ctx->occ_stage));
  ctx->msite = ctx->msite_buf;
  ctx->have_gauge = true;
  invalidate_graph(ctx);
  return TB_OK;
}

// ---- family B on REAL host vectors double[nchains][NT][NX] (tb_real.cu) ------------------------------------------
static int need_occupancy(tb_ctx *ctx) {
  if (!ctx->msite) {
    tb_set_error("family B entry point without an occupation field (call tb_set_occupancy first)");
    return TB_EINVAL;
  }
  return TB_OK;
}

// fM (TB_OP_M, vec_ops.c:96) / fM_transpose (TB_OP_MDAG, vec_ops.c:135) on real vectors, any lattice shape
extern "C" int tb_apply_real(tb_ctx *ctx, int op, const double *in_host, double *out_host) {
  if (!ctx || !in_host || !out_host || (op != TB_OP_M && op != TB_OP_MDAG)) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(need_occupancy(ctx));
  TB_CHECK(join_subs(ctx));
  double *din = ctx->stage, *dout = ctx->stage + ctx->nsite;   // stage holds 2 * nsite doubles
  TB_CUDA(cudaMemcpyAsync(din, in_host, ctx->nsite * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  TB_CHECK(tb_launch_real_apply(ctx, op == TB_OP_MDAG, din, dout));
  TB_CUDA(cudaMemcpyAsync(out_host, dout, ctx->nsite * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return sync_all(ctx);
}

// cg_MdM (propagator = 0, vec_ops.c:261) / cg_propagator (1, vec_ops.c:311) for every source of the batch.  A source
// whose solve diverges comes back filled with 1e50 (vec_ops.c:292-296) and status TB_CG_DIVERGED.  Lattices the
// on-chip real kernel serves (whole 8-row tiles, at most 4096 sites: 64 x 64, the size Thirring.h compiles in) run
// host -> H2D -> one kernel -> D2H per sub-batch of sources; other shapes go through the complex kernels.
extern "C" int tb_cg_real(tb_ctx *ctx, const double *b_host, double *x_host, int propagator, int *status, int *iters,
                          double *rr) {
  if (!ctx || !b_host || !x_host) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(need_occupancy(ctx));
  if (ctx->tune_solver != 1 && ctx->tune_solver != 5 && tb_real_cg_supported(ctx)) {
    TB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    TB_CHECK(fork_subs(ctx));
    double *db = ctx->stage, *dx = ctx->stage + ctx->nsite;
    for (int s = 0; s < ctx->nsub; s++) {
      int c0, n;
      sub_range(ctx, s, &c0, &n);
      if (n == 0) continue;
      const size_t off = (size_t)c0 * ctx->V, bytes = (size_t)n * ctx->V * sizeof(double);
      cudaStream_t st = ctx->sub_stream[s];
      TB_CUDA(cudaMemcpyAsync(db + off, b_host + off, bytes, cudaMemcpyHostToDevice, st));
      TB_CHECK(tb_run_cg_real(ctx, db, dx, propagator != 0, c0, n, st));
      TB_CUDA(cudaMemcpyAsync(x_host + off, dx + off, bytes, cudaMemcpyDeviceToHost, st));
    }
    TB_CHECK(sync_all(ctx));
    TB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    TB_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    TB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_solve_ms = ms;
    if (status || iters || rr) TB_CHECK(tb_cg_result(ctx, status, iters, rr));
    return TB_OK;
  }
  // other shapes: the complex kernels on vectors with zero imaginary parts
  const size_t n = ctx->nsite;
  double *cb = (double *)malloc(4 * n * sizeof(double));
  if (!cb) return TB_ENOMEM;
  double *cx = cb + 2 * n;
  for (size_t i = 0; i < n; i++) { cb[2 * i] = b_host[i]; cb[2 * i + 1] = 0.0; }
  int *st = (int *)malloc(ctx->C * sizeof(int));
  int rc = propagator ? tb_invert(ctx, cb, cx, st, iters, rr) : tb_cg(ctx, cb, cx, st, iters, rr);
  if (rc == TB_OK) {
    for (size_t i = 0; i < n; i++) x_host[i] = st[i / ctx->V] == TB_CG_DIVERGED ? 1e50 : cx[2 * i];
    if (status) memcpy(status, st, ctx->C * sizeof(int));
  }
  free(st);
  free(cb);
  return rc;
}

extern "C" int tb_apply(tb_ctx *ctx, int op, const double *in_host, double *out_host) {
  if (!ctx || !in_host || !out_host) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(need_gauge(ctx));
  if (single_stream_path(ctx)) {
    TB_CHECK(join_subs(ctx));
    TB_CHECK(h2d_vec(ctx, in_host, ctx->vin));
    TB_CHECK(tb_apply_dev(ctx, op, (const double *)ctx->vin, (double *)ctx->vout));
    TB_CHECK(d2h_vec(ctx, ctx->vout, out_host));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    return TB_OK;
  }
  TB_CHECK(upload_vec(ctx, in_host, ctx->vin));
  TB_CHECK(tb_apply_dev(ctx, op, (const double *)ctx->vin, (double *)ctx->vout));  // joins the sub-streams
  return download_vec(ctx, ctx->vout, out_host);
}

// Host buffers straight through the 64^2 on-chip kernel (tb_resident.cu, CANON): per sub-batch of chains one H2D copy
// of the source (and of the angles, when the call brings a new gauge field), ONE kernel that reads and writes the
// canonical layout and builds its own links, one D2H copy of the solution.  No re-layout or link kernel has to find a
// free SM between the solver CTAs of the other sub-batches.
static int solve_host_canon(tb_ctx *ctx, const double *A_host, const double *b_host, double *x_host) {
  TB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  TB_CHECK(fork_subs(ctx));
  double2 *Ac = (double2 *)ctx->stage, *bc = ctx->vin, *xc = ctx->stage_x;   // canonical device staging
  for (int s = 0; s < ctx->nsub; s++) {
    int c0, n;
    sub_range(ctx, s, &c0, &n);
    if (n == 0) continue;
    const size_t off = (size_t)c0 * ctx->V, bytes = (size_t)n * ctx->V * sizeof(double2);
    cudaStream_t st = ctx->sub_stream[s];
    if (A_host) TB_CUDA(cudaMemcpyAsync(Ac + off, (const double2 *)A_host + off, bytes, cudaMemcpyHostToDevice, st));
    TB_CUDA(cudaMemcpyAsync(bc + off, (const double2 *)b_host + off, bytes, cudaMemcpyHostToDevice, st));
    TB_CHECK(tb_run_cg_resident_canon(ctx, bc, xc, A_host ? Ac : nullptr, c0, n, st));
    TB_CUDA(cudaMemcpyAsync((double2 *)x_host + off, xc + off, bytes, cudaMemcpyDeviceToHost, st));
  }
  TB_CHECK(sync_all(ctx));
  TB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  TB_CUDA(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  TB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  ctx->last_solve_ms = ms;
  return TB_OK;
}

// the solve of a sub-batch starts as soon as ITS links and sources are on the device
static int solve_host(tb_ctx *ctx, bool with_conj, const double *b_host, double *x_host, int *status, int *iters,
                      double *rr) {
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(need_gauge(ctx));
  const int onchip = onchip_solver(ctx);
  if (ctx->tune_solver == 2 && !onchip) {
    tb_set_error("on-chip solver requested but %dx%d is not supported", ctx->nt, ctx->nx);
    return TB_EINVAL;
  }
  if (single_stream_path(ctx) && onchip) {
    cudaStream_t st = ctx->stream;
    const size_t c = ctx->C;
    TB_CHECK(join_subs(ctx));
    TB_CUDA(cudaEventRecord(ctx->ev0, st));
    TB_CHECK(h2d_vec(ctx, b_host, ctx->vin));
    const double2 *src = ctx->vin;
    if (with_conj) {   // fm_invert_cg, hmc.c:408-414
      TB_CHECK(tb_launch_dslash(ctx, tb_conj_is_dagger(ctx), ctx->vin, ctx->tmp, false));
      src = ctx->tmp;
    }
    TB_CHECK(run_onchip_slice(ctx, onchip, src, ctx->vout, 0, ctx->C, st));
    TB_CHECK(d2h_vec(ctx, ctx->vout, x_host));
    TB_CUDA(cudaMemcpyAsync(ctx->h_status, ctx->cg.status, c * sizeof(int), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaMemcpyAsync(ctx->h_iters, ctx->cg.iters, c * sizeof(int), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaMemcpyAsync(ctx->h_rr, ctx->cg.rr, c * sizeof(double), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaEventRecord(ctx->ev1, st));
    TB_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    TB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_solve_ms = ms;
    if (status) memcpy(status, ctx->h_status, c * sizeof(int));
    if (iters) memcpy(iters, ctx->h_iters, c * sizeof(int));
    if (rr) memcpy(rr, ctx->h_rr, c * sizeof(double));
    return TB_OK;
  }
  if (onchip == 1 && !with_conj && tb_resident_canon_supported(ctx)) {
    TB_CHECK(solve_host_canon(ctx, nullptr, b_host, x_host));
  } else if (!onchip || with_conj) {
    TB_CHECK(upload_vec(ctx, b_host, ctx->vin));
    if (with_conj) TB_CHECK(tb_invert_dev(ctx, (const double *)ctx->vin, (double *)ctx->vout));
    else TB_CHECK(tb_cg_dev(ctx, (const double *)ctx->vin, (double *)ctx->vout));
    TB_CHECK(download_vec(ctx, ctx->vout, x_host));
  } else {
    TB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    TB_CHECK(upload_vec(ctx, b_host, ctx->vin));
    for (int s = 0; s < ctx->nsub; s++) {
      int c0, n;
      sub_range(ctx, s, &c0, &n);
      if (n) TB_CHECK(run_onchip_slice(ctx, onchip, ctx->vin, ctx->vout, c0, n, ctx->sub_stream[s]));
    }
    TB_CHECK(download_vec(ctx, ctx->vout, x_host));
    TB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    TB_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    TB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_solve_ms = ms;
  }
  if (status || iters || rr) TB_CHECK(tb_cg_result(ctx, status, iters, rr));
  return TB_OK;
}

// tb_set_gauge + tb_cg in one call, interleaved per sub-batch: links and sources of sub-batch s travel back to back,
// so its solve starts after 1/nsub of the input has crossed PCIe instead of after the whole gauge field.
extern "C" int tb_cg_gauge(tb_ctx *ctx, const double *A_host, const double *b_host, double *x_host, int *status,
                           int *iters, double *rr) {
  if (!ctx || !A_host || !b_host || !x_host) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  if (ctx->nranks > 1) { tb_set_error("tb_cg_gauge: slab contexts use tb_set_gauge + tb_cg"); return TB_EINVAL; }
  ctx->msite = nullptr;
  ctx->have_gauge = true;
  const int onchip = onchip_solver(ctx);
  if (!onchip) {   // the streaming solver works on the whole batch at once: nothing to interleave
    TB_CHECK(tb_set_gauge(ctx, A_host));
    return solve_host(ctx, false, b_host, x_host, status, iters, rr);
  }
  if (onchip == 1 && tb_resident_canon_supported(ctx)) {
    TB_CHECK(solve_host_canon(ctx, A_host, b_host, x_host));
    if (status || iters || rr) TB_CHECK(tb_cg_result(ctx, status, iters, rr));
    return TB_OK;
  }
  TB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  TB_CHECK(fork_subs(ctx));
  // stage is shared by the two uploads of a sub-batch: the copy of b into it is ordered behind the pack of A on
  // the same stream
  for (int s = 0; s < ctx->nsub; s++) {
    int c0, n;
    sub_range(ctx, s, &c0, &n);
    if (n == 0) continue;
    TB_CHECK(upload_slice(ctx, s, A_host, ctx->Adev));
    TB_CHECK(tb_launch_links_slice(ctx, ctx->Adev, c0, n, ctx->sub_stream[s]));
    TB_CHECK(upload_slice(ctx, s, b_host, ctx->vin));
    TB_CHECK(run_onchip_slice(ctx, onchip, ctx->vin, ctx->vout, c0, n, ctx->sub_stream[s]));
  }
  TB_CHECK(download_vec(ctx, ctx->vout, x_host));
  TB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  TB_CUDA(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  TB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  ctx->last_solve_ms = ms;
  if (status || iters || rr) TB_CHECK(tb_cg_result(ctx, status, iters, rr));
  return TB_OK;
}

extern "C" int tb_cg(tb_ctx *ctx, const double *b_host, double *x_host, int *status, int *iters, double *rr) {
  if (!ctx || !b_host || !x_host) return TB_EINVAL;
  return solve_host(ctx, false, b_host, x_host, status, iters, rr);
}

extern "C" int tb_invert(tb_ctx *ctx, const double *v_host, double *x_host, int *status, int *iters,
                         double *rr) {
  if (!ctx || !v_host || !x_host) return TB_EINVAL;
  return solve_host(ctx, true, v_host, x_host, status, iters, rr);
}
