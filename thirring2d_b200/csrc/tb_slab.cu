// tb_slab.cu — slab decomposition plumbing: the IPC-exported exchange block of a rank and the peer mapping.
// The kernels that use it are in tb_stream.cu (SLAB template flag); see TbSlab in tb_common.cuh.
#include "tb_common.cuh"

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct SlabOffsets {
  size_t p, mp, w0, r, p1, red, red_flag, flags, red2, red3, bcast, total;
  int rep_stride, bcast_stride;
};

static SlabOffsets slab_offsets(const tb_ctx *ctx) {
  SlabOffsets o;
  const size_t vec = align_up(ctx->nsite * sizeof(double2), 256);
  o.p = 0;
  o.mp = vec;
  o.w0 = 2 * vec;
  o.r = 3 * vec;
  o.p1 = 4 * vec;
  o.red = 5 * vec;
  o.red_flag = o.red + align_up((size_t)TB_NRED * ctx->nranks * ctx->g.Cpad * sizeof(double), 256);
  o.flags = o.red_flag + align_up((size_t)TB_NRED * ctx->nranks * ctx->g.nctiles * sizeof(int), 256);
  o.red2 = o.flags + 256;
  o.red3 = o.red2 + align_up((size_t)TB_NRED * ctx->nranks * ctx->g.Cpad * sizeof(double2), 256);
  o.rep_stride = (int)align_up((size_t)TB_NRED * ctx->nranks * ctx->g.Cpad, 16);   // double2 elements: whole 256-byte blocks
  o.bcast = o.red3 + (size_t)TB_SLAB_NREP_MAX * o.rep_stride * sizeof(double2);
  o.bcast_stride = (int)align_up((size_t)TB_NRED * ctx->g.Cpad, 16);
  o.total = o.bcast + (size_t)TB_SLAB_NREP_MAX * o.bcast_stride * sizeof(double2);
  return o;
}

// allocate the exchange block and point ctx->p / Mp / W0 into it (called by tb_create_common, nranks > 1)
int tb_slab_layout(tb_ctx *ctx) {
  const SlabOffsets o = slab_offsets(ctx);
  cudaError_t e = cudaMalloc(&ctx->slab_block, o.total);
  if (e != cudaSuccess) {
    tb_set_error("cudaMalloc(slab block, %zu bytes) failed: %s", o.total, cudaGetErrorString(e));
    return TB_ENOMEM;
  }
  TB_CUDA(cudaMemset(ctx->slab_block, 0, o.total));
  ctx->slab_bytes = o.total;
  char *base = (char *)ctx->slab_block;
  ctx->p = (double2 *)(base + o.p);
  ctx->Mp = (double2 *)(base + o.mp);
  ctx->W0 = (double2 *)(base + o.w0);
  ctx->r = (double2 *)(base + o.r);
  ctx->p1 = (double2 *)(base + o.p1);
  int *local = nullptr;
  e = cudaMalloc((void **)&local, (2 + TB_NFLAGS + 6) * sizeof(int));
  if (e != cudaSuccess) {
    tb_set_error("cudaMalloc failed: %s", cudaGetErrorString(e));
    return TB_ENOMEM;
  }
  TB_CUDA(cudaMemset(local, 0, (2 + TB_NFLAGS + 6) * sizeof(int)));
  ctx->slab.seq = local;
  ctx->slab.done_ticket = (unsigned int *)(local + 1);
  ctx->slab.err = local + 1 + TB_NFLAGS;
  ctx->slab.gbar = (unsigned long long *)(local + 8);   // 8-byte aligned: ints 8, 9
  ctx->slab.go = local + 10;
  ctx->slab.P = ctx->nranks;
  ctx->slab.rank = ctx->rank;
  return TB_OK;
}

extern "C" int tb_create_slab(tb_ctx **out, int nt_global, int nx, int nchains, int mode, int device, int rank,
                              int nranks) {
  if (nranks < 2 || nranks > TB_SLAB_MAX_RANKS || rank < 0 || rank >= nranks || nt_global % nranks != 0 ||
      nt_global / nranks < 2) {
    tb_set_error("tb_create_slab: need 2 <= nranks <= %d, NT divisible by nranks and >= 2 rows per rank",
                 TB_SLAB_MAX_RANKS);
    return TB_EINVAL;
  }
  return tb_create_common(out, nt_global / nranks, nx, nchains, mode, device, rank, nranks, nt_global);
}

extern "C" int tb_slab_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int tb_slab_export(tb_ctx *ctx, void *handle_out) {
  if (!ctx || !handle_out || ctx->nranks < 2) {
    tb_set_error("tb_slab_export: not a slab context");
    return TB_EINVAL;
  }
  TB_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  TB_CUDA(cudaIpcGetMemHandle(&h, ctx->slab_block));
  memcpy(handle_out, &h, sizeof(h));
  return TB_OK;
}

// all_handles: nranks handles in rank order (e.g. from an all_gather).  Every rank must have created its
// context before any rank connects, and a host barrier must follow before the first collective call.
extern "C" int tb_slab_connect(tb_ctx *ctx, const void *all_handles) {
  if (!ctx || !all_handles || ctx->nranks < 2) {
    tb_set_error("tb_slab_connect: not a slab context");
    return TB_EINVAL;
  }
  TB_CUDA(cudaSetDevice(ctx->device));
  const SlabOffsets o = slab_offsets(ctx);
  const int P = ctx->nranks;
  for (int q = 0; q < P; q++) {
    if (q == ctx->rank) {
      ctx->peer_block[q] = ctx->slab_block;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)all_handles + (size_t)q * sizeof(h), sizeof(h));
    TB_CUDA(cudaIpcOpenMemHandle(&ctx->peer_block[q], h, cudaIpcMemLazyEnablePeerAccess));
  }
  const int prev = (ctx->rank + P - 1) % P, next = (ctx->rank + 1) % P;
  TbSlab &sl = ctx->slab;
  char *me = (char *)ctx->slab_block, *pv = (char *)ctx->peer_block[prev], *nx = (char *)ctx->peer_block[next];
  sl.p_prev = (const double2 *)(pv + o.p);
  sl.p_next = (const double2 *)(nx + o.p);
  sl.mp_prev = (const double2 *)(pv + o.mp);
  sl.mp_next = (const double2 *)(nx + o.mp);
  sl.W0_prev = (const double2 *)(pv + o.w0);
  sl.r_prev = (const double2 *)(pv + o.r);
  sl.r_next = (const double2 *)(nx + o.r);
  sl.p1_prev = (const double2 *)(pv + o.p1);
  sl.p1_next = (const double2 *)(nx + o.p1);
  sl.flags = (volatile int *)(me + o.flags);
  sl.sig_prev = (int *)(pv + o.flags) + 1;  // the previous rank sees me as its "next" neighbour
  sl.sig_next = (int *)(nx + o.flags) + 0;  // the next rank sees me as its "previous" neighbour
  sl.red = (double *)(me + o.red);
  sl.red_flag = (volatile int *)(me + o.red_flag);
  sl.red2 = (double2 *)(me + o.red2);
  sl.red3 = (double2 *)(me + o.red3);
  sl.rep_stride = o.rep_stride;
  sl.bcast = (double2 *)(me + o.bcast);
  sl.bcast_stride = o.bcast_stride;
  sl.nrep = TB_SLAB_NREP_MAX;
  for (int q = 0; q < P; q++) {
    sl.peer_red[q] = (double *)((char *)ctx->peer_block[q] + o.red);
    sl.peer_red_flag[q] = (int *)((char *)ctx->peer_block[q] + o.red_flag);
    sl.peer_red2[q] = (double2 *)((char *)ctx->peer_block[q] + o.red2);
    sl.peer_red3[q] = (double2 *)((char *)ctx->peer_block[q] + o.red3);
  }
  ctx->slab_connected = true;
  return TB_OK;
}

void tb_slab_release(tb_ctx *ctx) {
  if (ctx->nranks < 2) return;
  for (int q = 0; q < ctx->nranks; q++)
    if (q != ctx->rank && ctx->peer_block[q]) cudaIpcCloseMemHandle(ctx->peer_block[q]);
  if (ctx->slab.seq) cudaFree(ctx->slab.seq);
  if (ctx->slab.timeline) cudaFree(ctx->slab.timeline);
  if (ctx->slab_block) cudaFree(ctx->slab_block);
}
