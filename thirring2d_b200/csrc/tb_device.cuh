// tb_device.cuh — device helpers shared by the streaming and HMC kernels: block geometry, the deterministic
// per-chain reduction with last-block finalisation, and the slab-mode peer-flag primitives.
// Include inside an anonymous namespace of a .cu file.
#pragma once
#include "tb_common.cuh"

enum { FIN_DOT = 0, FIN_INIT = 1, FIN_PQ = 2, FIN_RR = 3 };

struct BlockPos {
  int ctile, xtile, ttile;
  int c_local, x_local;
  int c, x;
  bool valid;
};

__device__ __forceinline__ BlockPos block_pos(const TbGeom &g) {
  BlockPos b;
  b.ctile = blockIdx.x;
  b.xtile = blockIdx.y;
  b.ttile = blockIdx.z;
  b.c_local = threadIdx.x & (g.bc - 1);
  b.x_local = threadIdx.x >> g.bc_shift;
  b.c = b.ctile * g.bc + b.c_local;
  b.x = b.xtile * g.bx + b.x_local;
  b.valid = (b.c < g.C) && (b.x < g.nx);
  return b;
}

// ---- slab-mode primitives (peer flags over NVLink) ------------------------------------------------------
// spin until the flag written by the neighbour on `side` (0 = previous rank, 1 = next rank) reaches `need`
// A wait that outlives TB_SLAB_SPIN_CYCLES (seconds: a peer died or the collective call order differs between
// ranks) records an error instead of hanging the GPU; the host reports it at the next synchronisation.
#define TB_SLAB_SPIN_CYCLES 6000000000LL
__device__ __forceinline__ void slab_spin(volatile int *f, int need, int *err, int code) {
  const long long t0 = clock64();
  while (*f < need) {
    __nanosleep(64);
    if (clock64() - t0 > TB_SLAB_SPIN_CYCLES) {
      atomicExch(err, code);
      break;
    }
  }
  __threadfence_system();
}

__device__ __forceinline__ void slab_wait(const TbSlab &sl, int kind, int side, int need) {
  if (threadIdx.x == 0) slab_spin(sl.flags + kind * 2 + side, need, sl.err, 1 + kind);
  __syncthreads();
}

// every block calls this after its last global store; the last block of the grid publishes `value` into
// both neighbours' flag `kind`
__device__ __forceinline__ void slab_signal_done(const TbSlab &sl, int kind0, int kind1, int value) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int tk = atomicAdd(&sl.done_ticket[kind0], 1u);
    if (tk == nblocks - 1) {
      sl.done_ticket[kind0] = 0u;
      __threadfence_system();
      *(volatile int *)(sl.sig_prev + kind0 * 2) = value;
      *(volatile int *)(sl.sig_next + kind0 * 2) = value;
      if (kind1 >= 0) {
        *(volatile int *)(sl.sig_prev + kind1 * 2) = value;
        *(volatile int *)(sl.sig_next + kind1 * 2) = value;
      }
      __threadfence_system();
    }
  }
}

template <bool SLAB>
__device__ __forceinline__ double2 ld_halo(const double2 *p) {
  return SLAB ? __ldcv(p) : *p;   // peer memory: never serve a halo row from a stale L1 line
}

// CG scalar update for chain c from the global sum `total` (shared by the local and the slab path)
template <int FIN>
__device__ __forceinline__ void finalize_scalar(double total, int c, int ctile, const TbCgState &s) {
  if (FIN == FIN_DOT) {
    s.dot[c] = total;
  } else if (FIN == FIN_INIT) {
    s.rr[c] = total;
    s.rr_old[c] = total;
    s.rr_init[c] = total;
    if (total < s.accuracy) {  // hmc.c:359-361, x is already zero
      s.status[c] = TB_CG_ZERO_SOURCE;
      s.active[c] = 0;
      atomicSub(&s.tile_active[ctile], 1);
      atomicSub(s.n_active, 1);
    }
  } else if (FIN == FIN_PQ) {
    if (s.active[c]) {
      s.pq[c] = total;
      s.alpha[c] = s.rr_old[c] / total;  // hmc.c:371
    }
  } else if (FIN == FIN_RR) {
    if (s.active[c]) {
      const int it = s.iters[c] + 1;
      s.iters[c] = it;
      s.rr[c] = total;
      int st = -1;
      if (total < s.accuracy) st = TB_CG_CONVERGED;                                                  // hmc.c:381
      else if (!(total == total) || total / s.rr_init[c] > TB_DIVERGENCE_RATIO) st = TB_CG_DIVERGED;  // hmc.c:383
      else if (it >= s.max_iter - 1) st = TB_CG_MAXITER;                                             // hmc.c:364
      if (st >= 0) {
        s.status[c] = st;
        s.active[c] = 0;
        atomicSub(&s.tile_active[ctile], 1);
        atomicSub(s.n_active, 1);
      } else {
        s.beta[c] = total / s.rr_old[c];  // hmc.c:390
        s.rr_old[c] = total;              // hmc.c:394
      }
    }
  }
}

// Block-level deterministic reduction of `acc` over the x_local index for every chain of the tile, then
// ticket-based cross-block reduction by the last block of the chain tile, which either evaluates the CG scalar
// update (single GPU) or publishes this rank's partial to every rank's slot table (slab mode; RED = slot kind).
template <int FIN, bool SLAB, int RED>
__device__ __forceinline__ void reduce_finalize(double acc, const TbGeom &g, const TbCgState &s,
                                                const TbSlab &sl, const BlockPos &b, double *red) {
  __shared__ unsigned int s_last;
  const int idx = b.x_local * g.bc + b.c_local;
  red[idx] = acc;
  __syncthreads();
  for (int st = g.bx >> 1; st > 0; st >>= 1) {
    if (b.x_local < st) red[idx] += red[idx + st * g.bc];
    __syncthreads();
  }
  const int slot = b.ttile * g.nxtiles + b.xtile;
  const int cp = b.ctile * g.bc + b.c_local;  // < Cpad
  if (b.x_local == 0) {
    s.partial[(size_t)slot * g.Cpad + cp] = red[b.c_local];
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int tk = atomicAdd(&s.ticket[b.ctile], 1u);
    s_last = (tk == (unsigned int)(g.nslots - 1));
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // slot order is fixed (deterministic); the loads of a batch are independent so their L2 latency overlaps
  double sum = 0.0;
  {
    constexpr int NB = 16;
    const double *pp = s.partial + cp;
    const size_t stride = (size_t)g.bx * g.Cpad;
    int sl_ = b.x_local;
    for (; sl_ + (NB - 1) * g.bx < g.nslots; sl_ += NB * g.bx) {
      double v[NB];
      const double *q = pp + (size_t)sl_ * g.Cpad;
#pragma unroll
      for (int u = 0; u < NB; u++) v[u] = __ldcg(q + u * stride);
#pragma unroll
      for (int u = 0; u < NB; u++) sum += v[u];
    }
    for (; sl_ < g.nslots; sl_ += g.bx) sum += __ldcg(pp + (size_t)sl_ * g.Cpad);
  }
  red[idx] = sum;
  __syncthreads();
  for (int st = g.bx >> 1; st > 0; st >>= 1) {
    if (b.x_local < st) red[idx] += red[idx + st * g.bc];
    __syncthreads();
  }
  if (threadIdx.x == 0) s.ticket[b.ctile] = 0u;
  if (!SLAB) {
    if (b.x_local == 0 && b.c < g.C) finalize_scalar<FIN>(red[b.c_local], b.c, b.c >> g.ta_shift, s);
  } else {
    const int seq = *sl.seq;
    if (b.x_local == 0) {
      const double total = red[b.c_local];
      for (int q = 0; q < sl.P; q++) sl.peer_red[q][(size_t)(RED * sl.P + sl.rank) * g.Cpad + cp] = total;
      __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      for (int q = 0; q < sl.P; q++)
        *(volatile int *)(sl.peer_red_flag[q] + (RED * sl.P + sl.rank) * g.nctiles + b.ctile) = seq;
      __threadfence_system();
    }
  }
}


static inline dim3 grid_of(const TbGeom &g) { return dim3(g.nctiles, g.nxtiles, g.nttiles); }

#define TB_DISPATCH_TT(tt, EXPR)            \
  switch (tt) {                             \
    case 1: { constexpr int TT = 1; EXPR; } break;   \
    case 2: { constexpr int TT = 2; EXPR; } break;   \
    case 4: { constexpr int TT = 4; EXPR; } break;   \
    case 8: { constexpr int TT = 8; EXPR; } break;   \
    default: { constexpr int TT = 16; EXPR; } break; \
  }

