// tb_stream.cu — streaming (HBM/L2-bound) kernels of the batched fermion solve, sm_100a, FP64.
//
//   links      A -> W_mu = s * 1/2 * eta_mu * exp(iA_mu)            replaces hmc.c:140-141,152-153,163-164,173-174
//   dslash     out = M in | M^dagger in, optional fused Re<aux,out>  replaces fm_mul hmc.c:132-183,
//                                                                    fm_conjugate_mul hmc.c:197-248, dot hmc.c:368-370
//   axpy_norm  x += a p ; r -= a q ; ||r||^2                          replaces hmc.c:372-379
//   xpay       p = r + b p                                            replaces hmc.c:390-392
//   cg_init    x = 0 ; r = p = b ; ||b||^2                            replaces hmc.c:349-361
//
// Layout: double2 field[t][x][c], chain index c fastest.  A thread owns one (x,c) column position and
// marches over TT rows in t keeping psi(t-1), psi(t), psi(t+1) and W0(t-1) in registers, so the t-neighbours
// and the backward t-link cost no extra traffic; x-neighbours are the centre loads of the neighbouring
// threads of the same block (L1) or of the neighbouring block (L2).
//
// Reductions are per chain and deterministic: fixed-shape tree inside the block, one partial per
// (block, chain), and the LAST block of each chain tile (atomic ticket) sums the partials in slot order and
// evaluates the CG scalar update, so alpha/beta/convergence never leave the device and no FP64 atomics in
// arbitrary order are used.
#include <cuda.h>   // CUtensorMap and its enums only: the encoder is fetched through the runtime, libcuda is not linked
#include <cstdint>
#include <cstring>

#include "tb_common.cuh"

namespace {

#include "tb_device.cuh"

// ---------------------------------------------------------------------------------------------------------
// Dirac apply.  DAG: M^dagger instead of M.  DOT: accumulate Re<aux,out> (alpha via FIN_PQ when MASKED, plain dot
// otherwise).  MASKED: skip chains whose CG has finished.  SLAB: rows t-1 of the first local row and t+1 of the
// last one live in the neighbour ranks' HBM (in_prev / in_next, W0_prev).  Without SLAB the same pointers
// are the local arrays and the indexing is the periodic wrap.
//   SLAB protocol (epoch E = *sl.seq): wait flag `wait_ready` >= E on the sides this block touches, and
//   `wait_done` >= E-1 before overwriting `out` rows a neighbour may still be reading; when the whole grid
//   is done, publish `sig0`/`sig1` = E to both neighbours.
template <int TT, bool DAG, bool DOT, bool MASKED, bool SLAB>
__global__ void __launch_bounds__(TB_MAX_BLOCK)
dslash_kernel(const double2 *__restrict__ in, const double2 *in_prev, const double2 *in_next,
              double2 *__restrict__ out, const double2 *__restrict__ W0, const double2 *W0_prev,
              const double2 *__restrict__ W1, const double *__restrict__ mass,
              const double *__restrict__ msite, const double *__restrict__ emu, const double *__restrict__ emmu,
              const double2 *__restrict__ aux, const TbGeom g, const TbCgState s, const TbSlab sl,
              const int wait_ready, const int wait_done, const int sig0, const int sig1) {
  __shared__ double red[TB_MAX_BLOCK];
  const BlockPos b = block_pos(g);
  if (MASKED && s.tile_active[b.ctile] == 0) {
    // a finished chain tile still takes part in the grid-completion ticket while other tiles iterate
    if (SLAB && *s.n_active > 0) slab_signal_done(sl, sig0, sig1, *sl.seq);
    return;
  }
  int seq = 0;
  if (SLAB) {
    seq = *sl.seq;
    if (b.ttile == 0) {
      slab_wait(sl, wait_ready, 0, seq);
      if (wait_done >= 0) slab_wait(sl, wait_done, 0, seq - 1);
    }
    if (b.ttile == g.nttiles - 1) {
      slab_wait(sl, wait_ready, 1, seq);
      if (wait_done >= 0) slab_wait(sl, wait_done, 1, seq - 1);
    }
  }
  bool act = b.valid;
  if (MASKED && act) act = s.active[b.c] != 0;
  double acc = 0.0;
  if (act) {
    const double m = mass[b.c];
    const double af = DAG ? emmu[b.c] : emu[b.c];  // factor on the +t hop (hmc.c:144 / adjoint)
    const double ab = DAG ? emu[b.c] : emmu[b.c];  // factor on the -t hop (hmc.c:156 / adjoint)
    const size_t R = (size_t)g.R;
    const size_t j = (size_t)b.x * g.C + b.c;
    const size_t jp = (size_t)((b.x + 1 == g.nx) ? 0 : b.x + 1) * g.C + b.c;
    const size_t jm = (size_t)((b.x == 0) ? g.nx - 1 : b.x - 1) * g.C + b.c;
    const int t0 = b.ttile * TT;
    double2 pm, w0m;
    if (t0 == 0) {
      pm = ld_halo<SLAB>(&in_prev[(size_t)(g.nt - 1) * R + j]);
      w0m = ld_halo<SLAB>(&W0_prev[(size_t)(g.nt - 1) * R + j]);
    } else {
      pm = in[(size_t)(t0 - 1) * R + j];
      w0m = W0[(size_t)(t0 - 1) * R + j];
    }
    double2 pc = in[t0 * R + j];
#pragma unroll
    for (int i = 0; i < TT; i++) {
      const int t = t0 + i;
      if (t < g.nt) {
        const size_t row = t * R;
        const double2 pp = (t + 1 == g.nt) ? ld_halo<SLAB>(&in_next[j]) : in[row + R + j];
        const double2 pxp = in[row + jp];
        const double2 pxm = in[row + jm];
        const double2 w0c = W0[row + j];
        const double2 w1c = W1[row + j];
        const double2 w1m = W1[row + jm];
        // hops: +af W0(n) psi(n+t) - ab conj(W0(n-t)) psi(n-t) + W1(n) psi(n+x) - conj(W1(n-x)) psi(n-x)
        const double fr = w0c.x * af, fi = w0c.y * af;
        const double br = w0m.x * ab, bi = w0m.y * ab;
        double hr = fr * pp.x - fi * pp.y;
        double hi = fr * pp.y + fi * pp.x;
        hr -= br * pm.x + bi * pm.y;
        hi -= br * pm.y - bi * pm.x;
        hr += w1c.x * pxp.x - w1c.y * pxp.y;
        hi += w1c.x * pxp.y + w1c.y * pxp.x;
        hr -= w1m.x * pxm.x + w1m.y * pxm.y;
        hi -= w1m.x * pxm.y - w1m.y * pxm.x;
        // family B (vec_ops.c:107,130): an occupied site is an identity row -> per-site mass, links already masked
        const double ms = msite ? msite[row + j] : m;
        double2 o;
        if (DAG) {
          o.x = ms * pc.x - hr;
          o.y = ms * pc.y - hi;
        } else {
          o.x = ms * pc.x + hr;
          o.y = ms * pc.y + hi;
        }
        out[row + j] = o;
        if (DOT) {
          if (aux) {
            const double2 a = aux[row + j];
            acc += a.x * o.x + a.y * o.y;
          } else {
            acc += o.x * o.x + o.y * o.y;   // |out|^2 (fused variant: <p, M^dagger M p> = |M p|^2)
          }
        }
        pm = pc;
        pc = pp;
        w0m = w0c;
      }
    }
  }
  if (DOT) reduce_finalize<MASKED ? FIN_PQ : FIN_DOT, SLAB, TB_RED_PQ>(acc, g, s, sl, b, red);
  if (SLAB) slab_signal_done(sl, sig0, sig1, seq);
}

// x += alpha p ; r -= alpha q ; rr = ||r||^2 ; then beta / convergence (hmc.c:372-394)
template <int TT, bool SLAB>
__global__ void __launch_bounds__(TB_MAX_BLOCK)
axpy_norm_kernel(double2 *__restrict__ x, double2 *__restrict__ r, const double2 *__restrict__ p,
                 const double2 *__restrict__ q, const TbGeom g, const TbCgState s, const TbSlab sl) {
  __shared__ double red[TB_MAX_BLOCK];
  const BlockPos b = block_pos(g);
  if (s.tile_active[b.ctile] == 0) return;
  const bool act = b.valid && s.active[b.c] != 0;
  double acc = 0.0;
  if (act) {
    const double a = s.alpha[b.c];
    const size_t R = (size_t)g.R;
    const size_t j = (size_t)b.x * g.C + b.c;
    const int t0 = b.ttile * TT;
#pragma unroll
    for (int i = 0; i < TT; i++) {
      const int t = t0 + i;
      if (t < g.nt) {
        const size_t k = t * R + j;
        double2 xv = x[k], rv = r[k];
        const double2 pv = p[k], qv = q[k];
        xv.x += a * pv.x;
        xv.y += a * pv.y;
        rv.x -= a * qv.x;
        rv.y -= a * qv.y;
        x[k] = xv;
        r[k] = rv;
        acc += rv.x * rv.x + rv.y * rv.y;
      }
    }
  }
  reduce_finalize<FIN_RR, SLAB, TB_RED_RR>(acc, g, s, sl, b, red);
}

// Fused variant for M~ = M^dagger: q = M^dagger (Mp) is consumed in registers, never written:
//   x += alpha p ; r -= alpha q ; rr = ||r||^2 ; beta / convergence by the last block        (hmc.c:367-390)
// alpha = rr_old / |Mp|^2 was finalised by the preceding dslash (<p, M^dagger M p> = |M p|^2 exactly when
// M~ is the true adjoint).  One CG iteration = 240 B/site in 3 launches instead of 288 B/site in 4.
template <int TT, bool SLAB>
__global__ void __launch_bounds__(TB_MAX_BLOCK)
dslash_axpy_norm_kernel(const double2 *__restrict__ in, const double2 *in_prev, const double2 *in_next,
                        const double2 *__restrict__ W0, const double2 *W0_prev,
                        const double2 *__restrict__ W1, const double *__restrict__ mass,
                        const double *__restrict__ msite, const double *__restrict__ emu,
                        const double *__restrict__ emmu, const double2 *__restrict__ p, double2 *__restrict__ x,
                        double2 *__restrict__ r, const TbGeom g, const TbCgState s, const TbSlab sl) {
  __shared__ double red[TB_MAX_BLOCK];
  const BlockPos b = block_pos(g);
  if (s.tile_active[b.ctile] == 0) {
    if (SLAB && *s.n_active > 0) slab_signal_done(sl, TB_FLAG_MPDONE, -1, *sl.seq);
    return;
  }
  int seq = 0;
  if (SLAB) {
    seq = *sl.seq;
    if (b.ttile == 0) slab_wait(sl, TB_FLAG_MPREADY, 0, seq);
    if (b.ttile == g.nttiles - 1) slab_wait(sl, TB_FLAG_MPREADY, 1, seq);
  }
  const bool act = b.valid && s.active[b.c] != 0;
  double acc = 0.0;
  if (act) {
    const double m = mass[b.c];
    const double af = emmu[b.c], ab = emu[b.c];   // M^dagger: e^{-mu} on the +t hop, e^{+mu} on the -t hop
    const double a = s.alpha[b.c];   // single GPU: finalised by the preceding dslash; slab: by slab_scalars_kernel
    const size_t R = (size_t)g.R;
    const size_t j = (size_t)b.x * g.C + b.c;
    const size_t jp = (size_t)((b.x + 1 == g.nx) ? 0 : b.x + 1) * g.C + b.c;
    const size_t jm = (size_t)((b.x == 0) ? g.nx - 1 : b.x - 1) * g.C + b.c;
    const int t0 = b.ttile * TT;
    double2 pm, w0m;
    if (t0 == 0) {
      pm = ld_halo<SLAB>(&in_prev[(size_t)(g.nt - 1) * R + j]);
      w0m = ld_halo<SLAB>(&W0_prev[(size_t)(g.nt - 1) * R + j]);
    } else {
      pm = in[(size_t)(t0 - 1) * R + j];
      w0m = W0[(size_t)(t0 - 1) * R + j];
    }
    double2 pc = in[t0 * R + j];
#pragma unroll
    for (int i = 0; i < TT; i++) {
      const int t = t0 + i;
      if (t < g.nt) {
        const size_t row = t * R;
        const double2 pp = (t + 1 == g.nt) ? ld_halo<SLAB>(&in_next[j]) : in[row + R + j];
        const double2 pxp = in[row + jp];
        const double2 pxm = in[row + jm];
        const double2 w0c = W0[row + j];
        const double2 w1c = W1[row + j];
        const double2 w1m = W1[row + jm];
        const double2 pv = p[row + j];
        double2 xv = x[row + j], rv = r[row + j];
        const double fr = w0c.x * af, fi = w0c.y * af;
        const double br = w0m.x * ab, bi = w0m.y * ab;
        double hr = fr * pp.x - fi * pp.y;
        double hi = fr * pp.y + fi * pp.x;
        hr -= br * pm.x + bi * pm.y;
        hi -= br * pm.y - bi * pm.x;
        hr += w1c.x * pxp.x - w1c.y * pxp.y;
        hi += w1c.x * pxp.y + w1c.y * pxp.x;
        hr -= w1m.x * pxm.x + w1m.y * pxm.y;
        hi -= w1m.x * pxm.y - w1m.y * pxm.x;
        const double ms = msite ? msite[row + j] : m;
        const double qx = ms * pc.x - hr, qy = ms * pc.y - hi;   // q = M^dagger Mp
        xv.x += a * pv.x;
        xv.y += a * pv.y;
        rv.x -= a * qx;
        rv.y -= a * qy;
        x[row + j] = xv;
        r[row + j] = rv;
        acc += rv.x * rv.x + rv.y * rv.y;
        pm = pc;
        pc = pp;
        w0m = w0c;
      }
    }
  }
  reduce_finalize<FIN_RR, SLAB, TB_RED_RR>(acc, g, s, sl, b, red);
  if (SLAB) slab_signal_done(sl, TB_FLAG_MPDONE, -1, seq);
}

// ---------------------------------------------------------------------------------------------------------
// TMA-staged variants of the two stencil passes of the fused iteration (single GPU, family A).
//
// The register-marching kernels above keep one row of loads in flight per thread (about 40 KB per SM), which is
// at the edge of what 7 TB/s of HBM needs.  Here the rows of a block travel through a ring of NS shared-memory
// stages filled by bulk asynchronous copies (cp.async.bulk, completion counted on an mbarrier per stage): NS - 1
// rows of every resident block are in flight whatever the register budget, and the stencil reads its x-neighbours,
// its links and the t+1 row from shared memory.  Same block tiling (BC chains x BX sites), same per-chain reduction
// and last-block scalar update as dslash_kernel / dslash_axpy_norm_kernel; blocks are taller (gp.tt rows).
//   FUSED = false : out = M in, |out|^2 -> alpha                               (hmc.c:366,368-371)
//   FUSED = true  : q = M^dagger in on the fly; x += alpha p; r -= alpha q; ||r||^2 -> beta  (hmc.c:367,372-390)
//   CG = false    : a plain apply, out = M in or M^dagger in (`dagger`), no chain masks, no reduction (hmc.c:132-248)
// A stage holds, for one row of the tile, [site][chain] with one halo site either side where the stencil needs it:
//   P  (BX + 2) sites of the input field      W0  BX sites      W1  (BX + 1) sites (halo on the left)
//   FUSED: PV (the CG direction p), X, RR: BX sites each
constexpr int PIPE_NS_MAX = 4;   // stages per block; 3 when four of them would leave one block per SM

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// one 2-D box of a [site][chain] array (tensor map: dimension 0 = 2 C doubles, dimension 1 = sites) -> shared memory
__device__ __forceinline__ void tensor_g2s(uint32_t dst, const void *tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
      "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// tensor maps of the six arrays a staged block reads, in the order of the stage: P, W0, W1, PV, X, RR (TILED only)
struct PipeMaps {
  CUtensorMap m[6];
};

// Offsets in double2 from the start of a stage, for a tile of bc chains x bx sites (bc * bx = 256).  Regions in the order
//   P  (bx + 2) sites (one halo site either side)      [XP: then R, the same shape]
//   W0 bx sites                                         W1 (bx + 1) sites (halo site on the left)
//   [FUSED: PV, X, RR: bx sites each]
// A site is bc values, except in the link regions of a SHARED gauge field (SHW), where it is one value.  Every region
// that a tensor-map copy fills starts on a multiple of 8 double2 (128 bytes).
template <bool FUSED, bool XP, bool SHW>
struct PipeLay {
  __host__ __device__ static constexpr int al(int v) { return (v + 7) & ~7; }
  __host__ __device__ static constexpr int p(int) { return 0; }
  __host__ __device__ static constexpr int rx(int bc) { return 256 + 2 * bc; }
  __host__ __device__ static constexpr int w0(int bc) { return XP ? 512 + 4 * bc : 256 + 2 * bc; }
  __host__ __device__ static constexpr int w1(int bc) { return w0(bc) + (SHW ? al(256 / bc) : 256); }
  __host__ __device__ static constexpr int pv(int bc) { return w1(bc) + (SHW ? al(256 / bc + 1) : 256 + bc); }
  __host__ __device__ static constexpr int xx(int bc) { return pv(bc) + 256; }
  __host__ __device__ static constexpr int rr(int bc) { return pv(bc) + 512; }
  __host__ __device__ static constexpr int size(int bc) { return pv(bc) + (FUSED ? 768 : 0); }
  static size_t smem_bytes(int bc, int ns) { return (size_t)ns * size(bc) * sizeof(double2) + ns * sizeof(unsigned long long); }
};
static_assert(PipeLay<false, false, false>::size(16) == 768 + 48 && PipeLay<true, false, false>::rr(16) == 1280 + 48 &&
                  PipeLay<false, true, false>::size(16) == 1024 + 80,
              "the per-chain layouts of rounds 1 and 2");

// TILED: the tile holds 16 of the batch's chains (batches of more than 16 chains), so a row of the tile is bx runs of
// 256 bytes, one per site.  Issued as bx separate bulk copies they were 1.5x slower than the marching kernels (the copy
// engine retires a small copy every ~65 cycles); as ONE 2-D box per array and row, described by a tensor map
// (cp.async.bulk.tensor.2d), they cost the same instructions and bytes as the contiguous tiles of a 16-chain batch.
//
// XP (first pass of the CG iteration only): the direction update p = r + beta p (hmc.c:391-392) of the PREVIOUS iteration
// is folded into this pass instead of being a kernel of its own.  `in` is the previous direction, `r` the residual; the
// new direction is formed on the fly for every site the stencil touches (the tile, its halo rows and halo columns:
// the same fma as xpay_kernel, so the values are bit for bit those the separate kernel stores), written to `x` (a
// SECOND direction buffer: other blocks are still reading their halos from `in`), and M applied to it.  The iteration
// moves 224 B per site instead of 240 (r 16 + p 16 + links 32 + new p 16 + Mp 16 in this pass) in two launches.
//
// SHW: one gauge field shared by every chain of the batch (tb_set_gauge_shared: the sources of a multi-RHS solve).  W0 and
// W1 point at the compact fields [t][x]; a row of a tile stages bx (+ 1) links per field instead of bx * bc, and the
// threads of a site read them as shared-memory broadcasts: the iteration moves 180 B per site and chain instead of 240.
template <bool FUSED, int PIPE_NS, bool CG, bool TILED, bool XP, bool SHW>
__device__ __forceinline__ void
dslash_pipe_body(const double2 *__restrict__ in, double2 *__restrict__ out, const double2 *__restrict__ W0,
                 const double2 *__restrict__ W1, const double *__restrict__ mass, const double *__restrict__ emu,
                 const double *__restrict__ emmu, const double2 *__restrict__ pvec, double2 *__restrict__ x,
                 double2 *__restrict__ r, const TbGeom &g, const TbCgState &s, const TbSlab &sl, const int dagger,
                 const PipeMaps &maps) {
  using St = PipeLay<FUSED, XP, SHW>;
  static_assert(!XP || (CG && !FUSED), "XP is the first pass of the fused CG iteration");
  __shared__ double red[TB_MAX_BLOCK];
  extern __shared__ __align__(128) unsigned char pipe_smem[];
  const BlockPos b = block_pos(g);
  if (CG) {   // a tile of the staged geometry spans one or more of the chain tiles that tile_active counts in
    const int c0 = b.ctile * g.bc;
    int live = 0;
    for (int q = c0 >> g.ta_shift; q <= (c0 + g.bc - 1) >> g.ta_shift; q++) live += s.tile_active[q];
    if (live == 0) return;
  }
  const int bc = g.bc, bx = g.bx, TT = g.tt, tid = threadIdx.x;
  const int ssize = St::size(bc);
  double2 *stage0 = reinterpret_cast<double2 *>(pipe_smem);
  const uint32_t stage0_addr = smem_u32(stage0);
  const uint32_t bar0 = stage0_addr + (uint32_t)(PIPE_NS * ssize) * 16u;
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < PIPE_NS; q++) mbar_init(bar0 + 8u * q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  const int t0 = b.ttile * TT;
  const int c0 = b.ctile * bc, x0 = b.xtile * bx;
  const bool contiguous = TILED || bc == g.C;   // one copy per array and row: a contiguous run, or a 2-D box
  const int np = contiguous ? 1 : bx;                              // centre pieces per array and row
  const uint32_t plen = (uint32_t)(contiguous ? bx * bc : bc) * 16u;   // bytes per centre piece
  // bytes per halo site (a plain copy: in a ragged last tile only the chains that exist)
  const uint32_t hlen = (uint32_t)((TILED && c0 + bc > g.C) ? g.C - c0 : bc) * 16u;
  const size_t pstride = contiguous ? 0 : (size_t)g.C;             // double2 between centre pieces in global memory
  const uint32_t wlen = SHW ? (uint32_t)bx * 16u : plen;           // bytes of a row of links of the tile
  const uint32_t whalo = SHW ? 16u : hlen;                         // ... and of its left halo site
  const int xm = x0 == 0 ? g.nx - 1 : x0 - 1, xp = x0 + bx == g.nx ? 0 : x0 + bx;

  // row i of the block (-1 .. TT: one row of halo either side in t) -> stage (i + 1) % NS; issued by warp 0
  auto fill = [&](int i) {
    const int slot = (i + 1) % PIPE_NS;
    const uint32_t bar = bar0 + 8u * slot;
    const uint32_t dst0 = stage0_addr + (uint32_t)(slot * ssize) * 16u;
    int t = t0 + i;
    t = t < 0 ? t + g.nt : (t >= g.nt ? t - g.nt : t);
    const size_t row = (size_t)t * g.R + c0;
    const int lane = tid & 31;
    const int kind = i < 0 ? 0 : (i >= TT ? 1 : 2);   // 0: P and W0 centres; 1: P centre; 2: everything
    if constexpr (XP) {
      // pieces: 0 P, 1 R, 2 W0, 3 W1 (centres); 4, 5 P left / right; 6, 7 R left / right; 8 W1 left.  Row -1: 0..2, row TT: 0..1
      const int npc = kind == 0 ? 3 : kind == 1 ? 2 : 9;
      if (lane == 0) mbar_expect_tx(bar, kind == 0 ? 2 * plen + wlen : kind == 1 ? 2 * plen : 2 * plen + 2 * wlen + 4 * hlen + whalo);
      __syncwarp();
      if (lane < npc) {
        const int pc_ = lane;
        const int arr = pc_ < 4 ? pc_ : (pc_ < 6 ? 0 : (pc_ < 8 ? 1 : 3));     // 0 P, 1 R, 2 W0, 3 W1
        const int side = pc_ < 4 ? 0 : ((pc_ == 4 || pc_ == 6 || pc_ == 8) ? 1 : 2);   // 0 centre, 1 left, 2 right
        const double2 *base = arr == 0 ? in : (arr == 1 ? (const double2 *)r : (arr == 2 ? W0 : W1));
        const int off = arr == 0 ? St::p(bc) : (arr == 1 ? St::rx(bc) : (arr == 2 ? St::w0(bc) : St::w1(bc)));
        const bool halo_l = arr != 2;   // W0 has no halo site; P, R, W1 keep one on the left
        if (SHW && arr >= 2) {   // compact links: [site] of the row, W1 with its left halo site in front
          const size_t rs = (size_t)t * g.nx;
          if (side == 0) bulk_g2s(dst0 + (uint32_t)(off + (arr == 3 ? 1 : 0)) * 16u, base + rs + x0, wlen, bar);
          else bulk_g2s(dst0 + (uint32_t)off * 16u, base + rs + xm, 16u, bar);
        } else if (side == 0) {
          const uint32_t dst = dst0 + (uint32_t)(off + (halo_l ? bc : 0)) * 16u;
          if (TILED) tensor_g2s(dst, &maps.m[arr == 0 ? 0 : (arr == 1 ? 5 : (arr == 2 ? 1 : 2))], 2 * c0, t * g.nx + x0, bar);
          else bulk_g2s(dst, base + row + (size_t)x0 * g.C, plen, bar);
        } else if (side == 1) {
          bulk_g2s(dst0 + (uint32_t)off * 16u, base + row + (size_t)xm * g.C, hlen, bar);
        } else {
          bulk_g2s(dst0 + (uint32_t)(off + (bx + 1) * bc) * 16u, base + row + (size_t)xp * g.C, hlen, bar);
        }
      }
      return;
    }
    const int nseg = kind == 0 ? 2 * np : kind == 1 ? np : (FUSED ? 6 * np + 3 : 3 * np + 3);
    if (lane == 0) {
      const uint32_t bytes = kind == 0 ? np * (plen + wlen) : kind == 1 ? np * plen
                                                                          : np * ((FUSED ? 4 : 1) * plen + 2 * wlen) + 2 * hlen + whalo;
      mbar_expect_tx(bar, bytes);
    }
    __syncwarp();
    for (int sg = lane; sg < nseg; sg += 32) {
      // decode: arrays in the order P, W0, W1, PV, X, RR; the centre pieces of an array first, then the halo sites
      const double2 *src;
      uint32_t dst, bytes = plen;
      int arr = -1;   // centre piece of array `arr`, or -1: a halo site
      if (sg < np) { arr = 0; src = in + row + (size_t)x0 * g.C + sg * pstride; dst = (uint32_t)(St::p(bc) + bc) * 16u + sg * plen; }
      else if (kind == 1) { continue; }
      else if (sg < 2 * np) { arr = 1; const int k = sg - np; src = W0 + row + (size_t)x0 * g.C + k * pstride; dst = (uint32_t)St::w0(bc) * 16u + k * plen; }
      else if (kind == 0) { continue; }
      else if (sg < 3 * np) { arr = 2; const int k = sg - 2 * np; src = W1 + row + (size_t)x0 * g.C + k * pstride; dst = (uint32_t)(St::w1(bc) + bc) * 16u + k * plen; }
      else if (sg == 3 * np) { src = in + row + (size_t)xm * g.C; dst = (uint32_t)St::p(bc) * 16u; bytes = hlen; }
      else if (sg == 3 * np + 1) { src = in + row + (size_t)xp * g.C; dst = (uint32_t)(St::p(bc) + (bx + 1) * bc) * 16u; bytes = hlen; }
      else if (sg == 3 * np + 2) { src = W1 + row + (size_t)xm * g.C; dst = (uint32_t)St::w1(bc) * 16u; bytes = hlen; }
      else {
        const int k = sg - (3 * np + 3), a = k / np, kk = k - a * np;
        const double2 *base = a == 0 ? pvec : (a == 1 ? x : r);
        arr = 3 + a;
        src = base + row + (size_t)x0 * g.C + kk * pstride;
        dst = (uint32_t)(a == 0 ? St::pv(bc) : (a == 1 ? St::xx(bc) : St::rr(bc))) * 16u + kk * plen;
      }
      if (SHW && (arr == 1 || arr == 2 || sg == 3 * np + 2)) {   // compact links (W0 / W1 point at [t][x] fields)
        const size_t rs = (size_t)t * g.nx;
        if (arr == 1) bulk_g2s(dst0 + (uint32_t)St::w0(bc) * 16u, W0 + rs + x0, wlen, bar);
        else if (arr == 2) bulk_g2s(dst0 + (uint32_t)(St::w1(bc) + 1) * 16u, W1 + rs + x0, wlen, bar);
        else bulk_g2s(dst0 + (uint32_t)St::w1(bc) * 16u, W1 + rs + xm, 16u, bar);
      } else if (TILED && arr >= 0) {
        tensor_g2s(dst0 + dst, &maps.m[arr], 2 * c0, t * g.nx + x0, bar);
      } else {
        bulk_g2s(dst0 + dst, src, bytes, bar);
      }
    }
  };
  auto stage_of = [&](int i) { return stage0 + ((i + 1) % PIPE_NS) * ssize; };
  auto wait_row = [&](int i) { mbar_wait(bar0 + 8u * ((i + 1) % PIPE_NS), (uint32_t)(((i + 1) / PIPE_NS) & 1)); };

  if (tid < 32)
    for (int i = -1; i < PIPE_NS - 1 && i <= TT; i++) fill(i);

  // the missing chains of a ragged last tile compute on zero-filled boxes and store nothing
  const bool act = (!TILED || b.c < g.C) && (CG ? s.active[b.c] != 0 : true);
  const bool dag = FUSED || (!CG && dagger);
  const double m = mass[b.c];
  const double af = dag ? emmu[b.c] : emu[b.c];   // factor on the +t hop (M^dagger: e^{-mu})
  const double ab = dag ? emu[b.c] : emmu[b.c];   // factor on the -t hop
  const double a = FUSED ? s.alpha[b.c] : 0.0;
  const double be = XP ? s.beta[b.c] : 0.0;   // 0 in the first iteration (cg_reset_kernel): p = r = b
  const size_t j = (size_t)b.x * g.C + b.c;
  double acc = 0.0;
  // the field the stencil reads at offset `off` of a stage: the staged vector, or (XP) the new direction r + beta p
  auto field = [&](const double2 *S, int off) {
    const double2 pv = S[St::p(bc) + off];
    if constexpr (XP) {
      const double2 rv = S[St::rx(bc) + off];
      return make_double2(fma(be, pv.x, rv.x), fma(be, pv.y, rv.y));
    } else {
      return pv;
    }
  };

  wait_row(-1);
  double2 pm = field(stage_of(-1), bc + tid);
  const int wi = SHW ? b.x_local : tid;   // index of the thread's link in a stage row (compact: one per site)
  double2 w0m = stage_of(-1)[St::w0(bc) + wi];
  wait_row(0);
  double2 pc = field(stage_of(0), bc + tid);
  __syncthreads();   // stage of row -1 is free
  if (tid < 32 && PIPE_NS - 1 <= TT) fill(PIPE_NS - 1);

  for (int i = 0; i < TT; i++) {
    wait_row(i + 1);
    const double2 *S = stage_of(i);
    const double2 pp = field(stage_of(i + 1), bc + tid);
    const double2 pxm = field(S, tid), pxp = field(S, 2 * bc + tid);
    const double2 w0c = S[St::w0(bc) + wi];
    const double2 w1m = S[St::w1(bc) + wi], w1c = S[St::w1(bc) + (SHW ? 1 : bc) + wi];
    // hops: +af W0(n) psi(n+t) - ab conj(W0(n-t)) psi(n-t) + W1(n) psi(n+x) - conj(W1(n-x)) psi(n-x)
    const double fr = w0c.x * af, fi = w0c.y * af;
    const double br = w0m.x * ab, bi = w0m.y * ab;
    double hr = fr * pp.x - fi * pp.y;
    double hi = fr * pp.y + fi * pp.x;
    hr -= br * pm.x + bi * pm.y;
    hi -= br * pm.y - bi * pm.x;
    hr += w1c.x * pxp.x - w1c.y * pxp.y;
    hi += w1c.x * pxp.y + w1c.y * pxp.x;
    hr -= w1m.x * pxm.x + w1m.y * pxm.y;
    hi -= w1m.x * pxm.y - w1m.y * pxm.x;
    const size_t k = (size_t)(t0 + i) * g.R + j;
    if (FUSED) {
      const double qx = m * pc.x - hr, qy = m * pc.y - hi;   // q = M^dagger Mp
      const double2 pv = S[St::pv(bc) + tid];
      double2 xv = S[St::xx(bc) + tid], rv = S[St::rr(bc) + tid];
      xv.x += a * pv.x;
      xv.y += a * pv.y;
      rv.x -= a * qx;
      rv.y -= a * qy;
      if (act) {
        x[k] = xv;
        r[k] = rv;
        acc += rv.x * rv.x + rv.y * rv.y;
      }
    } else {
      double2 o;
      if (!CG && dag) {
        o.x = m * pc.x - hr;
        o.y = m * pc.y - hi;
      } else {
        o.x = m * pc.x + hr;
        o.y = m * pc.y + hi;
      }
      if (act) {
        out[k] = o;
        if (XP) x[k] = pc;              // the new direction, for the second pass and the next iteration
        acc += o.x * o.x + o.y * o.y;   // <p, M^dagger M p> = |M p|^2
      }
    }
    pm = pc;
    pc = pp;
    w0m = w0c;
    __syncthreads();   // every thread has read the stage of row i
    if (tid < 32 && i + PIPE_NS <= TT) fill(i + PIPE_NS);
  }
  if (CG) reduce_finalize<FUSED ? FIN_RR : FIN_PQ, false, FUSED ? TB_RED_RR : TB_RED_PQ>(acc, g, s, sl, b, red);
}

#define TB_PIPE_PARAMS                                                                                                  \
  const double2 *__restrict__ in, double2 *__restrict__ out, const double2 *__restrict__ W0, const double2 *__restrict__ W1, \
      const double *__restrict__ mass, const double *__restrict__ emu, const double *__restrict__ emmu,                 \
      const double2 *__restrict__ pvec, double2 *__restrict__ x, double2 *__restrict__ r, const TbGeom g, const TbCgState s, \
      const TbSlab sl, const int dagger, const __grid_constant__ PipeMaps maps
#define TB_PIPE_ARGS in, out, W0, W1, mass, emu, emmu, pvec, x, r, g, s, sl, dagger, maps

template <bool FUSED, int PIPE_NS, bool CG, bool TILED = false, bool XP = false, bool SHW = false>
__global__ void __launch_bounds__(TB_MAX_BLOCK) dslash_pipe_kernel(TB_PIPE_PARAMS) {
  dslash_pipe_body<FUSED, PIPE_NS, CG, TILED, XP, SHW>(TB_PIPE_ARGS);
}
// The plain first pass with whole-batch tiles, held to 64 registers = four blocks per SM.  That is what ptxas chose by
// itself until the body grew its variants; at the 70 registers it then took, 2048^2 ran at the marching kernels' 188 us
// per iteration instead of 166.  (Bounding every variant moved the others the wrong way: their register counts are
// ptxas's own.)
template <int PIPE_NS, bool CG>
__global__ void __launch_bounds__(TB_MAX_BLOCK, 4) dslash_pipe_plain_kernel(TB_PIPE_PARAMS) {
  dslash_pipe_body<false, PIPE_NS, CG, false, false, false>(TB_PIPE_ARGS);
}

// p = r + beta p (hmc.c:391-392).  SLAB: p is an exchange vector: wait until the neighbours have finished
// reading generation E-1 (PDONE), publish generation E (PREADY).
template <int TT, bool SLAB>
__global__ void __launch_bounds__(TB_MAX_BLOCK)
xpay_kernel(double2 *__restrict__ p, const double2 *__restrict__ r, const TbGeom g, const TbCgState s,
            const TbSlab sl) {
  const BlockPos b = block_pos(g);
  if (s.tile_active[b.ctile] == 0) {
    if (SLAB && *s.n_active > 0) slab_signal_done(sl, TB_FLAG_PREADY, -1, *sl.seq);
    return;
  }
  int seq = 0;
  if (SLAB) {
    seq = *sl.seq;
    if (b.ttile == 0) slab_wait(sl, TB_FLAG_PDONE, 0, seq - 1);
    if (b.ttile == g.nttiles - 1) slab_wait(sl, TB_FLAG_PDONE, 1, seq - 1);
  }
  if (b.valid && s.active[b.c] != 0) {
    const double be = s.beta[b.c];
    const size_t R = (size_t)g.R;
    const size_t j = (size_t)b.x * g.C + b.c;
    const int t0 = b.ttile * TT;
#pragma unroll
    for (int i = 0; i < TT; i++) {
      const int t = t0 + i;
      if (t < g.nt) {
        const size_t k = t * R + j;
        const double2 rv = r[k];
        double2 pv = p[k];
        pv.x = rv.x + be * pv.x;
        pv.y = rv.y + be * pv.y;
        p[k] = pv;
      }
    }
  }
  if (SLAB) slab_signal_done(sl, TB_FLAG_PREADY, -1, seq);
}

// x = 0 ; r = b ; p = b ; rr = ||b||^2 (hmc.c:349-361)
template <int TT, bool SLAB>
__global__ void __launch_bounds__(TB_MAX_BLOCK)
cg_init_kernel(const double2 *__restrict__ bsrc, double2 *__restrict__ x, double2 *__restrict__ r,
               double2 *__restrict__ p, const TbGeom g, const TbCgState s, const TbSlab sl) {
  __shared__ double red[TB_MAX_BLOCK];
  const BlockPos b = block_pos(g);
  int seq = 0;
  if (SLAB) {
    seq = *sl.seq;
    if (b.ttile == 0) slab_wait(sl, TB_FLAG_PDONE, 0, seq - 1);
    if (b.ttile == g.nttiles - 1) slab_wait(sl, TB_FLAG_PDONE, 1, seq - 1);
  }
  double acc = 0.0;
  if (b.valid) {
    const size_t R = (size_t)g.R;
    const size_t j = (size_t)b.x * g.C + b.c;
    const int t0 = b.ttile * TT;
#pragma unroll
    for (int i = 0; i < TT; i++) {
      const int t = t0 + i;
      if (t < g.nt) {
        const size_t k = t * R + j;
        const double2 v = bsrc[k];
        x[k] = make_double2(0.0, 0.0);
        r[k] = v;
        p[k] = v;
        acc += v.x * v.x + v.y * v.y;
      }
    }
  }
  reduce_finalize<FIN_INIT, SLAB, TB_RED_INIT>(acc, g, s, sl, b, red);
  if (SLAB) slab_signal_done(sl, TB_FLAG_PREADY, -1, seq);
}

// slab mode: copy a user vector into the exchange vector p as generation E (standalone applies)
template <int TT>
__global__ void __launch_bounds__(TB_MAX_BLOCK)
slab_stage_kernel(const double2 *__restrict__ src, double2 *__restrict__ p, const TbGeom g, const TbSlab sl) {
  const BlockPos b = block_pos(g);
  const int seq = *sl.seq;
  if (b.ttile == 0) slab_wait(sl, TB_FLAG_PDONE, 0, seq - 1);
  if (b.ttile == g.nttiles - 1) slab_wait(sl, TB_FLAG_PDONE, 1, seq - 1);
  if (b.valid) {
    const size_t R = (size_t)g.R;
    const size_t j = (size_t)b.x * g.C + b.c;
    const int t0 = b.ttile * TT;
#pragma unroll
    for (int i = 0; i < TT; i++) {
      const int t = t0 + i;
      if (t < g.nt) p[t * R + j] = src[t * R + j];
    }
  }
  slab_signal_done(sl, TB_FLAG_PREADY, -1, seq);
}

// slab mode: advance the epoch (a new generation of p is about to be written)
__global__ void slab_bump_kernel(const TbSlab sl) { *sl.seq += 1; }

// slab mode: one block.  Waits for every rank's partial of reduction RED at epoch E, sums them in rank order
// and evaluates the CG scalar update FIN for every chain.  After FIN_RR it opens the next epoch (unless every
// chain has finished); after FIN_INIT with nothing left to solve it releases the neighbours, who will never be
// asked to read this generation of p.
template <int FIN, int RED>
__global__ void __launch_bounds__(TB_MAX_BLOCK)
slab_scalars_kernel(const TbGeom g, const TbCgState s, const TbSlab sl) {
  if (FIN != FIN_INIT && *s.n_active == 0) return;
  const int seq = *sl.seq;
  for (int ctile = 0; ctile < g.nctiles; ctile++) {
    if (FIN != FIN_INIT && s.tile_active[ctile] == 0) continue;   // uniform: nobody published for this tile
    if ((int)threadIdx.x < sl.P)
      slab_spin(sl.red_flag + (RED * sl.P + threadIdx.x) * g.nctiles + ctile, seq, sl.err, 10 + RED);
    __syncthreads();
    for (int cl = threadIdx.x; cl < g.bc; cl += blockDim.x) {
      const int cp = ctile * g.bc + cl;
      double total = 0.0;
      for (int q = 0; q < sl.P; q++) total += __ldcv(&sl.red[(size_t)(RED * sl.P + q) * g.Cpad + cp]);
      if (cp < g.C) finalize_scalar<FIN>(total, cp, ctile, s);
    }
    __syncthreads();
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int left = *(volatile int *)s.n_active;
    if (FIN == FIN_RR && left > 0) *sl.seq = seq + 1;
    if (FIN == FIN_INIT && left == 0) {
      __threadfence_system();
      *(volatile int *)(sl.sig_prev + TB_FLAG_PDONE * 2) = seq;
      *(volatile int *)(sl.sig_next + TB_FLAG_PDONE * 2) = seq;
      *(volatile int *)(sl.sig_prev + TB_FLAG_MPDONE * 2) = seq;
      *(volatile int *)(sl.sig_next + TB_FLAG_MPDONE * 2) = seq;
      __threadfence_system();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Slab mode, whole solve in ONE launch per GPU (fused ADJOINT iteration, up to 32 chains).
//
// The multi-kernel slab iteration spends most of its time between kernels: five launches and four cross-GPU
// hand-shakes per iteration (2048^2 on 8 GPUs: 25 us of kernels in an 82 us iteration).  Here every GPU runs one
// persistent cooperative kernel; a block owns a fixed set of tiles for the whole solve and the phases of an iteration
// are separated by grid barriers (an atomic counter in L2) instead of kernel boundaries:
//   A  p = r + beta p on the fly (hmc.c:391-392), Mp = M p, |Mp|^2
//                               barrier   block 0: all-reduce over the GPUs -> alpha     (hmc.c:366,368-371)
//   B  q = M^dagger Mp on the fly, x += alpha p, r -= alpha q, ||r||^2
//                               barrier   block 0: all-reduce -> beta, convergence        (hmc.c:367,372-390)
// Two synchronisation points per iteration.  p is double-buffered and never exchanged: phase A recomputes
// p = r + beta p_old at the five points of its stencil (the same fma everywhere, so the copies agree to the bit with
// the stored one), reading the neighbours' rows of r and p_old.  The all-reduce is the one-shot peer-store exchange of
// the multi-kernel path (every rank stores its partial into every rank's slot table, sums in rank order => bitwise
// identical scalars and decisions on all ranks), and it doubles as the cross-GPU fence for the halo rows: a rank
// contributes to the |Mp|^2 sum only after its phase A, so whoever holds the sum knows that the neighbours' Mp rows
// are complete and that nobody reads r or p_old of this iteration any more; likewise ||r||^2 for r, the new p and the
// Mp rows.  No neighbour flag is left inside the solve.  Fields that are rewritten during the launch are loaded with
// ld.global.cg: L1 is not coherent across the phases of one kernel.
// ranks enter the launch seconds apart when one of them is still busy on the host: ~20 s before a wait gives up
#define TB_PERSIST_SPIN_CYCLES 40000000000LL
struct SlabCgArgs {
  const double2 *b;
  double2 *x, *r, *p, *p1, *Mp;
  const double2 *p_prev, *p_next, *p1_prev, *p1_next, *r_prev, *r_next, *mp_prev, *mp_next;
  const double2 *W0, *W0_prev, *W1;
  const double *mass, *emu, *emmu;
  int sysfence;   // 1: system-scope fences around the peer exchange (TB_SLAB_SYSFENCE=1), 0: device scope (see xgpu_fence)
};

__device__ __forceinline__ void spin_until(volatile int *f, int need) {
  const long long t0 = clock64();
  while (*f < need) {
    __nanosleep(64);
    if (clock64() - t0 > TB_PERSIST_SPIN_CYCLES) __trap();   // a launch error on this rank instead of a hung box
  }
}

__device__ __forceinline__ void grid_barrier(unsigned long long *bar, unsigned long long &target) {
  // device scope is enough for the halo rows too: a peer reads them out of THIS GPU's L2, after a flag that block 0
  // stores behind a system-scope fence (a system-scope fence in every thread cost ~5 us per barrier)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    atomicAdd(bar, 1ULL);
    const long long t0 = clock64();
    while (*(volatile unsigned long long *)bar < target)
      if (clock64() - t0 > TB_PERSIST_SPIN_CYCLES) __trap();
    __threadfence();
  }
  __syncthreads();
}

// The barrier in front of a reduction: every block arrives; block 0 alone waits for the arrivals (it then runs the
// all-reduce and releases the others through *sl.go, see block0_scalars / wait_go), so the other blocks poll one
// flag instead of the counter and then the flag.
__device__ __forceinline__ void grid_arrive(unsigned long long *bar, unsigned long long &target) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    atomicAdd(bar, 1ULL);
    if (blockIdx.x == 0) {
      const long long t0 = clock64();
      while (*(volatile unsigned long long *)bar < target)
        if (clock64() - t0 > TB_PERSIST_SPIN_CYCLES) __trap();
      __threadfence();
    }
  }
  if (blockIdx.x == 0) __syncthreads();
}

// tile `tile` of this rank's slab (one chain tile); rows at the two ends of the slab come last
__device__ __forceinline__ BlockPos tile_pos(const TbGeom &g, int tile) {
  BlockPos b;
  b.ctile = 0;
  b.xtile = tile % g.nxtiles;
  b.ttile = (tile / g.nxtiles + 1) % g.nttiles;
  b.c_local = threadIdx.x & (g.bc - 1);
  b.x_local = threadIdx.x >> g.bc_shift;
  b.c = b.c_local;
  b.x = b.xtile * g.bx + b.x_local;
  b.valid = (b.c < g.C) && (b.x < g.nx);
  return b;
}

// per-chain sum of `v` over the block (thread = x_local * bc + chain, bc <= 32): xor-shuffles between the lanes of a
// warp that hold the same chain, then the eight warp partials in warp order; valid in threads tid < bc.  Fixed shape:
// deterministic.
__device__ __forceinline__ double block_sum_chains(double v, const TbGeom &g, double *red) {
  for (int o = 16; o >= g.bc; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  __syncthreads();   // red may still be read by the previous call
  if (lane < g.bc) red[warp * g.bc + lane] = v;
  __syncthreads();
  double t = 0.0;
  if ((int)threadIdx.x < g.bc)
    for (int w = 0; w < nwarp; w++) t += red[w * g.bc + threadIdx.x];
  return t;
}

// per-chain sum of `acc` over the block -> partial[blockIdx.x][chain]
__device__ __forceinline__ void block_partial(double acc, const TbGeom &g, const TbCgState &s, double *red) {
  const double t = block_sum_chains(acc, g, red);
  if ((int)threadIdx.x < g.bc) s.partial[(size_t)blockIdx.x * g.Cpad + threadIdx.x] = t;
}

__device__ __forceinline__ void st_volatile_v2(double2 *p, long long x, long long y) {
  asm volatile("st.volatile.global.v2.b64 [%0], {%1, %2};\n" ::"l"(p), "l"(x), "l"(y) : "memory");
}
__device__ __forceinline__ void ld_volatile_v2(const double2 *p, long long &x, long long &y) {
  asm volatile("ld.volatile.global.v2.b64 {%0, %1}, [%2];\n" : "=l"(x), "=l"(y) : "l"(p) : "memory");
}

// block 0, after the grid barrier: sum the block partials in block order, all-reduce over the ranks, evaluate the CG
// scalars (finalize_scalar), then release the other blocks of this GPU through *sl.go = go_value.
// The exchange is one 16-byte store per (peer, chain): {bits of the partial, the same bits xor a tag of the sequence
// number}.  Value and tag travel together, so no fence (an NVLink round trip) separates them, and a reader that caught
// the two halves of different generations sees a tag that matches nothing and polls again; thread (q, chain) polls the
// slot that rank q writes into this GPU's memory.
template <int FIN, int RED>
__device__ __forceinline__ void block0_scalars(const TbGeom &g, const TbCgState &s, const TbSlab &sl, double *red,
                                               int seqv, int go_value, int sysfence) {
  const int c_local = threadIdx.x & (g.bc - 1), x_local = threadIdx.x >> g.bc_shift;
  double sum = 0.0;
  for (int blk = x_local; blk < (int)gridDim.x; blk += g.bx) sum += __ldcg(&s.partial[(size_t)blk * g.Cpad + c_local]);
  const double mine = block_sum_chains(sum, g, red);   // threads < bc
  __syncthreads();
  if ((int)threadIdx.x < g.bc) red[threadIdx.x] = mine;
  __syncthreads();
  const bool io = (int)threadIdx.x < sl.P * g.bc;   // thread (q, chain)
  const int q = threadIdx.x >> g.bc_shift;
  const size_t slot = (size_t)(RED * sl.P) * g.Cpad + c_local;
  double theirs = 0.0;
  if (io) {
    const double v = red[c_local];
    if (sysfence) __threadfence_system(); else __threadfence();   // this GPU's halo rows (visible device-wide since the grid barrier) before the tag
    const long long tag = (long long)((unsigned long long)seqv * 0x9E3779B97F4A7C15ULL);
    st_volatile_v2(sl.peer_red2[q] + slot + (size_t)sl.rank * g.Cpad, __double_as_longlong(v), __double_as_longlong(v) ^ tag);
    const double2 *in = sl.red2 + slot + (size_t)q * g.Cpad;
    const long long t0 = clock64();
    for (;;) {
      long long wx, wy;
      ld_volatile_v2(in, wx, wy);
      if ((wx ^ wy) == tag) { theirs = __longlong_as_double(wx); break; }
      if (clock64() - t0 > TB_PERSIST_SPIN_CYCLES) __trap();
    }
  }
  __syncthreads();
  if (io) red[threadIdx.x] = theirs;   // [q][chain]
  __syncthreads();
  if (x_local == 0 && c_local < g.C) {
    double total = 0.0;
    for (int r = 0; r < sl.P; r++) total += red[r * g.bc + c_local];   // rank order: the same bits on every rank
    finalize_scalar<FIN>(total, c_local, 0, s);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (sysfence) __threadfence_system(); else __threadfence();   // acquire side of the exchange: the neighbours' rows, for the blocks that wait on go
    *(volatile int *)sl.go = go_value;
  }
}

__device__ __forceinline__ void wait_go(const TbSlab &sl, int go_value) {
  if (threadIdx.x == 0) {
    spin_until((volatile int *)sl.go, go_value);
    __threadfence();
  }
  __syncthreads();
}

template <int TT>
__global__ void __launch_bounds__(TB_MAX_BLOCK, 3)
slab_cg_persistent_kernel(const SlabCgArgs a, const TbGeom g, const TbCgState s, const TbSlab sl) {
  __shared__ double red[TB_MAX_BLOCK];
  const int ntiles = g.nxtiles * g.nttiles;
  const size_t R = (size_t)g.R;
  unsigned long long bar_target = 0;
  const int E0 = *(volatile int *)sl.seq;   // block 0 rewrites it only after the last grid barrier
  const int c = threadIdx.x & (g.bc - 1);
  const bool chain = c < g.C;
  const double m = chain ? a.mass[c] : 0.0;
  const double e_p = chain ? a.emu[c] : 1.0, e_m = chain ? a.emmu[c] : 1.0;

  // the neighbours may still read the previous generation of p (an apply queued before this solve)
  auto wait_ends = [&](const BlockPos &b, int kind, int need) {
    if (b.ttile == 0 || b.ttile == g.nttiles - 1) {
      if (threadIdx.x == 0) {
        if (b.ttile == 0) spin_until(sl.flags + kind * 2 + 0, need);
        if (b.ttile == g.nttiles - 1) spin_until(sl.flags + kind * 2 + 1, need);
        __threadfence_system();
      }
      __syncthreads();
    }
  };

  // ---- x = 0, r = p = b, ||b||^2 (hmc.c:349-361): generation E0 + 1 of p
  double acc = 0.0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const BlockPos b = tile_pos(g, tile);
    wait_ends(b, TB_FLAG_PDONE, E0);
    if (b.valid) {
      const size_t j = (size_t)b.x * g.C + b.c;
      const int t0 = b.ttile * TT;
#pragma unroll
      for (int i = 0; i < TT; i++) {
        const int t = t0 + i;
        if (t < g.nt) {
          const size_t k = t * R + j;
          const double2 v = a.b[k];
          a.x[k] = make_double2(0.0, 0.0);
          a.r[k] = v;
          a.p[k] = v;
          acc += v.x * v.x + v.y * v.y;
        }
      }
    }
  }
  block_partial(acc, g, s, red);
  grid_arrive(sl.gbar, bar_target);
  int gen = E0 + 1;   // generation of p
  if (blockIdx.x == 0) block0_scalars<FIN_INIT, TB_RED_INIT>(g, s, sl, red, gen, 3 * gen, a.sysfence);
  wait_go(sl, 3 * gen);

  // p is double-buffered: odd iterations read p_old from a.p and write p_new into a.p1, even ones the other way round
  int n_act = *(volatile int *)s.n_active;
  for (int k = 1; n_act > 0; k++) {
    const bool act = chain && __ldcg(&s.active[c]) != 0;
    const double be = chain ? __ldcg(&s.beta[c]) : 0.0;   // 0 in the first iteration: p = r (hmc.c:352-353)
    const bool odd = k & 1;   // p_old of iteration 1 is a.p (= b, times beta = 0)
    const double2 *po = odd ? a.p : a.p1, *po_prev = odd ? a.p_prev : a.p1_prev, *po_next = odd ? a.p_next : a.p1_next;
    double2 *pn = odd ? a.p1 : a.p;
    // p = r + beta p (hmc.c:391-392) wherever the stencil needs it; the same fma everywhere, so the five copies of a
    // site's p agree to the bit with the one that is stored
    auto pnew = [&](const double2 rv, const double2 pv) {
      return make_double2(fma(be, pv.x, rv.x), fma(be, pv.y, rv.y));
    };
    // ---- A: p = r + beta p on the fly, Mp = M p, |Mp|^2.  The neighbours' rows of r and of the old p are complete:
    // their owners passed the ||r||^2 all-reduce of the previous iteration.
    acc = 0.0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const BlockPos b = tile_pos(g, tile);
      if (b.valid && act) {
        const size_t j = (size_t)b.x * g.C + b.c;
        const size_t jp = (size_t)((b.x + 1 == g.nx) ? 0 : b.x + 1) * g.C + b.c;
        const size_t jm = (size_t)((b.x == 0) ? g.nx - 1 : b.x - 1) * g.C + b.c;
        const int t0 = b.ttile * TT;
        double2 pm, w0m;
        if (t0 == 0) {
          const size_t kk = (size_t)(g.nt - 1) * R + j;
          pm = pnew(__ldcv(&a.r_prev[kk]), __ldcv(&po_prev[kk]));
          w0m = __ldcv(&a.W0_prev[kk]);
        } else {
          const size_t kk = (size_t)(t0 - 1) * R + j;
          pm = pnew(__ldcg(&a.r[kk]), __ldcg(&po[kk]));
          w0m = a.W0[kk];
        }
        double2 pc = pnew(__ldcg(&a.r[t0 * R + j]), __ldcg(&po[t0 * R + j]));
#pragma unroll
        for (int i = 0; i < TT; i++) {
          const int t = t0 + i;
          if (t < g.nt) {
            const size_t row = t * R;
            const double2 pp = (t + 1 == g.nt) ? pnew(__ldcv(&a.r_next[j]), __ldcv(&po_next[j]))
                                               : pnew(__ldcg(&a.r[row + R + j]), __ldcg(&po[row + R + j]));
            const double2 pxp = pnew(__ldcg(&a.r[row + jp]), __ldcg(&po[row + jp]));
            const double2 pxm = pnew(__ldcg(&a.r[row + jm]), __ldcg(&po[row + jm]));
            const double2 w0c = a.W0[row + j];
            const double2 w1c = a.W1[row + j];
            const double2 w1m = a.W1[row + jm];
            const double fr = w0c.x * e_p, fi = w0c.y * e_p;
            const double br = w0m.x * e_m, bi = w0m.y * e_m;
            double hr = fr * pp.x - fi * pp.y;
            double hi = fr * pp.y + fi * pp.x;
            hr -= br * pm.x + bi * pm.y;
            hi -= br * pm.y - bi * pm.x;
            hr += w1c.x * pxp.x - w1c.y * pxp.y;
            hi += w1c.x * pxp.y + w1c.y * pxp.x;
            hr -= w1m.x * pxm.x + w1m.y * pxm.y;
            hi -= w1m.x * pxm.y - w1m.y * pxm.x;
            double2 o;
            o.x = m * pc.x + hr;
            o.y = m * pc.y + hi;
            pn[row + j] = pc;
            a.Mp[row + j] = o;
            acc += o.x * o.x + o.y * o.y;   // <p, M^dagger M p> = |M p|^2
            pm = pc;
            pc = pp;
            w0m = w0c;
          }
        }
      }
    }
    block_partial(acc, g, s, red);
    grid_arrive(sl.gbar, bar_target);
    if (blockIdx.x == 0) block0_scalars<FIN_PQ, TB_RED_PQ>(g, s, sl, red, gen, 3 * gen + 1, a.sysfence);
    wait_go(sl, 3 * gen + 1);

    // ---- B: q = M^dagger Mp on the fly, x += alpha p, r -= alpha q, ||r||^2.  The neighbours' Mp rows are complete:
    // their owners contributed to the |Mp|^2 sum.
    acc = 0.0;
    const double al = chain ? __ldcg(&s.alpha[c]) : 0.0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const BlockPos b = tile_pos(g, tile);
      if (b.valid && act) {
        const size_t j = (size_t)b.x * g.C + b.c;
        const size_t jp = (size_t)((b.x + 1 == g.nx) ? 0 : b.x + 1) * g.C + b.c;
        const size_t jm = (size_t)((b.x == 0) ? g.nx - 1 : b.x - 1) * g.C + b.c;
        const int t0 = b.ttile * TT;
        double2 pm, w0m;
        if (t0 == 0) {
          pm = __ldcv(&a.mp_prev[(size_t)(g.nt - 1) * R + j]);
          w0m = __ldcv(&a.W0_prev[(size_t)(g.nt - 1) * R + j]);
        } else {
          pm = __ldcg(&a.Mp[(size_t)(t0 - 1) * R + j]);
          w0m = a.W0[(size_t)(t0 - 1) * R + j];
        }
        double2 pc = __ldcg(&a.Mp[t0 * R + j]);
#pragma unroll
        for (int i = 0; i < TT; i++) {
          const int t = t0 + i;
          if (t < g.nt) {
            const size_t row = t * R;
            const double2 pp = (t + 1 == g.nt) ? __ldcv(&a.mp_next[j]) : __ldcg(&a.Mp[row + R + j]);
            const double2 pxp = __ldcg(&a.Mp[row + jp]);
            const double2 pxm = __ldcg(&a.Mp[row + jm]);
            const double2 w0c = a.W0[row + j];
            const double2 w1c = a.W1[row + j];
            const double2 w1m = a.W1[row + jm];
            const double2 pv = __ldcg(&pn[row + j]);
            double2 xv = __ldcg(&a.x[row + j]), rv = __ldcg(&a.r[row + j]);
            const double fr = w0c.x * e_m, fi = w0c.y * e_m;   // M^dagger: e^{-mu} on the +t hop
            const double br = w0m.x * e_p, bi = w0m.y * e_p;
            double hr = fr * pp.x - fi * pp.y;
            double hi = fr * pp.y + fi * pp.x;
            hr -= br * pm.x + bi * pm.y;
            hi -= br * pm.y - bi * pm.x;
            hr += w1c.x * pxp.x - w1c.y * pxp.y;
            hi += w1c.x * pxp.y + w1c.y * pxp.x;
            hr -= w1m.x * pxm.x + w1m.y * pxm.y;
            hi -= w1m.x * pxm.y - w1m.y * pxm.x;
            const double qx = m * pc.x - hr, qy = m * pc.y - hi;
            xv.x += al * pv.x;
            xv.y += al * pv.y;
            rv.x -= al * qx;
            rv.y -= al * qy;
            a.x[row + j] = xv;
            a.r[row + j] = rv;
            acc += rv.x * rv.x + rv.y * rv.y;
            pm = pc;
            pc = pp;
            w0m = w0c;
          }
        }
      }
    }
    block_partial(acc, g, s, red);
    grid_arrive(sl.gbar, bar_target);
    if (blockIdx.x == 0) block0_scalars<FIN_RR, TB_RED_RR>(g, s, sl, red, gen, 3 * gen + 2, a.sysfence);
    wait_go(sl, 3 * gen + 2);
    n_act = *(volatile int *)s.n_active;
    gen++;
  }
  // every rank leaves in the same iteration (identical scalars).  Leave the epoch and the flags as the multi-kernel
  // protocol expects them: nobody reads p or Mp of this solve any more.
  grid_barrier(sl.gbar, bar_target);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *sl.seq = gen;
    __threadfence_system();
    for (int kind = 0; kind < TB_NFLAGS; kind++) {
      *(volatile int *)(sl.sig_prev + kind * 2) = gen;
      *(volatile int *)(sl.sig_next + kind * 2) = gen;
    }
    __threadfence_system();
  }
}

// ---------------------------------------------------------------------------------------------------------
// One-launch slab solve, second form: EVERY block takes part in the all-reduce (default; TB_SLAB_SYNC=0 selects the
// block-0 form above).  A synchronisation point of the form above is a chain of six hand-overs: block partials ->
// counter -> block 0 (which first has to notice the last arrival) -> peer stores -> block 0's scalar update in global
// memory -> go flag -> every block reloads the scalars; ~14 us at 2048^2.  Here:
//   * the block whose ticket is the LAST one of the grid gathers the block partials at once (nobody polls the counter)
//     and stores {partial, partial xor tag} into every rank's slot table, NREP replicas on separate L2 lines;
//   * every block polls the P slots of ITS replica in its own GPU's memory (one warp, one 128-byte line at C = 1), adds
//     them in rank order and evaluates alpha / beta / convergence itself, in registers: the CG scalars are bitwise
//     identical in every block of every rank, so there is neither a go flag nor a scalar reload, and the decisions
//     (stop, iterate) are taken everywhere in the same iteration.
// The all-reduce still doubles as the cross-GPU fence for the halo rows (a rank's partial is stored only after all its
// blocks have arrived behind a device-scope fence, and behind a system-scope fence of the storing threads).  A slot
// of kind K and generation g is overwritten at generation g + 1 of kind K only after the owner passed the reduction of
// the other kind in between, which needs every block of every rank to have arrived there, i.e. to have read slot g.
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}

// Release / acquire fences of the synchronisation points.  __threadfence() and __threadfence_system() are the
// sequentially consistent fences (MEMBAR.SC); the hand-overs below are plain release -> acquire chains, for which
// fence.acq_rel (MEMBAR.ALL) is what the PTX memory model asks for, and a synchronisation point is a chain of six of them.
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;\n" ::: "memory"); }
__device__ __forceinline__ void fence_sys() { asm volatile("fence.acq_rel.sys;\n" ::: "memory"); }

// The fences around the peer exchange.  Measured on this box (tools/nvlink_pingpong.cu, profiles/nvlink_pingpong_r02.txt):
// a flag store is visible on the other GPU after 1.13 us; a device-scope fence costs 0.23 us, a SYSTEM-scope fence
// 1.56 us -- two of them per all-reduce were a third of a synchronisation point.  Device scope is what the data path
// needs: halo rows never travel, the peer reads them out of THIS GPU's L2 over NVLink with ld.cv (no cached copy on its
// side), and every block's rows were ordered into that L2 by its own device-scope fence before its arrival was counted;
// the tag leaves this GPU after the fence below, so a peer that has seen the tag and then asks this L2 for a row finds
// it.  In the other direction value and tag travel in one 16-byte store.  (The PTX memory model asks for system scope
// between threads of two GPUs; TB_SLAB_SYSFENCE=1 restores it.)
__device__ __forceinline__ void xgpu_fence(int sysfence) {
  if (sysfence) fence_sys();
  else fence_gpu();
}

struct ChainScalars {   // CG scalars of the thread's chain (c = threadIdx.x & (bc - 1)), identical in every block
  double rr_old, rr_init, rr, alpha, beta;
  int active, iters, status;
};

template <int FIN>
__device__ __forceinline__ void update_scalars(double total, ChainScalars &cs, const TbCgState &s) {
  if (FIN == FIN_INIT) {
    cs.rr = cs.rr_old = cs.rr_init = total;
    if (total < s.accuracy) {   // hmc.c:359-361, x is already zero
      cs.status = TB_CG_ZERO_SOURCE;
      cs.active = 0;
    }
  } else if (FIN == FIN_PQ) {
    if (cs.active) cs.alpha = cs.rr_old / total;   // hmc.c:371
  } else if (FIN == FIN_RR) {
    if (cs.active) {
      const int it = ++cs.iters;
      cs.rr = total;
      int st = -1;
      if (total < s.accuracy) st = TB_CG_CONVERGED;                                                   // hmc.c:381
      else if (!(total == total) || total / cs.rr_init > TB_DIVERGENCE_RATIO) st = TB_CG_DIVERGED;   // hmc.c:383
      else if (it >= s.max_iter - 1) st = TB_CG_MAXITER;                                              // hmc.c:364
      if (st >= 0) {
        cs.status = st;
        cs.active = 0;
      } else {
        cs.beta = total / cs.rr_old;   // hmc.c:390
        cs.rr_old = total;             // hmc.c:394
      }
    }
  }
}

// Per-chain sum of `acc` over all blocks of all ranks; returns the total of the thread's chain (valid in every thread).
// mode 2: every block polls the peers' slots (NREP replicas).  mode 1: only the block that arrived last exchanges with
// the peers (one quiet slot per peer: pollers on a line delay the NVLink store they wait for); it then publishes the
// all-reduced totals {total, total xor tag} on NREP lines of its OWN GPU, which the other blocks poll.
// tl: optional time stamps of this synchronisation point (4 words: see TB_SLAB_TL_WORDS), nullptr when not recording.
template <int RED>
__device__ __forceinline__ double slab_allreduce(double acc, const TbGeom &g, const TbCgState &s, const TbSlab &sl,
                                                 double *red, unsigned long long &target, int gen, int mode,
                                                 int sysfence, unsigned long long *tl) {
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = (blockDim.x + 31) >> 5;
  const int c_local = tid & (g.bc - 1), x_local = tid >> g.bc_shift;
  const bool stamp = tl && tid == 0 && (blockIdx.x & 15) == 0;   // a sample of the blocks records
  const long long tag = (long long)((unsigned long long)gen * 0x9E3779B97F4A7C15ULL);
  for (int o = 16; o >= g.bc; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __syncthreads();   // red may still be read by the previous call
  if (lane < g.bc) red[warp * g.bc + lane] = acc;
  fence_gpu();   // this thread's field stores are visible device-wide before the block's arrival is counted
  __syncthreads();
  if (stamp) atomicMax(&tl[1], global_ns());
  if (warp == 0) {
    if (lane < g.bc) {
      double t = 0.0;
      for (int w = 0; w < nwarp; w++) t += red[w * g.bc + lane];
      __stcg(&s.partial[(size_t)blockIdx.x * g.Cpad + lane], t);
      fence_gpu();
    }
    __syncwarp();
    if (lane == 0) {
      target += gridDim.x;
      s_last = atomicAdd(sl.gbar, 1ULL) == target - 1;
    }
  }
  __syncthreads();
  const bool last = s_last != 0;   // block-uniform: this block arrived last, every other block's partial and rows are visible
  if (last) {
    fence_gpu();
    double sum = 0.0;
    for (int blk = x_local; blk < (int)gridDim.x; blk += g.bx) sum += __ldcg(&s.partial[(size_t)blk * g.Cpad + c_local]);
    const double mine = block_sum_chains(sum, g, red);   // threads < bc
    __syncthreads();
    if (tid < g.bc) red[tid] = mine;
    __syncthreads();
    const int nrep_out = mode == 2 ? sl.nrep : 1;
    const int nst = sl.P * nrep_out * g.bc;
    if (tid < nst) xgpu_fence(sysfence);   // this GPU's halo rows before the tag, for the peers
    for (int i = tid; i < nst; i += blockDim.x) {
      const int c = i & (g.bc - 1), k = i >> g.bc_shift, rep = k % nrep_out, q = k / nrep_out;
      const double v = red[c];
      st_volatile_v2(sl.peer_red3[q] + (size_t)rep * sl.rep_stride + (size_t)(RED * sl.P + sl.rank) * g.Cpad + c,
                     __double_as_longlong(v), __double_as_longlong(v) ^ tag);
    }
    if (tl && tid == 0) tl[2] = global_ns();
    __syncthreads();   // red is rewritten below
  }
  double total;
  if (mode == 2 || last) {
    const bool io = tid < sl.P * g.bc;   // thread (q, chain) polls the slot rank q writes into this GPU's memory
    if (io) {
      const int q = tid >> g.bc_shift, rep = mode == 2 ? (int)(blockIdx.x % sl.nrep) : 0;
      const double2 *in = sl.red3 + (size_t)rep * sl.rep_stride + (size_t)(RED * sl.P + q) * g.Cpad + c_local;
      const long long t0 = clock64();
      double theirs;
      for (;;) {
        long long wx, wy;
        ld_volatile_v2(in, wx, wy);
        if ((wx ^ wy) == tag) { theirs = __longlong_as_double(wx); break; }
        if (mode == 2) __nanosleep(20);
        if (clock64() - t0 > TB_PERSIST_SPIN_CYCLES) __trap();   // a launch error on this rank instead of a hung box
      }
      xgpu_fence(sysfence);   // acquire side: the neighbours' rows behind their tags (and this SM's L1)
      red[tid] = theirs;        // [q][chain]
    }
    __syncthreads();
    total = 0.0;
    for (int r = 0; r < sl.P; r++) total += red[r * g.bc + c_local];   // rank order: the same bits on every rank
    if (mode != 2 && tid < sl.nrep * g.bc) {   // publish the totals to the other blocks of this GPU
      fence_gpu();
      st_volatile_v2(sl.bcast + (size_t)(tid >> g.bc_shift) * sl.bcast_stride + (size_t)RED * g.Cpad + c_local,
                     __double_as_longlong(total), __double_as_longlong(total) ^ tag);
    }
  } else {
    if (tid < g.bc) {
      const double2 *in = sl.bcast + (size_t)(blockIdx.x % sl.nrep) * sl.bcast_stride + (size_t)RED * g.Cpad + tid;
      const long long t0 = clock64();
      double v;
      for (;;) {
        long long wx, wy;
        ld_volatile_v2(in, wx, wy);
        if ((wx ^ wy) == tag) { v = __longlong_as_double(wx); break; }
        __nanosleep(20);
        if (clock64() - t0 > TB_PERSIST_SPIN_CYCLES) __trap();
      }
      fence_gpu();   // acquire: this GPU's rows (the publisher saw every arrival) and, through it, the neighbours'
      red[tid] = v;
    }
    __syncthreads();
    total = red[c_local];
  }
  if (stamp) atomicMax(&tl[3], global_ns());
  return total;
}

// Work of one block: column strip `xtile` (bx sites wide) and rows [t_begin, t_end) of the slab.  The nt rows of a strip
// are cut into nseg segments as evenly as integer rows allow, nxtiles * nseg <= the number of co-resident blocks when the
// slab allows it: every block has work and none has more than one row above the average (2048^2 on 8 GPUs: 440 blocks
// of 4 or 5 rows instead of 256 tiles of 8 rows on 148 SMs).
struct SlabSeg { int xtile, t_begin, t_end; };
__device__ __forceinline__ SlabSeg slab_seg(const TbGeom &g, int w, int nseg) {
  SlabSeg sg;
  sg.xtile = w % g.nxtiles;
  const int j = w / g.nxtiles;
  sg.t_begin = (int)((long long)j * g.nt / nseg);
  sg.t_end = (int)((long long)(j + 1) * g.nt / nseg);
  return sg;
}

template <int TT>
__global__ void __launch_bounds__(TB_MAX_BLOCK, 3)
slab_cg_onelaunch_kernel(const SlabCgArgs a, const TbGeom g, const TbCgState s, const TbSlab sl, const int nseg,
                         const int mode) {
  __shared__ double red[TB_MAX_BLOCK];
  const int nwork = g.nxtiles * nseg;
  const int R = g.R;   // 32-bit indices: a slab of the one-launch solve has at most 2^27 elements per field
  unsigned long long bar_target = 0;
  const int E0 = *(volatile int *)sl.seq;   // rewritten only after the last grid barrier
  const int c = threadIdx.x & (g.bc - 1), x_local = threadIdx.x >> g.bc_shift;
  const bool chain = c < g.C;
  const double m = chain ? a.mass[c] : 0.0;
  const double e_p = chain ? a.emu[c] : 1.0, e_m = chain ? a.emmu[c] : 1.0;
  ChainScalars cs = {0.0, 0.0, 0.0, 0.0, 0.0, chain ? 1 : 0, 0, TB_CG_MAXITER};
  auto stamps = [&](int k, int half) -> unsigned long long * {
    return (sl.timeline && k < TB_SLAB_TL_ITERS) ? sl.timeline + ((size_t)k * 2 + half) * 4 : nullptr;
  };
  // rows [t0, t0 + nrows) of the chunks of a segment: at most TT rows each, as even as integer rows allow
  auto chunks = [&](const SlabSeg &sg) { return (sg.t_end - sg.t_begin + TT - 1) / TT; };
  auto chunk_lo = [&](const SlabSeg &sg, int nch, int ch) { return sg.t_begin + ch * (sg.t_end - sg.t_begin) / nch; };

  // ---- x = 0, r = p = b, ||b||^2 (hmc.c:349-361): generation E0 + 1 of p
  double acc = 0.0;
  for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
    const SlabSeg sg = slab_seg(g, w, nseg);
    // the neighbours may still read the previous generation of p (an apply queued before this solve)
    if (sg.t_begin == 0 || sg.t_end == g.nt) {
      if (threadIdx.x == 0) {
        if (sg.t_begin == 0) spin_until(sl.flags + TB_FLAG_PDONE * 2 + 0, E0);
        if (sg.t_end == g.nt) spin_until(sl.flags + TB_FLAG_PDONE * 2 + 1, E0);
        __threadfence_system();
      }
      __syncthreads();
    }
    const int x = sg.xtile * g.bx + x_local;
    if (chain && x < g.nx) {
      const int j = x * g.C + c;
      for (int t = sg.t_begin; t < sg.t_end; t++) {
        const int k = t * R + j;
        const double2 v = a.b[k];
        a.x[k] = make_double2(0.0, 0.0);
        a.r[k] = v;
        a.p[k] = v;
        acc += v.x * v.x + v.y * v.y;
      }
    }
  }
  int gen = E0 + 1;   // generation of p
  update_scalars<FIN_INIT>(slab_allreduce<TB_RED_INIT>(acc, g, s, sl, red, bar_target, gen, mode, a.sysfence, nullptr), cs, s);

  // p is double-buffered: odd iterations read p_old from a.p and write p_new into a.p1, even ones the other way round.
  // Fields of this GPU that other blocks rewrite between the phases are read with ordinary (L1-cached) loads: every
  // synchronisation point ends with an acquire fence and a CTA barrier in every block, which is what makes ordinary
  // loads after a grid-wide barrier see the other blocks' stores; the neighbours' rows come over NVLink with ld.cv.
  auto n_active = [&] { return __popc(__ballot_sync(0xffffffffu, (int)(threadIdx.x & 31) < g.bc && cs.active)); };
  for (int k = 1; n_active() > 0; k++) {
    const bool act = cs.active != 0;
    const double be = cs.beta;   // 0 in the first iteration: p = r (hmc.c:352-353)
    const bool odd = k & 1;      // p_old of iteration 1 is a.p (= b, times beta = 0)
    const double2 *po = odd ? a.p : a.p1, *po_prev = odd ? a.p_prev : a.p1_prev, *po_next = odd ? a.p_next : a.p1_next;
    double2 *pn = odd ? a.p1 : a.p;
    // p = r + beta p (hmc.c:391-392) wherever the stencil needs it; the same fma everywhere, so the five copies of a
    // site's p agree to the bit with the one that is stored
    auto pnew = [&](const double2 rv, const double2 pv) {
      return make_double2(fma(be, pv.x, rv.x), fma(be, pv.y, rv.y));
    };
    unsigned long long *tlA = stamps(k - 1, 0), *tlB = stamps(k - 1, 1);
    if (tlA && threadIdx.x == 0 && (blockIdx.x & 15) == 0) atomicMin(&tlA[0], global_ns());
    // ---- A: p = r + beta p on the fly, Mp = M p, |Mp|^2.  The neighbours' rows of r and of the old p are complete:
    // their owners passed the ||r||^2 all-reduce of the previous iteration.
    acc = 0.0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
      const SlabSeg sg = slab_seg(g, w, nseg);
      const int x = sg.xtile * g.bx + x_local;
      if (!(chain && x < g.nx && act)) continue;
      const int j = x * g.C + c;
      const int jp = ((x + 1 == g.nx) ? 0 : x + 1) * g.C + c;
      const int jm = ((x == 0) ? g.nx - 1 : x - 1) * g.C + c;
      const int nch = chunks(sg);
      for (int ch = 0; ch < nch; ch++) {
        const int t0 = chunk_lo(sg, nch, ch), nrows = chunk_lo(sg, nch, ch + 1) - t0;
        double2 pm, w0m;
        if (t0 == 0) {
          const int kk = (g.nt - 1) * R + j;
          pm = pnew(__ldcv(&a.r_prev[kk]), __ldcv(&po_prev[kk]));
          w0m = __ldcv(&a.W0_prev[kk]);
        } else {
          const int kk = (t0 - 1) * R + j;
          pm = pnew(a.r[kk], po[kk]);
          w0m = a.W0[kk];
        }
        double2 pc = pnew(a.r[t0 * R + j], po[t0 * R + j]);
#pragma unroll
        for (int i = 0; i < TT; i++) {
          if (i < nrows) {
            const int t = t0 + i;
            const int row = t * R;
            const double2 pp = (t + 1 == g.nt) ? pnew(__ldcv(&a.r_next[j]), __ldcv(&po_next[j]))
                                               : pnew(a.r[row + R + j], po[row + R + j]);
            const double2 pxp = pnew(a.r[row + jp], po[row + jp]);
            const double2 pxm = pnew(a.r[row + jm], po[row + jm]);
            const double2 w0c = a.W0[row + j];
            const double2 w1c = a.W1[row + j];
            const double2 w1m = a.W1[row + jm];
            const double fr = w0c.x * e_p, fi = w0c.y * e_p;
            const double br = w0m.x * e_m, bi = w0m.y * e_m;
            double hr = fr * pp.x - fi * pp.y;
            double hi = fr * pp.y + fi * pp.x;
            hr -= br * pm.x + bi * pm.y;
            hi -= br * pm.y - bi * pm.x;
            hr += w1c.x * pxp.x - w1c.y * pxp.y;
            hi += w1c.x * pxp.y + w1c.y * pxp.x;
            hr -= w1m.x * pxm.x + w1m.y * pxm.y;
            hi -= w1m.x * pxm.y - w1m.y * pxm.x;
            double2 o;
            o.x = m * pc.x + hr;
            o.y = m * pc.y + hi;
            pn[row + j] = pc;
            a.Mp[row + j] = o;
            acc += o.x * o.x + o.y * o.y;   // <p, M^dagger M p> = |M p|^2
            pm = pc;
            pc = pp;
            w0m = w0c;
          }
        }
      }
    }
    update_scalars<FIN_PQ>(slab_allreduce<TB_RED_PQ>(acc, g, s, sl, red, bar_target, gen, mode, a.sysfence, tlA), cs, s);

    // ---- B: q = M^dagger Mp on the fly, x += alpha p, r -= alpha q, ||r||^2.  The neighbours' Mp rows are complete:
    // their owners contributed to the |Mp|^2 sum.
    if (tlB && threadIdx.x == 0 && (blockIdx.x & 15) == 0) atomicMin(&tlB[0], global_ns());
    acc = 0.0;
    const double al = cs.alpha;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
      const SlabSeg sg = slab_seg(g, w, nseg);
      const int x = sg.xtile * g.bx + x_local;
      if (!(chain && x < g.nx && act)) continue;
      const int j = x * g.C + c;
      const int jp = ((x + 1 == g.nx) ? 0 : x + 1) * g.C + c;
      const int jm = ((x == 0) ? g.nx - 1 : x - 1) * g.C + c;
      const int nch = chunks(sg);
      // chunks in DESCENDING order: phase A went up, so this phase starts on the rows A wrote last (Mp, the new p) and
      // ends on chunk 0, where the next phase A starts with the r this phase just wrote: on slabs larger than L2 every
      // phase begins with what is still cached
      for (int ch = nch - 1; ch >= 0; ch--) {
        const int t0 = chunk_lo(sg, nch, ch), nrows = chunk_lo(sg, nch, ch + 1) - t0;
        double2 pm, w0m;
        if (t0 == 0) {
          pm = __ldcv(&a.mp_prev[(g.nt - 1) * R + j]);
          w0m = __ldcv(&a.W0_prev[(g.nt - 1) * R + j]);
        } else {
          pm = a.Mp[(t0 - 1) * R + j];
          w0m = a.W0[(t0 - 1) * R + j];
        }
        double2 pc = a.Mp[t0 * R + j];
#pragma unroll
        for (int i = 0; i < TT; i++) {
          if (i < nrows) {
            const int t = t0 + i;
            const int row = t * R;
            const double2 pp = (t + 1 == g.nt) ? __ldcv(&a.mp_next[j]) : a.Mp[row + R + j];
            const double2 pxp = a.Mp[row + jp];
            const double2 pxm = a.Mp[row + jm];
            const double2 w0c = a.W0[row + j];
            const double2 w1c = a.W1[row + j];
            const double2 w1m = a.W1[row + jm];
            const double2 pv = pn[row + j];
            double2 xv = a.x[row + j], rv = a.r[row + j];
            const double fr = w0c.x * e_m, fi = w0c.y * e_m;   // M^dagger: e^{-mu} on the +t hop
            const double br = w0m.x * e_p, bi = w0m.y * e_p;
            double hr = fr * pp.x - fi * pp.y;
            double hi = fr * pp.y + fi * pp.x;
            hr -= br * pm.x + bi * pm.y;
            hi -= br * pm.y - bi * pm.x;
            hr += w1c.x * pxp.x - w1c.y * pxp.y;
            hi += w1c.x * pxp.y + w1c.y * pxp.x;
            hr -= w1m.x * pxm.x + w1m.y * pxm.y;
            hi -= w1m.x * pxm.y - w1m.y * pxm.x;
            const double qx = m * pc.x - hr, qy = m * pc.y - hi;
            xv.x += al * pv.x;
            xv.y += al * pv.y;
            rv.x -= al * qx;
            rv.y -= al * qy;
            a.x[row + j] = xv;
            a.r[row + j] = rv;
            acc += rv.x * rv.x + rv.y * rv.y;
            pm = pc;
            pc = pp;
            w0m = w0c;
          }
        }
      }
    }
    update_scalars<FIN_RR>(slab_allreduce<TB_RED_RR>(acc, g, s, sl, red, bar_target, gen, mode, a.sysfence, tlB), cs, s);
    gen++;
  }
  // every block of every rank leaves in the same iteration (identical scalars).  Block 0 writes the outcome; then leave
  // the epoch and the flags as the multi-kernel protocol expects them: nobody reads p or Mp of this solve any more.
  if (blockIdx.x == 0 && (int)threadIdx.x < g.bc && chain) {
    s.status[c] = cs.status;
    s.iters[c] = cs.iters;
    s.rr[c] = cs.rr;
    s.rr_init[c] = cs.rr_init;
    s.rr_old[c] = cs.rr_old;
    s.alpha[c] = cs.alpha;
    s.beta[c] = cs.beta;
    s.active[c] = 0;
  }
  grid_barrier(sl.gbar, bar_target);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *s.n_active = 0;
    s.tile_active[0] = 0;
    *sl.seq = gen;
    __threadfence_system();
    for (int kind = 0; kind < TB_NFLAGS; kind++) {
      *(volatile int *)(sl.sig_prev + kind * 2) = gen;
      *(volatile int *)(sl.sig_next + kind * 2) = gen;
    }
    __threadfence_system();
  }
}

// plain per-chain Re<a,b>
template <int TT>
__global__ void __launch_bounds__(TB_MAX_BLOCK)
dot_kernel(const double2 *__restrict__ a, const double2 *__restrict__ bb, const TbGeom g, const TbCgState s,
           const TbSlab sl) {
  __shared__ double red[TB_MAX_BLOCK];
  const BlockPos b = block_pos(g);
  double acc = 0.0;
  if (b.valid) {
    const size_t R = (size_t)g.R;
    const size_t j = (size_t)b.x * g.C + b.c;
    const int t0 = b.ttile * TT;
#pragma unroll
    for (int i = 0; i < TT; i++) {
      const int t = t0 + i;
      if (t < g.nt) {
        const size_t k = t * R + j;
        const double2 u = a[k], v = bb[k];
        acc += u.x * v.x + u.y * v.y;
      }
    }
  }
  reduce_finalize<FIN_DOT, false, 0>(acc, g, s, sl, b, red);
}

__global__ void cg_reset_kernel(const TbGeom g, const TbCgState s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.Cpad) {
    const int valid = i < g.C;
    s.active[i] = valid;
    s.status[i] = TB_CG_MAXITER;
    s.iters[i] = 0;
    s.alpha[i] = 0.0;
    s.beta[i] = 0.0;
  }
  if (i < g.nctiles) {
    int n = g.C - i * g.bc;
    s.tile_active[i] = n > g.bc ? g.bc : (n < 0 ? 0 : n);
    s.ticket[i] = 0u;
  }
  if (i == 0) *s.n_active = g.C;
}

// W0 = s0(t) 1/2 eta0(x) (cos A0, sin A0) ; W1 = s1(x) 1/2 (cos A1, sin A1); device layout in and out, chains
// [c0, c0+n).  eta0 = (-1)^x (hmc.c:917-921), s = -1 on the wrap link (hmc.c:143-148,165-170); t_off / nt_global
// place a slab inside the global lattice.
__global__ void links_kernel(const double2 *__restrict__ A, double2 *__restrict__ W0, double2 *__restrict__ W1,
                             int nt, int nx, int C, int c0, int n, int t_off, int nt_global) {
  const size_t total = (size_t)nt * nx * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t site = i / n;
    const size_t k = site * C + c0 + (i % n);
    const int x = (int)(site % nx);
    const int t = (int)(site / nx) + t_off;
    const double2 a = A[k];
    double s0, c0v, s1, c1;
    sincos(a.x, &s0, &c0v);
    sincos(a.y, &s1, &c1);
    double f0 = (x & 1) ? -0.5 : 0.5;
    if (t == nt_global - 1) f0 = -f0;
    const double f1 = (x == nx - 1) ? -0.5 : 0.5;
    W0[k] = make_double2(f0 * c0v, f0 * s0);
    W1[k] = make_double2(f1 * c1, f1 * s1);
  }
}

// One field for the whole batch: A[t][x] -> the compact links Ws = (W0 [V], W1 [V]) and the per-chain arrays W0 / W1 /
// Adev filled with copies (links_kernel's arithmetic: the same links bit for bit)
__global__ void links_shared_kernel(const double2 *__restrict__ A, double2 *__restrict__ Ws, double2 *__restrict__ W0,
                                    double2 *__restrict__ W1, double2 *__restrict__ Adev, int nt, int nx, int C) {
  const size_t V = (size_t)nt * nx;
  for (size_t site = (size_t)blockIdx.x * blockDim.y + threadIdx.y; site < V; site += (size_t)gridDim.x * blockDim.y) {
    const int x = (int)(site % nx);
    const int t = (int)(site / nx);
    const double2 a = A[site];
    double s0, c0v, s1, c1;
    sincos(a.x, &s0, &c0v);
    sincos(a.y, &s1, &c1);
    double f0 = (x & 1) ? -0.5 : 0.5;
    if (t == nt - 1) f0 = -f0;
    const double f1 = (x == nx - 1) ? -0.5 : 0.5;
    const double2 w0 = make_double2(f0 * c0v, f0 * s0), w1 = make_double2(f1 * c1, f1 * s1);
    if (threadIdx.x == 0) {
      Ws[site] = w0;
      Ws[V + site] = w1;
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      W0[site * C + c] = w0;
      W1[site * C + c] = w1;
      Adev[site * C + c] = a;
    }
  }
}

// Family B (vec_ops.c:96-172): links are real constants masked by the occupation field, W_mu(n) = s 1/2 eta_mu
// if both n and n+mu^ are free, else 0 (hops into or out of an occupied site are dropped, vec_ops.c:110-128); the
// site mass is m on free sites and 1 on occupied ones (identity row, vec_ops.c:130).  occ: [t][x][c] ints.
__global__ void occupancy_links_kernel(const int *__restrict__ occ, const double *__restrict__ mass,
                                       double2 *__restrict__ W0, double2 *__restrict__ W1,
                                       double *__restrict__ msite, int nt, int nx, int C, int bc) {
  const size_t total = (size_t)nt * nx * C;
  const size_t R = (size_t)nx * C;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(k % C);
    const size_t site = k / C;
    const int x = (int)(site % nx);
    const int t = (int)(site / nx);
    const size_t kt = (size_t)((t + 1 == nt) ? 0 : t + 1) * R + (size_t)x * C + c;
    const size_t kx = (size_t)t * R + (size_t)((x + 1 == nx) ? 0 : x + 1) * C + c;
    const bool free_n = occ[k] == 0;
    double f0 = (x & 1) ? -0.5 : 0.5;
    if (t == nt - 1) f0 = -f0;
    // the x link across the boundary: ANTISYMMETRIC -1/2, SYMMETRIC +1/2 (vec_ops.c:201-207), OPENX none
    // (fermionbag.c:713-717,761-765)
    const double f1 = (x == nx - 1) ? (bc == TB_BC_SYMMETRIC ? 0.5 : (bc == TB_BC_OPENX ? 0.0 : -0.5)) : 0.5;
    W0[k] = make_double2((free_n && occ[kt] == 0) ? f0 : 0.0, 0.0);
    W1[k] = make_double2((free_n && occ[kx] == 0) ? f1 : 0.0, 0.0);
    msite[k] = free_n ? mass[c] : 1.0;
  }
}

__global__ void transpose_int_kernel(const int *__restrict__ src, int *__restrict__ dst, int rows, int cols) {
  // src[rows][cols] -> dst[cols][rows]
  __shared__ int tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = src[(size_t)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

// src[r * ld_src + c], r < rows, c < cols   ->   dst[c * ld_dst + r]   (tiled transpose of double2)
__global__ void transpose_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, int rows, int cols,
                                 size_t ld_src, size_t ld_dst) {
  __shared__ double2 tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = src[(size_t)r * ld_src + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[(size_t)c * ld_dst + r] = tile[threadIdx.x][i];
  }
}


}  // namespace

// ---------------------------------------------------------------------------------------------------------
// host-side launch wrappers

#ifndef TB_PIPE_TILED_DEFAULT
#define TB_PIPE_TILED_DEFAULT true
#endif

// cuTensorMapEncodeTiled, through the runtime's driver entry-point query (the library does not link libcuda)
typedef CUresult (*TmapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return (TmapEncodeFn)p;
  }();
  return fn;
}

// Tensor map of one device-layout vector [t][x][chain] of the context for the tiled staged kernels: dimension 0 = the
// 2 C doubles of a site, dimension 1 = the nt * nx sites, box = (16 chains, 16 sites).  Cached by base pointer.
static int tmap_for(tb_ctx *ctx, const void *base, CUtensorMap *out) {
  for (int i = 0; i < TB_TMAP_CACHE; i++)
    if (ctx->tmap_ptr[i] == base) {
      memcpy(out, ctx->tmap_store[i], sizeof(CUtensorMap));
      return TB_OK;
    }
  static_assert(sizeof(CUtensorMap) == 128, "tb_ctx::tmap_store");
  const TbGeom &g = ctx->gp;
  const cuuint64_t dims[2] = {(cuuint64_t)2 * ctx->C, (cuuint64_t)ctx->nt * ctx->nx};
  const cuuint64_t strides[1] = {(cuuint64_t)ctx->C * sizeof(double2)};
  const cuuint32_t box[2] = {(cuuint32_t)2 * g.bc, (cuuint32_t)g.bx};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = tmap_encoder()(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(base), dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    tb_set_error("cuTensorMapEncodeTiled failed (%d) for a %d x %d lattice, %d chains", (int)rc, ctx->nt, ctx->nx, ctx->C);
    return TB_ECUDA;
  }
  const int slot = ctx->tmap_next;
  ctx->tmap_next = (slot + 1) % TB_TMAP_CACHE;
  ctx->tmap_ptr[slot] = base;
  memcpy(ctx->tmap_store[slot], out, sizeof(CUtensorMap));
  return TB_OK;
}

int tb_choose_geom(tb_ctx *ctx) {
  TbGeom &g = ctx->g;
  g.nt = ctx->nt;
  g.nx = ctx->nx;
  g.C = ctx->C;
  g.R = ctx->nx * ctx->C;
  int bc = 1;
  while (bc < ctx->C && bc < 32) bc <<= 1;
  int bx = TB_MAX_BLOCK / bc;
  int nxp = 1;
  while (nxp < ctx->nx) nxp <<= 1;
  if (bx > nxp) bx = nxp;
  if (bc * bx < 32) bx = 32 / bc;  // at least one warp
  g.bc = bc;
  g.bx = bx;
  g.bc_shift = 0;
  while ((1 << g.bc_shift) < bc) g.bc_shift++;
  g.nctiles = (ctx->C + bc - 1) / bc;
  g.nxtiles = (ctx->nx + bx - 1) / bx;
  g.Cpad = g.nctiles * bc;
  int tt = ctx->tune_tt;
  if (const char *e = getenv("TB_FORCE_ROWS")) tt = atoi(e);   // development override, wins over tb_set_tuning
  if (tt != 1 && tt != 2 && tt != 4 && tt != 8 && tt != 16) {
    tt = 16;
    const long nb1 = (long)g.nctiles * g.nxtiles;
    while (tt > 1 && nb1 * ((ctx->nt + tt - 1) / tt) < 6L * TB_NUM_SMS_B200) tt >>= 1;
  }
  g.tt = tt;
  g.nttiles = (ctx->nt + tt - 1) / tt;
  g.nslots = g.nxtiles * g.nttiles;
  g.ta_shift = g.bc_shift;
  // TMA-staged kernels: a tile holds EVERY chain of its sites, so the sites of a tile are one contiguous run of a row
  // and a row of the tile is one bulk copy per array (4 KB) plus three halo sites.  Measured per CG iteration against
  // the marching kernels (B200, gpurun_out/probe_pipe_r01s.txt): 2048^2 x 1 chain 189 -> 167 us, 1024^2 x 4 188 -> 171,
  // 512^2 x 16 191 -> 169; 256^2 x 32 equal; 256^2 x 64 (4 sites per tile: the halo sites are a quarter of the
  // traffic into shared memory) 191 -> 196; 256^2 x 128 360 -> 389.  Tiles of 32 out of 64 chains (a 512-byte copy per
  // site) were 1.5x SLOWER: the copy engine of an SM retires a small copy every ~65 cycles.  So: batches of up to 16
  // chains (tiles at least 16 sites wide), blocks at least 16 rows tall (a ring of stages has a fill latency) with
  // >= 2.5 blocks per SM.  Larger batches (a multiple of 16 chains): tiles of 16 chains x 16 sites moved as 2-D boxes
  // through tensor maps, which cost what the contiguous tiles of a 16-chain batch cost (profiles/pipe_r02n_probe.txt, per
  // CG iteration): 256^2 x 64 191 -> 172 us, 128^2 x 2048 1353 -> 1259, 512^2 x 64 691 -> 653, 128^2 x 512 356 -> 325.
  // Everything else keeps the marching kernels.
  ctx->gp = g;
  ctx->pipe_ok = false;
  ctx->pipe_tiled = false;
  ctx->tmap_next = 0;
  for (int i = 0; i < TB_TMAP_CACHE; i++) ctx->tmap_ptr[i] = nullptr;
  const char *et = getenv("TB_PIPE_TEST");   // tests: stage every shape the kernels can handle
  const int max_chains = et ? 128 : 16, min_rows = et ? 4 : 16;
  // Batches of more than 16 chains: tiles of 16 chains x 16 sites whose rows are 2-D boxes of the [site][chain] arrays,
  // one tensor-map copy per array and row (dslash_pipe_kernel<TILED>).  A batch that is not a multiple of 16 chains ends
  // in a ragged tile: the tensor copy zero-fills the chains that do not exist, their threads are inactive.
  // TB_PIPE_TILED=0 keeps the marching kernels.
  const char *etl = getenv("TB_PIPE_TILED");
  const bool tiled_on = etl ? atoi(etl) != 0 : TB_PIPE_TILED_DEFAULT;
  const bool tiled = tiled_on && ctx->nranks == 1 && ctx->C > 16 && ctx->nx % 16 == 0 &&
                     (size_t)ctx->nt * ctx->nx < (1ull << 31) && tmap_encoder() != nullptr;
  const bool whole = ctx->nranks == 1 && ctx->C <= max_chains && (ctx->C & (ctx->C - 1)) == 0 &&
                     ctx->nx % (TB_MAX_BLOCK / ctx->C) == 0;
  if (tiled || whole) {
    TbGeom &p = ctx->gp;
    p.bc = tiled ? 16 : ctx->C;
    p.bx = TB_MAX_BLOCK / p.bc;
    p.bc_shift = 0;
    while ((1 << p.bc_shift) < p.bc) p.bc_shift++;
    p.nctiles = (ctx->C + p.bc - 1) / p.bc;   // tiled: the last tile may be ragged (its missing chains are zero-filled boxes)
    p.nxtiles = ctx->nx / p.bx;
    p.Cpad = p.nctiles * p.bc;                // <= the marching geometry's Cpad, which sizes the per-chain arrays
    const long min_blocks = et ? 1 : 370;
    const char *er = getenv("TB_PIPE_ROWS");   // development: rows per staged block
    for (int ttp = 64; ttp >= min_rows; ttp--) {
      if (er && ttp != atoi(er)) continue;
      if (ctx->nt % ttp != 0 || (long)p.nctiles * p.nxtiles * (ctx->nt / ttp) < min_blocks) continue;
      if ((size_t)p.nxtiles * (ctx->nt / ttp) * p.Cpad > (size_t)g.nxtiles * ctx->nt * g.Cpad) continue;   // partial sums buffer
      p.tt = ttp;
      p.nttiles = ctx->nt / ttp;
      p.nslots = p.nxtiles * p.nttiles;
      ctx->gp_tt0 = ttp;
      ctx->pipe_ok = true;
      ctx->pipe_tiled = tiled;
      break;
    }
  }
  tb_gauge_sharing(ctx, ctx->gauge_shared);
  return TB_OK;
}

static bool use_pipe(const tb_ctx *ctx) {
  return ctx->pipe_ok && !ctx->msite && ctx->tune_tt == 0 && getenv("TB_NO_PIPE") == nullptr;
}

// The two-launch iteration (direction update folded into the first staged pass) saves 16 of 240 B per site and a launch,
// but the separate xpay kernel finds r and the next pass finds p in the 126 MB L2 while a vector is not much larger than
// that.  Measured per CG iteration (B200, profiles/pipe_r02o_probe.txt), three -> two launches: vectors of 134 MB
// (2048^2 x 2, 1024^2 x 8, 128^2 x 512) 323 -> 333, 323 -> 330, 325 -> 343 us; 67 MB (2048^2) 166 -> 178; vectors of
// 268 MB (4096^2, 1024^2 x 16, 512^2 x 64, 256^2 x 256) 646 -> 621, 648 -> 618, 653 -> 631, 657 -> 631; 537 MB
// (128^2 x 2048) 1259 -> 1190.  So: from 200 MB per vector.  TB_PIPE_XPAY=0 / 1 overrides.
static bool use_pipe_xp(const tb_ctx *ctx) {
  if (!use_pipe(ctx) || ctx->nranks > 1) return false;
  const char *e = getenv("TB_PIPE_XPAY");
  return e ? atoi(e) != 0 : ctx->nsite * sizeof(double2) >= (200ull << 20);
}

// Which streaming kernels the next apply / CG iteration of a family-A field would use: 0 register-marching, 1 staged
// with whole-batch tiles (bulk copies), 2 staged with 16-chain tiles (tensor-map copies); and the staged tile shape.
int tb_stream_kernels(const tb_ctx *ctx, int *tile_chains, int *tile_sites, int *rows_per_block) {
  const bool on = use_pipe(ctx);
  const TbGeom &g = on ? ctx->gp : ctx->g;
  if (tile_chains) *tile_chains = g.bc;
  if (tile_sites) *tile_sites = g.bx;
  if (rows_per_block) *rows_per_block = g.tt;
  return on ? (ctx->pipe_tiled ? 2 : 1) : 0;
}

// in / out as in the kernel; x: the solution (FUSED) or the buffer the new direction goes to (XP); pcur: the direction
// the second pass reads (FUSED; default the context's p)
template <bool FUSED, bool CG = true, bool XP = false>
static int launch_pipe(tb_ctx *ctx, const double2 *in, double2 *out, double2 *x, bool dagger = false,
                       const double2 *pcur = nullptr) {
  const TbGeom &g = ctx->gp;
  if (!pcur) pcur = ctx->p;
  const bool shw = ctx->gauge_shared && ctx->Ws != nullptr;
  const double2 *w0 = shw ? ctx->Ws : ctx->W0, *w1 = shw ? ctx->Ws + ctx->V : ctx->W1;
  PipeMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (ctx->pipe_tiled) {
    TB_CHECK(tmap_for(ctx, in, &maps.m[0]));
    if (!shw) {
      TB_CHECK(tmap_for(ctx, ctx->W0, &maps.m[1]));
      TB_CHECK(tmap_for(ctx, ctx->W1, &maps.m[2]));
    }
    maps.m[3] = maps.m[4] = maps.m[5] = maps.m[0];
    if (FUSED) {
      TB_CHECK(tmap_for(ctx, pcur, &maps.m[3]));
      TB_CHECK(tmap_for(ctx, x, &maps.m[4]));
    }
    if (FUSED || XP) TB_CHECK(tmap_for(ctx, ctx->r, &maps.m[5]));
  }
  size_t smem;
  void (*kern)(const double2 *, double2 *, const double2 *, const double2 *, const double *, const double *, const double *,
               const double2 *, double2 *, double2 *, const TbGeom, const TbCgState, const TbSlab, const int, const PipeMaps);
  if (shw) {
    // the stages of a shared field are small (a row of links is bx values): a ring of 8 keeps as many bytes in flight in
    // the passes that stage one vector as a ring of 4 does with per-chain links
    constexpr int NS = (FUSED || XP) ? 4 : 8;
    smem = PipeLay<FUSED, XP, true>::smem_bytes(g.bc, NS);
    kern = ctx->pipe_tiled ? dslash_pipe_kernel<FUSED, NS, CG, true, XP, true> : dslash_pipe_kernel<FUSED, NS, CG, false, XP, true>;
  } else {
    using St = PipeLay<FUSED, XP, false>;
    int ns = PIPE_NS_MAX;
    while (ns > 2 && 2 * (St::smem_bytes(g.bc, ns) + 4096) > 227 * 1024) ns--;   // two blocks per SM at least
    if (ns < 3) ns = 3;
    smem = St::smem_bytes(g.bc, ns);
    if (ctx->pipe_tiled) {
      kern = ns == 4 ? dslash_pipe_kernel<FUSED, 4, CG, true, XP> : dslash_pipe_kernel<FUSED, 3, CG, true, XP>;
    } else if constexpr (!FUSED && !XP) {
      kern = ns == 4 ? dslash_pipe_plain_kernel<4, CG> : dslash_pipe_plain_kernel<3, CG>;
    } else {
      kern = ns == 4 ? dslash_pipe_kernel<FUSED, 4, CG, false, XP> : dslash_pipe_kernel<FUSED, 3, CG, false, XP>;
    }
  }
  TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid_of(g), TB_MAX_BLOCK, smem, ctx->stream>>>(in, out, w0, w1, ctx->d_mass, ctx->d_emu, ctx->d_emmu,
                                                         pcur, x, ctx->r, g, ctx->cg, ctx->slab, dagger ? 1 : 0, maps);
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

// Chain-slice variants: operate on chains [c0, c0+n) of the context on stream st (the host-buffer path
// pipelines sub-batches of chains on separate streams).
int tb_launch_links_slice(tb_ctx *ctx, const double2 *d_A_dev_layout, int c0, int n, cudaStream_t st) {
  ctx->msite = nullptr;  // complex U(1) links: family A, no occupation mask
  tb_gauge_sharing(ctx, false);
  const size_t total = ctx->V * (size_t)n;
  int blocks = (int)((total + 255) / 256);
  if (blocks > TB_NUM_SMS_B200 * 16) blocks = TB_NUM_SMS_B200 * 16;
  links_kernel<<<blocks, 256, 0, st>>>(d_A_dev_layout, ctx->W0, ctx->W1, ctx->nt, ctx->nx, ctx->C, c0, n,
                                       ctx->t_off, ctx->nt_global);
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

// canonical slice [n chains][V] (contiguous) -> device layout columns [c0, c0+n) of d_vec[V][C]
int tb_launch_pack_slice(tb_ctx *ctx, const double2 *d_canonical_slice, double2 *d_vec, int c0, int n,
                         cudaStream_t st) {
  if (ctx->C == 1) {
    if ((const void *)d_canonical_slice != (const void *)d_vec)
      TB_CUDA(cudaMemcpyAsync(d_vec, d_canonical_slice, ctx->nsite * sizeof(double2), cudaMemcpyDeviceToDevice, st));
    return TB_OK;
  }
  const int rows = n, cols = (int)ctx->V;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, st>>>(d_canonical_slice, d_vec + c0, rows, cols, ctx->V, (size_t)ctx->C);
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

int tb_launch_unpack_slice(tb_ctx *ctx, const double2 *d_vec, double2 *d_canonical_slice, int c0, int n,
                           cudaStream_t st) {
  if (ctx->C == 1) {
    if ((const void *)d_canonical_slice != (const void *)d_vec)
      TB_CUDA(cudaMemcpyAsync(d_canonical_slice, d_vec, ctx->nsite * sizeof(double2), cudaMemcpyDeviceToDevice, st));
    return TB_OK;
  }
  const int rows = (int)ctx->V, cols = n;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, st>>>(d_vec + c0, d_canonical_slice, rows, cols, (size_t)ctx->C, ctx->V);
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

// occupation field in canonical layout [chain][t][x] (device ints) -> masked links + site masses
int tb_launch_occupancy(tb_ctx *ctx, const int *d_field_canonical) {
  tb_gauge_sharing(ctx, false);
  cudaStream_t st = ctx->stream;
  const int *occ = d_field_canonical;
  if (ctx->C > 1) {
    dim3 grid(((int)ctx->V + 31) / 32, (ctx->C + 31) / 32), block(32, 8);
    transpose_int_kernel<<<grid, block, 0, st>>>(d_field_canonical, ctx->occ_dev, ctx->C, (int)ctx->V);
    occ = ctx->occ_dev;
    ctx->launches++;
  }
  int blocks = (int)((ctx->nsite + 255) / 256);
  if (blocks > TB_NUM_SMS_B200 * 16) blocks = TB_NUM_SMS_B200 * 16;
  occupancy_links_kernel<<<blocks, 256, 0, st>>>(occ, ctx->d_mass, ctx->W0, ctx->W1, ctx->msite_buf, ctx->nt,
                                                 ctx->nx, ctx->C, ctx->occ_bc);
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

// With a shared field the second pass keeps three blocks per SM instead of two and the first pass a ring of 8 stages:
// shorter blocks fill the waves better.  Measured per CG iteration, replicated -> shared (profiles/shared_r02x_probe.txt):
// 1024^2 x 16: 64 rows 621 -> 504, 32 rows 598 -> 482, 16 rows 633 -> 512; 512^2 x 32 (16-chain tiles): 64 rows 318 ->
// 313, 32 rows 334 -> 286, 16 rows 340 -> 282; 256^2 x 64: 32 rows 169 -> 169, 16 rows 177 -> 154.
void tb_gauge_sharing(tb_ctx *ctx, bool shared) {
  ctx->gauge_shared = shared;
  if (!ctx->pipe_ok) return;
  TbGeom &p = ctx->gp;
  int tt = ctx->gp_tt0;
  const int want = ctx->pipe_tiled ? 16 : 32;
  if (shared && want < tt && ctx->nt % want == 0 &&
      (size_t)p.nxtiles * (ctx->nt / want) * p.Cpad <= (size_t)ctx->g.nxtiles * ctx->nt * ctx->g.Cpad)
    tt = want;
  p.tt = tt;
  p.nttiles = ctx->nt / tt;
  p.nslots = p.nxtiles * p.nttiles;
}

int tb_launch_links_shared(tb_ctx *ctx, const double2 *d_A_one_field) {
  if (ctx->nranks > 1) { tb_set_error("a shared gauge field is not available in slab mode"); return TB_EINVAL; }
  if (!ctx->Ws) TB_CUDA(cudaMalloc((void **)&ctx->Ws, 2 * ctx->V * sizeof(double2)));
  const dim3 block(32, 8);
  int blocks = (int)((ctx->V + 7) / 8);
  if (blocks > TB_NUM_SMS_B200 * 16) blocks = TB_NUM_SMS_B200 * 16;
  links_shared_kernel<<<blocks, block, 0, ctx->stream>>>(d_A_one_field, ctx->Ws, ctx->W0, ctx->W1, ctx->Adev, ctx->nt, ctx->nx,
                                                        ctx->C);
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  ctx->msite = nullptr;
  tb_gauge_sharing(ctx, true);
  return TB_OK;
}

int tb_launch_links(tb_ctx *ctx, const double *d_A_dev_layout) {
  return tb_launch_links_slice(ctx, (const double2 *)d_A_dev_layout, 0, ctx->C, ctx->stream);
}

int tb_launch_pack(tb_ctx *ctx, const double *d_canonical, double2 *d_vec) {
  return tb_launch_pack_slice(ctx, (const double2 *)d_canonical, d_vec, 0, ctx->C, ctx->stream);
}

int tb_launch_unpack(tb_ctx *ctx, const double2 *d_vec, double *d_canonical) {
  return tb_launch_unpack_slice(ctx, d_vec, (double2 *)d_canonical, 0, ctx->C, ctx->stream);
}

// Generic dslash launch.  masked/dot select the CG variants; in slab mode the halo pointers and the flag protocol
// (wait_ready, wait_done, sig0, sig1) come from the caller.
struct DslashArgs {
  const double2 *in, *in_prev, *in_next;
  double2 *out;
  const double2 *aux;
  bool dagger, dot, masked;
  int wait_ready, wait_done, sig0, sig1;
};

template <bool SLAB>
static int launch_dslash_t(tb_ctx *ctx, const DslashArgs &a) {
  const TbGeom &g = ctx->g;
  const dim3 grid = grid_of(g);
  const int block = g.bc * g.bx;
  const double2 *w0p = SLAB ? ctx->slab.W0_prev : ctx->W0;
#define ARGS a.in, a.in_prev, a.in_next, a.out, ctx->W0, w0p, ctx->W1, ctx->d_mass, ctx->msite, ctx->d_emu, ctx->d_emmu, a.aux, g, \
             ctx->cg, ctx->slab, a.wait_ready, a.wait_done, a.sig0, a.sig1
#define L(DAG, DOT, MASKED) \
  TB_DISPATCH_TT(g.tt, (dslash_kernel<TT, DAG, DOT, MASKED, SLAB><<<grid, block, 0, ctx->stream>>>(ARGS)))
  if (a.dot) {
    if (a.masked) { if (a.dagger) { L(true, true, true) } else { L(false, true, true) } }
    else { if (a.dagger) { L(true, true, false) } else { L(false, true, false) } }
  } else {
    if (a.masked) { if (a.dagger) { L(true, false, true) } else { L(false, false, true) } }
    else { if (a.dagger) { L(true, false, false) } else { L(false, false, false) } }
  }
#undef L
#undef ARGS
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

int tb_launch_dslash(tb_ctx *ctx, bool dagger, const double2 *in, double2 *out, bool masked) {
  if (!masked && use_pipe(ctx)) return launch_pipe<false, false>(ctx, in, out, nullptr, dagger);
  DslashArgs a = {in, in, in, out, nullptr, dagger, false, masked, 0, -1, 0, -1};
  return launch_dslash_t<false>(ctx, a);
}

int tb_launch_dot(tb_ctx *ctx, const double2 *a, const double2 *b, double *d_out) {
  const TbGeom &g = ctx->g;
  TbCgState s = ctx->cg;
  s.dot = d_out;
  TB_DISPATCH_TT(g.tt, (dot_kernel<TT><<<grid_of(g), g.bc * g.bx, 0, ctx->stream>>>(a, b, g, s, ctx->slab)))
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

template <bool SLAB>
static int cg_iteration(tb_ctx *ctx, double2 *x) {
  const TbGeom &g = ctx->g;
  const dim3 grid = grid_of(g);
  const int block = g.bc * g.bx;
  const TbSlab &sl = ctx->slab;
  cudaStream_t st = ctx->stream;
  const bool dag = tb_conj_is_dagger(ctx);
  // Mp = M p (hmc.c:366).  slab: needs generation E of the neighbours' p, overwrites Mp (neighbours must have
  // finished reading generation E-1), then publishes Mp and releases p.
  DslashArgs k1 = {ctx->p, SLAB ? sl.p_prev : ctx->p, SLAB ? sl.p_next : ctx->p, ctx->Mp, nullptr, false, false, true,
                   TB_FLAG_PREADY, TB_FLAG_MPDONE, TB_FLAG_MPREADY, TB_FLAG_PDONE};
  TB_CHECK(launch_dslash_t<SLAB>(ctx, k1));
  // q = M~ Mp with the fused Re<p,q> (hmc.c:367-371)
  DslashArgs k2 = {ctx->Mp, SLAB ? sl.mp_prev : ctx->Mp, SLAB ? sl.mp_next : ctx->Mp, ctx->q, ctx->p, dag, true, true,
                   TB_FLAG_MPREADY, -1, TB_FLAG_MPDONE, -1};
  TB_CHECK(launch_dslash_t<SLAB>(ctx, k2));
  if (SLAB) {
    slab_scalars_kernel<FIN_PQ, TB_RED_PQ><<<1, TB_MAX_BLOCK, 0, st>>>(g, ctx->cg, sl);
    ctx->launches++;
  }
  TB_DISPATCH_TT(g.tt, (axpy_norm_kernel<TT, SLAB><<<grid, block, 0, st>>>(x, ctx->r, ctx->p, ctx->q, g, ctx->cg, sl)))
  ctx->launches++;
  if (SLAB) {
    slab_scalars_kernel<FIN_RR, TB_RED_RR><<<1, TB_MAX_BLOCK, 0, st>>>(g, ctx->cg, sl);
    ctx->launches++;
  }
  TB_CUDA(cudaGetLastError());
  TB_DISPATCH_TT(g.tt, (xpay_kernel<TT, SLAB><<<grid, block, 0, st>>>(ctx->p, ctx->r, g, ctx->cg, sl)))
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

// Fused ADJOINT iteration (single GPU): dslash + |Mp|^2 -> alpha ; dslash^dagger fused with the x/r update and
// ||r||^2 -> beta ; xpay.
static int cg_iteration_fused(tb_ctx *ctx, double2 *x) {
  const TbGeom &g = ctx->g;
  const dim3 grid = grid_of(g);
  const int block = g.bc * g.bx;
  cudaStream_t st = ctx->stream;
  if (use_pipe_xp(ctx)) {
    // two launches: the direction update of the previous iteration is part of the first pass, which reads the old
    // direction from one buffer and leaves the new one in the other (p and q: the fused iteration never stores q)
    double2 *pold = (ctx->xp_parity & 1) ? ctx->q : ctx->p, *pnew = (ctx->xp_parity & 1) ? ctx->p : ctx->q;
    ctx->xp_parity ^= 1;
    TB_CHECK((launch_pipe<false, true, true>(ctx, pold, ctx->Mp, pnew)));
    TB_CHECK((launch_pipe<true>(ctx, ctx->Mp, nullptr, x, false, pnew)));
    return TB_OK;
  }
  if (use_pipe(ctx)) {   // rows staged through shared memory by bulk asynchronous copies
    TB_CHECK(launch_pipe<false>(ctx, ctx->p, ctx->Mp, nullptr));
    TB_CHECK(launch_pipe<true>(ctx, ctx->Mp, nullptr, x));
  } else {
    DslashArgs k1 = {ctx->p, ctx->p, ctx->p, ctx->Mp, nullptr, false, true, true, 0, -1, 0, -1};
    TB_CHECK(launch_dslash_t<false>(ctx, k1));
    TB_DISPATCH_TT(g.tt, (dslash_axpy_norm_kernel<TT, false><<<grid, block, 0, st>>>(ctx->Mp, ctx->Mp, ctx->Mp, ctx->W0,
        ctx->W0, ctx->W1, ctx->d_mass, ctx->msite, ctx->d_emu, ctx->d_emmu, ctx->p, x, ctx->r, g, ctx->cg, ctx->slab)))
    ctx->launches++;
  }
  TB_DISPATCH_TT(g.tt, (xpay_kernel<TT, false><<<grid, block, 0, st>>>(ctx->p, ctx->r, g, ctx->cg, ctx->slab)))
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

// Fused ADJOINT iteration in slab mode: 240 B/site; the all-reduced scalars are evaluated by the one-block
// slab_scalars_kernel between the passes.  (Letting every block of the consumer kernel poll the peers' flags was
// measured slower: thousands of pollers on one L2 line delay the NVLink store they are waiting for.)
static int cg_iteration_fused_slab(tb_ctx *ctx, double2 *x) {
  const TbGeom &g = ctx->g;
  const dim3 grid = grid_of(g);
  const int block = g.bc * g.bx;
  const TbSlab &sl = ctx->slab;
  cudaStream_t st = ctx->stream;
  DslashArgs k1 = {ctx->p, sl.p_prev, sl.p_next, ctx->Mp, nullptr, false, true, true,
                   TB_FLAG_PREADY, TB_FLAG_MPDONE, TB_FLAG_MPREADY, TB_FLAG_PDONE};
  TB_CHECK(launch_dslash_t<true>(ctx, k1));
  slab_scalars_kernel<FIN_PQ, TB_RED_PQ><<<1, TB_MAX_BLOCK, 0, st>>>(g, ctx->cg, sl);
  TB_DISPATCH_TT(g.tt, (dslash_axpy_norm_kernel<TT, true><<<grid, block, 0, st>>>(ctx->Mp, sl.mp_prev, sl.mp_next, ctx->W0,
      sl.W0_prev, ctx->W1, ctx->d_mass, ctx->msite, ctx->d_emu, ctx->d_emmu, ctx->p, x, ctx->r, g, ctx->cg, sl)))
  slab_scalars_kernel<FIN_RR, TB_RED_RR><<<1, TB_MAX_BLOCK, 0, st>>>(g, ctx->cg, sl);
  TB_DISPATCH_TT(g.tt, (xpay_kernel<TT, true><<<grid, block, 0, st>>>(ctx->p, ctx->r, g, ctx->cg, sl)))
  ctx->launches += 4;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

// slab mode: out = Op in on the distributed lattice (collective: every rank calls it with its slab)
int tb_slab_apply(tb_ctx *ctx, int op, const double2 *in, double2 *out) {
  const TbGeom &g = ctx->g;
  const TbSlab &sl = ctx->slab;
  cudaStream_t st = ctx->stream;
  const bool dag = tb_conj_is_dagger(ctx);
  slab_bump_kernel<<<1, 1, 0, st>>>(sl);
  TB_DISPATCH_TT(g.tt, (slab_stage_kernel<TT><<<grid_of(g), g.bc * g.bx, 0, st>>>(in, ctx->p, g, sl)))
  ctx->launches += 2;
  TB_CUDA(cudaGetLastError());
  if (op == TB_OP_MDM) {
    DslashArgs k1 = {ctx->p, sl.p_prev, sl.p_next, ctx->Mp, nullptr, false, false, false,
                     TB_FLAG_PREADY, TB_FLAG_MPDONE, TB_FLAG_MPREADY, TB_FLAG_PDONE};
    TB_CHECK(launch_dslash_t<true>(ctx, k1));
    DslashArgs k2 = {ctx->Mp, sl.mp_prev, sl.mp_next, out, nullptr, dag, false, false,
                     TB_FLAG_MPREADY, -1, TB_FLAG_MPDONE, -1};
    return launch_dslash_t<true>(ctx, k2);
  }
  const bool d = (op == TB_OP_MDAG) || (op == TB_OP_MCONJ && dag);
  // a single apply reads only p: release both exchange vectors for the next epoch
  DslashArgs k = {ctx->p, sl.p_prev, sl.p_next, out, nullptr, d, false, false,
                  TB_FLAG_PREADY, -1, TB_FLAG_PDONE, TB_FLAG_MPDONE};
  return launch_dslash_t<true>(ctx, k);
}

// slab mode: the whole solve as one persistent cooperative launch per GPU (slab_cg_persistent_kernel)
// Measured on 2 GPUs (us per CG iteration, multi-kernel -> one launch): 2048^2 142.6 -> 120.5; 1024^2 (the per-GPU size
// of 2048^2 on 8 GPUs) 78.4 -> 33.5; 512^2 68.0 -> 29.0.  Slabs of more than 4M sites are HBM-bound, where the
// multi-kernel path (L1-cached neighbour loads, four blocks per SM) is at 0.85 of the HBM peak: it keeps those.
// Measured (profiles/slab_r02f_8gpu.txt): with 8M sites per GPU (4096^2 on 2 GPUs) the one-launch solve still beats the
// multi-kernel form, 358 vs 435 us per iteration (phase B walks its chunks downwards, so every phase starts on what the
// previous one left in L2); the 32-bit indices of the kernel hold up to 2^27 elements per field.
static size_t persist_max_sites() {
  if (const char *e = getenv("TB_PERSIST_MAX_SITES")) return (size_t)atoll(e);
  return (size_t)16 << 20;
}

static bool use_persistent_slab(const tb_ctx *ctx) {
  return ctx->nranks > 1 && ctx->g.nctiles == 1 && !ctx->msite && tb_conj_is_dagger(ctx) && ctx->cg_variant != 4 &&
         ctx->g.bx >= ctx->nranks &&   // block 0 exchanges with thread (rank, chain): needs nranks * bc threads
         ctx->nsite <= persist_max_sites() && getenv("TB_NO_PERSIST") == nullptr;
}

// mode 0: block 0 runs the all-reduce (slab_cg_persistent_kernel); 1: the last-arriving block exchanges with the peers
// and publishes the totals locally; 2: every block polls the peers' slots (slab_cg_onelaunch_kernel, see slab_allreduce)
static int launch_persistent_slab(tb_ctx *ctx, const double2 *b, int mode, int *nblocks_out) {
  constexpr int TT = 8;
  TbGeom g = ctx->g;   // same column strips; mode 0: tiles of TT rows, else balanced row segments (slab_seg)
  g.tt = TT;
  g.nttiles = (ctx->nt + TT - 1) / TT;
  g.nslots = g.nxtiles * g.nttiles;
  TbSlab &sl = ctx->slab;
  SlabCgArgs a = {b, ctx->xw, ctx->r, ctx->p, ctx->p1, ctx->Mp, sl.p_prev, sl.p_next, sl.p1_prev, sl.p1_next,
                  sl.r_prev, sl.r_next, sl.mp_prev, sl.mp_next,
                  ctx->W0, sl.W0_prev, ctx->W1, ctx->d_mass, ctx->d_emu, ctx->d_emmu,
                  (getenv("TB_SLAB_SYSFENCE") && atoi(getenv("TB_SLAB_SYSFENCE"))) ? 1 : 0};
  const void *kern = mode == 0 ? (const void *)slab_cg_persistent_kernel<TT> : (const void *)slab_cg_onelaunch_kernel<TT>;
  int per_sm = 0, nsm = TB_NUM_SMS_B200, coop = 0;
  TB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
  TB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device));
  TB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, g.bc * g.bx, 0));
  if (!coop || per_sm < 1) { *nblocks_out = 0; return TB_OK; }
  const int capacity = per_sm * nsm;
  int nseg = 1, nblocks;
  if (mode == 0) {
    nblocks = g.nxtiles * g.nttiles;
  } else {
    nseg = capacity / g.nxtiles;
    if (nseg > ctx->nt) nseg = ctx->nt;
    if (nseg < 1) nseg = 1;
    nblocks = g.nxtiles * nseg;
  }
  if (nblocks > capacity) nblocks = capacity;
  if ((size_t)nblocks * g.Cpad > (size_t)ctx->g.nxtiles * ctx->nt * ctx->g.Cpad) { *nblocks_out = 0; return TB_OK; }   // partial[]
  TB_CUDA(cudaMemsetAsync(sl.gbar, 0, sizeof(unsigned long long), ctx->stream));
  sl.nrep = TB_SLAB_NREP_MAX;   // the same on every rank: a rank polls the replicas its peers write
  if (const char *e = getenv("TB_SLAB_NREP")) {
    const int n = atoi(e);
    sl.nrep = n < 1 ? 1 : (n > TB_SLAB_NREP_MAX ? TB_SLAB_NREP_MAX : n);
  }
  if (sl.nrep > g.bx) sl.nrep = g.bx;   // a thread per (replica, chain) publishes the totals
  // optional timeline of the first TB_SLAB_TL_ITERS iterations (TB_SLAB_TIMELINE=<file prefix>): per synchronisation
  // point {first block starts the phase, last block finishes it, partial stored to the peers, last block holds the total}
  const char *tlpath = mode != 0 ? getenv("TB_SLAB_TIMELINE") : nullptr;
  const size_t tlwords = (size_t)TB_SLAB_TL_ITERS * TB_SLAB_TL_WORDS;
  unsigned long long *tl_host = nullptr;
  if (tlpath) {
    if (!sl.timeline) TB_CUDA(cudaMalloc((void **)&sl.timeline, tlwords * sizeof(unsigned long long)));
    tl_host = (unsigned long long *)malloc(tlwords * sizeof(unsigned long long));
    for (size_t i = 0; i < tlwords; i++) tl_host[i] = (i % 4 == 0) ? ~0ULL : 0ULL;
    TB_CUDA(cudaMemcpyAsync(sl.timeline, tl_host, tlwords * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  TbGeom gg = g;
  TbCgState ss = ctx->cg;
  TbSlab sls = sl;
  if (!tlpath) sls.timeline = nullptr;
  int nseg_arg = nseg, mode_arg = mode;
  void *args[] = {&a, &gg, &ss, &sls, &nseg_arg, &mode_arg};   // the block-0 kernel takes the first four
  TB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(nblocks), dim3(g.bc * g.bx), args, 0, ctx->stream));
  ctx->launches++;
  *nblocks_out = nblocks;
  if (tlpath) {
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    TB_CUDA(cudaMemcpy(tl_host, sl.timeline, tlwords * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    char name[512];
    snprintf(name, sizeof(name), "%s.rank%d.txt", tlpath, ctx->rank);
    if (FILE *f = fopen(name, "w")) {
      fprintf(f, "# %d blocks, %d row segments per column strip, sync mode %d\n", nblocks, nseg, mode);
      fprintf(f, "# globaltimer ns since the first stamp of this rank; per CG iteration and phase (A: Mp = M p, B: M^dagger + "
                 "update): start = first block enters, end = last block has its partial, stored = partial sent to the peers, "
                 "total = last block holds the all-reduced sum (every 16th block records)\n"
                 "# iter A_start A_end A_stored A_total B_start B_end B_stored B_total\n");
      unsigned long long t0 = tl_host[0];
      for (int k = 0; k < TB_SLAB_TL_ITERS; k++) {
        if (tl_host[(size_t)k * 8] == ~0ULL) break;
        fprintf(f, "%d", k + 1);
        for (int w = 0; w < 8; w++) fprintf(f, " %lld", (long long)(tl_host[(size_t)k * 8 + w] - t0));
        fprintf(f, "\n");
      }
      fclose(f);
    }
    free(tl_host);
  }
  return TB_OK;
}

// Which form (measured, us per CG iteration, profiles/slab_r02*.txt): with device-scope fences around the peer exchange
// the hybrid form wins at every size on 2 GPUs (1024^2: 29.7 vs 32.2 for block 0's form, 2048^2: 100.8 vs 116.8,
// 4096^2: 352 vs 367); TB_SLAB_SYNC=0 / 2 select the other forms.
static int launch_persistent_slab_auto(tb_ctx *ctx, const double2 *b, int *nblocks_out) {
  int mode = 1;
  if (const char *e = getenv("TB_SLAB_SYNC")) { const int v = atoi(e); if (v >= 0 && v <= 2) mode = v; }
  return launch_persistent_slab(ctx, b, mode, nblocks_out);
}

// Streaming CG driver: the whole solve stays on the device; the host only polls the number of chains
// still iterating, one graph launch (tune_chunk iterations) behind the device.
int tb_run_cg_stream(tb_ctx *ctx, const double2 *b, double2 *x) {
  const TbGeom &g = ctx->g;
  const dim3 grid = grid_of(g);
  const int block = g.bc * g.bx;
  cudaStream_t st = ctx->stream;
  const bool slab = ctx->nranks > 1;
  // the fused 3-kernel iteration needs M~ = M^dagger; TB_CG_VARIANT=4 (or tune) forces the 4-kernel form
  const bool fused = tb_conj_is_dagger(ctx) && ctx->cg_variant != 4;
  cg_reset_kernel<<<(g.Cpad + 255) / 256, 256, 0, st>>>(g, ctx->cg);
  ctx->launches++;
  if (use_persistent_slab(ctx)) {
    int nblocks = 0;
    TB_CHECK(launch_persistent_slab_auto(ctx, b, &nblocks));
    if (nblocks > 0) {
      TB_CUDA(cudaMemcpyAsync(x, ctx->xw, ctx->nsite * sizeof(double2), cudaMemcpyDeviceToDevice, st));
      return TB_OK;
    }
  }
  if (slab) {
    slab_bump_kernel<<<1, 1, 0, st>>>(ctx->slab);
    TB_DISPATCH_TT(g.tt, (cg_init_kernel<TT, true><<<grid, block, 0, st>>>(b, ctx->xw, ctx->r, ctx->p, g, ctx->cg, ctx->slab)))
    slab_scalars_kernel<FIN_INIT, TB_RED_INIT><<<1, TB_MAX_BLOCK, 0, st>>>(g, ctx->cg, ctx->slab);
    ctx->launches += 3;
  } else {
    TB_DISPATCH_TT(g.tt, (cg_init_kernel<TT, false><<<grid, block, 0, st>>>(b, ctx->xw, ctx->r, ctx->p, g, ctx->cg, ctx->slab)))
    ctx->launches++;
  }
  TB_CUDA(cudaGetLastError());

  int chunk = ctx->tune_chunk > 0 ? ctx->tune_chunk : 16;
  // the two-launch iteration alternates between two direction buffers: a graph must hold an even number of iterations
  const bool xp = fused && !slab && use_pipe_xp(ctx);
  if (xp) chunk += chunk & 1;
  ctx->xp_parity = 0;
  const bool use_graph = getenv("TB_NO_GRAPH") == nullptr;
  // the graph holds its kernel arguments by value: everything a later call may change is part of the key (the per-site
  // mass pointer of family B is one of them; CG tolerances and tile shapes invalidate the graph where they are set)
  const int graph_key = chunk * 32 + (fused ? 1 : 0) + (use_pipe(ctx) ? 2 : 0) + (ctx->msite ? 4 : 0) + (xp ? 8 : 0) +
                        (ctx->gauge_shared ? 16 : 0);
  if (use_graph && (ctx->cg_graph == nullptr || ctx->cg_graph_chunk != graph_key)) {
    if (ctx->cg_graph) { cudaGraphExecDestroy(ctx->cg_graph); ctx->cg_graph = nullptr; }
    cudaStream_t cap;
    TB_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    cudaStream_t saved = ctx->stream;
    const long long saved_launches = ctx->launches;
    ctx->stream = cap;
    TB_CUDA(cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed));
    int rc = TB_OK;
    for (int i = 0; i < chunk && rc == TB_OK; i++)
      rc = slab ? (fused ? cg_iteration_fused_slab(ctx, ctx->xw) : cg_iteration<true>(ctx, ctx->xw))
                : (fused ? cg_iteration_fused(ctx, ctx->xw) : cg_iteration<false>(ctx, ctx->xw));
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(cap, &graph);
    ctx->stream = saved;
    ctx->launches = saved_launches;
    if (rc != TB_OK) { cudaStreamDestroy(cap); return rc; }
    TB_CUDA(e);
    TB_CUDA(cudaGraphInstantiate(&ctx->cg_graph, graph, 0));
    cudaGraphDestroy(graph);
    cudaStreamDestroy(cap);
    ctx->cg_graph_chunk = graph_key;
  }

  const long max_chunks = ((long)ctx->cg.max_iter + chunk - 1) / chunk + 1;
  for (long i = 0; i < max_chunks; i++) {
    if (use_graph) {
      TB_CUDA(cudaGraphLaunch(ctx->cg_graph, st));
      ctx->launches += (slab ? (fused ? 5LL : 6LL) : (fused ? (xp ? 2LL : 3LL) : 4LL)) * chunk;
    } else {
      for (int k = 0; k < chunk; k++)
        TB_CHECK(slab ? (fused ? cg_iteration_fused_slab(ctx, ctx->xw) : cg_iteration<true>(ctx, ctx->xw))
                      : (fused ? cg_iteration_fused(ctx, ctx->xw) : cg_iteration<false>(ctx, ctx->xw)));
    }
    const int slot = (int)(i & 1);
    TB_CUDA(cudaMemcpyAsync(&ctx->h_flag[slot], ctx->cg.n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaEventRecord(ctx->ev_flag[slot], st));
    if (i > 0) {
      TB_CUDA(cudaEventSynchronize(ctx->ev_flag[slot ^ 1]));
      if (ctx->h_flag[slot ^ 1] == 0) break;
    }
  }
  TB_CUDA(cudaMemcpyAsync(x, ctx->xw, ctx->nsite * sizeof(double2), cudaMemcpyDeviceToDevice, st));
  return TB_OK;
}
