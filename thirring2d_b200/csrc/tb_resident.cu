// tb_resident.cu — on-chip resident batched CG for small lattices (16^2, 32^2, 64^2), sm_100a, FP64.
//
// One CTA owns one Markov chain for the WHOLE solve (fmdm_invert_cg, hmc.c:341-404): because chains are
// independent, every CG reduction is a block-level reduction and no grid-wide synchronisation or kernel
// boundary is needed.  The CG state never leaves the SM between iterations:
//
//   registers      r, p (persistent) and Mp, q (transient) for the TT x TX site tile of the thread
//   shared memory  one exchange field F (p, then Mp: what the stencil neighbours read) 16 B/site
//                  the two link fields W0, W1                                           32 B/site
//   tensor memory  x, thread-private columns (tcgen05.ld/st 32x32b): x += alpha p costs no LSU wavefronts; the
//                  fallback (TB_RESIDENT_X_TMEM=0) is a 128-bit L2 read-modify-write of a chain-major workspace
//
// 64^2: 48 B/site * 4096 = 192 KB of the 227 KB shared memory, one CTA of 256 threads (2 x 8 sites each, 255
// registers) per SM, 148 chains in flight per B200.  HBM is touched only to load b and the links once and to
// store x once per solve.  Larger lattices (128^2, 256^2) use one thread-block CLUSTER per chain: tb_cluster.cu.
//
// A thread owns a TT x TX tile of sites, so most stencil neighbours are its own registers; the tile's halo comes
// from F.  Reductions are fixed-shape (shuffle tree, then warp partials summed in warp order) => run-to-run
// deterministic.
#include <cstdint>

#include "tb_common.cuh"

namespace {

#include "tb_onchip.cuh"

template <int NWARPS>
__device__ __forceinline__ double block_sum(double v, double *scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < NWARPS; w++) s += scratch[w];
  return s;
}

// Same, with the warp partials added as a balanced tree (three dependent additions instead of NWARPS after the
// barrier); NWARPS must be 8.  A different but equally fixed summation order.
__device__ __forceinline__ double block_sum8(double v, double *scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  const double2 *s2 = reinterpret_cast<const double2 *>(scratch);
  const double2 a = s2[0], b = s2[1], c = s2[2], d = s2[3];
  return ((a.x + a.y) + (b.x + b.y)) + ((c.x + c.y) + (d.x + d.y));
}

// block_sum8 split at its barrier, so that independent work can sit in front of it
__device__ __forceinline__ void block_sum8_post(double v, double *scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
}
__device__ __forceinline__ double block_sum8_total(const double *scratch) {
  __syncthreads();
  const double2 *s2 = reinterpret_cast<const double2 *>(scratch);
  const double2 a = s2[0], b = s2[1], c = s2[2], d = s2[3];
  return ((a.x + a.y) + (b.x + b.y)) + ((c.x + c.y) + (d.x + d.y));
}

// Shared-memory site index: the TX x-sites of a thread's tile live in TX separate sub-planes of a row, so
// that for a fixed tile column j consecutive threads (x-groups) touch consecutive double2 (conflict-free
// LDS.128 / STS.128 whatever TX is).
template <int NX, int TX>
__device__ __forceinline__ int sidx(int t, int x) {
  return t * NX + (x % TX) * (NX / TX) + x / TX;
}

// out = m f +- hops on the thread's TT x TX tile.  f: own tile (registers); F: the same field in shared memory
// (only the tile's halo is read from it).  DAG: M^dagger instead of M.  HAS_MU: the t-links carry e^{+-mu}
// (af on the +t hop, ab on the -t hop); otherwise both are 1 and the scaling is skipped.
// Shared-memory traffic per site and apply: 16 B * (2 + 3/TX + 3/TT)  (halo of f, W0 rows, W1 columns).
template <int NT, int NX, int TX, int TT, bool DAG, bool HAS_MU>
__device__ __forceinline__ void tile_apply(const double2 (&f)[TT][TX], double2 (&out)[TT][TX], const double2 *F,
                                           const double2 *W0s, const double2 *W1s, int t0, int g,
                                           double m, double af, double ab) {
  constexpr int SF = DAG ? -1 : 1;  // sign of the forward hops (hmc.c:144,166 / adjoint)
  constexpr int SB = -SF;           // sign of the backward hops (hmc.c:158,179 / adjoint)
  constexpr int NG = NX / TX;       // x-groups per row == stride between sub-planes
  const int gl = (g + NG - 1) % NG, gr = (g + 1) % NG;
  const int tm = (t0 + NT - 1) % NT, te = (t0 + TT) % NT;
  double2 w0m[TX];                  // W0(t-1, x0+j): slides down the tile
#pragma unroll
  for (int j = 0; j < TX; j++) w0m[j] = W0s[tm * NX + j * NG + g];
#pragma unroll
  for (int i = 0; i < TT; i++) {
    const int row = (t0 + i) * NX;
    const double2 fL = F[row + (TX - 1) * NG + gl];    // f(t, x0-1)
    const double2 fR = F[row + gr];                    // f(t, x0+TX)
    double2 w1m = W1s[row + (TX - 1) * NG + gl];       // W1(t, x0-1): slides along the row
#pragma unroll
    for (int j = 0; j < TX; j++) {
      const double2 w0c = W0s[row + j * NG + g];
      const double2 w1c = W1s[row + j * NG + g];
      const double2 up = (i == TT - 1) ? F[te * NX + j * NG + g] : f[(i + 1) % TT][j];
      const double2 dn = (i == 0) ? F[tm * NX + j * NG + g] : f[(i + TT - 1) % TT][j];
      const double2 rt = (j == TX - 1) ? fR : f[i][(j + 1) % TX];
      const double2 lf = (j == 0) ? fL : f[i][(j + TX - 1) % TX];
      // m f(n) + af W0(n) f(n+t) - ab conj(W0(n-t)) f(n-t) + W1(n) f(n+x) - conj(W1(n-x)) f(n-x)  (hmc.c:137-180)
      double2 o = make_double2(m * f[i][j].x, m * f[i][j].y);
      if (HAS_MU) {
        hop_acc<SF>(o, make_double2(w0c.x * af, w0c.y * af), up);
        hopc_acc<SB>(o, make_double2(w0m[j].x * ab, w0m[j].y * ab), dn);
      } else {
        hop_acc<SF>(o, w0c, up);
        hopc_acc<SB>(o, w0m[j], dn);
      }
      hop_acc<SF>(o, w1c, rt);
      hopc_acc<SB>(o, w1m, lf);
      out[i][j] = o;
      w0m[j] = w0c;
      w1m = w1c;
    }
  }
}

template <int NT, int NX, int TX, int TT>
struct ResidentCfg {
  static constexpr int V = NT * NX;
  static constexpr int NTHREADS = (NT / TT) * (NX / TX);
  static constexpr int NWARPS = NTHREADS / 32;
  static constexpr int REGS = (TX * TT >= 16) ? 255 : 128;
  static constexpr int MINBLOCKS_RAW = 65536 / (NTHREADS * (REGS + 1));
  static constexpr int MINBLOCKS = MINBLOCKS_RAW < 1 ? 1 : (MINBLOCKS_RAW > 16 ? 16 : MINBLOCKS_RAW);
  static constexpr size_t SMEM = (size_t)V * 48 + 64 * sizeof(double);
  // TMEM: 4 words (one double2) per site of the thread's tile; warps of the same lane quarter stack in columns
  static constexpr int TMEM_WORDS = TX * TT * 4;
  static constexpr int TMEM_NEED = TMEM_WORDS * ((NWARPS + 3) / 4);
  static constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : (TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512)));
  static constexpr bool TMEM_OK = (TMEM_WORDS % 16 == 0) && TMEM_NEED <= 512;
};

template <int NT, int NX, int TX, int TT, bool DAG, bool HAS_MU, bool XT>
__global__ void __launch_bounds__(ResidentCfg<NT, NX, TX, TT>::NTHREADS, ResidentCfg<NT, NX, TX, TT>::MINBLOCKS)
resident_cg_kernel(const double2 *__restrict__ bsrc, double2 *__restrict__ xout,
                   const double2 *__restrict__ W0g, const double2 *__restrict__ W1g,
                   const double *__restrict__ mass, const double *__restrict__ emu,
                   const double *__restrict__ emmu, double2 *__restrict__ xw, const TbCgState s, const int C,
                   const int c_first) {
  using Cfg = ResidentCfg<NT, NX, TX, TT>;
  constexpr int V = Cfg::V, NTHREADS = Cfg::NTHREADS, NWARPS = Cfg::NWARPS, NG = NX / TX;
  static_assert(NTHREADS % 32 == 0, "a CTA must be whole warps");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *F = reinterpret_cast<double2 *>(smem_raw);
  double2 *W0s = F + V;
  double2 *W1s = W0s + V;
  double *scrA = reinterpret_cast<double *>(W1s + V);
  double *scrB = scrA + 32;
  __shared__ uint32_t tmem_base_s;
  uint32_t xaddr = 0;   // this thread's x columns in tensor memory (XT)
  if (XT) {
    if (threadIdx.x < 32) {   // one fully active warp allocates (and later frees) the columns
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst), "r"(Cfg::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t warp = threadIdx.x >> 5;
    xaddr = tmem_base_s + (((warp & 3u) * 32u) << 16) + (warp >> 2) * (uint32_t)Cfg::TMEM_WORDS;
  }

  const int c = c_first + blockIdx.x;
  const int tid = threadIdx.x;
  const int g = tid % NG;            // x-group: the tile covers x in [g*TX, g*TX+TX)
  const int t0 = (tid / NG) * TT;    // and t in [t0, t0+TT)
  const double m = mass[c];
  const double e_p = emu[c], e_m = emmu[c];

  // links: device layout [site][chain] -> shared memory
  for (int k = tid; k < V; k += NTHREADS) {
    const int t = k / NX, x = k % NX;
    W0s[sidx<NX, TX>(t, x)] = W0g[(size_t)k * C + c];
    W1s[sidx<NX, TX>(t, x)] = W1g[(size_t)k * C + c];
  }
  double2 r[TT][TX], p[TT][TX];
  double2 *xwc = xw + (size_t)c * V;   // chain-major workspace, indexed like shared memory
  double rr = 0.0;
#pragma unroll
  for (int i = 0; i < TT; i++)
#pragma unroll
    for (int j = 0; j < TX; j++) {
      const int k = (t0 + i) * NX + g * TX + j;
      const int ks = (t0 + i) * NX + j * NG + g;
      r[i][j] = bsrc[(size_t)k * C + c];
      p[i][j] = r[i][j];
      rr = fma(r[i][j].x, r[i][j].x, rr);
      rr = fma(r[i][j].y, r[i][j].y, rr);
      F[ks] = p[i][j];
    }
  rr = block_sum<NWARPS>(rr, scrA);   // hmc.c:354-356
  const double rr_init = rr;
  double rr_old = rr;
  int status = TB_CG_MAXITER, iters = 0;
  __syncthreads();  // F and the links are in place

  if (rr_old < s.accuracy) {  // hmc.c:359-361
    status = TB_CG_ZERO_SOURCE;
  } else {
    for (int k = 1; k < s.max_iter; k++) {  // hmc.c:364
      double2 mp[TT][TX], q[TT][TX];
      tile_apply<NT, NX, TX, TT, false, HAS_MU>(p, mp, F, W0s, W1s, t0, g, m, e_p, e_m);   // Mp = M p, hmc.c:366
      double pq = 0.0;
      if (DAG) {
        // M~ = M^dagger: <p, M^dagger M p> = |M p|^2, so alpha is known before q exists and the pq reduction
        // overlaps the barrier that publishes Mp
#pragma unroll
        for (int i = 0; i < TT; i++)
#pragma unroll
          for (int j = 0; j < TX; j++) {
            pq = fma(mp[i][j].x, mp[i][j].x, pq);
            pq = fma(mp[i][j].y, mp[i][j].y, pq);
          }
      }
      __syncthreads();  // everyone has read p from F
#pragma unroll
      for (int i = 0; i < TT; i++)
#pragma unroll
        for (int j = 0; j < TX; j++) F[(t0 + i) * NX + j * NG + g] = mp[i][j];
      if (DAG) pq = block_sum<NWARPS>(pq, scrB);   // its barrier also publishes Mp
      else __syncthreads();
      // q = M~ Mp, hmc.c:367 (M^dagger swaps the roles of e^{mu} and e^{-mu})
      tile_apply<NT, NX, TX, TT, DAG, HAS_MU>(mp, q, F, W0s, W1s, t0, g, m, DAG ? e_m : e_p, DAG ? e_p : e_m);
      if (!DAG) {
#pragma unroll
        for (int i = 0; i < TT; i++)
#pragma unroll
          for (int j = 0; j < TX; j++) {   // hmc.c:368-370
            pq = fma(p[i][j].x, q[i][j].x, pq);
            pq = fma(p[i][j].y, q[i][j].y, pq);
          }
        pq = block_sum<NWARPS>(pq, scrB);
      }
      const double a = rr_old / pq;   // hmc.c:371
      rr = 0.0;
#pragma unroll
      for (int i = 0; i < TT; i++)
#pragma unroll
        for (int j = 0; j < TX; j++) {
          r[i][j].x = fma(-a, q[i][j].x, r[i][j].x);   // hmc.c:374-375
          r[i][j].y = fma(-a, q[i][j].y, r[i][j].y);
          rr = fma(r[i][j].x, r[i][j].x, rr);          // hmc.c:377-379
          rr = fma(r[i][j].y, r[i][j].y, rr);
        }
      if (XT) {
        // x += a p (hmc.c:372-373) in tensor memory: 4 sites (16 words) per tcgen05.ld / tcgen05.st
        rr = block_sum<NWARPS>(rr, scrA);
#pragma unroll
        for (int ch = 0; ch < Cfg::TMEM_WORDS / 16; ch++) {
          uint32_t v[16];
          if (k > 1) tmem_ld16(v, xaddr + ch * 16);
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int f = ch * 4 + u, i = f / TX, j = f % TX;
            double xr = (k > 1) ? __hiloint2double((int)v[4 * u + 1], (int)v[4 * u]) : 0.0;   // hmc.c:351: x0 = 0
            double xi = (k > 1) ? __hiloint2double((int)v[4 * u + 3], (int)v[4 * u + 2]) : 0.0;
            xr += a * p[i][j].x;
            xi += a * p[i][j].y;
            v[4 * u] = (uint32_t)__double2loint(xr);
            v[4 * u + 1] = (uint32_t)__double2hiint(xr);
            v[4 * u + 2] = (uint32_t)__double2loint(xi);
            v[4 * u + 3] = (uint32_t)__double2hiint(xi);
          }
          tmem_st16(xaddr + ch * 16, v);
        }
        tmem_wait_st();
      } else {
      // x += a p (hmc.c:372-373): x lives in L2 (chain-major workspace, one owner thread per element); the
      // loads are issued before the ||r||^2 reduction so that their latency hides behind its barrier
      double2 xv[TT][TX];
      if (k > 1) {
#pragma unroll
        for (int i = 0; i < TT; i++)
#pragma unroll
          for (int j = 0; j < TX; j++) xv[i][j] = __ldcg(&xwc[(t0 + i) * NX + j * NG + g]);
      }
      rr = block_sum<NWARPS>(rr, scrA);
#pragma unroll
      for (int i = 0; i < TT; i++)
#pragma unroll
        for (int j = 0; j < TX; j++) {
          double2 v = (k > 1) ? xv[i][j] : make_double2(0.0, 0.0);   // hmc.c:351: x0 = 0
          v.x += a * p[i][j].x;
          v.y += a * p[i][j].y;
          __stcg(&xwc[(t0 + i) * NX + j * NG + g], v);
        }
      }
      iters = k;
      if (rr < s.accuracy) { status = TB_CG_CONVERGED; break; }                                        // hmc.c:381
      if (!(rr == rr) || rr / rr_init > TB_DIVERGENCE_RATIO) { status = TB_CG_DIVERGED; break; }      // hmc.c:383
      const double be = rr / rr_old;   // hmc.c:390
#pragma unroll
      for (int i = 0; i < TT; i++)
#pragma unroll
        for (int j = 0; j < TX; j++) {
          p[i][j].x = fma(be, p[i][j].x, r[i][j].x);   // hmc.c:391-392
          p[i][j].y = fma(be, p[i][j].y, r[i][j].y);
          F[(t0 + i) * NX + j * NG + g] = p[i][j];     // all reads of Mp finished before the pq reduction's barrier
        }
      rr_old = rr;
      __syncthreads();
    }
  }
  if (XT) {
#pragma unroll
    for (int ch = 0; ch < Cfg::TMEM_WORDS / 16; ch++) {
      uint32_t v[16];
      if (iters > 0) tmem_ld16(v, xaddr + ch * 16);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int f = ch * 4 + u, i = f / TX, j = f % TX;
        const int kk = (t0 + i) * NX + g * TX + j;
        xout[(size_t)kk * C + c] = (iters > 0) ? make_double2(__hiloint2double((int)v[4 * u + 1], (int)v[4 * u]),
                                                               __hiloint2double((int)v[4 * u + 3], (int)v[4 * u + 2]))
                                                : make_double2(0.0, 0.0);
      }
    }
    __syncthreads();   // every warp has read its columns
    if (threadIdx.x < 32)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base_s), "r"(Cfg::TMEM_COLS) : "memory");
  } else {
  // every element of the workspace was written and is read back by the same thread
#pragma unroll
  for (int i = 0; i < TT; i++)
#pragma unroll
    for (int j = 0; j < TX; j++) {
      const int k = (t0 + i) * NX + g * TX + j;
      xout[(size_t)k * C + c] = (iters > 0) ? __ldcg(&xwc[(t0 + i) * NX + j * NG + g]) : make_double2(0.0, 0.0);
    }
  }
  if (tid == 0) {
    s.status[c] = status;
    s.iters[c] = iters;
    s.rr[c] = rr;
    s.rr_init[c] = rr_init;
    s.active[c] = 0;
  }
}

// ---- links in tensor memory (64^2) -------------------------------------------------------------------------------
// The same solve with the two link fields kept as thread-private TMEM columns next to x (tb_onchip.cuh): the
// stencil's shared-memory traffic drops from 62 to 20 B per site and apply (only the tile's field halo), and the
// 128 KB of shared memory this frees hold p and Mp in SEPARATE exchange buffers, which removes the write-after-read
// barrier: three CTA barriers per CG iteration instead of four.  8 x 2 sites per thread, 256 threads.
template <int NT, int NX>
struct ResidentWtCfg {
  static constexpr int V = NT * NX, TX = 2, TT = 8;
  static constexpr int NTHREADS = (NT / TT) * (NX / TX);
  static constexpr int NWARPS = NTHREADS / 32;
  static constexpr size_t SMEM = (size_t)V * 32 + 64 * sizeof(double);
  static_assert(NWARPS == 8, "two warps per TMEM lane quarter; block_sum8");
};

// CANON: the host-buffer path.  bsrc and xout are in the CANONICAL layout [chain][t][x] (what the reference's
// row-pointer vectors flatten to), so a sub-batch of chains goes H2D copy -> this kernel -> D2H copy with no re-layout
// kernel in between, and a chain's 64 KB are one contiguous, fully coalesced run.  With canon.A set the links are built
// here too, from the canonical angles (the arithmetic of links_kernel, bit for bit), and written back in the device
// layout together with the angles, so the context's W0 / W1 / Adev are what tb_set_gauge would have left.
struct TbCanon {
  const double2 *A;          // canonical angles [chain][t][x] = (A_t, A_x), or nullptr: links are read from W0g / W1g
  double2 *W0, *W1, *Adev;   // device-layout outputs (only with A)
};

template <int NT, int NX, bool DAG, bool HAS_MU, bool MASKED, bool PLAN, bool CANON = false>
__global__ void __launch_bounds__(ResidentWtCfg<NT, NX>::NTHREADS, 1)
resident_wt_kernel(const double2 *__restrict__ bsrc, double2 *__restrict__ xout, const double2 *__restrict__ W0g,
                   const double2 *__restrict__ W1g, const double *__restrict__ mass, const double *__restrict__ msite,
                   const double *__restrict__ emu, const double *__restrict__ emmu, const TbCgState s, const int C,
                   const int c_first, const TbPlan plan, const TbCanon canon) {
  using Cfg = ResidentWtCfg<NT, NX>;
  constexpr int V = Cfg::V, TX = 2, TT = 8, NG = NX / TX;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *Fp = reinterpret_cast<double2 *>(smem_raw);   // p, as the stencil neighbours read it
  double2 *Fm = Fp + V;                                  // Mp
  double *scrA = reinterpret_cast<double *>(Fm + V);
  double *scrB = scrA + 32;
  __shared__ uint32_t tmem_base_s;
  if (threadIdx.x < 32) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst), "r"(TM_COLS_WT) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t xaddr = tmem_base_s + (((warp & 3u) * 32u) << 16) + (warp >> 2) * (uint32_t)TM_SPAN;

  const int tid = threadIdx.x;
  const int g = tid % NG;
  const int t0 = (tid / NG) * TT;
  const int tm = (t0 + NT - 1) % NT, te = (t0 + TT) % NT;   // rows above and below the tile (periodic)
  // The segment bookkeeping of a planned launch lives in shared memory and is re-read (volatile) where it is needed:
  // the CG loop has no register to spare (255), and four more live values cost it 16 spill instructions per
  // iteration and 17 % of its speed.
  __shared__ int seg_s[4];   // chain (-1: no more work), k_begin, k_end, index of the CTA's next segment
  volatile int *const sv = seg_s;
  if (PLAN && tid == 0) sv[3] = plan.seg_lo[blockIdx.x];
  for (bool more = true; more; more = PLAN) {
  int c = c_first + blockIdx.x, k_begin = 1;
  if (PLAN) {
    __syncthreads();
    if (tid == 0) {
      const int sg = sv[3];
      int4 q = make_int4(-1, 1, 0, 0);
      if (sg < plan.seg_hi[blockIdx.x]) {
        q = plan.segs[sg];
        if (q.y > 1) {   // the tail of a split chain: its head was the first job of a CTA with a lower index
          const int h = plan_wait_hand(&plan.hand[q.x]);
          __threadfence();
          if (h < 0) q.y = -1;   // the chain ended inside its head
        }
      }
      sv[0] = q.x; sv[1] = q.y; sv[2] = q.z; sv[3] = sg + 1;
    }
    __syncthreads();
    c = sv[0];
    k_begin = sv[1];
    if (c < 0) break;
    if (k_begin < 0) continue;
  }
  const double m = mass[c];
  const double e_p = emu[c], e_m = emmu[c];

  if (CANON && canon.A) {
    // links from the canonical angles: W0 = s0(t) 1/2 eta0(x) e^{iA_t}, W1 = s1(x) 1/2 e^{iA_x} (links_kernel,
    // hmc.c:140-174), for the tile and its backward halo; the tile's own links and angles also go to the context's
    // device-layout arrays
    const double2 *Ac = canon.A + (size_t)c * V;
    const int xm = (g * TX + NX - 1) % NX, tmr = (t0 + NT - 1) % NT;
    auto link0 = [&](double a, int t, int x) {
      double sn, cs;
      sincos(a, &sn, &cs);
      double f0 = (x & 1) ? -0.5 : 0.5;
      if (t == NT - 1) f0 = -f0;
      return make_double2(f0 * cs, f0 * sn);
    };
    auto link1 = [&](double a, int x) {
      double sn, cs;
      sincos(a, &sn, &cs);
      const double f1 = (x == NX - 1) ? -0.5 : 0.5;
      return make_double2(f1 * cs, f1 * sn);
    };
#pragma unroll 1
    for (int i = 0; i < TT; i++) {
      const int t = t0 + i;
#pragma unroll
      for (int j = 0; j < TX; j++) {
        const int x = g * TX + j, k = t * NX + x;
        const double2 a = Ac[k];
        const double2 w0 = link0(a.x, t, x), w1 = link1(a.y, x);
        tmem_st_d2(xaddr + TM_ROW + 16 * i + 4 * j, w0);
        tmem_st_d2(xaddr + TM_ROW + 16 * i + 8 + 4 * j, w1);
        canon.W0[(size_t)k * C + c] = w0;
        canon.W1[(size_t)k * C + c] = w1;
        canon.Adev[(size_t)k * C + c] = a;
      }
      tmem_st_d2(xaddr + TM_W1M + 4 * i, link1(Ac[t * NX + xm].y, xm));
    }
#pragma unroll
    for (int j = 0; j < TX; j++) tmem_st_d2(xaddr + TM_W0M + 4 * j, link0(Ac[tmr * NX + g * TX + j].x, tmr, g * TX + j));
    tmem_wait_st();
  } else
  // links of the tile and of its backward halo: device layout [site][chain] -> tensor memory
  {
    const int xm = (g * TX + NX - 1) % NX, tmr = (t0 + NT - 1) % NT;
#pragma unroll
    for (int i = 0; i < TT; i++) {
      const size_t row = (size_t)(t0 + i) * NX;
#pragma unroll
      for (int j = 0; j < TX; j++) {
        tmem_st_d2(xaddr + TM_ROW + 16 * i + 4 * j, W0g[(row + g * TX + j) * C + c]);
        tmem_st_d2(xaddr + TM_ROW + 16 * i + 8 + 4 * j, W1g[(row + g * TX + j) * C + c]);
      }
      tmem_st_d2(xaddr + TM_W1M + 4 * i, W1g[(row + xm) * C + c]);
    }
#pragma unroll
    for (int j = 0; j < TX; j++) tmem_st_d2(xaddr + TM_W0M + 4 * j, W0g[((size_t)tmr * NX + g * TX + j) * C + c]);
    tmem_wait_st();
  }
  double2 r[TT][TX], p[TT][TX];
  double rr = 0.0, rr_init, rr_old;
  uint32_t occ = 0;   // family B: occupied sites of the tile (identity rows, vec_ops.c:130)
  if (PLAN && k_begin > 1) {
    // resume: r, p from the stored state, x back into tensor memory, p published for the stencil
    const size_t sb = (size_t)c * V + tid;
#pragma unroll
    for (int i = 0; i < TT; i++)
#pragma unroll
      for (int j = 0; j < TX; j++) {
        const int f = i * TX + j;
        r[i][j] = __ldcg(&plan.sr[sb + (size_t)f * Cfg::NTHREADS]);
        p[i][j] = __ldcg(&plan.sp[sb + (size_t)f * Cfg::NTHREADS]);
        tmem_st_d2(xaddr + TM_X + 4 * f, __ldcg(&plan.sx[sb + (size_t)f * Cfg::NTHREADS]));
        Fp[(t0 + i) * NX + j * NG + g] = p[i][j];
      }
    tmem_wait_st();
    rr_old = __ldcg(&s.rr_old[c]);
    rr_init = __ldcg(&s.rr_init[c]);
    rr = rr_old;
    __syncthreads();
  } else {
#pragma unroll
  for (int i = 0; i < TT; i++)
#pragma unroll
    for (int j = 0; j < TX; j++) {
      const int k = (t0 + i) * NX + g * TX + j;
      r[i][j] = CANON ? bsrc[(size_t)c * V + k] : bsrc[(size_t)k * C + c];
      p[i][j] = r[i][j];
      rr = fma(r[i][j].x, r[i][j].x, rr);
      rr = fma(r[i][j].y, r[i][j].y, rr);
      Fp[(t0 + i) * NX + j * NG + g] = p[i][j];
      if (MASKED && msite[(size_t)k * C + c] != m) occ |= 1u << (i * TX + j);
    }
  rr = block_sum8(rr, scrA);   // hmc.c:354-356; its barrier publishes p
  rr_init = rr;
  rr_old = rr;
  }
  int status = TB_CG_MAXITER, iters = k_begin - 1;

  if (rr_old < s.accuracy && k_begin == 1) {  // hmc.c:359-361
    status = TB_CG_ZERO_SOURCE;
  } else {
    for (int k = k_begin; k < s.max_iter && (!PLAN || k < sv[2]); k++) {  // hmc.c:364
      // Mp = M p (hmc.c:366): every site is published to Fm as soon as it is finished (the last readers of Fm
      // passed the ||r||^2 barrier), and <p, M^dagger M p> = |M p|^2 is accumulated on the way
      double2 mp[TT][TX];
      double pq = 0.0;
      tile_apply_wt<NX, false, HAS_MU, MASKED>(
          p, Fp, Fp + tm * NX, Fp + te * NX, false, xaddr, t0, g, m, occ, e_p, e_m,
          [&](int i, int j, const double2 o) {
            mp[i][j] = o;
            Fm[(t0 + i) * NX + j * NG + g] = o;
            if (DAG) {
              pq = fma(o.x, o.x, pq);
              pq = fma(o.y, o.y, pq);
            }
          },
          [] {});
      if (DAG) pq = block_sum8(pq, scrB);   // its barrier also publishes Mp
      else __syncthreads();
      rr = 0.0;
      double a;
      if (DAG) {
        // alpha is known before q = M^dagger Mp exists (hmc.c:367,371): q is consumed site by site,
        // r -= alpha q and ||r||^2 (hmc.c:374-379), and never stored
        a = rr_old / pq;
        tile_apply_wt<NX, true, HAS_MU, MASKED>(
            mp, Fm, Fm + tm * NX, Fm + te * NX, false, xaddr, t0, g, m, occ, e_m, e_p,
            [&](int i, int j, const double2 o) {
              r[i][j].x = fma(-a, o.x, r[i][j].x);
              r[i][j].y = fma(-a, o.y, r[i][j].y);
              rr = fma(r[i][j].x, r[i][j].x, rr);
              rr = fma(r[i][j].y, r[i][j].y, rr);
            },
            [] {});
      } else {
        double2 q[TT][TX];
        tile_apply_wt<NX, false, HAS_MU, MASKED>(
            mp, Fm, Fm + tm * NX, Fm + te * NX, false, xaddr, t0, g, m, occ, e_p, e_m,
            [&](int i, int j, const double2 o) {
              q[i][j] = o;
              pq = fma(p[i][j].x, o.x, pq);   // hmc.c:368-370
              pq = fma(p[i][j].y, o.y, pq);
            },
            [] {});
        pq = block_sum8(pq, scrB);
        a = rr_old / pq;   // hmc.c:371
#pragma unroll
        for (int i = 0; i < TT; i++)
#pragma unroll
          for (int j = 0; j < TX; j++) {
            r[i][j].x = fma(-a, q[i][j].x, r[i][j].x);   // hmc.c:374-375
            r[i][j].y = fma(-a, q[i][j].y, r[i][j].y);
            rr = fma(r[i][j].x, r[i][j].x, rr);          // hmc.c:377-379
            rr = fma(r[i][j].y, r[i][j].y, rr);
          }
      }
      // ||r||^2 with x += a p (hmc.c:372-373) between the posting of the warp partial and the barrier: warps that are
      // ahead update their tensor-memory columns instead of waiting
      block_sum8_post(rr, scrA);
      tmem_x_axpy(xaddr + TM_X, p, a, k == 1);
      rr = block_sum8_total(scrA);
      iters = k;
      if (rr < s.accuracy) { status = TB_CG_CONVERGED; break; }                                        // hmc.c:381
      if (!(rr == rr) || rr / rr_init > TB_DIVERGENCE_RATIO) { status = TB_CG_DIVERGED; break; }      // hmc.c:383
      const double be = rr / rr_old;   // hmc.c:390
#pragma unroll
      for (int i = 0; i < TT; i++)
#pragma unroll
        for (int j = 0; j < TX; j++) {
          p[i][j].x = fma(be, p[i][j].x, r[i][j].x);   // hmc.c:391-392
          p[i][j].y = fma(be, p[i][j].y, r[i][j].y);
          Fp[(t0 + i) * NX + j * NG + g] = p[i][j];    // the last readers of Fp passed the |Mp|^2 barrier
        }
      rr_old = rr;
      __syncthreads();
    }
  }
  if (PLAN) c = sv[0];
  if (PLAN && status == TB_CG_MAXITER && sv[2] < s.max_iter) {
    // the head of a split chain ends here: store r, p, x and the two scalars, then raise the chain's flag
    const size_t sb = (size_t)c * V + tid;
#pragma unroll
    for (int ch = 0; ch < 4; ch++) {
      uint32_t v[16];
      tmem_ld16(v, xaddr + TM_X + ch * 16);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int f = ch * 4 + u, i = f / TX, j = f % TX;
        __stcg(&plan.sr[sb + (size_t)f * Cfg::NTHREADS], r[i][j]);
        __stcg(&plan.sp[sb + (size_t)f * Cfg::NTHREADS], p[i][j]);
        __stcg(&plan.sx[sb + (size_t)f * Cfg::NTHREADS],
               make_double2(__hiloint2double((int)v[4 * u + 1], (int)v[4 * u]),
                            __hiloint2double((int)v[4 * u + 3], (int)v[4 * u + 2])));
      }
    }
    if (tid == 0) {
      s.rr_old[c] = rr_old;
      s.rr_init[c] = rr_init;
    }
    __threadfence();
    __syncthreads();   // every thread's state is out (and every warp has read its tensor-memory columns)
    if (tid == 0) {
      __threadfence();
      *(volatile int *)&plan.hand[c] = sv[2];
    }
    continue;
  }
#pragma unroll
  for (int ch = 0; ch < 4; ch++) {
    uint32_t v[16];
    if (iters > 0) tmem_ld16(v, xaddr + TM_X + ch * 16);
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int f = ch * 4 + u, i = f / TX, j = f % TX;
      const int kk = (t0 + i) * NX + g * TX + j;
      xout[CANON ? (size_t)c * V + kk : (size_t)kk * C + c] =
          (iters > 0) ? make_double2(__hiloint2double((int)v[4 * u + 1], (int)v[4 * u]),
                                     __hiloint2double((int)v[4 * u + 3], (int)v[4 * u + 2]))
                      : make_double2(0.0, 0.0);
    }
  }
  __syncthreads();   // every warp has read its columns
  if (tid == 0) {
    s.status[c] = status;
    s.iters[c] = iters;
    s.rr[c] = rr;
    s.rr_init[c] = rr_init;
    s.active[c] = 0;
    if (PLAN && sv[2] != 0x7fffffff) {   // the chain ended inside its head: the CTA that holds the tail skips it
      __threadfence();
      *(volatile int *)&plan.hand[c] = -1;
    }
  }
  }   // segments
  if (PLAN) __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base_s), "r"(TM_COLS_WT) : "memory");
}

template <int NT, int NX>
int launch_resident_wt(tb_ctx *ctx, const double2 *b, double2 *x, int c0, int n, cudaStream_t st, const TbCanon *canon) {
  using Cfg = ResidentWtCfg<NT, NX>;
  const bool dag = tb_conj_is_dagger(ctx);
  // whole batches larger than the SM count are balanced over the SMs (see TbPlan)
  int nsm = TB_NUM_SMS_B200;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
  const bool plan = !canon && plan_pays(ctx, c0, n, nsm);
  TbPlan pl = {};
  TbCanon cn = {};
  if (plan) {
    TB_CHECK(plan_prepare(ctx, nsm, st, &pl));
    auto kern = resident_wt_kernel<NT, NX, false, false, false, true>;
    if (dag) kern = ctx->has_mu ? resident_wt_kernel<NT, NX, true, true, false, true> : resident_wt_kernel<NT, NX, true, false, false, true>;
    else if (ctx->has_mu) kern = resident_wt_kernel<NT, NX, false, true, false, true>;
    TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    kern<<<nsm, Cfg::NTHREADS, Cfg::SMEM, st>>>(b, x, ctx->W0, ctx->W1, ctx->d_mass, ctx->msite, ctx->d_emu, ctx->d_emmu,
                                                ctx->cg, ctx->C, 0, pl, cn);
    ctx->launches++;
    TB_CUDA(cudaGetLastError());
    return TB_OK;
  }
  auto kern = resident_wt_kernel<NT, NX, false, false, false, false>;
  if (canon) {   // host-buffer path: canonical source / solution (and angles), family A only
    cn = *canon;
    if (dag) kern = ctx->has_mu ? resident_wt_kernel<NT, NX, true, true, false, false, true> : resident_wt_kernel<NT, NX, true, false, false, false, true>;
    else kern = ctx->has_mu ? resident_wt_kernel<NT, NX, false, true, false, false, true> : resident_wt_kernel<NT, NX, false, false, false, false, true>;
  } else if (ctx->msite) kern = ctx->has_mu ? resident_wt_kernel<NT, NX, true, true, true, false> : resident_wt_kernel<NT, NX, true, false, true, false>;
  else if (dag) kern = ctx->has_mu ? resident_wt_kernel<NT, NX, true, true, false, false> : resident_wt_kernel<NT, NX, true, false, false, false>;
  else if (ctx->has_mu) kern = resident_wt_kernel<NT, NX, false, true, false, false>;
  TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  kern<<<n, Cfg::NTHREADS, Cfg::SMEM, st>>>(b, x, ctx->W0, ctx->W1, ctx->d_mass, ctx->msite, ctx->d_emu, ctx->d_emmu,
                                            ctx->cg, ctx->C, c0, pl, cn);
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

template <int NT, int NX, int TX, int TT>
int launch_resident(tb_ctx *ctx, const double2 *b, double2 *x, int c0, int n, cudaStream_t st) {
  using Cfg = ResidentCfg<NT, NX, TX, TT>;
  const bool dag = tb_conj_is_dagger(ctx);
  constexpr bool T = Cfg::TMEM_OK;
  const bool xt = T && ctx->resident_x_tmem;
  auto kern = resident_cg_kernel<NT, NX, TX, TT, false, false, false>;
#define PICK(D, M)                                                                               \
  kern = xt ? resident_cg_kernel<NT, NX, TX, TT, D, M, T> : resident_cg_kernel<NT, NX, TX, TT, D, M, false>;
  if (dag) { if (ctx->has_mu) { PICK(true, true) } else { PICK(true, false) } }
  else { if (ctx->has_mu) { PICK(false, true) } else { PICK(false, false) } }
#undef PICK
  TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  if (getenv("TB_DEBUG")) {
    int per_sm = -1;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Cfg::NTHREADS, Cfg::SMEM);
    fprintf(stderr, "resident kernel %dx%d tile %dx%d x_tmem=%d: %d CTAs/SM by occupancy, %d threads, %d regs, %zu B smem\n", NT, NX,
            TT, TX, (int)xt, per_sm, Cfg::NTHREADS, fa.numRegs, (size_t)Cfg::SMEM);
  }
  kern<<<n, Cfg::NTHREADS, Cfg::SMEM, st>>>(b, x, ctx->W0, ctx->W1, ctx->d_mass, ctx->d_emu, ctx->d_emmu, ctx->xw,
                                            ctx->cg, ctx->C, c0);
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

}  // namespace

bool tb_resident_supported(const tb_ctx *ctx) {
  if (ctx->nranks != 1 || ctx->nt != ctx->nx) return false;
  // family B (occupation mask): the 64^2 kernel with the links in tensor memory handles it, for M~ = M^T
  if (ctx->msite) return ctx->nt == 64 && tb_conj_is_dagger(ctx) && ctx->resident_x_tmem && ctx->tune_tt == 0;
  return ctx->nt == 16 || ctx->nt == 32 || ctx->nt == 64;
}

// One kernel launch per (sub-)batch of chains [c0, c0+n).  The tile shape per thread is a tuning knob
// (tb_set_tuning rows_per_thread): 0 = default (64^2: 2x8 sites with the links in tensor memory; 32^2: 2x8; 16^2: 2x4),
// 44 = 4x4, 18 = 1x8, 28 = 2x8 with the links in shared memory, 24 = 2x4.
int tb_run_cg_resident_slice(tb_ctx *ctx, const double2 *b, double2 *x, int c0, int n, cudaStream_t st) {
  if (b == x) {
    tb_set_error("tb_run_cg_resident: in-place solve is not supported");
    return TB_EINVAL;
  }
  const int shape = ctx->tune_tt;
  switch (ctx->nt) {
    case 16:
      if (shape == 18) return launch_resident<16, 16, 1, 8>(ctx, b, x, c0, n, st);
      return launch_resident<16, 16, 2, 4>(ctx, b, x, c0, n, st);
    case 32:
      if (shape == 18) return launch_resident<32, 32, 1, 8>(ctx, b, x, c0, n, st);
      if (shape == 44) return launch_resident<32, 32, 4, 4>(ctx, b, x, c0, n, st);
      return launch_resident<32, 32, 2, 8>(ctx, b, x, c0, n, st);
    case 64:
      if (shape == 18) return launch_resident<64, 64, 1, 8>(ctx, b, x, c0, n, st);
      if (shape == 44) return launch_resident<64, 64, 4, 4>(ctx, b, x, c0, n, st);
      if (shape == 24) return launch_resident<64, 64, 2, 4>(ctx, b, x, c0, n, st);
      if (shape == 28 || !ctx->resident_x_tmem) return launch_resident<64, 64, 2, 8>(ctx, b, x, c0, n, st);
      return launch_resident_wt<64, 64>(ctx, b, x, c0, n, st, nullptr);   // default: links and x in tensor memory
    default: tb_set_error("resident solver: unsupported lattice %dx%d", ctx->nt, ctx->nx); return TB_EINVAL;
  }
}

// Host-buffer path of the 64^2 kernel: chains [c0, c0 + n) with source and solution in the canonical layout
// (b_canon / x_canon point at chain 0 of the context) and, when A_canon is not null, the links built inside the kernel
// from the canonical angles.  Returns TB_EINVAL when the context is not served by that kernel (the caller falls back to
// the re-layout kernels).
bool tb_resident_canon_supported(const tb_ctx *ctx) {
  return tb_resident_supported(ctx) && ctx->nt == 64 && !ctx->msite && ctx->resident_x_tmem && ctx->tune_tt == 0 &&
         getenv("TB_NO_CANON") == nullptr;
}

int tb_run_cg_resident_canon(tb_ctx *ctx, const double2 *b_canon, double2 *x_canon, const double2 *A_canon, int c0, int n,
                             cudaStream_t st) {
  if (!tb_resident_canon_supported(ctx)) return TB_EINVAL;
  TbCanon cn = {A_canon, ctx->W0, ctx->W1, ctx->Adev};
  if (A_canon) tb_gauge_sharing(ctx, false);   // the kernel writes per-chain links
  return launch_resident_wt<64, 64>(ctx, b_canon, x_canon, c0, n, st, &cn);
}

int tb_run_cg_resident(tb_ctx *ctx, const double2 *b, double2 *x) {
  return tb_run_cg_resident_slice(ctx, b, x, 0, ctx->C, ctx->stream);
}

// The schedule plan_kernel would write for these iteration estimates, computed on the host by the same code (no GPU
// involved): segs4[4 * s] = (chain, first iteration, end iteration, 0), CTA b owns segments [seg_lo[b], seg_hi[b]).
// segs4 holds up to C + M - 1 segments.  For tests and for inspecting a plan.
extern "C" int tb_plan_schedule(const int *est, const int *status, int nchains, int machines, int *segs4, int *seg_lo,
                                int *seg_hi) {
  if (!est || !status || !segs4 || !seg_lo || !seg_hi || nchains < 1 || machines < 1) {
    tb_set_error("tb_plan_schedule: invalid arguments");
    return TB_EINVAL;
  }
  long long W = 0;
  int mx = 0;
  bool bad = false;
  for (int c = 0; c < nchains; c++) {
    if (est[c] <= 0 || status[c] != TB_CG_CONVERGED) bad = true;
    W += est[c];
    mx = est[c] > mx ? est[c] : mx;
  }
  int4 *segs = reinterpret_cast<int4 *>(segs4);
  if (bad) {
    for (int b = 0; b < machines; b++) plan_deal(b, nchains, machines, segs, seg_lo, seg_hi);
  } else {
    plan_fill(est, nchains, machines, W, mx, segs, seg_lo, seg_hi);
  }
  return TB_OK;
}
