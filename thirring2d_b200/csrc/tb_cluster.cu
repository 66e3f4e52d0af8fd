// tb_cluster.cu — on-chip resident batched CG for lattices that do not fit one SM: one THREAD-BLOCK CLUSTER per
// Markov chain (128^2 = 4 CTAs, 256^2 = 16 CTAs, and the rectangles in between), sm_100a, FP64.
//
// Same algorithm and per-thread arithmetic as the 64^2 kernel of tb_resident.cu (fmdm_invert_cg, hmc.c:341-404),
// with the lattice cut into t-slabs of 4096 sites (LT = 4096/NX rows of NX sites), one per CTA of the cluster.
// The CG state of a slab never leaves its SM:
//
//   registers      r, p (persistent), Mp (transient): 8 x 2 sites per thread, 256 threads
//   shared memory  exchange fields Fp (p) and Fm (Mp) of the slab                       32 B/site = 128 KB
//                  halo rows t = -1 and t = LT of p and of Mp, pushed by the neighbours  4 x NX x 16 B
//   tensor memory  x and the links W0, W1 of the thread's tile, thread-private columns (tcgen05.ld/st 32x32b)
//
// Distributed shared memory carries everything that crosses a CTA boundary:
//   * halos are PUSHED: the first / last row of p or Mp of a slab goes into the neighbour CTA's halo buffer as one bulk
//     copy shared -> peer shared out of the exchange field; the stencil reads local memory only;
//   * both CG reductions are all-to-all over the cluster: every CTA sends its block sum into every CTA's slot table
//     (8-byte st.async), and all CTAs add the slots with the same tree => bitwise identical alpha, beta and stopping
//     decisions on every CTA with no further communication (and run-to-run deterministic);
//   * every hand-over is counted in bytes on an mbarrier of the CONSUMER, which waits where the latency is hidden: a
//     stencil does the first four rows of every tile (which need only the CTA's own data, except one hop into the first
//     row of the slab that is added afterwards), THEN waits, and finishes with the halo rows in place; the x += alpha p
//     update in tensor memory runs while the ||r||^2 partials travel.  Producers never wait (round 1's
//     st.shared::cluster + barrier.cluster.arrive.release cost them the round trip of their stores, three times per
//     iteration).
// No cluster barrier inside the CG loop; they remain in the prologue and at the segment boundaries of a planned launch.
// HBM is touched once per solve.
#include <cstdint>

#include "tb_common.cuh"

namespace {

#include "tb_onchip.cuh"

#ifndef TB_CLUSTER_P2P
#define TB_CLUSTER_P2P 1   // 0: the three split cluster barriers per iteration of round 1 (kept for comparison builds)
#endif
constexpr bool P2P = TB_CLUSTER_P2P != 0;
constexpr int TX = 2, TT = 8;              // sites per thread: TT rows (t) x TX columns (x)
constexpr int VL = 4096;                   // sites per CTA
constexpr int NTHREADS = VL / (TX * TT);   // 256
constexpr int NWARPS = NTHREADS / 32;      // 8
// Geometry and shared-memory map (double2 units from the start of dynamic shared memory) of an NX-wide slab.
template <int NX>
struct Slab {
  static constexpr int LT = VL / NX;         // rows per CTA
  static constexpr int NGX = NX / TX;        // tiles per row; consecutive threads = consecutive tiles of a row
  static_assert(LT % TT == 0 && LT * NX == VL && NGX % 32 == 0, "slab shape");
  static constexpr int OFF_FP = 0, OFF_FM = VL;
  static constexpr int OFF_HP = 2 * VL;              // halo rows of p : [dn (t = -1) | up (t = LT)], row layout
  static constexpr int OFF_HM = OFF_HP + 2 * NX;     // halo rows of Mp
  static constexpr int OFF_END = OFF_HM + 2 * NX;
  static constexpr int H_DN = 0, H_UP = NX;
  // doubles after OFF_END: warp partials A, B [NWARPS each], cluster slots A, B [16 each], then four mbarriers
  // (p halos, Mp halos, slots A, slots B)
  static constexpr size_t SMEM = (size_t)OFF_END * sizeof(double2) + (2 * NWARPS + 32 + 4) * sizeof(double);
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, const double2 v) {
  asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};\n" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void st_cluster(uint32_t addr, const double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;\n" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }

// ---- point-to-point hand-overs inside the CG loop --------------------------------------------------------------------
// A store into a peer's shared memory followed by barrier.cluster.arrive.release costs the PRODUCER the round trip of
// its stores (ncu: 19 % of the 256^2 kernel's warp time in "membar" stalls, three times per iteration).  Here the data
// travels with its own completion: halo rows as one bulk copy shared -> peer shared per row, reduction partials as 8-byte
// st.async, both counted in bytes on an mbarrier of the CONSUMER, which waits for the bytes it expects (two rows; one
// partial per CTA of the cluster) exactly where it used to wait for the cluster barrier.  Producers never wait.
// Flow control needs nothing extra: both reductions are all-to-all, so no CTA can be a whole phase ahead of a peer
// (the data of phase n + 1 of a barrier is only sent after every CTA has contributed to a reduction that it joins after
// its own wait for phase n).
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
// every thread; a wait that outlives ~1e7 polls (seconds: a protocol error) traps instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .u32 n;\n"
      "mov.u32 n, 0;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "add.u32 n, n, 1;\n"
      "setp.gt.u32 p, n, 0x1000000;\n"
      "@p trap;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// 8 bytes into a peer's shared memory, counted on the peer's mbarrier (both addresses shared::cluster)
__device__ __forceinline__ void st_async(uint32_t addr, const double v, uint32_t bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];\n" ::"r"(addr), "d"(v), "r"(bar)
               : "memory");
}
// bytes of my shared memory -> a peer's (dst and bar shared::cluster, src shared::cta; multiples of 16)
__device__ __forceinline__ void bulk_s2c(uint32_t dst, uint32_t src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "r"(src), "r"(bytes), "r"(bar) : "memory");
}

// Block sum -> every CTA's slot table; ends with a CTA barrier BEFORE the remote stores, so the caller's shared-
// memory writes that precede it are visible CTA-wide afterwards.  After the caller's cluster barrier,
// cluster_total adds the CS slots with the same pairwise tree on every CTA.
// The same pairwise tree on every CTA (=> bitwise identical totals everywhere): four dependent additions for 16 slots
// instead of sixteen; every warp of every CTA walks this chain twice per iteration.
template <int N>
__device__ __forceinline__ double tree_sum(const double *v_in) {
  static_assert((N & (N - 1)) == 0, "power of two");
  double v[N];
#pragma unroll
  for (int q = 0; q < N; q++) v[q] = v_in[q];
#pragma unroll
  for (int w = 1; w < N; w <<= 1)
#pragma unroll
    for (int q = 0; q + w < N; q += 2 * w) v[q] += v[q + w];
  return v[0];
}
// bar_saddr != 0: the partial travels as st.async counted on the receiver's mbarrier (the CG loop); 0: a plain store,
// to be followed by a cluster barrier (the prologue)
template <int CS>
__device__ __forceinline__ void cluster_sum_post(double v, double *wscr, uint32_t slots_saddr, uint32_t my_rank,
                                                 uint32_t bar_saddr = 0) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) wscr[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    const double s = tree_sum<NWARPS>(wscr);
    if (threadIdx.x < CS) {
      const uint32_t dst = mapa_shared(slots_saddr + my_rank * 8u, threadIdx.x);
      if (bar_saddr) st_async(dst, s, mapa_shared(bar_saddr, threadIdx.x));
      else st_cluster(dst, s);
    }
  }
}
template <int CS>
__device__ __forceinline__ double cluster_total(const double *slots) {
  return tree_sum<CS>(slots);
}

// first row of the slab -> the "up" halo of the CTA that owns the rows before mine; last row -> the "dn" halo of
// the CTA that owns the rows after mine (periodic in the cluster rank; the antiperiodic sign lives in the links)
template <int NX>
__device__ __forceinline__ void push_rows(const double2 (&v)[TT][TX], int hset, uint32_t smem_base, uint32_t rank_m,
                                          uint32_t rank_p, bool top, bool bot, int g) {
  using G = Slab<NX>;
  if (top) {
    const uint32_t a = mapa_shared(smem_base + (uint32_t)(hset + G::H_UP + g) * 16u, rank_m);
#pragma unroll
    for (int j = 0; j < TX; j++) st_cluster(a + j * G::NGX * 16u, v[0][j]);
  }
  if (bot) {
    const uint32_t a = mapa_shared(smem_base + (uint32_t)(hset + G::H_DN + g) * 16u, rank_p);
#pragma unroll
    for (int j = 0; j < TX; j++) st_cluster(a + j * G::NGX * 16u, v[TT - 1][j]);
  }
}

// The same two rows as bulk copies out of the exchange field (row 0 and row LT - 1 of F, already published inside the
// CTA and laid out like the halo buffers), counted on the receivers' halo mbarrier.  One thread.
template <int NX>
__device__ __forceinline__ void push_rows_bulk(uint32_t f_saddr, int hset, uint32_t smem_base, uint32_t bar_saddr,
                                               uint32_t rank_m, uint32_t rank_p) {
  using G = Slab<NX>;
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // the rows were written through the generic proxy
  bulk_s2c(mapa_shared(smem_base + (uint32_t)(hset + G::H_UP) * 16u, rank_m), f_saddr, NX * 16u, mapa_shared(bar_saddr, rank_m));
  bulk_s2c(mapa_shared(smem_base + (uint32_t)(hset + G::H_DN) * 16u, rank_p), f_saddr + (uint32_t)((G::LT - 1) * NX) * 16u,
           NX * 16u, mapa_shared(bar_saddr, rank_p));
}

// PLAN: the grid is one cluster per "machine" of a planned launch (tb_onchip.cuh: TbPlan) and every cluster works
// through its list of segments; a chain that is split is stored by the cluster that ran its head (r, p, x of every
// slab, the flag raised by rank 0 after a cluster barrier) and picked up by the cluster that runs its tail.  Every
// CTA of a cluster reads the same segment list and the same flag, so the cluster takes every decision as one.
template <int NX, int CS, bool DAG, bool HAS_MU, bool MASKED, bool PLAN>
__global__ void __launch_bounds__(NTHREADS, 1)
cluster_cg_kernel(const double2 *__restrict__ bsrc, double2 *__restrict__ xout, const double2 *__restrict__ W0g,
                  const double2 *__restrict__ W1g, const double *__restrict__ mass, const double *__restrict__ msite,
                  const double *__restrict__ emu, const double *__restrict__ emmu, const TbCgState s, const int C,
                  const int c_first, const TbPlan plan) {
  using G = Slab<NX>;
  constexpr int LT = G::LT, NGX = G::NGX, NT = CS * LT;
  static_assert(CS >= 2 && CS <= 16, "slot tables hold 16 ranks");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *S = reinterpret_cast<double2 *>(smem_raw);
  double2 *Fp = S + G::OFF_FP, *Fm = S + G::OFF_FM;
  double *wscrA = reinterpret_cast<double *>(S + G::OFF_END);
  double *wscrB = wscrA + NWARPS;
  double *slotA = wscrB + NWARPS;
  double *slotB = slotA + 16;
  __shared__ uint32_t tmem_base_s;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t slotA_addr = (uint32_t)__cvta_generic_to_shared(slotA);
  const uint32_t slotB_addr = (uint32_t)__cvta_generic_to_shared(slotB);
  // mbarriers of the point-to-point hand-overs: p halos, Mp halos, slots A (||r||^2), slots B (|Mp|^2 or <p, q>)
  const uint32_t barHP = slotB_addr + 16 * 8, barHM = barHP + 8, barA = barHP + 16, barB = barHP + 24;
  const uint32_t fp_saddr = smem_base + (uint32_t)G::OFF_FP * 16u, fm_saddr = smem_base + (uint32_t)G::OFF_FM * 16u;
  auto init_bars = [&] {   // one thread; made visible to the peers by the cluster barrier every solve starts with
    mbar_init(barHP, 1); mbar_init(barHM, 1); mbar_init(barA, 1); mbar_init(barB, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  };

  if (threadIdx.x < 32) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst), "r"(TM_COLS_WT) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (P2P && !PLAN && threadIdx.x == 0) init_bars();
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t xaddr = tmem_base_s + (((warp & 3u) * 32u) << 16) + (warp >> 2) * (uint32_t)TM_SPAN;

  const uint32_t rank = cluster_ctarank();
  const uint32_t rank_m = (rank + CS - 1) % CS, rank_p = (rank + 1) % CS;
  const int tid = threadIdx.x;
  const int g = tid % NGX;
  const int t0 = (tid / NGX) * TT;
  const bool top = t0 == 0, bot = t0 + TT == LT;   // warp-uniform
  const int tg0 = (int)rank * LT;   // first global row of the slab
  // rows above / below the tile: own exchange field, or the halo rows the neighbour CTAs push
  const double2 *p_dn = top ? S + G::OFF_HP + G::H_DN : Fp + (t0 - 1) * NX;
  const double2 *p_up = bot ? S + G::OFF_HP + G::H_UP : Fp + (t0 + TT) * NX;
  const double2 *m_dn = top ? S + G::OFF_HM + G::H_DN : Fm + (t0 - 1) * NX;
  const double2 *m_up = bot ? S + G::OFF_HM + G::H_UP : Fm + (t0 + TT) * NX;
  // segment bookkeeping of a planned launch: in shared memory, re-read where needed (the CG loop has no register to
  // spare, see resident_wt_kernel)
  __shared__ int seg_s[4];   // chain (-1: no more work), k_begin, k_end, index of the cluster's next segment
  volatile int *const sv = seg_s;
  if (PLAN && tid == 0) sv[3] = plan.seg_lo[blockIdx.x / CS];
  for (bool more = true; more; more = PLAN) {
  int c = c_first + (int)(blockIdx.x / CS), k_begin = 1;
  if (PLAN) {
    __syncthreads();
    if (tid == 0) {
      if (P2P) init_bars();   // every phase of the previous segment was waited for: nothing is in flight
      const int sg = sv[3];
      int4 q = make_int4(-1, 1, 0, 0);
      if (sg < plan.seg_hi[blockIdx.x / CS]) {
        q = plan.segs[sg];
        if (q.y > 1) {   // the tail of a split chain: its head was the first job of a cluster with a lower index
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

  // links of the tile and of its backward halo: device layout [site][chain] -> tensor memory
  {
    const int xm = (g * TX + NX - 1) % NX, tmr = (tg0 + t0 + NT - 1) % NT;
#pragma unroll
    for (int i = 0; i < TT; i++) {
      const size_t row = (size_t)(tg0 + t0 + i) * NX;
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
    const size_t sb = ((size_t)c * CS + rank) * VL + tid;
#pragma unroll
    for (int i = 0; i < TT; i++)
#pragma unroll
      for (int j = 0; j < TX; j++) {
        const int f = i * TX + j;
        r[i][j] = __ldcg(&plan.sr[sb + (size_t)f * NTHREADS]);
        p[i][j] = __ldcg(&plan.sp[sb + (size_t)f * NTHREADS]);
        tmem_st_d2(xaddr + TM_X + 4 * f, __ldcg(&plan.sx[sb + (size_t)f * NTHREADS]));
        Fp[(t0 + i) * NX + j * NGX + g] = p[i][j];
      }
    tmem_wait_st();
    rr_old = __ldcg(&s.rr_old[c]);
    rr_init = __ldcg(&s.rr_init[c]);
    rr = rr_old;
    // every CTA of the cluster has left its previous segment before anyone stores into a peer's shared memory
    cluster_arrive();
    cluster_wait();
    if (!P2P) push_rows<NX>(p, G::OFF_HP, smem_base, rank_m, rank_p, top, bot, g);
    __syncthreads();   // p is published inside the CTA
  } else {
#pragma unroll
  for (int i = 0; i < TT; i++)
#pragma unroll
    for (int j = 0; j < TX; j++) {
      const size_t gs = (size_t)(tg0 + t0 + i) * NX + g * TX + j;
      r[i][j] = bsrc[gs * C + c];
      p[i][j] = r[i][j];
      rr = fma(r[i][j].x, r[i][j].x, rr);
      rr = fma(r[i][j].y, r[i][j].y, rr);
      Fp[(t0 + i) * NX + j * NGX + g] = p[i][j];
      if (MASKED && msite[gs * C + c] != m) occ |= 1u << (i * TX + j);
    }
  // every CTA of the cluster is running (planned launch: has left its previous segment) before anyone stores into a
  // peer's shared memory
  cluster_arrive();
  cluster_wait();
  if (!P2P) push_rows<NX>(p, G::OFF_HP, smem_base, rank_m, rank_p, top, bot, g);
  cluster_sum_post<CS>(rr, wscrA, slotA_addr, rank);   // hmc.c:354-356; its CTA barrier publishes Fp
  cluster_arrive();
  cluster_wait();
  rr = cluster_total<CS>(slotA);
  rr_init = rr;
  rr_old = rr;
  }
  int status = TB_CG_MAXITER, iters = k_begin - 1;

  if (rr_old < s.accuracy && k_begin == 1) {  // hmc.c:359-361
    status = TB_CG_ZERO_SOURCE;
  } else {
    // parity of the mbarrier phase iteration k completes: every barrier completes one phase per iteration, from 0
    auto par = [&](int k) { return (uint32_t)(k ^ (PLAN ? sv[1] : 1)) & 1u; };
    if (P2P) {
      // the halo rows of the first direction, if an iteration will read them (a hand-over nobody waits for would land
      // in the barriers of the cluster's next segment)
      if (tid == 32 && k_begin < s.max_iter && (!PLAN || k_begin < sv[2]))
        push_rows_bulk<NX>(fp_saddr, G::OFF_HP, smem_base, barHP, rank_m, rank_p);
    } else {
      cluster_arrive();   // pairs with the wait inside the first stencil (p and its halos are already published)
      if (k_begin >= s.max_iter || (PLAN && k_begin >= sv[2])) cluster_wait();   // no iteration will run: close the barrier
    }
    for (int k = k_begin; k < s.max_iter && (!PLAN || k < sv[2]); k++) {  // hmc.c:364
      // ---- Mp = M p (hmc.c:366).  Finished sites go to Fm at once (its last readers passed the ||r||^2
      // barrier); the first row of the slab waits for its hop from the halo row.
      double2 mp[TT][TX];
      double pq = 0.0;
      tile_apply_wt<NX, false, HAS_MU, MASKED>(
          p, Fp, p_dn, p_up, top, xaddr, t0, g, m, occ, e_p, e_m,
          [&](int i, int j, const double2 o) {
            mp[i][j] = o;
            if (!(i == 0 && top)) {
              Fm[(t0 + i) * NX + j * NGX + g] = o;
              if (DAG) {   // <p, M^dagger M p> = |M p|^2
                pq = fma(o.x, o.x, pq);
                pq = fma(o.y, o.y, pq);
              }
            }
          },
          [&] {   // the p halos of this iteration are in place
            if (P2P) {
              if (tid == 0) mbar_expect_tx(barHP, 2u * NX * 16u);
              mbar_wait(barHP, par(k));
            } else {
              cluster_wait();
            }
          });
      if (top) {
        tile_fixup_dn<NX, false, HAS_MU>(mp[0], p_dn, xaddr, g, e_m);
#pragma unroll
        for (int j = 0; j < TX; j++) {
          Fm[t0 * NX + j * NGX + g] = mp[0][j];
          if (DAG) {
            pq = fma(mp[0][j].x, mp[0][j].x, pq);
            pq = fma(mp[0][j].y, mp[0][j].y, pq);
          }
        }
      }
      if (!P2P) push_rows<NX>(mp, G::OFF_HM, smem_base, rank_m, rank_p, top, bot, g);
      if (DAG) cluster_sum_post<CS>(pq, wscrB, slotB_addr, rank, P2P ? barB : 0u);   // its CTA barrier publishes Fm inside the CTA
      else __syncthreads();
      if (P2P) {
        if (tid == 32) push_rows_bulk<NX>(fm_saddr, G::OFF_HM, smem_base, barHM, rank_m, rank_p);
      } else {
        cluster_arrive();   // Mp halos and |Mp|^2 partials are on their way
      }
      rr = 0.0;
      double a = 0.0;
      if (DAG) {
        // ---- q = M^dagger Mp (hmc.c:367) is consumed site by site: r -= alpha q, ||r||^2 (hmc.c:371-379).  alpha
        // needs the cluster-wide |Mp|^2, so the first four rows of q are parked until the wait in the middle.
        double2 qh[TT / 2][TX];
        auto consume = [&](int i, int j, const double2 o) {
          r[i][j].x = fma(-a, o.x, r[i][j].x);
          r[i][j].y = fma(-a, o.y, r[i][j].y);
          rr = fma(r[i][j].x, r[i][j].x, rr);
          rr = fma(r[i][j].y, r[i][j].y, rr);
        };
        tile_apply_wt<NX, true, HAS_MU, MASKED>(
            mp, Fm, m_dn, m_up, top, xaddr, t0, g, m, occ, e_m, e_p,
            [&](int i, int j, const double2 o) {
              if (i < TT / 2) qh[i % (TT / 2)][j] = o;
              else consume(i, j, o);
            },
            [&] {
              if (P2P) {
                if (tid == 0) {
                  mbar_expect_tx(barHM, 2u * NX * 16u);
                  mbar_expect_tx(barB, CS * 8u);
                }
                mbar_wait(barHM, par(k));
                mbar_wait(barB, par(k));
              } else {
                cluster_wait();
              }
              a = rr_old / cluster_total<CS>(slotB);   // hmc.c:371
              if (top) tile_fixup_dn<NX, true, HAS_MU>(qh[0], m_dn, xaddr, g, e_p);
#pragma unroll
              for (int i = 0; i < TT / 2; i++)
#pragma unroll
                for (int j = 0; j < TX; j++) consume(i, j, qh[i][j]);
            });
      } else {
        // ---- M~ = M (REF_COMPAT): alpha needs <p, q> of the whole lattice, q is kept
        double2 q[TT][TX];
        tile_apply_wt<NX, false, HAS_MU, MASKED>(
            mp, Fm, m_dn, m_up, top, xaddr, t0, g, m, occ, e_p, e_m, [&](int i, int j, const double2 o) { q[i][j] = o; },
            [&] {
              if (P2P) {
                if (tid == 0) mbar_expect_tx(barHM, 2u * NX * 16u);
                mbar_wait(barHM, par(k));
              } else {
                cluster_wait();
              }
            });
        if (top) tile_fixup_dn<NX, false, HAS_MU>(q[0], m_dn, xaddr, g, e_m);
#pragma unroll
        for (int i = 0; i < TT; i++)
#pragma unroll
          for (int j = 0; j < TX; j++) {   // hmc.c:368-370
            pq = fma(p[i][j].x, q[i][j].x, pq);
            pq = fma(p[i][j].y, q[i][j].y, pq);
          }
        cluster_sum_post<CS>(pq, wscrB, slotB_addr, rank, P2P ? barB : 0u);
        if (P2P) {
          if (tid == 0) mbar_expect_tx(barB, CS * 8u);
          mbar_wait(barB, par(k));
        } else {
          cluster_arrive();
          cluster_wait();
        }
        a = rr_old / cluster_total<CS>(slotB);   // hmc.c:371
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
      cluster_sum_post<CS>(rr, wscrA, slotA_addr, rank, P2P ? barA : 0u);   // its CTA barrier: every local read of Fm has finished
      if (!P2P) cluster_arrive();
      tmem_x_axpy(xaddr + TM_X, p, a, k == 1);   // x += a p (hmc.c:372-373) while the partials cross the cluster
      if (P2P) {
        if (tid == 0) mbar_expect_tx(barA, CS * 8u);
        mbar_wait(barA, par(k));
      } else {
        cluster_wait();
      }
      rr = cluster_total<CS>(slotA);
      iters = k;
      // identical rr on every CTA => the whole cluster leaves the loop together, no barrier half-open
      if (rr < s.accuracy) { status = TB_CG_CONVERGED; break; }                                        // hmc.c:381
      if (!(rr == rr) || rr / rr_init > TB_DIVERGENCE_RATIO) { status = TB_CG_DIVERGED; break; }      // hmc.c:383
      const double be = rr / rr_old;   // hmc.c:390
#pragma unroll
      for (int i = 0; i < TT; i++)
#pragma unroll
        for (int j = 0; j < TX; j++) {
          p[i][j].x = fma(be, p[i][j].x, r[i][j].x);   // hmc.c:391-392
          p[i][j].y = fma(be, p[i][j].y, r[i][j].y);
          Fp[(t0 + i) * NX + j * NGX + g] = p[i][j];   // the last readers of Fp passed the |Mp|^2 barrier
        }
      if (!P2P) push_rows<NX>(p, G::OFF_HP, smem_base, rank_m, rank_p, top, bot, g);
      rr_old = rr;
      __syncthreads();    // p is published inside the CTA
      if (P2P) {          // its halo rows leave if another iteration will read them
        if (tid == 32 && k + 1 < s.max_iter && (!PLAN || k + 1 < sv[2]))
          push_rows_bulk<NX>(fp_saddr, G::OFF_HP, smem_base, barHP, rank_m, rank_p);
      } else {
        cluster_arrive();   // and its halos are on their way; the wait is in the middle of the next stencil
        if (k + 1 >= s.max_iter || (PLAN && k + 1 >= sv[2])) cluster_wait();   // loop ends here: close the barrier
      }
    }
  }
  if (PLAN) c = sv[0];
  if (PLAN && status == TB_CG_MAXITER && sv[2] < s.max_iter) {
    // the head of a split chain ends here: every slab stores r, p, x, rank 0 the two scalars; after a cluster
    // barrier rank 0 raises the chain's flag
    const size_t sb = ((size_t)c * CS + rank) * VL + tid;
#pragma unroll
    for (int ch = 0; ch < 4; ch++) {
      uint32_t v[16];
      tmem_ld16(v, xaddr + TM_X + ch * 16);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int f = ch * 4 + u, i = f / TX, j = f % TX;
        __stcg(&plan.sr[sb + (size_t)f * NTHREADS], r[i][j]);
        __stcg(&plan.sp[sb + (size_t)f * NTHREADS], p[i][j]);
        __stcg(&plan.sx[sb + (size_t)f * NTHREADS],
               make_double2(__hiloint2double((int)v[4 * u + 1], (int)v[4 * u]),
                            __hiloint2double((int)v[4 * u + 3], (int)v[4 * u + 2])));
      }
    }
    if (tid == 0 && rank == 0) {
      s.rr_old[c] = rr_old;
      s.rr_init[c] = rr_init;
    }
    __threadfence();
    cluster_arrive();   // every thread of every slab has its state out (and has read its tensor-memory columns)
    cluster_wait();
    if (tid == 0 && rank == 0) {
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
      const size_t gs = (size_t)(tg0 + t0 + i) * NX + g * TX + j;
      xout[gs * C + c] = (iters > 0) ? make_double2(__hiloint2double((int)v[4 * u + 1], (int)v[4 * u]),
                                                     __hiloint2double((int)v[4 * u + 3], (int)v[4 * u + 2]))
                                      : make_double2(0.0, 0.0);
    }
  }
  __syncthreads();   // every warp has read its columns
  if (tid == 0 && rank == 0) {
    s.status[c] = status;
    s.iters[c] = iters;
    s.rr[c] = rr;
    s.rr_init[c] = rr_init;
    s.active[c] = 0;
    if (PLAN && sv[2] != 0x7fffffff) {   // the chain ended inside its head: the cluster that holds the tail skips it
      __threadfence();
      *(volatile int *)&plan.hand[c] = -1;
    }
  }
  }   // segments
  if (PLAN) __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base_s), "r"(TM_COLS_WT) : "memory");
}

template <int NX, int CS>
struct ClusterLaunch {
  using Kern = void (*)(const double2 *, double2 *, const double2 *, const double2 *, const double *, const double *,
                        const double *, const double *, const TbCgState, const int, const int, const TbPlan);
  static Kern pick(bool dag, bool has_mu, bool masked = false) {
    if (masked) return has_mu ? cluster_cg_kernel<NX, CS, true, true, true, false> : cluster_cg_kernel<NX, CS, true, false, true, false>;
    if (dag) return has_mu ? cluster_cg_kernel<NX, CS, true, true, false, false> : cluster_cg_kernel<NX, CS, true, false, false, false>;
    return has_mu ? cluster_cg_kernel<NX, CS, false, true, false, false> : cluster_cg_kernel<NX, CS, false, false, false, false>;
  }
  // planned launches exist for M~ = M^dagger without an occupation mask (the production mode of the HMC solve)
  static Kern pick_planned(bool has_mu) {
    return has_mu ? cluster_cg_kernel<NX, CS, true, true, false, true> : cluster_cg_kernel<NX, CS, true, false, false, true>;
  }
  static int config(Kern kern, cudaLaunchConfig_t *cfg, cudaLaunchAttribute *attr, int nclusters, cudaStream_t st) {
    TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Slab<NX>::SMEM));
    if (CS > 8) TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = dim3((unsigned)(nclusters * CS), 1, 1);
    cfg->blockDim = dim3(NTHREADS, 1, 1);
    cfg->dynamicSmemBytes = Slab<NX>::SMEM;
    cfg->stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg->attrs = attr;
    cfg->numAttrs = 1;
    return TB_OK;
  }
  // how many clusters of this shape the device can hold at once (0 = cannot be scheduled at all)
  static int max_active(Kern kern) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    if (config(kern, &cfg, attr, 1, nullptr) != TB_OK) return 0;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    return n;
  }
  static int max_active(bool dag, bool has_mu) { return max_active(pick(dag, has_mu)); }
  static int launch(tb_ctx *ctx, const double2 *b, double2 *x, int c0, int n, cudaStream_t st) {
    const bool dag = tb_conj_is_dagger(ctx);
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    TbPlan pl = {};
    // a whole batch that fills its last wave of clusters badly (256^2: 8 chains on 7 clusters of 16) is cut into one
    // equal share of CG iterations per co-resident cluster; the grid of a planned launch is exactly one wave, so the
    // cluster that waits for a hand-over and the cluster that produces it are always on the device together
    if (dag && plan_pays(ctx, c0, n, tb_cluster_capacity(ctx))) {
      Kern kp = pick_planned(ctx->has_mu);
      int cap = max_active(kp);
      if (cap > tb_cluster_capacity(ctx)) cap = tb_cluster_capacity(ctx);
      if (cap > 0 && plan_pays(ctx, c0, n, cap)) {
        TB_CHECK(plan_prepare(ctx, cap, st, &pl));
        TB_CHECK(config(kp, &cfg, attr, cap, st));
        TB_CUDA(cudaLaunchKernelEx(&cfg, kp, b, x, (const double2 *)ctx->W0, (const double2 *)ctx->W1,
                                   (const double *)ctx->d_mass, (const double *)ctx->msite, (const double *)ctx->d_emu,
                                   (const double *)ctx->d_emmu, ctx->cg, ctx->C, 0, pl));
        ctx->launches++;
        return TB_OK;
      }
    }
    Kern kern = pick(dag, ctx->has_mu, ctx->msite != nullptr);
    TB_CHECK(config(kern, &cfg, attr, n, st));
    TB_CUDA(cudaLaunchKernelEx(&cfg, kern, b, x, (const double2 *)ctx->W0, (const double2 *)ctx->W1,
                               (const double *)ctx->d_mass, (const double *)ctx->msite, (const double *)ctx->d_emu,
                               (const double *)ctx->d_emmu, ctx->cg, ctx->C, c0, pl));
    ctx->launches++;
    return TB_OK;
  }
};

// (row length NX, CTAs per cluster): lattices of CS * 4096/NX rows by NX sites
#define TB_CLUSTER_SHAPES(X) X(64, 2) X(64, 4) X(128, 2) X(128, 4) X(128, 8) X(256, 4) X(256, 8) X(256, 16)

}  // namespace

// Clusters of the context's lattice shape that fit on the device at once; 0 when the lattice has no cluster shape
// or the device cannot co-schedule one (then the streaming solver is used).  Cached per context.
int tb_cluster_capacity(tb_ctx *ctx) {
  if (ctx->cluster_capacity >= 0) return ctx->cluster_capacity;
  int cap = 0;
  if (ctx->nranks == 1 && ctx->nx >= 64 && VL % ctx->nx == 0 && ctx->nt % (VL / ctx->nx) == 0) {
    const int cs = ctx->nt / (VL / ctx->nx);
#define X(NX_, CS_) \
  if (ctx->nx == NX_ && cs == CS_) cap = ClusterLaunch<NX_, CS_>::max_active(tb_conj_is_dagger(ctx), ctx->has_mu);
    TB_CLUSTER_SHAPES(X)
#undef X
  }
  ctx->cluster_capacity = cap;
  return cap;
}

bool tb_cluster_supported(tb_ctx *ctx) {
  if (ctx->msite && !tb_conj_is_dagger(ctx)) return false;   // family B (occupation mask) needs M~ = M^T
  return tb_cluster_capacity(ctx) > 0;
}

int tb_run_cg_cluster_slice(tb_ctx *ctx, const double2 *b, double2 *x, int c0, int n, cudaStream_t st) {
  if (b == x) {
    tb_set_error("tb_run_cg_cluster: in-place solve is not supported");
    return TB_EINVAL;
  }
  const int cs = ctx->nx >= 64 && VL % ctx->nx == 0 ? ctx->nt / (VL / ctx->nx) : 0;
#define X(NX_, CS_) \
  if (ctx->nx == NX_ && cs == CS_) return ClusterLaunch<NX_, CS_>::launch(ctx, b, x, c0, n, st);
  TB_CLUSTER_SHAPES(X)
#undef X
  tb_set_error("cluster solver: unsupported lattice %dx%d", ctx->nt, ctx->nx);
  return TB_EINVAL;
}
