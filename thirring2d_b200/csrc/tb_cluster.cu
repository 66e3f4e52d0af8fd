// tb_cluster.cu — on-chip resident batched CG for lattices that do not fit one SM: one THREAD-BLOCK CLUSTER per
// Markov chain (128^2 = 2x2 CTAs, 256^2 = 4x4 CTAs, and the rectangles in between), sm_100a, FP64.
//
// Same algorithm and per-thread arithmetic as tb_resident.cu (fmdm_invert_cg, hmc.c:341-404), with the lattice
// cut into 64 x 64 sub-lattices, one per CTA of the cluster.  The CG state of a sub-lattice never leaves its SM:
//
//   registers      r, p (persistent), Mp, q (transient): 8 x 2 sites per thread, 256 threads
//   shared memory  exchange field F (p, then Mp), links W0, W1 of the sub-lattice       48 B/site = 192 KB
//                  halo rows/columns of p and of Mp, pushed by the four neighbour CTAs  2 x 4 x 64 x 16 B
//                  halo row of W0 and halo column of W1 (constant during the solve)     2 x 64 x 16 B
//   tensor memory  x, thread-private columns (tcgen05.ld/st 32x32b)
//
// Distributed shared memory carries everything that crosses a CTA boundary:
//   * halos are PUSHED: the thread that produces a boundary site of p or Mp also stores it into the neighbour
//     CTA's halo buffer (st.shared::cluster); the stencil itself only ever reads local shared memory;
//   * both CG reductions are all-to-all over the cluster: every CTA stores its block sum into every CTA's slot
//     table, and all CTAs add the slots in rank order => bitwise identical alpha, beta and stopping decisions on
//     every CTA with no further communication (and run-to-run deterministic);
//   * barrier.cluster (arrive.release / wait.acquire) publishes both; the x += alpha p update in tensor memory
//     runs between the arrive and the wait of the ||r||^2 reduction.
// Three cluster barriers and one CTA barrier per CG iteration.  HBM is touched once per solve.
#include <cstdint>

#include "tb_common.cuh"

namespace {

#include "tb_onchip.cuh"

constexpr int LT = 64, LX = 64;            // sub-lattice of one CTA
constexpr int TX = 2, TT = 8;              // sites per thread: TT rows (t) x TX columns (x)
constexpr int NG = LX / TX;                // x-groups per row = lanes of a warp
constexpr int NTHREADS = (LT / TT) * NG;   // 256
constexpr int NWARPS = NTHREADS / 32;      // 8
constexpr int VL = LT * LX;
static_assert(NG == 32, "a warp is one row of tiles");

// shared-memory map, in double2 units from the start of dynamic shared memory
constexpr int OFF_F = 0, OFF_W0 = VL, OFF_W1 = 2 * VL;
constexpr int OFF_HP = 3 * VL;             // halos of p : [dn | up | left | right], 64 each
constexpr int OFF_HM = OFF_HP + 4 * 64;    // halos of Mp
constexpr int OFF_W0H = OFF_HM + 4 * 64;   // W0(t = -1, x)   row layout
constexpr int OFF_W1H = OFF_W0H + 64;      // W1(t, x = -1)   indexed by t
constexpr int OFF_END = OFF_W1H + 64;
constexpr int H_DN = 0, H_UP = 64, H_L = 128, H_R = 192;
// doubles after OFF_END: warp partials A, B [NWARPS each], cluster slots A, B [16 each]
constexpr size_t CL_SMEM = (size_t)OFF_END * sizeof(double2) + (2 * NWARPS + 32) * sizeof(double);
constexpr int TMEM_WORDS = TX * TT * 4;                    // 64 words of x per thread
constexpr int TMEM_COLS = TMEM_WORDS * (NWARPS / 4);       // 128 columns

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

// Block sum -> every CTA's slot table.  After the caller's cluster barrier, cluster_total adds the CS slots.
template <int CS>
__device__ __forceinline__ void cluster_sum_post(double v, double *wscr, uint32_t slots_saddr, uint32_t my_rank) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) wscr[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NWARPS; w++) s += wscr[w];   // warp order, like block_sum of tb_resident.cu
    if (threadIdx.x < CS) st_cluster(mapa_shared(slots_saddr + my_rank * 8u, threadIdx.x), s);
  }
}
template <int CS>
__device__ __forceinline__ double cluster_total(const double *slots) {
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < CS; q++) s += slots[q];
  return s;
}

// The thread's boundary sites of v -> the halo buffers (set HSET) of the four neighbour CTAs.
struct Nbr {
  uint32_t tm, tp, xm, xp;   // cluster ranks of the CTAs at (ct-1), (ct+1), (cx-1), (cx+1), periodic
};
__device__ __forceinline__ void push_halos(const double2 (&v)[TT][TX], int hset, uint32_t smem_base, const Nbr nb,
                                           int t0, int g) {
  const uint32_t hb = smem_base + (uint32_t)hset * 16u;
  if (t0 == 0) {             // my first row is the (t+1) halo of the CTA above
    const uint32_t a = mapa_shared(hb + (H_UP + g) * 16u, nb.tm);
#pragma unroll
    for (int j = 0; j < TX; j++) st_cluster(a + j * NG * 16u, v[0][j]);
  }
  if (t0 + TT == LT) {       // my last row is the (t-1) halo of the CTA below
    const uint32_t a = mapa_shared(hb + (H_DN + g) * 16u, nb.tp);
#pragma unroll
    for (int j = 0; j < TX; j++) st_cluster(a + j * NG * 16u, v[TT - 1][j]);
  }
  if (g == 0) {              // my first column is the (x+1) halo of the CTA to the left
    const uint32_t a = mapa_shared(hb + (H_R + t0) * 16u, nb.xm);
#pragma unroll
    for (int i = 0; i < TT; i++) st_cluster(a + i * 16u, v[i][0]);
  }
  if (g == NG - 1) {         // my last column is the (x-1) halo of the CTA to the right
    const uint32_t a = mapa_shared(hb + (H_L + t0) * 16u, nb.xp);
#pragma unroll
    for (int i = 0; i < TT; i++) st_cluster(a + i * 16u, v[i][TX - 1]);
  }
}

// out = m f +- hops on the thread's tile of the 64 x 64 sub-lattice; S = shared memory as double2[], hset = the
// halo set that belongs to the field in F.  Same arithmetic and order as tile_apply of tb_resident.cu.
template <bool DAG, bool HAS_MU>
__device__ __forceinline__ void tile_apply_cl(const double2 (&f)[TT][TX], double2 (&out)[TT][TX], const double2 *S,
                                              int hset, int t0, int g, double m, double af, double ab) {
  constexpr int SF = DAG ? -1 : 1;
  constexpr int SB = -SF;
  // rows t0-1 and t0+TT of the field, and row t0-1 of W0: own F / W0 or the halo rows (warp-uniform selects)
  const int rowm = (t0 == 0) ? hset + H_DN : OFF_F + (t0 - 1) * LX;
  const int rowe = (t0 + TT == LT) ? hset + H_UP : OFF_F + (t0 + TT) * LX;
  const int w0rm = (t0 == 0) ? OFF_W0H : OFF_W0 + (t0 - 1) * LX;
  // columns x0-1 and x0+TX: own F (stride LX) or the halo columns (stride 1); same for W1(t, x0-1)
  const int colL = (g == 0) ? hset + H_L + t0 : OFF_F + t0 * LX + (TX - 1) * NG + g - 1;
  const int colR = (g == NG - 1) ? hset + H_R + t0 : OFF_F + t0 * LX + g + 1;
  const int w1L = (g == 0) ? OFF_W1H + t0 : OFF_W1 + t0 * LX + (TX - 1) * NG + g - 1;
  const int sL = (g == 0) ? 1 : LX, sR = (g == NG - 1) ? 1 : LX;
  double2 w0m[TX];
#pragma unroll
  for (int j = 0; j < TX; j++) w0m[j] = S[w0rm + j * NG + g];
#pragma unroll
  for (int i = 0; i < TT; i++) {
    const int row = (t0 + i) * LX;
    const double2 fL = S[colL + i * sL];
    const double2 fR = S[colR + i * sR];
    double2 w1m = S[w1L + i * sL];
#pragma unroll
    for (int j = 0; j < TX; j++) {
      const double2 w0c = S[OFF_W0 + row + j * NG + g];
      const double2 w1c = S[OFF_W1 + row + j * NG + g];
      const double2 up = (i == TT - 1) ? S[rowe + j * NG + g] : f[(i + 1) % TT][j];
      const double2 dn = (i == 0) ? S[rowm + j * NG + g] : f[(i + TT - 1) % TT][j];
      const double2 rt = (j == TX - 1) ? fR : f[i][(j + 1) % TX];
      const double2 lf = (j == 0) ? fL : f[i][(j + TX - 1) % TX];
      double2 o = make_double2(m * f[i][j].x, m * f[i][j].y);   // hmc.c:137-180
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

template <int CT, int CX, bool DAG, bool HAS_MU>
__global__ void __launch_bounds__(NTHREADS, 1)
cluster_cg_kernel(const double2 *__restrict__ bsrc, double2 *__restrict__ xout, const double2 *__restrict__ W0g,
                  const double2 *__restrict__ W1g, const double *__restrict__ mass, const double *__restrict__ emu,
                  const double *__restrict__ emmu, const TbCgState s, const int C, const int c_first) {
  constexpr int CS = CT * CX, NT = CT * LT, NX = CX * LX;
  static_assert(CS <= 16, "slot tables hold 16 ranks");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *S = reinterpret_cast<double2 *>(smem_raw);
  double *wscrA = reinterpret_cast<double *>(S + OFF_END);
  double *wscrB = wscrA + NWARPS;
  double *slotA = wscrB + NWARPS;
  double *slotB = slotA + 16;
  __shared__ uint32_t tmem_base_s;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t slotA_addr = (uint32_t)__cvta_generic_to_shared(slotA);
  const uint32_t slotB_addr = (uint32_t)__cvta_generic_to_shared(slotB);

  if (threadIdx.x < 32) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t xaddr = tmem_base_s + (((warp & 3u) * 32u) << 16) + (warp >> 2) * (uint32_t)TMEM_WORDS;

  const uint32_t rank = cluster_ctarank();
  const int ct = (int)rank / CX, cx = (int)rank % CX;
  Nbr nb;
  nb.tm = (uint32_t)(((ct + CT - 1) % CT) * CX + cx);
  nb.tp = (uint32_t)(((ct + 1) % CT) * CX + cx);
  nb.xm = (uint32_t)(ct * CX + (cx + CX - 1) % CX);
  nb.xp = (uint32_t)(ct * CX + (cx + 1) % CX);
  const int c = c_first + (int)(blockIdx.x / CS);
  const int tid = threadIdx.x;
  const int g = tid % NG;
  const int t0 = (tid / NG) * TT;
  const double m = mass[c];
  const double e_p = emu[c], e_m = emmu[c];
  const int tg0 = ct * LT, xg0 = cx * LX;   // global origin of the sub-lattice

  // links of the sub-lattice and their halos: device layout [site][chain] -> shared memory
  for (int k = tid; k < VL; k += NTHREADS) {
    const int t = k / LX, x = k % LX;
    const size_t gs = (size_t)(tg0 + t) * NX + xg0 + x;
    const int ks = t * LX + (x % TX) * NG + x / TX;
    S[OFF_W0 + ks] = W0g[gs * C + c];
    S[OFF_W1 + ks] = W1g[gs * C + c];
  }
  if (tid < LX) {
    const int x = tid, tgm = (tg0 + NT - 1) % NT;
    S[OFF_W0H + (x % TX) * NG + x / TX] = W0g[((size_t)tgm * NX + xg0 + x) * C + c];
  } else if (tid < LX + LT) {
    const int t = tid - LX, xgm = (xg0 + NX - 1) % NX;
    S[OFF_W1H + t] = W1g[((size_t)(tg0 + t) * NX + xgm) * C + c];
  }
  double2 r[TT][TX], p[TT][TX];
  double rr = 0.0;
#pragma unroll
  for (int i = 0; i < TT; i++)
#pragma unroll
    for (int j = 0; j < TX; j++) {
      const size_t gs = (size_t)(tg0 + t0 + i) * NX + xg0 + g * TX + j;
      r[i][j] = bsrc[gs * C + c];
      p[i][j] = r[i][j];
      rr = fma(r[i][j].x, r[i][j].x, rr);
      rr = fma(r[i][j].y, r[i][j].y, rr);
      S[OFF_F + (t0 + i) * LX + j * NG + g] = p[i][j];
    }
  // every CTA of the cluster is running before anyone stores into a peer's shared memory
  cluster_arrive();
  cluster_wait();
  push_halos(p, OFF_HP, smem_base, nb, t0, g);
  cluster_sum_post<CS>(rr, wscrA, slotA_addr, rank);   // hmc.c:354-356
  cluster_arrive();
  cluster_wait();
  rr = cluster_total<CS>(slotA);
  const double rr_init = rr;
  double rr_old = rr;
  int status = TB_CG_MAXITER, iters = 0;

  if (rr_old < s.accuracy) {  // hmc.c:359-361
    status = TB_CG_ZERO_SOURCE;
  } else {
    for (int k = 1; k < s.max_iter; k++) {  // hmc.c:364
      double2 mp[TT][TX], q[TT][TX];
      tile_apply_cl<false, HAS_MU>(p, mp, S, OFF_HP, t0, g, m, e_p, e_m);   // Mp = M p, hmc.c:366
      double pq = 0.0;
      if (DAG) {   // <p, M^dagger M p> = |M p|^2
#pragma unroll
        for (int i = 0; i < TT; i++)
#pragma unroll
          for (int j = 0; j < TX; j++) {
            pq = fma(mp[i][j].x, mp[i][j].x, pq);
            pq = fma(mp[i][j].y, mp[i][j].y, pq);
          }
      }
      __syncthreads();  // this CTA has read p from F (peers read only their own halo copies)
#pragma unroll
      for (int i = 0; i < TT; i++)
#pragma unroll
        for (int j = 0; j < TX; j++) S[OFF_F + (t0 + i) * LX + j * NG + g] = mp[i][j];
      push_halos(mp, OFF_HM, smem_base, nb, t0, g);
      if (DAG) cluster_sum_post<CS>(pq, wscrB, slotB_addr, rank);
      cluster_arrive();   // publishes Mp, its halos and the |Mp|^2 partials
      cluster_wait();
      if (DAG) pq = cluster_total<CS>(slotB);
      // q = M~ Mp, hmc.c:367
      tile_apply_cl<DAG, HAS_MU>(mp, q, S, OFF_HM, t0, g, m, DAG ? e_m : e_p, DAG ? e_p : e_m);
      if (!DAG) {
#pragma unroll
        for (int i = 0; i < TT; i++)
#pragma unroll
          for (int j = 0; j < TX; j++) {   // hmc.c:368-370
            pq = fma(p[i][j].x, q[i][j].x, pq);
            pq = fma(p[i][j].y, q[i][j].y, pq);
          }
        cluster_sum_post<CS>(pq, wscrB, slotB_addr, rank);
        cluster_arrive();
        cluster_wait();
        pq = cluster_total<CS>(slotB);
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
      cluster_sum_post<CS>(rr, wscrA, slotA_addr, rank);
      cluster_arrive();
      // x += a p (hmc.c:372-373) in tensor memory while the ||r||^2 partials cross the cluster
#pragma unroll
      for (int ch = 0; ch < TMEM_WORDS / 16; ch++) {
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
      cluster_wait();
      rr = cluster_total<CS>(slotA);
      iters = k;
      // identical rr on every CTA => the whole cluster leaves the loop together
      if (rr < s.accuracy) { status = TB_CG_CONVERGED; break; }                                        // hmc.c:381
      if (!(rr == rr) || rr / rr_init > TB_DIVERGENCE_RATIO) { status = TB_CG_DIVERGED; break; }      // hmc.c:383
      const double be = rr / rr_old;   // hmc.c:390
#pragma unroll
      for (int i = 0; i < TT; i++)
#pragma unroll
        for (int j = 0; j < TX; j++) {
          p[i][j].x = fma(be, p[i][j].x, r[i][j].x);   // hmc.c:391-392
          p[i][j].y = fma(be, p[i][j].y, r[i][j].y);
          S[OFF_F + (t0 + i) * LX + j * NG + g] = p[i][j];   // local reads of Mp ended before the ||r||^2 barrier
        }
      push_halos(p, OFF_HP, smem_base, nb, t0, g);
      rr_old = rr;
      cluster_arrive();   // publishes p and its halos
      cluster_wait();
    }
  }
#pragma unroll
  for (int ch = 0; ch < TMEM_WORDS / 16; ch++) {
    uint32_t v[16];
    if (iters > 0) tmem_ld16(v, xaddr + ch * 16);
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int f = ch * 4 + u, i = f / TX, j = f % TX;
      const size_t gs = (size_t)(tg0 + t0 + i) * NX + xg0 + g * TX + j;
      xout[gs * C + c] = (iters > 0) ? make_double2(__hiloint2double((int)v[4 * u + 1], (int)v[4 * u]),
                                                     __hiloint2double((int)v[4 * u + 3], (int)v[4 * u + 2]))
                                      : make_double2(0.0, 0.0);
    }
  }
  __syncthreads();   // every warp has read its columns
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base_s), "r"(TMEM_COLS) : "memory");
  if (tid == 0 && rank == 0) {
    s.status[c] = status;
    s.iters[c] = iters;
    s.rr[c] = rr;
    s.rr_init[c] = rr_init;
    s.active[c] = 0;
  }
}

template <int CT, int CX>
struct ClusterLaunch {
  using Kern = void (*)(const double2 *, double2 *, const double2 *, const double2 *, const double *, const double *,
                        const double *, const TbCgState, const int, const int);
  static Kern pick(bool dag, bool has_mu) {
    if (dag) return has_mu ? cluster_cg_kernel<CT, CX, true, true> : cluster_cg_kernel<CT, CX, true, false>;
    return has_mu ? cluster_cg_kernel<CT, CX, false, true> : cluster_cg_kernel<CT, CX, false, false>;
  }
  static int config(Kern kern, cudaLaunchConfig_t *cfg, cudaLaunchAttribute *attr, int nclusters, cudaStream_t st) {
    constexpr int CS = CT * CX;
    TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CL_SMEM));
    if (CS > 8) TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    memset(cfg, 0, sizeof(*cfg));
    cfg->gridDim = dim3((unsigned)(nclusters * CS), 1, 1);
    cfg->blockDim = dim3(NTHREADS, 1, 1);
    cfg->dynamicSmemBytes = CL_SMEM;
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
  static int max_active(bool dag, bool has_mu) {
    Kern kern = pick(dag, has_mu);
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
  static int launch(tb_ctx *ctx, const double2 *b, double2 *x, int c0, int n, cudaStream_t st) {
    Kern kern = pick(tb_conj_is_dagger(ctx), ctx->has_mu);
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    TB_CHECK(config(kern, &cfg, attr, n, st));
    TB_CUDA(cudaLaunchKernelEx(&cfg, kern, b, x, (const double2 *)ctx->W0, (const double2 *)ctx->W1,
                               (const double *)ctx->d_mass, (const double *)ctx->d_emu, (const double *)ctx->d_emmu,
                               ctx->cg, ctx->C, c0));
    ctx->launches++;
    return TB_OK;
  }
};

// lattice -> cluster shape (CT x CX sub-lattices of 64 x 64)
#define TB_CLUSTER_SHAPES(X) X(1, 2) X(2, 1) X(2, 2) X(2, 4) X(4, 2) X(4, 4)

}  // namespace

// Clusters of the context's lattice shape that fit on the device at once; 0 when the lattice has no cluster shape
// or the device cannot co-schedule one (then the streaming solver is used).  Cached per context.
int tb_cluster_capacity(tb_ctx *ctx) {
  if (ctx->cluster_capacity >= 0) return ctx->cluster_capacity;
  int cap = 0;
  if (ctx->nranks == 1 && ctx->msite == nullptr && ctx->nt % LT == 0 && ctx->nx % LX == 0) {
    const int ctn = ctx->nt / LT, cxn = ctx->nx / LX;
#define X(CT_, CX_) \
  if (ctn == CT_ && cxn == CX_) cap = ClusterLaunch<CT_, CX_>::max_active(tb_conj_is_dagger(ctx), ctx->has_mu);
    TB_CLUSTER_SHAPES(X)
#undef X
  }
  ctx->cluster_capacity = cap;
  return cap;
}

bool tb_cluster_supported(tb_ctx *ctx) { return ctx->msite == nullptr && tb_cluster_capacity(ctx) > 0; }

int tb_run_cg_cluster_slice(tb_ctx *ctx, const double2 *b, double2 *x, int c0, int n, cudaStream_t st) {
  if (b == x) {
    tb_set_error("tb_run_cg_cluster: in-place solve is not supported");
    return TB_EINVAL;
  }
  const int ctn = ctx->nt / LT, cxn = ctx->nx / LX;
#define X(CT_, CX_) \
  if (ctn == CT_ && cxn == CX_) return ClusterLaunch<CT_, CX_>::launch(ctx, b, x, c0, n, st);
  TB_CLUSTER_SHAPES(X)
#undef X
  tb_set_error("cluster solver: unsupported lattice %dx%d", ctx->nt, ctx->nx);
  return TB_EINVAL;
}

int tb_run_cg_cluster(tb_ctx *ctx, const double2 *b, double2 *x);
int tb_run_cg_cluster(tb_ctx *ctx, const double2 *b, double2 *x) {
  return tb_run_cg_cluster_slice(ctx, b, x, 0, ctx->C, ctx->stream);
}
