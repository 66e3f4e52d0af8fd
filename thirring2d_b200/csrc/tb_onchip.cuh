// tb_onchip.cuh — device helpers shared by the on-chip CG kernels (tb_resident.cu: one CTA per chain,
// tb_cluster.cu: one thread-block cluster per chain).  Include inside an anonymous namespace.
#pragma once

// o += sgn * (w * f)  and  o += sgn * (conj(w) * f), as four FMAs each (operand negation is free in SASS)
template <int SGN>
__device__ __forceinline__ void hop_acc(double2 &o, const double2 w, const double2 f) {
  if (SGN > 0) {
    o.x = fma(w.x, f.x, o.x);  o.x = fma(-w.y, f.y, o.x);
    o.y = fma(w.x, f.y, o.y);  o.y = fma(w.y, f.x, o.y);
  } else {
    o.x = fma(-w.x, f.x, o.x); o.x = fma(w.y, f.y, o.x);
    o.y = fma(-w.x, f.y, o.y); o.y = fma(-w.y, f.x, o.y);
  }
}
template <int SGN>
__device__ __forceinline__ void hopc_acc(double2 &o, const double2 w, const double2 f) {
  if (SGN > 0) {
    o.x = fma(w.x, f.x, o.x);  o.x = fma(w.y, f.y, o.x);
    o.y = fma(w.x, f.y, o.y);  o.y = fma(-w.y, f.x, o.y);
  } else {
    o.x = fma(-w.x, f.x, o.x); o.x = fma(-w.y, f.y, o.x);
    o.y = fma(-w.x, f.y, o.y); o.y = fma(w.y, f.x, o.y);
  }
}

// ---- tensor memory (TMEM) as a thread-private home for the solution vector x ------------------------------
// 256 KB per SM that nothing else on this path uses.  A warp can only reach the 32 TMEM lanes of its own quarter
// (warp % 4); within them every thread owns its lane, so "32x32b" loads/stores are exactly a per-thread scratch
// array: WORDS 32-bit columns per thread, warps that share a quarter stacked along the columns.  x += alpha p then
// costs no LSU wavefronts and no L2 round trip.
__device__ __forceinline__ void tmem_ld16(uint32_t (&v)[16], uint32_t taddr) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32"
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32"
               "[%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
               :
               : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }


// ---- links in tensor memory ------------------------------------------------------------------------------------
// The link fields are read-only during a solve and every thread needs exactly the links of its own tile plus the
// tile's backward halo (W0 of the row above, W1 of the column to the left): 42 double2 for an 8 x 2 tile.  Kept as
// thread-private TMEM columns they cost no shared-memory bandwidth (the LSU data pipe is what bounds the on-chip
// kernels).  Asynchronous variants: issue, compute on something else, then tmem_wait_ld before the first use.
// A row of links (20 words) is fetched as five independent x4 loads: one x16 load needs a 16-register aligned
// landing block, of which ptxas finds only one under this register pressure and then copies every value out of
// it (a quarter of the stencil's instructions were those moves); x4 blocks can live anywhere.
__device__ __forceinline__ void tmem_ld4_at(uint32_t (&v)[20], int k, uint32_t taddr) {   // fills v[4k .. 4k+3]
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(v[4 * k]), "=r"(v[4 * k + 1]), "=r"(v[4 * k + 2]), "=r"(v[4 * k + 3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_async(uint32_t (&v)[20], uint32_t taddr) {
#pragma unroll
  for (int k = 0; k < 4; k++) tmem_ld4_at(v, k, taddr + 4 * k);
}
__device__ __forceinline__ void tmem_ld4_async(uint32_t (&v)[20], uint32_t taddr) { tmem_ld4_at(v, 4, taddr); }
__device__ __forceinline__ void tmem_ld8_async(uint32_t (&v)[8], uint32_t taddr) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
// tcgen05.wait::ld is what PTX requires between a tensor-memory load and the first use of its registers; ptxas
// turns it into scoreboard waits on exactly those registers (no instruction of its own in the SASS), so it is
// stated once per stage without tying the registers to the asm (that forced a copy of every loaded value).
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void tmem_st_d2(uint32_t taddr, const double2 v) {
  tmem_st4(taddr, (uint32_t)__double2loint(v.x), (uint32_t)__double2hiint(v.x), (uint32_t)__double2loint(v.y),
           (uint32_t)__double2hiint(v.y));
}
__device__ __forceinline__ double2 d2_from_words(const uint32_t *w) {
  return make_double2(__hiloint2double((int)w[1], (int)w[0]), __hiloint2double((int)w[3], (int)w[2]));
}

// Per-thread TMEM map of the 8 x 2 tile (32-bit columns): x, then the links.
//   [0, 64)      x(i, j)                       4 words per site, site f = i*2 + j
//   [64, 192)    row i: W0(i,0) W0(i,1) W1(i,0) W1(i,1)      16 words per row
//   [192, 224)   W1(i, x0-1)                   4 words per row
//   [224, 232)   W0(t0-1, x0+j)                4 words per column
// 240 columns per thread (16-aligned); the two warps of a lane quarter stack along the columns: 480 of 512.
constexpr int TM_X = 0, TM_ROW = 64, TM_W1M = 192, TM_W0M = 224, TM_SPAN = 240, TM_COLS_WT = 512;

// m f +- hops on the thread's 8 x 2 tile with the links streamed from tensor memory (software-pipelined one row
// ahead).  F: the field in shared memory (rows of NX sites in sub-plane layout); only the tile's halo is read from
// it: columns x0-1 and x0+2 of the tile's rows, and the rows above and below the tile, given as row pointers
// rowdn (t0-1) and rowup (t0+8) so that a cluster kernel can point them at its halo buffers.  Same arithmetic and
// hop order as tile_apply.
//   skip_dn   the hop from row t0-1 is left out (its data is not there yet); the caller adds it afterwards
//   epi(i, j, value)   receives every finished site (i, j compile-time after unrolling): the caller stores it,
//                      publishes it or folds it into the CG update at once, so no output tile need stay live
//   mid()     runs between rows 3 and 4; rowup is first read after it (a cluster kernel waits for its halos there)
//   MASKED    family B (vec_ops.c:107,130): bit i*2+j of occ set = occupied site = identity row (mass 1; its links
//             are already zero)
template <int NX, bool DAG, bool HAS_MU, bool MASKED, typename Epi, typename Mid>
__device__ __forceinline__ void tile_apply_wt(const double2 (&f)[8][2], const double2 *F, const double2 *rowdn,
                                              const double2 *rowup, bool skip_dn, uint32_t wb, int t0, int g,
                                              double m, uint32_t occ, double af, double ab, Epi epi, Mid mid) {
  constexpr int TX = 2, TT = 8;
  constexpr int SF = DAG ? -1 : 1;
  constexpr int SB = -SF;
  constexpr int NG = NX / TX;
  const int gl = (g + NG - 1) % NG, gr = (g + 1) % NG;
  const double2 zero = make_double2(0.0, 0.0);
  uint32_t h[8], buf[2][20];
  tmem_ld8_async(h, wb + TM_W0M);
  tmem_ld16_async(buf[0], wb + TM_ROW);
  tmem_ld4_async(buf[0], wb + TM_W1M);
  double2 fdn[TX];   // halo of the field while the links are in flight
#pragma unroll
  for (int j = 0; j < TX; j++) fdn[j] = skip_dn ? zero : rowdn[j * NG + g];
  tmem_wait_ld();
  double2 w0m[TX];
#pragma unroll
  for (int j = 0; j < TX; j++) w0m[j] = d2_from_words(&h[4 * j]);
#pragma unroll
  for (int i = 0; i < TT; i++) {
    if (i == TT / 2) mid();
    uint32_t(&cur)[20] = buf[i & 1];
    if (i + 1 < TT) {
      tmem_ld16_async(buf[(i + 1) & 1], wb + TM_ROW + 16 * (i + 1));
      tmem_ld4_async(buf[(i + 1) & 1], wb + TM_W1M + 4 * (i + 1));
    }
    const int row = (t0 + i) * NX;
    const double2 fL = F[row + (TX - 1) * NG + gl];    // f(t, x0-1)
    const double2 fR = F[row + gr];                    // f(t, x0+TX)
    double2 w1m = d2_from_words(&cur[16]);
#pragma unroll
    for (int j = 0; j < TX; j++) {
      const double2 w0c = d2_from_words(&cur[4 * j]);
      const double2 w1c = d2_from_words(&cur[8 + 4 * j]);
      const double2 up = (i == TT - 1) ? rowup[j * NG + g] : f[(i + 1) % TT][j];
      const double2 dn = (i == 0) ? fdn[j] : f[(i + TT - 1) % TT][j];
      const double2 rt = (j == TX - 1) ? fR : f[i][(j + 1) % TX];
      const double2 lf = (j == 0) ? fL : f[i][(j + TX - 1) % TX];
      const double ms = (MASKED && ((occ >> (i * TX + j)) & 1u)) ? 1.0 : m;
      double2 o = make_double2(ms * f[i][j].x, ms * f[i][j].y);   // hmc.c:137-180
      if (HAS_MU) {
        hop_acc<SF>(o, make_double2(w0c.x * af, w0c.y * af), up);
        hopc_acc<SB>(o, make_double2(w0m[j].x * ab, w0m[j].y * ab), dn);
      } else {
        hop_acc<SF>(o, w0c, up);
        hopc_acc<SB>(o, w0m[j], dn);
      }
      hop_acc<SF>(o, w1c, rt);
      hopc_acc<SB>(o, w1m, lf);
      epi(i, j, o);
      w0m[j] = w0c;
      w1m = w1c;
    }
    if (i + 1 < TT) tmem_wait_ld();
  }
}

// the hop tile_apply_wt left out with skip_dn: -+ ab conj(W0(t0-1, x)) f(t0-1, x) into the tile's first row
template <int NX, bool DAG, bool HAS_MU>
__device__ __forceinline__ void tile_fixup_dn(double2 (&row0)[2], const double2 *rowdn, uint32_t wb, int g, double ab) {
  constexpr int SB = DAG ? 1 : -1;
  uint32_t h[8];
  tmem_ld8_async(h, wb + TM_W0M);
  tmem_wait_ld();
#pragma unroll
  for (int j = 0; j < 2; j++) {
    double2 w = d2_from_words(&h[4 * j]);
    if (HAS_MU) w = make_double2(w.x * ab, w.y * ab);
    hopc_acc<SB>(row0[j], w, rowdn[j * (NX / 2) + g]);
  }
}

// x += a p on the thread's tile in tensor memory (hmc.c:372-373); first = x is still the zero start vector
__device__ __forceinline__ void tmem_x_axpy(uint32_t xaddr, const double2 (&p)[8][2], double a, bool first) {
#pragma unroll
  for (int ch = 0; ch < 4; ch++) {
    uint32_t v[16];
    if (!first) tmem_ld16(v, xaddr + ch * 16);
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int f = ch * 4 + u, i = f / 2, j = f % 2;
      double xr = first ? 0.0 : __hiloint2double((int)v[4 * u + 1], (int)v[4 * u]);   // hmc.c:351: x0 = 0
      double xi = first ? 0.0 : __hiloint2double((int)v[4 * u + 3], (int)v[4 * u + 2]);
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
}

// ---- planned launches: balancing a batch that is not a multiple of the SM count --------------------------------------
// One CTA (or cluster) per chain leaves SMs idle in the last wave (256 chains on 148 SMs: 1.73 waves cost 2).  A solve cannot be
// made faster by adding SMs, but it can be PAUSED: r, p (registers) and x (tensor memory) of a chain are 192 KB.
// plan_kernel cuts the concatenated iteration ranges of all chains into one equal share per SM (McNaughton's wrap-
// around rule for preemptive scheduling; the expected iteration counts are those of the previous solve of the
// context): every CTA owns a list of segments (chain, first iteration, end iteration).  At most one chain per CTA is
// split: the CTA that holds its head runs it FIRST (from iteration 1, then stores the state and raises hand[chain]);
// the CTA that holds its tail runs it LAST, after a spin-wait on hand[chain] that normally finds the flag long set.
// The head's CTA has the lower block index, so it is dispatched no later than the CTA that waits for it.  The
// arithmetic of a chain is unaffected (the state round-trips bit for bit): x, the iteration count and the status are
// those of the one-CTA-per-chain launch.
// The wait for a hand-over normally finds the flag set.  It can only last if blocks were not dispatched in index
// order (the head's block has the lower index); after ~10 s the kernel traps rather than hang the device.
#define TB_PLAN_SPIN_CYCLES 20000000000LL
struct TbPlan {
  const int4 *segs;      // (chain, k_begin, k_end, -); k_begin == 1: fresh start; k_end == INT_MAX: to the end
  const int *seg_lo, *seg_hi;   // [gridDim.x] segment range of a CTA
  int *hand;             // [C] 0 = head not finished, k > 0 = state stored, resume at iteration k, -1 = chain finished
  double2 *sr, *sp, *sx; // stored state, [chain][16 tile sites][256 threads]
};

// one thread: wait until the head of a split chain has been stored (k > 0) or the chain has ended (-1).  Not inlined:
// the solver kernels sit at 255 registers and the register allocation of their CG loop must not see this code.
__device__ __noinline__ int plan_wait_hand(const int *hand) {
  int h;
  const long long t_spin = clock64();
  while ((h = *(const volatile int *)hand) == 0) {
    __nanosleep(200);
    if (clock64() - t_spin > TB_PLAN_SPIN_CYCLES) __trap();   // a launch error instead of a hung device
  }
  return h;
}


// The schedule of a planned launch (one thread: C is a few hundred).  est = iteration counts of the context's previous
// solve.  Machines are filled one after the other with T = ceil(sum est / M) iterations each; the chain that straddles
// the boundary between machine j (its end) and j + 1 (its start) is split: head [1, 1 + first) FIRST on machine
// j + 1, tail [1 + first, end) LAST on machine j.  Machine j is CTA M - 1 - j, so the head's CTA has the lower index.
// Without usable estimates (first solve of a context, or a chain that did not converge) chains are dealt out whole,
// round-robin, which is what the hardware does with one CTA per chain.
// Host and device: tb_plan_schedule (tb_resident.cu) runs the same code on the CPU for tests/test_plan_schedule.py.
__host__ __device__ inline void plan_fill(const int *est, int C, int M, long long W, int mx, int4 *segs, int *seg_lo,
                                          int *seg_hi) {
  // A hand-over is not worth fewer than MINS iterations on either side.  A machine may therefore run over its share by
  // a few iterations (a chain that overshoots by less than MINS stays whole; a tail of MINS/2 .. MINS iterations is
  // stretched to MINS) or stay up to MINS/2 - 1 under it, and the share of the machines still to fill is recomputed from
  // the work still to place: nothing piles up on the last machine (149 equal chains on 148 machines: 285 iterations on
  // the busiest one, not 554).
  const int INF = 0x7fffffff, MINS = 8;
  long long left = W;   // iterations not placed yet
  int nseg = 0, j = 0;
  auto share = [&](int jj) {
    const long long t = (left + (M - jj) - 1) / (M - jj);
    return t < mx ? (long long)mx : t;   // >= the longest chain: a head always ends before its tail's turn comes
  };
  long long rem = share(0);
  seg_lo[M - 1] = 0;
  for (int c = 0; c < C; c++) {
    const int n = est[c];
    if (rem < MINS / 2 && j < M - 1) {   // machine j is full (up to 3 iterations short: the later shares absorb them)
      seg_hi[M - 1 - j] = nseg;
      j++;
      seg_lo[M - 1 - j] = nseg;
      rem = share(j);
    }
    if (j == M - 1 || n - rem < MINS || n < 2 * MINS) {   // whole
      segs[nseg++] = make_int4(c, 1, INF, 0);
      rem -= n;
      left -= n;
    } else {
      const long long piece = rem > MINS ? rem : MINS;   // iterations of c that machine j runs: the tail, its last job
      const int first = (int)(n - piece);                // >= MINS; the head: first job of machine j + 1
      segs[nseg++] = make_int4(c, 1 + first, INF, 0);
      left -= piece;
      seg_hi[M - 1 - j] = nseg;
      j++;
      seg_lo[M - 1 - j] = nseg;
      rem = share(j);
      segs[nseg++] = make_int4(c, 1, 1 + first, 0);
      rem -= first;
      left -= first;
    }
  }
  seg_hi[M - 1 - j] = nseg;
  for (j++; j < M; j++) seg_lo[M - 1 - j] = seg_hi[M - 1 - j] = nseg;
}

// chains dealt out whole: CTA b gets chains b, b + M, ...
__host__ __device__ inline void plan_deal(int b, int C, int M, int4 *segs, int *seg_lo, int *seg_hi) {
  const int per = C / M, extra = C % M;
  const int lo = b * per + (b < extra ? b : extra), cnt = per + (b < extra ? 1 : 0);
  seg_lo[b] = lo;
  seg_hi[b] = lo + cnt;
  for (int i = 0; i < cnt; i++) segs[lo + i] = make_int4(b + i * M, 1, 0x7fffffff, 0);
}

__global__ void plan_kernel(const int *__restrict__ est, const int *__restrict__ status, int C, int M, int4 *segs,
                            int *seg_lo, int *seg_hi, int *hand) {
  extern __shared__ int est_s[];   // [C]: one thread walks the chains, out of shared memory
  __shared__ long long W_s;
  __shared__ int mx_s, bad_s;
  if (threadIdx.x == 0) { W_s = 0; mx_s = 0; bad_s = 0; }
  __syncthreads();
  long long w = 0;
  int mx = 0, bad = 0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    hand[c] = 0;
    const int n = est[c];
    est_s[c] = n;
    if (n <= 0 || status[c] != TB_CG_CONVERGED) bad = 1;
    w += n;
    mx = n > mx ? n : mx;
  }
  atomicAdd((unsigned long long *)&W_s, (unsigned long long)w);
  atomicMax(&mx_s, mx);
  if (bad) atomicOr(&bad_s, 1);
  __syncthreads();
  if (bad_s) {
    for (int b = threadIdx.x; b < M; b += blockDim.x) plan_deal(b, C, M, segs, seg_lo, seg_hi);
    return;
  }
  if (threadIdx.x == 0) plan_fill(est_s, C, M, W_s, mx_s, segs, seg_lo, seg_hi);
}


// plan_kernel on the context's stream; M = CTAs (64^2) or clusters (128^2, 256^2) of the planned launch
static int plan_prepare(tb_ctx *ctx, int M, cudaStream_t st, TbPlan *pl) {
  const int C = ctx->C;
  if (!ctx->plan_buf) {
    const int mmax = TB_NUM_SMS_B200 > M ? TB_NUM_SMS_B200 : M;
    TB_CUDA(cudaMalloc((void **)&ctx->plan_buf, ((size_t)(C + mmax) * 4 + 2 * mmax + C) * sizeof(int)));
    ctx->plan_m = mmax;
  }
  if (M > ctx->plan_m) { tb_set_error("plan_prepare: %d machines, buffer holds %d", M, ctx->plan_m); return TB_EINVAL; }
  int4 *segs = (int4 *)ctx->plan_buf;
  int *lo = ctx->plan_buf + (size_t)(C + ctx->plan_m) * 4, *hi = lo + ctx->plan_m, *hand = hi + ctx->plan_m;
  if (C > 12000) TB_CUDA(cudaFuncSetAttribute(plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C * (int)sizeof(int)));
  plan_kernel<<<1, 256, C * sizeof(int), st>>>(ctx->cg.iters, ctx->cg.status, C, M, segs, lo, hi, hand);
  ctx->launches++;
  pl->segs = segs; pl->seg_lo = lo; pl->seg_hi = hi; pl->hand = hand;
  pl->sr = ctx->r; pl->sp = ctx->p; pl->sx = ctx->q;
  return TB_OK;
}

// worth it when the last wave of a plain launch is less than ~3/4 full: a planned launch costs about 4 % (a third
// segment per CTA: links and state in and out once more); measured at 64^2: 256 chains +10 %, 200 +38 %, 296 -4 %,
// 1000 -1 %
static bool plan_pays(const tb_ctx *ctx, int c0, int n, int M) {
  return c0 == 0 && n == ctx->C && n > M && n <= 40000 && !ctx->msite && !getenv("TB_NO_PLAN") &&
         (double)((n + M - 1) / M) * M >= 1.06 * n;
}
