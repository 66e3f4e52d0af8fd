// tb_resident.cu — on-chip resident batched CG for small lattices (16^2, 32^2, 64^2), sm_100a, FP64.
//
// One CTA owns one Markov chain for the WHOLE solve (fmdm_invert_cg, hmc.c:341-404): because chains are
// independent, every CG reduction is a block-level reduction and no grid-wide synchronisation or kernel
// boundary is needed.  The CG state never leaves the SM between iterations:
//
//   registers      r, p (persistent) and Mp, q (transient) for the TS sites of the thread's t-column
//   shared memory  one exchange field F (p, then Mp: what the stencil neighbours read) 16 B/site
//                  the two link fields W0, W1                                           32 B/site
//   L2 (RED.ADD)   x += alpha p, fire-and-forget reductions into a chain-major workspace
//
// 64^2: 48 B/site * 4096 = 192 KB of the 227 KB shared memory, one CTA of 512 threads per SM, 148 chains in
// flight per B200.  HBM is touched only to load b and the links once and to store x once per solve.
//
// A thread owns TS consecutive t-sites of one x column, so the t-neighbours of both stencils are its own
// registers; x-neighbours and the column ends come from F.  Reductions are fixed-shape (shuffle tree, then
// warp partials summed in warp order) => run-to-run deterministic.
#include "tb_common.cuh"

namespace {

__device__ __forceinline__ double2 cmul(const double2 w, const double2 p) {
  return make_double2(w.x * p.x - w.y * p.y, w.x * p.y + w.y * p.x);
}
__device__ __forceinline__ double2 cmulc(const double2 w, const double2 p) {  // conj(w) * p
  return make_double2(w.x * p.x + w.y * p.y, w.x * p.y - w.y * p.x);
}

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

// out[i] = m f[i] +- hops, for the TS sites of this thread's column.  f: own column values (registers),
// F: the same field in shared memory (neighbours), DAG: apply M^dagger instead of M.
template <int NT, int NX, int TS, bool DAG>
__device__ __forceinline__ void column_apply(const double2 (&f)[TS], double2 (&out)[TS], const double2 *F,
                                             const double2 *W0s, const double2 *W1s, int t0, int x,
                                             double m, double af, double ab) {
  const int xp = (x + 1) & (NX - 1), xm = (x - 1) & (NX - 1);
  const int tm = (t0 - 1) & (NT - 1), te = (t0 + TS) & (NT - 1);
  const double2 fU = F[tm * NX + x];   // f(t0-1, x)
  const double2 fD = F[te * NX + x];   // f(t0+TS, x)
  double2 w0m = W0s[tm * NX + x];      // W0(t-1, x), slides down the column
#pragma unroll
  for (int i = 0; i < TS; i++) {
    const int t = t0 + i;
    const double2 w0c = W0s[t * NX + x];
    const double2 w1c = W1s[t * NX + x];
    const double2 w1m = W1s[t * NX + xm];
    const double2 fxp = F[t * NX + xp];
    const double2 fxm = F[t * NX + xm];
    const double2 up = (i == TS - 1) ? fD : f[(i + 1) % TS];
    const double2 dn = (i == 0) ? fU : f[(i + TS - 1) % TS];
    // +af W0(n) f(n+t) - ab conj(W0(n-t)) f(n-t) + W1(n) f(n+x) - conj(W1(n-x)) f(n-x)   (hmc.c:140-180)
    const double2 a = cmul(make_double2(w0c.x * af, w0c.y * af), up);
    const double2 b = cmulc(make_double2(w0m.x * ab, w0m.y * ab), dn);
    const double2 c = cmul(w1c, fxp);
    const double2 d = cmulc(w1m, fxm);
    const double hr = (a.x - b.x) + (c.x - d.x);
    const double hi = (a.y - b.y) + (c.y - d.y);
    if (DAG) out[i] = make_double2(m * f[i].x - hr, m * f[i].y - hi);
    else out[i] = make_double2(m * f[i].x + hr, m * f[i].y + hi);
    w0m = w0c;
  }
}

template <int NT, int NX, int TS, bool DAG>
__global__ void __launch_bounds__((NT / TS) * NX, (512 / ((NT / TS) * NX)) > 0 ? (512 / ((NT / TS) * NX)) : 1)
resident_cg_kernel(const double2 *__restrict__ bsrc, double2 *__restrict__ xout,
                   const double2 *__restrict__ W0g, const double2 *__restrict__ W1g,
                   const double *__restrict__ mass, const double *__restrict__ emu,
                   const double *__restrict__ emmu, double2 *__restrict__ xw, const TbCgState s, const int C) {
  constexpr int V = NT * NX;
  constexpr int NTHREADS = (NT / TS) * NX;
  constexpr int NWARPS = NTHREADS / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *F = reinterpret_cast<double2 *>(smem_raw);
  double2 *W0s = F + V;
  double2 *W1s = W0s + V;
  double *scrA = reinterpret_cast<double *>(W1s + V);
  double *scrB = scrA + 32;

  const int c = blockIdx.x;
  const int tid = threadIdx.x;
  const int x = tid % NX;
  const int t0 = (tid / NX) * TS;
  const double m = mass[c];
  const double e_p = emu[c], e_m = emmu[c];

  // links and source: device layout [site][chain] -> on-chip
  for (int k = tid; k < V; k += NTHREADS) {
    W0s[k] = W0g[(size_t)k * C + c];
    W1s[k] = W1g[(size_t)k * C + c];
  }
  double2 r[TS], p[TS];
  double rr = 0.0;
#pragma unroll
  for (int i = 0; i < TS; i++) {
    const int k = (t0 + i) * NX + x;
    r[i] = bsrc[(size_t)k * C + c];
    p[i] = r[i];
    rr += r[i].x * r[i].x + r[i].y * r[i].y;
    F[k] = p[i];
    xw[(size_t)c * V + k] = make_double2(0.0, 0.0);  // hmc.c:351
  }
  rr = block_sum<NWARPS>(rr, scrA);   // hmc.c:354-356
  const double rr_init = rr;
  double rr_old = rr;
  int status = TB_CG_MAXITER, iters = 0;
  __syncthreads();  // F, links and the zeroed workspace are in place

  if (rr_old < s.accuracy) {  // hmc.c:359-361
    status = TB_CG_ZERO_SOURCE;
  } else {
    for (int k = 1; k < s.max_iter; k++) {  // hmc.c:364
      double2 mp[TS], q[TS];
      column_apply<NT, NX, TS, false>(p, mp, F, W0s, W1s, t0, x, m, e_p, e_m);   // Mp = M p, hmc.c:366
      __syncthreads();  // everyone has read p from F
#pragma unroll
      for (int i = 0; i < TS; i++) F[(t0 + i) * NX + x] = mp[i];
      __syncthreads();
      // q = M~ Mp, hmc.c:367 (M^dagger swaps the roles of e^{mu} and e^{-mu})
      column_apply<NT, NX, TS, DAG>(mp, q, F, W0s, W1s, t0, x, m, DAG ? e_m : e_p, DAG ? e_p : e_m);
      double pq = 0.0;
#pragma unroll
      for (int i = 0; i < TS; i++) pq += p[i].x * q[i].x + p[i].y * q[i].y;   // hmc.c:368-370
      pq = block_sum<NWARPS>(pq, scrB);
      const double a = rr_old / pq;   // hmc.c:371
      rr = 0.0;
#pragma unroll
      for (int i = 0; i < TS; i++) {
        double2 *xk = &xw[(size_t)c * V + (t0 + i) * NX + x];
        atomicAdd(&xk->x, a * p[i].x);   // x += a p, hmc.c:372-373 (RED.ADD at L2, one writer per address)
        atomicAdd(&xk->y, a * p[i].y);
        r[i].x -= a * q[i].x;            // hmc.c:374-375
        r[i].y -= a * q[i].y;
        rr += r[i].x * r[i].x + r[i].y * r[i].y;   // hmc.c:377-379
      }
      rr = block_sum<NWARPS>(rr, scrA);
      iters = k;
      if (rr < s.accuracy) { status = TB_CG_CONVERGED; break; }                                        // hmc.c:381
      if (!(rr == rr) || rr / rr_init > TB_DIVERGENCE_RATIO) { status = TB_CG_DIVERGED; break; }      // hmc.c:383
      const double be = rr / rr_old;   // hmc.c:390
#pragma unroll
      for (int i = 0; i < TS; i++) {
        p[i].x = r[i].x + be * p[i].x;   // hmc.c:391-392
        p[i].y = r[i].y + be * p[i].y;
        F[(t0 + i) * NX + x] = p[i];     // all reads of Mp finished before the pq reduction's barrier
      }
      rr_old = rr;
      __syncthreads();
    }
  }
  // the RED.ADDs of this CTA must have landed before x is read back
  __threadfence();
  __syncthreads();
#pragma unroll
  for (int i = 0; i < TS; i++) {
    const int k = (t0 + i) * NX + x;
    xout[(size_t)k * C + c] = __ldcg(&xw[(size_t)c * V + k]);
  }
  if (tid == 0) {
    s.status[c] = status;
    s.iters[c] = iters;
    s.rr[c] = rr;
    s.rr_init[c] = rr_init;
    s.active[c] = 0;
  }
}

template <int NT, int NX, int TS>
int launch_resident(tb_ctx *ctx, const double2 *b, double2 *x) {
  constexpr int V = NT * NX;
  constexpr int NTHREADS = (NT / TS) * NX;
  const size_t smem = (size_t)V * 48 + 64 * sizeof(double);
  const bool dag = tb_conj_is_dagger(ctx);
  auto kern = dag ? resident_cg_kernel<NT, NX, TS, true> : resident_cg_kernel<NT, NX, TS, false>;
  TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<ctx->C, NTHREADS, smem, ctx->stream>>>(b, x, ctx->W0, ctx->W1, ctx->d_mass, ctx->d_emu, ctx->d_emmu,
                                                ctx->xw, ctx->cg, ctx->C);
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

}  // namespace

bool tb_resident_supported(const tb_ctx *ctx) {
  return ctx->nt == ctx->nx && (ctx->nt == 16 || ctx->nt == 32 || ctx->nt == 64);
}

// One kernel launch per solve.
int tb_run_cg_resident(tb_ctx *ctx, const double2 *b, double2 *x) {
  if (b == x) {
    tb_set_error("tb_run_cg_resident: in-place solve is not supported");
    return TB_EINVAL;
  }
  switch (ctx->nt) {
    case 16: return launch_resident<16, 16, 8>(ctx, b, x);
    case 32: return launch_resident<32, 32, 8>(ctx, b, x);
    case 64: return launch_resident<64, 64, 8>(ctx, b, x);
    default: tb_set_error("resident solver: unsupported lattice %dx%d", ctx->nt, ctx->nx); return TB_EINVAL;
  }
}
