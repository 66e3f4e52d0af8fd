// tb_strict.cu — the reference's fmdm_invert_cg (hmc.c:341-404) in the reference's OWN floating-point evaluation order,
// on the GPU, for parity work (tb_set_tuning solver = 5).  Compiled with -fmad=false (see the Makefile): no product is
// contracted into an FMA, as gcc does for the reference under -std=c99.
//
// The fast solvers differ from the reference's arithmetic in two ways that cannot change a result beyond rounding but do
// move the iteration at which ||r||^2 crosses 1e-30 in a 700+ iteration solve by 1-3: FMA contraction, and dot products
// summed as trees instead of sequentially.  This solver removes both:
//   * per-site arithmetic in the order hmc.c:137-180 evaluates it: v = m c, then the four hops added one after the other
//     (+t, -t, +x, -x), each hop as (link * e^{+-mu}) * psi with separately rounded products;
//   * every sum over the lattice (hmc.c:354-356, 368-370, 377-379) is accumulated by ONE thread in (t, x) order.
// With the links' cos / sin supplied by the host's libm (tb_set_links_trig) every bit of the recursion is the
// reference's: same iteration count, same residuals, same solution.  With the device's sincos (tb_set_gauge*) a link may
// differ from glibc's in its last bit.
// One CTA per chain; vectors stay in the device layout [site][chain].  It is slow by design (two sequential sums of
// NT*NX terms per iteration) and is not used unless asked for.
#include "tb_common.cuh"

namespace {

// out = M in (dagger = false) or M^dagger in, restating hmc.c:132-183 / the adjoint of it.  W0, W1 carry the factor
// s * 1/2 * eta (exact scalings by +-1/2, so (W * e) equals the reference's 0.5 * cos * eta * e bit for bit).
__device__ __forceinline__ void strict_apply(const double2 *__restrict__ in, double2 *__restrict__ out,
                                             const double2 *__restrict__ W0, const double2 *__restrict__ W1, int nt,
                                             int nx, int C, int c, double m, double e_fwd, double e_bwd, bool dagger) {
  const int V = nt * nx;
  for (int k = threadIdx.x; k < V; k += blockDim.x) {
    const int t = k / nx, x = k - t * nx;
    const int tp = (t + 1 == nt) ? 0 : t + 1, tm = (t == 0) ? nt - 1 : t - 1;
    const int xp = (x + 1 == nx) ? 0 : x + 1, xm = (x == 0) ? nx - 1 : x - 1;
    const double2 cv = in[(size_t)k * C + c];
    double vr = m * cv.x, vi = m * cv.y;   // hmc.c:137
    double lr, li, pr, pi;
    double2 w, psi;
    // +t, hmc.c:140-148
    w = W0[(size_t)k * C + c];
    psi = in[((size_t)tp * nx + x) * C + c];
    lr = w.x * e_fwd;  li = w.y * e_fwd;
    pr = lr * psi.x - li * psi.y;
    pi = lr * psi.y + li * psi.x;
    if (dagger) { vr -= pr; vi -= pi; } else { vr += pr; vi += pi; }
    // -t, hmc.c:151-159: conjugate link of the site below
    w = W0[((size_t)tm * nx + x) * C + c];
    psi = in[((size_t)tm * nx + x) * C + c];
    lr = w.x * e_bwd;  li = w.y * e_bwd;
    pr = lr * psi.x + li * psi.y;
    pi = lr * psi.y - li * psi.x;
    if (dagger) { vr += pr; vi += pi; } else { vr -= pr; vi -= pi; }
    // +x, hmc.c:162-170
    w = W1[(size_t)k * C + c];
    psi = in[((size_t)t * nx + xp) * C + c];
    pr = w.x * psi.x - w.y * psi.y;
    pi = w.x * psi.y + w.y * psi.x;
    if (dagger) { vr -= pr; vi -= pi; } else { vr += pr; vi += pi; }
    // -x, hmc.c:172-180
    w = W1[((size_t)t * nx + xm) * C + c];
    psi = in[((size_t)t * nx + xm) * C + c];
    pr = w.x * psi.x + w.y * psi.y;
    pi = w.x * psi.y - w.y * psi.x;
    if (dagger) { vr += pr; vi += pi; } else { vr -= pr; vi -= pi; }
    out[(size_t)k * C + c] = make_double2(vr, vi);   // hmc.c:182
  }
}

// sum of term[k * C + c], k = 0 .. V-1, in that order, by thread 0; broadcast through shared memory
__device__ __forceinline__ double strict_sum(const double *term, int V, int C, int c, double *bc) {
  __syncthreads();   // the terms are written
  if (threadIdx.x == 0) {
    double s = 0.0;
    int k = 0;
    for (; k + 8 <= V; k += 8) {   // the loads are independent of the running sum
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = __ldcg(&term[(size_t)(k + u) * C + c]);
#pragma unroll
      for (int u = 0; u < 8; u++) s += v[u];
    }
    for (; k < V; k++) s += __ldcg(&term[(size_t)k * C + c]);
    *bc = s;
  }
  __syncthreads();
  const double s = *bc;
  __syncthreads();   // bc may be rewritten
  return s;
}

__global__ void __launch_bounds__(256)
strict_cg_kernel(const double2 *__restrict__ b, double2 *__restrict__ x, double2 *r, double2 *p, double2 *Mp,
                 double2 *q, double *term, const double2 *__restrict__ W0, const double2 *__restrict__ W1,
                 const double *__restrict__ mass, const double *__restrict__ emu, const double *__restrict__ emmu,
                 const TbCgState s, const int nt, const int nx, const int C, const int adjoint) {
  __shared__ double bc;
  const int c = blockIdx.x, V = nt * nx;
  const double m = mass[c], e_p = emu[c], e_m = emmu[c];
  for (int k = threadIdx.x; k < V; k += blockDim.x) {   // hmc.c:349-356
    const size_t i = (size_t)k * C + c;
    const double2 v = b[i];
    x[i] = make_double2(0.0, 0.0);
    r[i] = v;
    p[i] = v;
    term[i] = v.x * v.x + v.y * v.y;
  }
  double rr_old = strict_sum(term, V, C, c, &bc);
  const double rr_init = rr_old;
  double rr = rr_old;
  int status = TB_CG_MAXITER, iters = 0;
  if (rr_old < s.accuracy) {   // hmc.c:359-361
    status = TB_CG_ZERO_SOURCE;
  } else {
    for (int k = 1; k < s.max_iter; k++) {   // hmc.c:364
      strict_apply(p, Mp, W0, W1, nt, nx, C, c, m, e_p, e_m, false);   // hmc.c:366
      __syncthreads();
      strict_apply(Mp, q, W0, W1, nt, nx, C, c, m, adjoint ? e_m : e_p, adjoint ? e_p : e_m, adjoint != 0);   // :367
      __syncthreads();
      for (int j = threadIdx.x; j < V; j += blockDim.x) {   // hmc.c:368-370
        const size_t i = (size_t)j * C + c;
        const double2 pv = p[i], qv = q[i];
        term[i] = pv.x * qv.x + pv.y * qv.y;
      }
      const double pq = strict_sum(term, V, C, c, &bc);
      const double a = rr_old / pq;   // hmc.c:371
      for (int j = threadIdx.x; j < V; j += blockDim.x) {   // hmc.c:372-379
        const size_t i = (size_t)j * C + c;
        double2 xv = x[i], rv = r[i];
        const double2 pv = p[i], qv = q[i];
        xv.x += a * pv.x;
        xv.y += a * pv.y;
        rv.x -= a * qv.x;
        rv.y -= a * qv.y;
        x[i] = xv;
        r[i] = rv;
        term[i] = rv.x * rv.x + rv.y * rv.y;
      }
      rr = strict_sum(term, V, C, c, &bc);
      iters = k;
      if (rr < s.accuracy) { status = TB_CG_CONVERGED; break; }                   // hmc.c:381
      if (rr / rr_init > TB_DIVERGENCE_RATIO) { status = TB_CG_DIVERGED; break; }   // hmc.c:383
      const double be = rr / rr_old;   // hmc.c:390
      for (int j = threadIdx.x; j < V; j += blockDim.x) {   // hmc.c:391-392
        const size_t i = (size_t)j * C + c;
        const double2 rv = r[i];
        double2 pv = p[i];
        pv.x = rv.x + be * pv.x;
        pv.y = rv.y + be * pv.y;
        p[i] = pv;
      }
      rr_old = rr;   // hmc.c:394
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    s.status[c] = status;
    s.iters[c] = iters;
    s.rr[c] = rr;
    s.rr_init[c] = rr_init;
    s.active[c] = 0;
  }
}

// W0 = s0(t) 1/2 eta0(x) (cos A0, sin A0), W1 = s1(x) 1/2 (cos A1, sin A1) from cos / sin computed by the caller
__global__ void links_from_trig_kernel(const double2 *__restrict__ T0, const double2 *__restrict__ T1,
                                       double2 *__restrict__ W0, double2 *__restrict__ W1, int nt, int nx, int C,
                                       int t_off, int nt_global) {
  const size_t total = (size_t)nt * nx * C;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    const size_t site = k / C;
    const int x = (int)(site % nx);
    const int t = (int)(site / nx) + t_off;
    double f0 = (x & 1) ? -0.5 : 0.5;   // eta0 = (-1)^x, hmc.c:917-921
    if (t == nt_global - 1) f0 = -f0;   // antiperiodic wrap, hmc.c:143-148
    const double f1 = (x == nx - 1) ? -0.5 : 0.5;
    const double2 a = T0[k], bb = T1[k];
    W0[k] = make_double2(f0 * a.x, f0 * a.y);
    W1[k] = make_double2(f1 * bb.x, f1 * bb.y);
  }
}

}  // namespace

int tb_run_cg_strict(tb_ctx *ctx, const double2 *b, double2 *x) {
  if (ctx->nranks > 1 || ctx->msite) {
    tb_set_error("strict solver: single-GPU family A contexts only");
    return TB_EINVAL;
  }
  if (b == x) {
    tb_set_error("strict solver: in-place solve is not supported");
    return TB_EINVAL;
  }
  strict_cg_kernel<<<ctx->C, 256, 0, ctx->stream>>>(b, x, ctx->r, ctx->p, ctx->Mp, ctx->q, ctx->stage, ctx->W0,
                                                   ctx->W1, ctx->d_mass, ctx->d_emu, ctx->d_emmu, ctx->cg, ctx->nt,
                                                   ctx->nx, ctx->C, tb_conj_is_dagger(ctx) ? 1 : 0);
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

int tb_launch_links_from_trig(tb_ctx *ctx, const double2 *T0, const double2 *T1) {
  tb_gauge_sharing(ctx, false);
  int blocks = (int)((ctx->nsite + 255) / 256);
  if (blocks > TB_NUM_SMS_B200 * 16) blocks = TB_NUM_SMS_B200 * 16;
  links_from_trig_kernel<<<blocks, 256, 0, ctx->stream>>>(T0, T1, ctx->W0, ctx->W1, ctx->nt, ctx->nx, ctx->C,
                                                          ctx->t_off, ctx->nt_global);
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}
