// tb_real.cu — family B of the reference (vec_ops.c behind Thirring.h) as REAL 8-byte kernels, sm_100a, FP64.
//
//   fM / fM_transpose   vec_ops.c:96-172 (ANTISYMMETRIC | OPENX), :176-249 (SYMMETRIC)      real_apply_kernel
//   cg_MdM              vec_ops.c:261-307   CG on M^T M, x0 = 0, stop ||r||^2 < 1e-30, 1e50 on divergence
//   cg_propagator       vec_ops.c:311-321   (M^T M)^-1 M^T source                              real_cg_kernel
//   vec_dot, vec_dmul_add  vec_ops.c:51-62  batched, device-resident                          real_dot / real_dmul_add
//
// The operator is real: M = m + D on the free sites (field == 0, vec_ops.c:107) and the identity on occupied ones, D
// the staggered hop matrix with hops into or out of an occupied site dropped.  The complex kernels carry it with
// imaginary parts that are zero (twice the bytes, four times the multiplications per hop); here a site is one double.
// Vectors are in the CANONICAL layout double[source][t][x] on the host and on the device: a multi-RHS batch (the 2 NX
// point sources of measure_propagator, fermionbag.c:389-435, or the V/2 of calc_Dinv_cg,
// fluctuation_determinant.c:973-1003) is nsrc contiguous lattices and goes host -> H2D -> one kernel -> D2H.
//
// On-chip CG (real_cg_kernel): one CTA per source, the whole solve in one launch.  A thread owns 8 consecutive t-rows of
// one x-column: r, p, x and the transient M p live in registers (32 doubles); what the stencil neighbours read goes
// through two shared-memory exchange fields (p and M p, 8 bytes per site) that hold ZERO on occupied sites, so a hop
// into an occupied site adds an exact zero and no per-hop mask is needed; the occupied sites themselves are identity
// rows by an 8-bit mask per thread.  No link field exists at all: the hop coefficients of a thread are five doubles
// (+-1/2 eta e^{+-mu}, +-1/2 with the boundary rule of its column).  64 x 64: 512 threads (16 warps per SM), 64 KB of
// shared memory, HBM touched once per solve.
#include "tb_common.cuh"

namespace {

// sign / presence of the x hop across the boundary: ANTISYMMETRIC -1, SYMMETRIC +1, OPENX 0 (Thirring.h:27-29)
__device__ __forceinline__ double wrap_factor(int bc) { return bc == TB_BC_SYMMETRIC ? 1.0 : (bc == TB_BC_OPENX ? 0.0 : -1.0); }

// out = M in (transpose = 0) or M^T in, one thread per site, any lattice shape; canonical real vectors
__global__ void real_apply_kernel(const double *__restrict__ in, double *__restrict__ out, const int *__restrict__ field,
                                  const double *__restrict__ mass, const double *__restrict__ emu,
                                  const double *__restrict__ emmu, int nt, int nx, int nchains, int bc, int transpose) {
  const size_t V = (size_t)nt * nx, total = V * nchains;
  const double wf = wrap_factor(bc);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i / V);
    const int k = (int)(i - (size_t)c * V), t = k / nx, x = k - t * nx;
    const double *v = in + (size_t)c * V;
    const int *f = field + (size_t)c * V;
    double o = v[k];
    if (f[k] == 0) {   // vec_ops.c:107
      const double eta = (x & 1) ? -1.0 : 1.0;
      const double eu = transpose ? emmu[c] : emu[c], ed = transpose ? emu[c] : emmu[c];
      const double sg = transpose ? -1.0 : 1.0;
      const int tp = t + 1 == nt ? 0 : t + 1, tm = t == 0 ? nt - 1 : t - 1;
      const int xp = x + 1 == nx ? 0 : x + 1, xm = x == 0 ? nx - 1 : x - 1;
      o = mass[c] * v[k];
      if (f[tp * nx + x] == 0) o = fma((tp > t ? sg : -sg) * 0.5 * eta * eu, v[tp * nx + x], o);    // vec_ops.c:110-113
      if (f[tm * nx + x] == 0) o = fma((tm > t ? sg : -sg) * 0.5 * eta * ed, v[tm * nx + x], o);    // vec_ops.c:115-118
      if (f[t * nx + xp] == 0) o = fma((xp > x ? sg : sg * wf) * 0.5, v[t * nx + xp], o);           // :120-123, :201-203
      if (f[t * nx + xm] == 0) o = fma((xm > x ? sg * -wf : -sg) * 0.5, v[t * nx + xm], o);         // :125-128, :205-207
    }
    out[i] = o;
  }
}

template <int NW>
__device__ __forceinline__ double real_block_sum(double v, double *scratch, int nwarps) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < nwarps; w++) s += scratch[w];   // warp order: deterministic
  return s;
}

// The hop coefficients of one thread (column x, rows t0 .. t0+7) for M or M^T
struct RealCoef {
  double up, dn, up_top, dn_bot, xp, xm;   // up_top: the +t hop of the tile's last row, dn_bot: the -t hop of its first
};

__device__ __forceinline__ RealCoef real_coef(int x, int t0, int nt, int nx, double eu, double ed, int bc, bool transpose) {
  const double eta = (x & 1) ? -1.0 : 1.0, wf = wrap_factor(bc), sg = transpose ? -1.0 : 1.0;
  RealCoef k;
  // M: +1/2 eta e^{mu} psi(t+1) - 1/2 eta e^{-mu} psi(t-1), opposite sign across the antiperiodic t boundary;
  // M^T: all hop signs flipped and e^{mu} <-> e^{-mu} (vec_ops.c:144-171)
  const double fu = transpose ? ed : eu, fd = transpose ? eu : ed;
  k.up = sg * 0.5 * eta * fu;
  k.dn = -sg * 0.5 * eta * fd;
  k.up_top = (t0 + 8 == nt) ? -k.up : k.up;
  k.dn_bot = (t0 == 0) ? -k.dn : k.dn;
  k.xp = (x == nx - 1) ? sg * wf * 0.5 : sg * 0.5;
  k.xm = (x == 0) ? -sg * wf * 0.5 : -sg * 0.5;
  return k;
}

// o[i] = m f[i] + hops, i = 0..7, for the thread's column; F: the field with zeros on occupied sites (shared memory).
// base = t0 * nx + x; the row stride nx is a compile-time constant in the 64 x 64 instantiation (immediate offsets).
template <typename Epi>
__device__ __forceinline__ void real_tile_apply(const double (&f)[8], const double *F, int base, int off_below,
                                                int off_above, int dxm, int dxp, int nx, double m, const RealCoef &k,
                                                unsigned occ, Epi epi) {
  const double below = F[base + off_below], above = F[base + off_above];
  // neighbours in t inside the tile come from registers: they must be the MASKED values the exchange field holds
  double fz[8];
#pragma unroll
  for (int i = 0; i < 8; i++) fz[i] = ((occ >> i) & 1u) ? 0.0 : f[i];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int row = base + i * nx;
    const double l = F[row + dxm], r = F[row + dxp];
    const double up = i == 7 ? above : fz[(i + 1) & 7];
    const double dn = i == 0 ? below : fz[(i + 7) & 7];
    double o = m * f[i];
    o = fma(i == 7 ? k.up_top : k.up, up, o);
    o = fma(i == 0 ? k.dn_bot : k.dn, dn, o);
    o = fma(k.xp, r, o);
    o = fma(k.xm, l, o);
    if ((occ >> i) & 1u) o = f[i];   // identity row (vec_ops.c:130)
    epi(i, o);
  }
}

// NTT, NXT: compile-time lattice shape (64 x 64, the size Thirring.h compiles in), or 0, 0: the shape of the arguments
template <bool PROP, int NTT, int NXT>
__global__ void __launch_bounds__(512, 1)
real_cg_kernel(const double *__restrict__ bsrc, double *__restrict__ xout, const int *__restrict__ field,
               const double *__restrict__ mass, const double *__restrict__ emu, const double *__restrict__ emmu,
               const TbCgState s, const int nt_rt, const int nx_rt, const int bc) {
  extern __shared__ __align__(16) unsigned char real_smem[];
  const int nt = NTT ? NTT : nt_rt, nx = NXT ? NXT : nx_rt;
  const int V = nt * nx, c = blockIdx.x, tid = threadIdx.x, nwarps = blockDim.x >> 5;
  double *Fp = reinterpret_cast<double *>(real_smem), *Fm = Fp + V, *scrA = Fm + V, *scrB = scrA + 32;
  const int x = tid % nx, t0 = (tid / nx) * 8, base = t0 * nx + x;
  // offsets of the four halo neighbours relative to a site of the tile (periodic indices; the boundary rules are in
  // the coefficients)
  const int off_below = (t0 == 0 ? nt - 1 : -1) * nx, off_above = (t0 + 8 == nt ? 8 - nt : 8) * nx;
  const int dxm = x == 0 ? nx - 1 : -1, dxp = x + 1 == nx ? 1 - nx : 1;
  const double m = mass[c];
  const RealCoef kM = real_coef(x, t0, nt, nx, emu[c], emmu[c], bc, false);
  const RealCoef kT = real_coef(x, t0, nt, nx, emu[c], emmu[c], bc, true);
  const double *bc_ = bsrc + (size_t)c * V;
  const int *fc = field + (size_t)c * V;
  unsigned occ = 0;
  double r[8], p[8], xv[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int k = base + i * nx;
    if (fc[k] != 0) occ |= 1u << i;
    r[i] = bc_[k];
    xv[i] = 0.0;   // vec_zero(inv), vec_ops.c:268
  }
  if (PROP) {   // cg_propagator: the source is M^T source (vec_ops.c:316)
#pragma unroll
    for (int i = 0; i < 8; i++) Fp[base + i * nx] = ((occ >> i) & 1u) ? 0.0 : r[i];
    __syncthreads();
    double tsrc[8];
    real_tile_apply(r, Fp, base, off_below, off_above, dxm, dxp, nx, m, kT, occ, [&](int i, double o) { tsrc[i] = o; });
    __syncthreads();   // every thread has read Fp
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = tsrc[i];
  }
  double rr = 0.0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    p[i] = r[i];
    rr = fma(r[i], r[i], rr);
    Fp[base + i * nx] = ((occ >> i) & 1u) ? 0.0 : p[i];
  }
  rr = real_block_sum<0>(rr, scrA, nwarps);   // its barrier publishes p
  const double rr_init = rr;
  double rr_old = rr;
  int status = TB_CG_MAXITER, iters = 0;
  if (rr_old < s.accuracy) {   // vec_ops.c:275-277
    status = TB_CG_ZERO_SOURCE;
  } else {
    for (int k = 1; k < s.max_iter; k++) {   // vec_ops.c:280
      double mp[8], pq = 0.0;
      // M p (vec_ops.c:282), published masked as it is produced; <p, M^T M p> = |M p|^2
      real_tile_apply(p, Fp, base, off_below, off_above, dxm, dxp, nx, m, kM, occ, [&](int i, double o) {
        mp[i] = o;
        Fm[base + i * nx] = ((occ >> i) & 1u) ? 0.0 : o;
        pq = fma(o, o, pq);
      });
      pq = real_block_sum<1>(pq, scrB, nwarps);   // its barrier publishes M p
      const double a = rr_old / pq;                // vec_ops.c:285
      rr = 0.0;
      // q = M^T M p consumed on the fly: r -= a q (vec_ops.c:287), ||r||^2
      real_tile_apply(mp, Fm, base, off_below, off_above, dxm, dxp, nx, m, kT, occ, [&](int i, double o) {
        r[i] = fma(-a, o, r[i]);
        rr = fma(r[i], r[i], rr);
      });
#pragma unroll
      for (int i = 0; i < 8; i++) xv[i] = fma(a, p[i], xv[i]);   // vec_ops.c:286
      rr = real_block_sum<2>(rr, scrA, nwarps);
      iters = k;
      if (rr < s.accuracy) { status = TB_CG_CONVERGED; break; }                     // vec_ops.c:290
      if (rr / rr_init > TB_DIVERGENCE_RATIO) { status = TB_CG_DIVERGED; break; }   // vec_ops.c:292
      const double be = rr / rr_old;   // vec_ops.c:298
#pragma unroll
      for (int i = 0; i < 8; i++) {
        p[i] = fma(be, p[i], r[i]);    // vec_ops.c:299
        Fp[base + i * nx] = ((occ >> i) & 1u) ? 0.0 : p[i];   // the last readers of Fp passed the |Mp|^2 barrier
      }
      rr_old = rr;
      __syncthreads();
    }
  }
  double *xc = xout + (size_t)c * V;
#pragma unroll
  for (int i = 0; i < 8; i++) xc[base + i * nx] = status == TB_CG_DIVERGED ? 1e50 : xv[i];   // vec_ops.c:294
  if (tid == 0) {
    s.status[c] = status;
    s.iters[c] = iters;
    s.rr[c] = rr;
    s.rr_init[c] = rr_init;
    s.active[c] = 0;
  }
}

// per-vector dot product of a batch of real lattices (vec_dot, vec_ops.c:56-62): one CTA per vector, fixed tree
__global__ void __launch_bounds__(256) real_dot_kernel(const double *__restrict__ a, const double *__restrict__ b,
                                                       double *__restrict__ out, int V) {
  __shared__ double scratch[32];
  const double *av = a + (size_t)blockIdx.x * V, *bv = b + (size_t)blockIdx.x * V;
  double acc = 0.0;
  for (int k = threadIdx.x; k < V; k += blockDim.x) acc = fma(av[k], bv[k], acc);
  const double t = real_block_sum<3>(acc, scratch, blockDim.x >> 5);
  if (threadIdx.x == 0) out[blockIdx.x] = t;
}

// a = b + e d with a per-vector scalar e (vec_dmul_add, vec_ops.c:51-55), batched
__global__ void real_dmul_add_kernel(double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ d,
                                     const double *__restrict__ e, size_t V, size_t total) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    a[i] = fma(e[i / V], d[i], b[i]);
}

}  // namespace

// the on-chip real CG serves lattices of whole 8-row tiles whose nt/8 * nx threads are whole warps, at most 512
bool tb_real_cg_supported(const tb_ctx *ctx) {
  if (!ctx->msite || ctx->nranks != 1 || ctx->nt % 8 != 0) return false;
  const int threads = ctx->nt / 8 * ctx->nx;
  return threads % 32 == 0 && threads <= 512 && tb_conj_is_dagger(ctx) && getenv("TB_NO_REAL") == nullptr;
}

int tb_launch_real_apply(tb_ctx *ctx, bool transpose, const double *d_in, double *d_out) {
  int blocks = (int)((ctx->nsite + 255) / 256);
  if (blocks > TB_NUM_SMS_B200 * 16) blocks = TB_NUM_SMS_B200 * 16;
  real_apply_kernel<<<blocks, 256, 0, ctx->stream>>>(d_in, d_out, ctx->occ_stage, ctx->d_mass, ctx->d_emu, ctx->d_emmu,
                                                     ctx->nt, ctx->nx, ctx->C, ctx->occ_bc, transpose ? 1 : 0);
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

// sources [c0, c0 + n) of a batch in the canonical real layout (pointers at source 0)
int tb_run_cg_real(tb_ctx *ctx, const double *d_b, double *d_x, bool propagator, int c0, int n, cudaStream_t st) {
  const int V = (int)ctx->V, threads = ctx->nt / 8 * ctx->nx;
  const size_t smem = (size_t)(2 * V + 64) * sizeof(double);
  const bool fixed = ctx->nt == 64 && ctx->nx == 64;
  auto kern = fixed ? (propagator ? real_cg_kernel<true, 64, 64> : real_cg_kernel<false, 64, 64>)
                    : (propagator ? real_cg_kernel<true, 0, 0> : real_cg_kernel<false, 0, 0>);
  TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TbCgState s = ctx->cg;   // per-chain outputs of sources c0.. land at their own index
  s.status += c0; s.iters += c0; s.rr += c0; s.rr_init += c0; s.active += c0;
  kern<<<n, threads, smem, st>>>(d_b + (size_t)c0 * V, d_x + (size_t)c0 * V, ctx->occ_stage + (size_t)c0 * V,
                                 ctx->d_mass + c0, ctx->d_emu + c0, ctx->d_emmu + c0, s, ctx->nt, ctx->nx, ctx->occ_bc);
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

extern "C" int tb_vec_dot_real_dev(tb_ctx *ctx, const double *d_a, const double *d_b, double *out_host) {
  if (!ctx || !d_a || !d_b || !out_host) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  real_dot_kernel<<<ctx->C, 256, 0, ctx->stream>>>(d_a, d_b, ctx->cg.dot, (int)ctx->V);
  ctx->launches++;
  TB_CUDA(cudaMemcpyAsync(ctx->h_rr, ctx->cg.dot, ctx->C * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  TB_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(out_host, ctx->h_rr, ctx->C * sizeof(double));
  return TB_OK;
}

extern "C" int tb_vec_dmul_add_real_dev(tb_ctx *ctx, double *d_a, const double *d_b, const double *d_d,
                                        const double *e_host) {
  if (!ctx || !d_a || !d_b || !d_d || !e_host) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CUDA(cudaMemcpyAsync(ctx->cg.dot, e_host, ctx->C * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  int blocks = (int)((ctx->nsite + 255) / 256);
  if (blocks > TB_NUM_SMS_B200 * 16) blocks = TB_NUM_SMS_B200 * 16;
  real_dmul_add_kernel<<<blocks, 256, 0, ctx->stream>>>(d_a, d_b, d_d, ctx->cg.dot, ctx->V, ctx->nsite);
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}
