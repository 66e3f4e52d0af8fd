// tb_hmc.cu — device-resident batched HMC trajectory (the caller of the hot path).
//
// One call = update_gauge (hmc.c:671-746) for every chain of the context, as coded, quirks included
// (SURVEY Appendix A/D): N(0,1)+iN(0,1) heat-bath vectors with S = sum |.|^2, kinetic term sum p^2 without 1/2,
// position step A += 2 eps p, the extra "stochastic M~" action term and its force, per-chain Metropolis test.
// Nothing crosses PCIe during a trajectory except the per-chain observables at the end.
//
//   fill_gauss / fill_momentum   Box-Muller from a counter-based Philox stream     hmc.c:418-447, 483-499
//   gauge_action_kernel          sum (1 - cos A)                                     hmc.c:73-79
//   gauge_step_links_kernel      A += 2 eps p fused with the link rebuild            hmc.c:97-101
//   force_kernel                 gauge + pseudofermion + "M_conjugate" forces        hmc.c:504-661
//   accept_kernel                dS, exp(-dS) > u, A <- new_A per chain              hmc.c:729-744
//   heatbath_kernel              quenched heat-bath sweeps, start configuration      hmc.c:82-93
#include "tb_common.cuh"

namespace {
#include "tb_device.cuh"

#define TB_NF 2.0  /* "#define Nf 2", hmc.c:28 */

// ---- Philox4x32-10: counter (site index, global chain index, trajectory | sweep, stream), key (seed lo, seed hi) ----
// Every consumer has its own range of the stream word (TB_RNG_*): trajectory fields 1..4, heat bath, measure() sources
// and condensate sources cannot collide whatever nsrc is; the chain sits in the counter, not in the key, so two
// (seed, chain) pairs never share a stream.
__device__ __forceinline__ uint4 philox(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

// uniform in (0,1) with the 32-bit resolution of the reference's mersenne() (mersenne_inline.c:108); the half
// offset keeps log() finite (the reference has a 2^-32 chance of log(0))
__device__ __forceinline__ double u01(unsigned int r) { return ((double)r + 0.5) * 2.3283064365386963e-10; }

// stream word: consumer in the top byte, source index below (nsrc < 2^24)
#define TB_RNG_MEASURE 0x01000000u
#define TB_RNG_CONDENSATE 0x02000000u
#define TB_RNG_HEATBATH 0x03000000u
#define TB_RNG_MAX_SOURCES 0x01000000

struct RngKey {
  unsigned long long seed;
  unsigned int traj, stream;
  unsigned int chain0;   // global index of the context's first chain (tb_hmc_set_chain_offset)
};

// (sqrt(-2 ln x1) cos(2 pi x2), sqrt(-2 ln x1) sin(2 pi x2)), hmc.c:425-426 / 490-491
__device__ __forceinline__ double2 box_muller(size_t site, int c, const RngKey k) {
  const uint4 r = philox(make_uint4((unsigned int)site, (unsigned int)c + k.chain0, k.traj, k.stream),
                         make_uint2((unsigned int)k.seed, (unsigned int)(k.seed >> 32)));
  const double x1 = u01(r.x), x2 = u01(r.y);
  const double rad = sqrt(-2.0 * log(x1));
  double s, co;
  sincospi(2.0 * x2, &s, &co);
  return make_double2(rad * co, rad * s);
}

__global__ void fill_gauss_kernel(double2 *__restrict__ v, size_t nsite_total, int C, RngKey k) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nsite_total; i += (size_t)gridDim.x * blockDim.x)
    v[i] = box_muller(i / C, (int)(i % C), k);
}

__global__ void fill_uniform_kernel(double *__restrict__ u, int C, RngKey k) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const uint4 r = philox(make_uint4(0u, (unsigned int)c + k.chain0, k.traj, k.stream),
                           make_uint2((unsigned int)k.seed, (unsigned int)(k.seed >> 32)));
    u[c] = u01(r.x);
  }
}

// per chain: sum over sites and directions of (1 - cos A); the caller multiplies by Nf/g (hmc.c:73-79)
template <int TT>
__global__ void __launch_bounds__(TB_MAX_BLOCK)
gauge_action_kernel(const double2 *__restrict__ A, const TbGeom g, const TbCgState s, const TbSlab sl) {
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
        const double2 a = A[t * R + j];
        acc += 1.0 - cos(a.x);
        acc += 1.0 - cos(a.y);
      }
    }
  }
  reduce_finalize<FIN_DOT, false, 0>(acc, g, s, sl, b, red);
}

// A += 2 eps mom (hmc.c:97-101) and W <- links(A) in one pass
__global__ void gauge_step_links_kernel(double2 *__restrict__ A, const double2 *__restrict__ mom,
                                        double2 *__restrict__ W0, double2 *__restrict__ W1, double two_eps,
                                        int nt, int nx, int C, int t_off, int nt_global) {
  const size_t total = (size_t)nt * nx * C;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    const size_t site = k / C;
    const int x = (int)(site % nx);
    const int t = (int)(site / nx) + t_off;
    double2 a = A[k];
    const double2 p = mom[k];
    a.x += two_eps * p.x;
    a.y += two_eps * p.y;
    A[k] = a;
    double s0, c0, s1, c1;
    sincos(a.x, &s0, &c0);
    sincos(a.y, &s1, &c1);
    double f0 = (x & 1) ? -0.5 : 0.5;
    if (t == nt_global - 1) f0 = -f0;
    const double f1 = (x == nx - 1) ? -0.5 : 0.5;
    W0[k] = make_double2(f0 * c0, f0 * s0);
    W1[k] = make_double2(f1 * c1, f1 * s1);
  }
}

// momentum_step (hmc.c:504-661) after the CG: mom -= eps * (gauge force), then -= eps * (pseudofermion force),
// then -= eps * ("M_conjugate" force), in the reference's order.  chi = (M~M)^-1 psi, phi = M chi, st = the
// stochastic vector.  s = +1 on interior links, -1 on the wrap link ("if (t2 > t) ... else", hmc.c:523-530).
__global__ void force_kernel(double2 *__restrict__ mom, const double2 *__restrict__ A,
                             const double2 *__restrict__ chi, const double2 *__restrict__ phi,
                             const double2 *__restrict__ st, const double *__restrict__ nf_over_g,
                             const double *__restrict__ emu, const double *__restrict__ emmu, double eps,
                             int nt, int nx, int C) {
  const size_t total = (size_t)nt * nx * C;
  const size_t R = (size_t)nx * C;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(k % C);
    const size_t site = k / C;
    const int x = (int)(site % nx);
    const int t = (int)(site / nx);
    const size_t kt = (size_t)((t + 1 == nt) ? 0 : t + 1) * R + (size_t)x * C + c;     // n + t^
    const size_t kx = (size_t)t * R + (size_t)((x + 1 == nx) ? 0 : x + 1) * C + c;     // n + x^
    const double st_sign = (t + 1 == nt) ? -1.0 : 1.0, sx_sign = (x + 1 == nx) ? -1.0 : 1.0;
    const double eta0 = (x & 1) ? -1.0 : 1.0;   // hmc.c:917-921
    const double2 a = A[k];
    double2 p = mom[k];
    double sA0, cA0, sA1, cA1;
    sincos(a.x, &sA0, &cA0);
    sincos(a.y, &sA1, &cA1);
    // gauge force, hmc.c:507-510
    const double nfg = nf_over_g[c];
    p.x -= eps * nfg * sA0;
    p.y -= eps * nfg * sA1;
    // pseudofermion force, hmc.c:517-577:  F = -s eta (e^{mu} X1 - e^{-mu} X2),
    //   X_i = -Re(c_i) sin A - Im(c_i) cos A,  c1 = conj(phi(n)) chi(n+mu^),  c2 = conj(chi(n)) phi(n+mu^)
    const double2 ch = chi[k], ph = phi[k];
    {
      const double2 cht = chi[kt], pht = phi[kt];
      const double c1r = ph.x * cht.x + ph.y * cht.y, c1i = ph.x * cht.y - ph.y * cht.x;
      const double c2r = ch.x * pht.x + ch.y * pht.y, c2i = ch.x * pht.y - ch.y * pht.x;
      const double X1 = -c1r * sA0 - c1i * cA0, X2 = -c2r * sA0 - c2i * cA0;
      const double F = -st_sign * eta0 * (emu[c] * X1 - emmu[c] * X2);
      p.x -= eps * F;
    }
    {
      const double2 chx = chi[kx], phx = phi[kx];
      const double c1r = ph.x * chx.x + ph.y * chx.y, c1i = ph.x * chx.y - ph.y * chx.x;
      const double c2r = ch.x * phx.x + ch.y * phx.y, c2i = ch.x * phx.y - ch.y * phx.x;
      const double X1 = -c1r * sA1 - c1i * cA1, X2 = -c2r * sA1 - c2i * cA1;
      const double F = -sx_sign * (X1 - X2);
      p.y -= eps * F;
    }
    // "M_conjugate" force, hmc.c:608-661: t-links only (the two x terms cancel exactly)
    {
      const double2 s0 = st[k], s1 = st[kt];
      const double cr = s0.x * s1.x + s0.y * s1.y, ci = s0.x * s1.y - s0.y * s1.x;
      const double X = cr * sA0 + ci * cA0;
      const double F = -st_sign * 0.5 * eta0 * (emu[c] - emmu[c]) * X;
      p.x -= eps * F;
    }
    mom[k] = p;
  }
}

// obs layout per chain (doubles): 0 Sg 1 Smdm 2 Smd 3 Smom | 4 Sg' 5 Smdm' 6 Smd' 7 Smom' | 8 dS 9 accepted
// the raw sums arrive in sums[k*Cpad + c]; Sg sums are multiplied by Nf/g here (hmc.c:78)
__global__ void accept_kernel(double *__restrict__ obs, const double *__restrict__ sums,
                              const double *__restrict__ nf_over_g, const double *__restrict__ u,
                              const int *__restrict__ cg_failed, int *__restrict__ accept, int C, int Cpad) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double o[10];
  for (int k = 0; k < 8; k++) o[k] = sums[(size_t)k * Cpad + c];
  o[0] *= nf_over_g[c];
  o[4] *= nf_over_g[c];
  // hmc.c:729-732
  const double dS = (o[7] - o[3]) + (o[5] - o[1]) + (o[6] - o[2]) + (o[4] - o[0]);
  const int acc = (cg_failed[c] == 0) && (exp(-dS) > u[c]);   // hmc.c:738
  o[8] = dS;
  o[9] = acc;
  accept[c] = acc;
  for (int k = 0; k < 10; k++) obs[(size_t)c * 10 + k] = o[k];
}

// A <- new_A for the accepted chains (hmc.c:740-741)
__global__ void commit_kernel(double2 *__restrict__ A, const double2 *__restrict__ newA,
                              const int *__restrict__ accept, size_t total, int C) {
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x)
    if (accept[k % C]) A[k] = newA[k];
}

// after every solve of a trajectory: remember chains whose solve failed, add the iteration counts to the device-side
// total (the host reads both once, at the end of the trajectory)
__global__ void or_failed_kernel(int *__restrict__ failed, const int *__restrict__ status, const int *__restrict__ iters,
                                 unsigned long long *__restrict__ iter_sum, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C && status[c] != TB_CG_CONVERGED && status[c] != TB_CG_ZERO_SOURCE) failed[c] |= 1 << status[c];
  unsigned int n = c < C ? (unsigned int)iters[c] : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(iter_sum, (unsigned long long)n);
}

// update_puregauge_hb (hmc.c:82-93): links decouple in the quenched action, so every link runs its own
// Metropolis chain of `sweeps` proposals new = 2 pi u - pi, accepted with exp((Nf/g)(cos new - cos old))
__global__ void heatbath_kernel(double2 *__restrict__ A, const double *__restrict__ nf_over_g, int sweeps,
                                size_t total, int C, RngKey k) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t site = i / C;
    const double nfg = nf_over_g[c];
    double2 a = A[i];
    for (int s = 0; s < sweeps; s++) {
      const uint4 r = philox(make_uint4((unsigned int)site, (unsigned int)c + k.chain0, (unsigned int)s, k.stream),
                             make_uint2((unsigned int)k.seed, (unsigned int)(k.seed >> 32)));
      const double n0 = 2.0 * M_PI * u01(r.x) - M_PI, n1 = 2.0 * M_PI * u01(r.z) - M_PI;
      if (u01(r.y) < exp(nfg * (cos(n0) - cos(a.x)))) a.x = n0;
      if (u01(r.w) < exp(nfg * (cos(n1) - cos(a.y)))) a.y = n1;
    }
    A[i] = a;
  }
}

// per chain: sum of both link angles (Magnetisation = sum A / V, hmc.c:831-833,839)
template <int TT>
__global__ void __launch_bounds__(TB_MAX_BLOCK)
magnetisation_kernel(const double2 *__restrict__ A, const TbGeom g, const TbCgState s, const TbSlab sl) {
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
        const double2 a = A[t * R + j];
        acc += a.x + a.y;
      }
    }
  }
  reduce_finalize<FIN_DOT, false, 0>(acc, g, s, sl, b, red);
}

// per chain: Im<a,b> = sum re(a) im(b) - im(a) re(b)   (fermion_phase, hmc.c:805-808)
template <int TT>
__global__ void __launch_bounds__(TB_MAX_BLOCK)
im_dot_kernel(const double2 *__restrict__ a, const double2 *__restrict__ bb, const TbGeom g, const TbCgState s,
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
        const double2 u = a[t * R + j], v = bb[t * R + j];
        acc += u.x * v.y - u.y * v.x;
      }
    }
  }
  reduce_finalize<FIN_DOT, false, 0>(acc, g, s, sl, b, red);
}

int ew_blocks(size_t n) {
  size_t b = (n + 255) / 256;
  return (int)(b > (size_t)TB_NUM_SMS_B200 * 16 ? (size_t)TB_NUM_SMS_B200 * 16 : b);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
int tb_run_cg_any(tb_ctx *ctx, const double2 *b, double2 *x);  // tb_api.cu: solver dispatch + timing
int tb_run_cg_async(tb_ctx *ctx, const double2 *b, double2 *x);  // the same, no host synchronisation (on-chip solvers)

static int hmc_alloc(tb_ctx *ctx) {
  if (ctx->hmc.mom) return TB_OK;
  if (ctx->nranks > 1) {
    tb_set_error("the device-resident trajectory is chain-parallel; slab contexts are not supported");
    return TB_EINVAL;
  }
  const size_t n = ctx->nsite, cp = ctx->g.Cpad;
  double2 **vecs[] = {&ctx->hmc.mom, &ctx->hmc.newA, &ctx->hmc.psi, &ctx->hmc.st, &ctx->hmc.chi, &ctx->hmc.phi,
                      &ctx->hmc.gauss};
  for (double2 **v : vecs) {
    cudaError_t e = cudaMalloc((void **)v, n * sizeof(double2));
    if (e != cudaSuccess) {
      tb_set_error("cudaMalloc(HMC vector) failed: %s", cudaGetErrorString(e));
      return TB_ENOMEM;
    }
  }
  TB_CUDA(cudaMalloc((void **)&ctx->hmc.sums, 12 * cp * sizeof(double)));
  TB_CUDA(cudaMalloc((void **)&ctx->hmc.obs, 10 * cp * sizeof(double)));
  TB_CUDA(cudaMalloc((void **)&ctx->hmc.u, cp * sizeof(double)));
  TB_CUDA(cudaMalloc((void **)&ctx->hmc.accept, 2 * cp * sizeof(int)));
  ctx->hmc.failed = ctx->hmc.accept + cp;
  TB_CUDA(cudaMalloc((void **)&ctx->hmc.iter_sum, sizeof(unsigned long long)));
  if (!ctx->hmc.nf_over_g) {
    TB_CUDA(cudaMalloc((void **)&ctx->hmc.nf_over_g, cp * sizeof(double)));
    double *h = (double *)malloc(cp * sizeof(double));
    for (size_t c = 0; c < cp; c++) h[c] = TB_NF / 1.0;
    TB_CUDA(cudaMemcpy(ctx->hmc.nf_over_g, h, cp * sizeof(double), cudaMemcpyHostToDevice));
    free(h);
  }
  return TB_OK;
}

void tb_hmc_release(tb_ctx *ctx) {
  void *p[] = {ctx->hmc.mom, ctx->hmc.newA, ctx->hmc.psi, ctx->hmc.st, ctx->hmc.chi, ctx->hmc.phi, ctx->hmc.gauss,
               ctx->hmc.sums, ctx->hmc.obs, ctx->hmc.u, ctx->hmc.accept, ctx->hmc.nf_over_g, ctx->hmc.iter_sum};
  for (void *q : p)
    if (q) cudaFree(q);
}

extern "C" int tb_hmc_set_coupling(tb_ctx *ctx, const double *g, int n) {
  if (!ctx || !g || (n != 1 && n != ctx->C)) {
    tb_set_error("tb_hmc_set_coupling: n must be 1 or nchains");
    return TB_EINVAL;
  }
  TB_CUDA(cudaSetDevice(ctx->device));
  const size_t cp = ctx->g.Cpad;
  if (!ctx->hmc.nf_over_g) TB_CUDA(cudaMalloc((void **)&ctx->hmc.nf_over_g, cp * sizeof(double)));
  double *h = (double *)malloc(cp * sizeof(double));
  for (size_t c = 0; c < cp; c++) h[c] = TB_NF / g[(n == 1 || c >= (size_t)ctx->C) ? 0 : c];   // Nf/g, hmc.c:78
  TB_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaError_t e = cudaMemcpy(ctx->hmc.nf_over_g, h, cp * sizeof(double), cudaMemcpyHostToDevice);
  free(h);
  TB_CUDA(e);
  return TB_OK;
}

// Global index of this context's first chain: enters the Philox key, so an ensemble sharded over several contexts
// (GPUs) draws the same random numbers per chain whatever the number of shards.
extern "C" int tb_hmc_set_chain_offset(tb_ctx *ctx, unsigned int first_chain) {
  if (!ctx) return TB_EINVAL;
  ctx->hmc_chain_offset = first_chain;
  return TB_OK;
}

static int dot_to(tb_ctx *ctx, const double2 *a, const double2 *b, int slot) {
  return tb_launch_dot(ctx, a, b, ctx->hmc.sums + (size_t)slot * ctx->g.Cpad);
}

static int gauge_action_to(tb_ctx *ctx, const double2 *A, int slot) {
  const TbGeom &g = ctx->g;
  TbCgState s = ctx->cg;
  s.dot = ctx->hmc.sums + (size_t)slot * g.Cpad;
  TB_DISPATCH_TT(g.tt, (gauge_action_kernel<TT><<<grid_of(g), g.bc * g.bx, 0, ctx->stream>>>(A, g, s, ctx->slab)))
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  return TB_OK;
}

static int links_from(tb_ctx *ctx, const double2 *A) { return tb_launch_links(ctx, (const double *)A); }

extern "C" int tb_hmc_heatbath(tb_ctx *ctx, int sweeps, unsigned long long seed) {
  if (!ctx || sweeps < 0) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(hmc_alloc(ctx));
  if (!ctx->have_gauge) TB_CUDA(cudaMemsetAsync(ctx->Adev, 0, ctx->nsite * sizeof(double2), ctx->stream));  // hmc.c:915
  const RngKey k = {seed, 0u, TB_RNG_HEATBATH, ctx->hmc_chain_offset};
  heatbath_kernel<<<ew_blocks(ctx->nsite), 256, 0, ctx->stream>>>(ctx->Adev, ctx->hmc.nf_over_g, sweeps, ctx->nsite,
                                                                 ctx->C, k);
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  TB_CHECK(links_from(ctx, ctx->Adev));
  ctx->have_gauge = true;
  return TB_OK;
}

// One trajectory for every chain.  Random inputs are either drawn on the device (Philox keyed by seed, chain,
// trajectory index) or supplied by the caller in the canonical host layout for parity tests:
//   xi_host, st_host  complex [chain][t][x]   Gaussian vectors of random_pseudofermion / stochastic_vector
//   mom_host          real    [chain][t][x][2] momenta,     u_host [chain]  Metropolis uniforms
// obs_host (may be NULL): 10 doubles per chain, see accept_kernel.
extern "C" int tb_hmc_trajectory(tb_ctx *ctx, int nsteps, double traj_length, unsigned long long seed,
                                 unsigned int traj_index, const double *xi_host, const double *mom_host,
                                 const double *st_host, const double *u_host, double *obs_host, int *accepted_host,
                                 long long *cg_iters_host) {
  if (!ctx || nsteps < 1 || !(traj_length > 0)) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  tb_gauge_sharing(ctx, false);   // a trajectory evolves one field per chain (the per-chain copies are what it starts from)
  if (!ctx->have_gauge) {
    tb_set_error("tb_hmc_trajectory: no gauge field (tb_set_gauge / tb_hmc_heatbath first)");
    return TB_EINVAL;
  }
  TB_CHECK(hmc_alloc(ctx));
  TB_CHECK(tb_synchronize(ctx));  // joins the host-path sub-streams
  cudaStream_t st = ctx->stream;
  const size_t n = ctx->nsite, cp = ctx->g.Cpad;
  const int C = ctx->C;
  const bool dag = tb_conj_is_dagger(ctx);
  auto &H = ctx->hmc;
  const int eb = ew_blocks(n);
  long long cg_iters = 0;
  auto upload = [&](const double *host, double2 *dst) -> int {
    TB_CUDA(cudaMemcpyAsync(ctx->stage, host, n * sizeof(double2), cudaMemcpyHostToDevice, st));
    return tb_launch_pack(ctx, ctx->stage, dst);
  };
  // nothing in the trajectory waits for the host: failures and iteration counts are collected on the device
  auto after_solve = [&]() -> int {
    or_failed_kernel<<<(C + 255) / 256, 256, 0, st>>>(H.failed, ctx->cg.status, ctx->cg.iters, H.iter_sum, C);
    ctx->launches++;
    TB_CUDA(cudaGetLastError());
    return TB_OK;
  };
  TB_CUDA(cudaMemsetAsync(H.failed, 0, cp * sizeof(int), st));
  TB_CUDA(cudaMemsetAsync(H.iter_sum, 0, sizeof(unsigned long long), st));

  // links of the current configuration (a previous rejected trajectory leaves them consistent, but be safe)
  TB_CHECK(links_from(ctx, ctx->Adev));
  // random_pseudofermion, hmc.c:418-436: Smdm = |xi|^2, psi = M~ xi
  if (xi_host) TB_CHECK(upload(xi_host, H.gauss));
  else { fill_gauss_kernel<<<eb, 256, 0, st>>>(H.gauss, n, C, RngKey{seed, traj_index, 1u, ctx->hmc_chain_offset}); ctx->launches++; }
  TB_CHECK(dot_to(ctx, H.gauss, H.gauss, 1));
  TB_CHECK(tb_launch_dslash(ctx, dag, H.gauss, H.psi, false));
  // random_momentum, hmc.c:483-499: Smom = sum p^2
  if (mom_host) TB_CHECK(upload(mom_host, H.mom));
  else { fill_gauss_kernel<<<eb, 256, 0, st>>>(H.mom, n, C, RngKey{seed, traj_index, 2u, ctx->hmc_chain_offset}); ctx->launches++; }
  TB_CHECK(dot_to(ctx, H.mom, H.mom, 3));
  // calc_gauge_action, hmc.c:697
  TB_CHECK(gauge_action_to(ctx, ctx->Adev, 0));
  // stochastic_vector + stochastic_md_action, hmc.c:698-699: Smd = Re<st, M~ st>
  if (st_host) TB_CHECK(upload(st_host, H.st));
  else { fill_gauss_kernel<<<eb, 256, 0, st>>>(H.st, n, C, RngKey{seed, traj_index, 3u, ctx->hmc_chain_offset}); ctx->launches++; }
  TB_CHECK(tb_launch_dslash(ctx, dag, H.st, ctx->tmp, false));
  TB_CHECK(dot_to(ctx, H.st, ctx->tmp, 2));
  if (u_host) TB_CUDA(cudaMemcpyAsync(H.u, u_host, C * sizeof(double), cudaMemcpyHostToDevice, st));
  else { fill_uniform_kernel<<<(C + 255) / 256, 256, 0, st>>>(H.u, C, RngKey{seed, traj_index, 4u, ctx->hmc_chain_offset}); ctx->launches++; }
  // new_A = A, hmc.c:703-705
  TB_CUDA(cudaMemcpyAsync(H.newA, ctx->Adev, n * sizeof(double2), cudaMemcpyDeviceToDevice, st));

  const double eps_q = traj_length * 0.5 / nsteps;   // hmc.c:712,714
  const double eps_p = traj_length / nsteps;         // hmc.c:713
  for (int i = 0; i < nsteps; i++) {
    gauge_step_links_kernel<<<eb, 256, 0, st>>>(H.newA, H.mom, ctx->W0, ctx->W1, 2.0 * eps_q, ctx->nt, ctx->nx, C,
                                                ctx->t_off, ctx->nt_global);
    ctx->launches++;
    // momentum_step, hmc.c:504: chi = (M~M)^-1 psi, phi = M chi, forces
    TB_CHECK(tb_run_cg_async(ctx, H.psi, H.chi));
    TB_CHECK(after_solve());
    TB_CHECK(tb_launch_dslash(ctx, false, H.chi, H.phi, false));
    force_kernel<<<eb, 256, 0, st>>>(H.mom, H.newA, H.chi, H.phi, H.st, H.nf_over_g, ctx->d_emu, ctx->d_emmu, eps_p,
                                     ctx->nt, ctx->nx, C);
    gauge_step_links_kernel<<<eb, 256, 0, st>>>(H.newA, H.mom, ctx->W0, ctx->W1, 2.0 * eps_q, ctx->nt, ctx->nx, C,
                                                ctx->t_off, ctx->nt_global);
    ctx->launches += 2;
    TB_CUDA(cudaGetLastError());
  }
  // pseudofermion_action on the proposed field, hmc.c:719
  TB_CHECK(tb_run_cg_async(ctx, H.psi, H.chi));
  TB_CHECK(after_solve());
  TB_CHECK(dot_to(ctx, H.psi, H.chi, 5));
  TB_CHECK(dot_to(ctx, H.mom, H.mom, 7));               // hmc.c:721-724
  TB_CHECK(gauge_action_to(ctx, H.newA, 4));            // hmc.c:725
  TB_CHECK(tb_launch_dslash(ctx, dag, H.st, ctx->tmp, false));
  TB_CHECK(dot_to(ctx, H.st, ctx->tmp, 6));             // hmc.c:726
  accept_kernel<<<(C + 255) / 256, 256, 0, st>>>(H.obs, H.sums, H.nf_over_g, H.u, H.failed, H.accept, C, (int)cp);
  commit_kernel<<<eb, 256, 0, st>>>(ctx->Adev, H.newA, H.accept, n, C);
  ctx->launches += 2;
  TB_CHECK(links_from(ctx, ctx->Adev));
  TB_CUDA(cudaGetLastError());
  if (obs_host || accepted_host) {
    double *tmp = (double *)malloc((size_t)C * 10 * sizeof(double));
    TB_CUDA(cudaMemcpyAsync(tmp, H.obs, (size_t)C * 10 * sizeof(double), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    if (obs_host) memcpy(obs_host, tmp, (size_t)C * 10 * sizeof(double));
    if (accepted_host)
      for (int c = 0; c < C; c++) accepted_host[c] = (int)tmp[(size_t)c * 10 + 9];
    free(tmp);
  } else {
    TB_CUDA(cudaStreamSynchronize(st));
  }
  if (cg_iters_host) {
    unsigned long long total = 0;
    TB_CUDA(cudaMemcpy(&total, H.iter_sum, sizeof(total), cudaMemcpyDeviceToHost));
    *cg_iters_host = cg_iters + (long long)total;
  }
  return TB_OK;
}

// The force of one momentum step, exposed for checking: dS/dA for every link of every chain on the context's current
// field, exactly what momentum_step subtracts (times eps) from the momenta (hmc.c:504-661): the gauge force
// (Nf/g) sin A, the pseudofermion force of Re<psi, (M~M)^-1 psi> through chi = (M~M)^-1 psi and M chi, and the force
// of the stochastic Re<st, M~ st> term as coded.  psi_host: complex [chain][t][x]; st_host may be NULL (no such term).
// force_host: real [chain][t][x][2].  This is what the reference's disabled CHECK_FORCE block (hmc.c:502,535-559)
// compares with a finite difference of pseudofermion_action.
extern "C" int tb_hmc_force(tb_ctx *ctx, const double *psi_host, const double *st_host, double *force_host) {
  if (!ctx || !psi_host || !force_host) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->have_gauge) { tb_set_error("tb_hmc_force: no gauge field"); return TB_EINVAL; }
  TB_CHECK(hmc_alloc(ctx));
  TB_CHECK(tb_synchronize(ctx));
  cudaStream_t st = ctx->stream;
  const size_t n = ctx->nsite;
  auto &H = ctx->hmc;
  auto upload = [&](const double *host, double2 *dst) -> int {
    TB_CUDA(cudaMemcpyAsync(ctx->stage, host, n * sizeof(double2), cudaMemcpyHostToDevice, st));
    return tb_launch_pack(ctx, ctx->stage, dst);
  };
  TB_CHECK(links_from(ctx, ctx->Adev));
  TB_CHECK(upload(psi_host, H.psi));
  if (st_host) TB_CHECK(upload(st_host, H.st));
  else TB_CUDA(cudaMemsetAsync(H.st, 0, n * sizeof(double2), st));
  TB_CUDA(cudaMemsetAsync(H.mom, 0, n * sizeof(double2), st));
  TB_CUDA(cudaMemcpyAsync(H.newA, ctx->Adev, n * sizeof(double2), cudaMemcpyDeviceToDevice, st));
  TB_CHECK(tb_run_cg_any(ctx, H.psi, H.chi));                          // hmc.c:515
  TB_CHECK(tb_launch_dslash(ctx, false, H.chi, H.phi, false));         // hmc.c:516
  force_kernel<<<ew_blocks(n), 256, 0, st>>>(H.mom, H.newA, H.chi, H.phi, H.st, H.nf_over_g, ctx->d_emu, ctx->d_emmu, 1.0,
                                              ctx->nt, ctx->nx, ctx->C);   // mom = 0 - 1 * force
  ctx->launches++;
  TB_CUDA(cudaGetLastError());
  TB_CHECK(tb_launch_unpack(ctx, H.mom, (double *)ctx->stage_x));
  TB_CUDA(cudaMemcpyAsync(force_host, ctx->stage_x, n * sizeof(double2), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  for (size_t i = 0; i < 2 * n; i++) force_host[i] = -force_host[i];
  return TB_OK;
}

// per chain: bit TB_CG_MAXITER / TB_CG_DIVERGED set when a solve of the last trajectory ended that way
extern "C" int tb_hmc_cg_failures(tb_ctx *ctx, int *mask_host) {
  if (!ctx || !mask_host) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->hmc.failed) { tb_set_error("tb_hmc_cg_failures: no trajectory has run"); return TB_EINVAL; }
  TB_CUDA(cudaMemcpyAsync(mask_host, ctx->hmc.failed, ctx->C * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  TB_CUDA(cudaStreamSynchronize(ctx->stream));
  return TB_OK;
}

// current gauge field (angles) to the host, canonical layout [chain][t][x][2]
extern "C" int tb_get_gauge(tb_ctx *ctx, double *A_host) {
  if (!ctx || !A_host) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CHECK(tb_synchronize(ctx));
  TB_CHECK(tb_launch_unpack(ctx, ctx->Adev, (double *)ctx->stage_x));
  TB_CUDA(cudaMemcpyAsync(A_host, ctx->stage_x, ctx->nsite * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
  TB_CUDA(cudaStreamSynchronize(ctx->stream));
  return TB_OK;
}

// measure() of hmc.c:823-842 for every chain: Magnetisation = sum A / V and Phase = (1/nsrc) sum_i Im<c_i, M~ c_i>
// over nsrc fresh stochastic vectors (fermion_phase, hmc.c:794-815; the reference uses 20).
// sources_host (optional, parity tests): complex [nsrc][chain][t][x].
extern "C" int tb_hmc_measure(tb_ctx *ctx, int nsrc, unsigned long long seed, unsigned int meas_index,
                              const double *sources_host, double *magnetisation_host, double *phase_host) {
  if (!ctx || nsrc < 0 || nsrc >= TB_RNG_MAX_SOURCES) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->have_gauge) { tb_set_error("tb_hmc_measure: no gauge field"); return TB_EINVAL; }
  TB_CHECK(hmc_alloc(ctx));
  TB_CHECK(tb_synchronize(ctx));
  cudaStream_t st = ctx->stream;
  const TbGeom &g = ctx->g;
  const size_t n = ctx->nsite, cp = g.Cpad;
  const int C = ctx->C;
  auto &H = ctx->hmc;
  TbCgState s = ctx->cg;
  s.dot = H.sums + 8 * cp;
  TB_DISPATCH_TT(g.tt, (magnetisation_kernel<TT><<<grid_of(g), g.bc * g.bx, 0, st>>>(ctx->Adev, g, s, ctx->slab)))
  ctx->launches++;
  double *hm = (double *)malloc(2 * (size_t)C * sizeof(double)), *hp = hm + C;
  TB_CUDA(cudaMemcpyAsync(hm, H.sums + 8 * cp, C * sizeof(double), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  for (int c = 0; c < C; c++) {
    if (magnetisation_host) magnetisation_host[c] = hm[c] / (double)ctx->V;   // hmc.c:839
    if (phase_host) phase_host[c] = 0.0;
  }
  for (int i = 0; i < nsrc; i++) {
    if (sources_host) {
      TB_CUDA(cudaMemcpyAsync(ctx->stage, sources_host + (size_t)i * 2 * n, n * sizeof(double2), cudaMemcpyHostToDevice, st));
      TB_CHECK(tb_launch_pack(ctx, ctx->stage, H.gauss));
    } else {
      fill_gauss_kernel<<<ew_blocks(n), 256, 0, st>>>(H.gauss, n, C, RngKey{seed, meas_index, TB_RNG_MEASURE + (unsigned int)i, ctx->hmc_chain_offset});
      ctx->launches++;
    }
    TB_CHECK(tb_launch_dslash(ctx, tb_conj_is_dagger(ctx), H.gauss, ctx->tmp, false));
    s.dot = H.sums + 9 * cp;
    TB_DISPATCH_TT(g.tt, (im_dot_kernel<TT><<<grid_of(g), g.bc * g.bx, 0, st>>>(H.gauss, ctx->tmp, g, s, ctx->slab)))
    ctx->launches++;
    TB_CUDA(cudaMemcpyAsync(hp, H.sums + 9 * cp, C * sizeof(double), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    if (phase_host)
      for (int c = 0; c < C; c++) phase_host[c] += hp[c];
  }
  if (phase_host && nsrc > 0)
    for (int c = 0; c < C; c++) phase_host[c] /= (double)nsrc;   // hmc.c:814
  free(hm);
  return TB_OK;
}

// Chiral condensate (absent from the reference's measure(), F6): per chain
//   <psibar psi> = (1/V) Tr M^-1  ~  (1/(2 V nsrc)) sum_i Re <eta_i, M^-1 eta_i>,   M^-1 eta = (M~M)^-1 M~ eta
// through fm_invert_cg (hmc.c:408-414) on nsrc stochastic vectors eta_i (stochastic_vector, hmc.c:439-447: real
// and imaginary parts N(0,1), so E|eta|^2 = 2 per site, hence the 1/2).  One batched solve per source covers
// every chain / (g, m) point of the context.  sources_host (optional): complex [nsrc][chain][t][x].
// cg_iters_host (optional): CG iterations summed over chains and sources.
extern "C" int tb_hmc_condensate(tb_ctx *ctx, int nsrc, unsigned long long seed, unsigned int meas_index,
                                 const double *sources_host, double *condensate_host, long long *cg_iters_host) {
  if (!ctx || nsrc < 1 || nsrc >= TB_RNG_MAX_SOURCES || !condensate_host) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->have_gauge) { tb_set_error("tb_hmc_condensate: no gauge field"); return TB_EINVAL; }
  TB_CHECK(hmc_alloc(ctx));
  TB_CHECK(tb_synchronize(ctx));
  cudaStream_t st = ctx->stream;
  const size_t n = ctx->nsite, cp = ctx->g.Cpad;
  const int C = ctx->C;
  auto &H = ctx->hmc;
  long long cg_iters = 0;
  double *acc = (double *)calloc((size_t)C, sizeof(double));
  double *hd = (double *)malloc((size_t)C * sizeof(double));
  int rc = TB_OK;
  for (int i = 0; i < nsrc && rc == TB_OK; i++) {
    auto one = [&]() -> int {
      if (sources_host) {
        TB_CUDA(cudaMemcpyAsync(ctx->stage, sources_host + (size_t)i * 2 * n, n * sizeof(double2), cudaMemcpyHostToDevice, st));
        TB_CHECK(tb_launch_pack(ctx, ctx->stage, H.gauss));
      } else {
        fill_gauss_kernel<<<ew_blocks(n), 256, 0, st>>>(H.gauss, n, C, RngKey{seed, meas_index, TB_RNG_CONDENSATE + (unsigned int)i, ctx->hmc_chain_offset});
        ctx->launches++;
      }
      TB_CHECK(tb_launch_dslash(ctx, tb_conj_is_dagger(ctx), H.gauss, ctx->tmp, false));   // hmc.c:410
      TB_CHECK(tb_run_cg_any(ctx, ctx->tmp, H.chi));                                        // hmc.c:411
      TB_CHECK(dot_to(ctx, H.gauss, H.chi, 10));
      TB_CUDA(cudaMemcpyAsync(hd, H.sums + 10 * cp, C * sizeof(double), cudaMemcpyDeviceToHost, st));
      TB_CUDA(cudaMemcpyAsync(ctx->h_iters, ctx->cg.iters, C * sizeof(int), cudaMemcpyDeviceToHost, st));
      TB_CUDA(cudaMemcpyAsync(ctx->h_status, ctx->cg.status, C * sizeof(int), cudaMemcpyDeviceToHost, st));
      TB_CUDA(cudaStreamSynchronize(st));
      return TB_OK;
    };
    rc = one();
    if (rc != TB_OK) break;
    for (int c = 0; c < C; c++) {
      // a chain whose solve failed poisons its estimate instead of silently biasing it
      acc[c] += (ctx->h_status[c] == TB_CG_CONVERGED || ctx->h_status[c] == TB_CG_ZERO_SOURCE) ? hd[c] : NAN;
      cg_iters += ctx->h_iters[c];
    }
  }
  if (rc == TB_OK)
    for (int c = 0; c < C; c++) condensate_host[c] = acc[c] / (2.0 * (double)ctx->V * (double)nsrc);
  if (cg_iters_host) *cg_iters_host = cg_iters;
  free(acc);
  free(hd);
  return rc;
}

// ---- on-disk format -----------------------------------------------------------------------
// hmc.c never writes its configuration; the checkpoint mirrors fermionbag's raw dump idea (fermionbag.c:125-161):
// a 64-byte header (magic, NT, NX, nchains, mode, index of the next trajectory) followed by the raw FP64 angles
// A[chain][t][x][dir].  The trajectory index keys the device random stream (momenta, pseudofermion noise, Metropolis
// uniforms, measurement sources): a resumed run must continue it, or it replays the noise of its first leg.
struct TbCkptHeader {
  char magic[16];
  int nt, nx, nchains, mode;
  unsigned int next_traj;   // 0 = not recorded (files written before the field existed)
  char pad[28];
};
static_assert(sizeof(TbCkptHeader) == 64, "checkpoint header is 64 bytes");

extern "C" int tb_checkpoint_set_next_trajectory(tb_ctx *ctx, unsigned int next_traj) {
  if (!ctx) return TB_EINVAL;
  ctx->ckpt_next_traj = next_traj;
  return TB_OK;
}

extern "C" int tb_checkpoint_next_trajectory(const tb_ctx *ctx, unsigned int *next_traj) {
  if (!ctx || !next_traj) return TB_EINVAL;
  *next_traj = ctx->ckpt_next_traj;
  return TB_OK;
}

extern "C" int tb_checkpoint_write(tb_ctx *ctx, const char *path) {
  if (!ctx || !path) return TB_EINVAL;
  if (!ctx->have_gauge) { tb_set_error("tb_checkpoint_write: no gauge field"); return TB_EINVAL; }
  double *A = (double *)malloc(ctx->nsite * 2 * sizeof(double));
  if (!A) return TB_ENOMEM;
  int rc = tb_get_gauge(ctx, A);
  if (rc == TB_OK) {
    FILE *f = fopen(path, "wb");
    if (!f) { tb_set_error("tb_checkpoint_write: cannot open %s", path); rc = TB_EINVAL; }
    else {
      TbCkptHeader h;
      memset(&h, 0, sizeof(h));
      memcpy(h.magic, "THIRRING2D-A-V1", 15);
      h.nt = ctx->nt; h.nx = ctx->nx; h.nchains = ctx->C; h.mode = ctx->mode;
      h.next_traj = ctx->ckpt_next_traj;
      if (fwrite(&h, sizeof(h), 1, f) != 1 || fwrite(A, sizeof(double), ctx->nsite * 2, f) != ctx->nsite * 2) {
        tb_set_error("tb_checkpoint_write: short write to %s", path);
        rc = TB_EINVAL;
      }
      fclose(f);
    }
  }
  free(A);
  return rc;
}

extern "C" int tb_checkpoint_read(tb_ctx *ctx, const char *path) {
  if (!ctx || !path) return TB_EINVAL;
  FILE *f = fopen(path, "rb");
  if (!f) { tb_set_error("tb_checkpoint_read: cannot open %s", path); return TB_EINVAL; }
  TbCkptHeader h;
  int rc = TB_OK;
  if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "THIRRING2D-A-V1", 15) != 0) {
    tb_set_error("tb_checkpoint_read: %s is not a Thirring2D gauge checkpoint", path);
    rc = TB_EINVAL;
  } else if (h.nt != ctx->nt || h.nx != ctx->nx || h.nchains != ctx->C) {
    tb_set_error("tb_checkpoint_read: %s holds %d chains of %dx%d, the context %d of %dx%d", path, h.nchains, h.nt,
                 h.nx, ctx->C, ctx->nt, ctx->nx);
    rc = TB_EINVAL;
  }
  double *A = nullptr;
  if (rc == TB_OK) {
    A = (double *)malloc(ctx->nsite * 2 * sizeof(double));
    if (fread(A, sizeof(double), ctx->nsite * 2, f) != ctx->nsite * 2) {
      tb_set_error("tb_checkpoint_read: %s is truncated", path);
      rc = TB_EINVAL;
    }
  }
  fclose(f);
  if (rc == TB_OK) rc = tb_set_gauge(ctx, A);
  if (rc == TB_OK) rc = tb_synchronize(ctx);
  if (rc == TB_OK) ctx->ckpt_next_traj = h.next_traj;
  free(A);
  return rc;
}
