// tb_probe.cu — measured denominators for the roofline of the on-chip solvers: the FP64 FMA issue rate of this GPU.
// The batched 64^2 solve is bound by FP64 issue, not by HBM; MEASURED_PEAKS.json has a copy bandwidth and
// a bf16 GEMM rate but no FP64 figure, so bench.py measures one with this kernel in the same run.
#include "tb_common.cuh"

namespace {

// ILP independent DFMA chains per thread, `iters` rounds: nothing but FP64 FMAs in the loop (the loop overhead is one
// integer add and a branch per 4 * ILP FMAs)
template <int ILP>
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b) {
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) v[i] = (double)(threadIdx.x + i) * 1e-3;
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) v[i] = fma(v[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += v[i];
  if (s == 123.456) out[0] = s;   // keeps the chains alive; never true for the arguments used
}

// The same with every source operand in its own register (8 accumulators, 8 + 8 multiplicands, rotated so that no two
// consecutive FMAs share a source): what a stencil's FMAs look like to the register file.  `iters` rounds of 32 FMAs.
__global__ void __launch_bounds__(256) dfma_distinct_kernel(double *out, int iters, double a0, double b0) {
  double v[8], a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    v[i] = (double)(threadIdx.x + i) * 1e-3;
    a[i] = a0 + 1e-7 * (double)(i + (threadIdx.x & 3));
    b[i] = b0 * (double)(i + 1);
  }
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = fma(a[(i + u) & 7], v[i], b[(i + 2 * u + 1) & 7]);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += v[i];
  if (s == 123.456) out[0] = s;
}

}  // namespace

// FP64 FMA throughput of the context's device in TFLOP/s (2 flops per FMA), best of `repeats` launches of about a
// half a millisecond each, CUDA events on the context's stream.
static int measure_fp64(tb_ctx *ctx, int kind, int repeats, double *tflops_out);

extern "C" int tb_measure_fp64_peak(tb_ctx *ctx, int repeats, double *tflops_out) { return measure_fp64(ctx, 0, repeats, tflops_out); }

// kind 0: the peak above (two of the three sources shared by all FMAs); kind 1: every source operand in its own register
extern "C" int tb_measure_fp64_rate(tb_ctx *ctx, int kind, int repeats, double *tflops_out) {
  if (kind != 0 && kind != 1) return TB_EINVAL;
  return measure_fp64(ctx, kind, repeats, tflops_out);
}

static int measure_fp64(tb_ctx *ctx, int kind, int repeats, double *tflops_out) {
  if (!ctx || !tflops_out || repeats < 1) return TB_EINVAL;
  TB_CUDA(cudaSetDevice(ctx->device));
  constexpr int ILP = 8;
  int nsm = TB_NUM_SMS_B200;
  TB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device));
  const int blocks = nsm * 8, threads = 256, iters = 1024;   // ~0.6 ms per launch
  double *d = nullptr;
  TB_CUDA(cudaMalloc((void **)&d, sizeof(double)));
  cudaEvent_t e0, e1;
  TB_CUDA(cudaEventCreate(&e0));
  TB_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int r = 0; r < repeats + 1; r++) {   // the first launch is a warm-up
    TB_CUDA(cudaEventRecord(e0, ctx->stream));
    if (kind == 0) dfma_peak_kernel<ILP><<<blocks, threads, 0, ctx->stream>>>(d, iters, 0.999999, 1e-9);
    else dfma_distinct_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters, 0.999999, 1e-9);
    TB_CUDA(cudaEventRecord(e1, ctx->stream));
    TB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    TB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * ILP * 4.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  ctx->launches += repeats + 1;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops_out = best;
  return TB_OK;
}
