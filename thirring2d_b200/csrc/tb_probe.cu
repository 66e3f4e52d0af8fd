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

}  // namespace

// FP64 FMA throughput of the context's device in TFLOP/s (2 flops per FMA), best of `repeats` launches of about a
// half a millisecond each, CUDA events on the context's stream.
extern "C" int tb_measure_fp64_peak(tb_ctx *ctx, int repeats, double *tflops_out) {
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
    dfma_peak_kernel<ILP><<<blocks, threads, 0, ctx->stream>>>(d, iters, 0.999999, 1e-9);
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
