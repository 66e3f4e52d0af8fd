// nvlink_pingpong.cu — what a cross-GPU hand-over costs on this box (development aid; the floor under the slab solve's
// all-reduce).  GPU 0 and GPU 1 bounce a counter through peer stores: each side spins on a flag in ITS memory that the
// other side writes over NVLink.  Variants: a bare volatile store; the release pattern of the solver (system-scope
// fence in front of the store, another behind the load that saw it).  Prints the one-way latency = round trip / 2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/nvlink_pingpong tools/nvlink_pingpong.cu
#include <cuda_runtime.h>

#include <cstdio>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// F: 0 none, 1 fence.acq_rel.gpu, 2 fence.acq_rel.sys, 3 __threadfence() (fence.sc.gpu), 4 __threadfence_system() (fence.sc.sys)
template <int F>
__device__ __forceinline__ void fence() {
  if (F == 1) asm volatile("fence.acq_rel.gpu;" ::: "memory");
  if (F == 2) asm volatile("fence.acq_rel.sys;" ::: "memory");
  if (F == 3) __threadfence();
  if (F == 4) __threadfence_system();
}

// one fence in front of every store (release) and one behind every load that saw the flag (acquire): 4 per round trip
template <int F>
__global__ void bounce(volatile int *mine, volatile int *theirs, int first, int rounds) {
  for (int k = 1; k <= rounds; k++) {
    if (first) {
      fence<F>();
      *theirs = k;
      while (*mine < k) {}
      fence<F>();
    } else {
      while (*mine < k) {}
      fence<F>();
      fence<F>();
      *theirs = k;
    }
  }
}

template <int F>
static void launch(volatile int *mine, volatile int *theirs, int first, int rounds, cudaStream_t st) {
  bounce<F><<<1, 1, 0, st>>>(mine, theirs, first, rounds);
}

int main() {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
  int *f[2];
  cudaStream_t st[2];
  cudaEvent_t e0, e1;
  for (int d = 0; d < 2; d++) {
    CK(cudaSetDevice(d));
    CK(cudaDeviceEnablePeerAccess(1 - d, 0));
    CK(cudaMalloc((void **)&f[d], 256));
    CK(cudaStreamCreate(&st[d]));
  }
  CK(cudaSetDevice(0));
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int rounds = 20000;
  const char *names[5] = {"bare volatile store / load", "fence.acq_rel.gpu", "fence.acq_rel.sys", "__threadfence() = fence.sc.gpu",
                          "__threadfence_system() = fence.sc.sys"};
  void (*fn[5])(volatile int *, volatile int *, int, int, cudaStream_t) = {launch<0>, launch<1>, launch<2>, launch<3>, launch<4>};
  float base = 0;
  for (int fenced = 0; fenced < 5; fenced++) {
    for (int rep = 0; rep < 2; rep++) {
      for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaMemset(f[d], 0, 256)); CK(cudaDeviceSynchronize()); }
      CK(cudaSetDevice(1));
      fn[fenced](f[1], f[0], 0, rounds, st[1]);
      CK(cudaSetDevice(0));
      CK(cudaEventRecord(e0, st[0]));
      fn[fenced](f[0], f[1], 1, rounds, st[0]);
      CK(cudaEventRecord(e1, st[0]));
      CK(cudaEventSynchronize(e1));
      CK(cudaSetDevice(1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep) {
        const float rt = ms * 1e3f / rounds;
        if (fenced == 0) base = rt;
        printf("%-40s round trip %6.2f us, one way %5.2f us, per fence %5.2f us\n", names[fenced], rt, rt / 2,
               fenced ? (rt - base) / 4 : 0.f);
      }
    }
  }
  return 0;
}
