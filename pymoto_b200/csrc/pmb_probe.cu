// libpmb: in-run FP64 peak probe.  bench.py normalises the matrix-free kernel's achieved FLOP/s by the rate THIS GPU
// sustains on a register-only stream of FP64 work, measured in the same process (MEASURED_PEAKS.json carries no FP64
// figure): kind 0 = DFMA (16 independent accumulators per thread), kind 1 = DMMA.8x8x4 (mma.sync m8n8k4 f64, 8 independent
// accumulator fragments per warp).  Not on the product path.
#include "pmb_common.cuh"

__device__ __forceinline__ void probe_dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) probe_dmma_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3 + 1.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) probe_dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) probe_dfma_kernel(double* out, int iters) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = 0.0;
  const double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3 + 1.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(a, b, c[i]);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static int probe_blocks() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms * 2;  // 2 x 256 threads = 16 warps per SM
}

extern "C" long long pmb_probe_fp64_out_doubles(void) { return 256LL * probe_blocks(); }

// tflops_out (HOST) = best of 3 launches; out: pmb_probe_fp64_out_doubles() device doubles (scratch)
extern "C" int pmb_probe_fp64(int kind, int iters, double* out, double* tflops_out, void* stream) {
  PMB_REQUIRE((kind == 0 || kind == 1) && iters > 0 && out && tflops_out, "pmb_probe_fp64: invalid argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return pmb_set_error("pmb_probe_fp64: cudaEventCreate failed");
  const int blocks = probe_blocks();
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0, st);
    if (kind == 1) probe_dmma_kernel<<<blocks, 256, 0, st>>>(out, iters);
    else probe_dfma_kernel<<<blocks, 256, 0, st>>>(out, iters);
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) {
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      return pmb_set_error("pmb_probe_fp64: %s", cudaGetErrorString(cudaGetLastError()));
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double fma = kind == 1 ? (double)blocks * 8 * iters * 8 * 256 : (double)blocks * 256 * iters * 16;
  *tflops_out = 2.0 * fma / (best * 1e-3) / 1e12;
  return 0;
}
