// Probe: FP64 tensor-core (mma.sync m8n8k4.f64 -> SASS DMMA.8x8x4) vs DFMA throughput on one GPU, 4..32 resident warps per SM,
// 8 independent accumulator fragments per warp.  Build and run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/probe_dmma.cu -o /tmp/probe_dmma && /tmp/probe_dmma
// Result on the round-1 B200 (profiles/probe_dmma_r1.txt): DMMA 37.1-37.2 TFLOP/s, DFMA 32.7-34.2 TFLOP/s -- the FP64 tensor
// path has the same peak as the vector path but needs 8x fewer issue slots (basis of elem_kernel_mma, pmb_elem.cu).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k_dmma(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3 + 1.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(double* out, int iters) {
  double c[16];
  for (int i = 0; i < 16; ++i) c[i] = 0.0;
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3 + 1.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(a, b, c[i]);
  }
  double s = 0; for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 4; warps <= 32; warps *= 2) {
    int threads = 256, blocks = 148 * warps * 32 / threads; if (blocks < 148) { blocks = 148; threads = warps * 32; }
    int iters = 20000;
    for (int which = 0; which < 2; ++which) {
      float best = 1e30f;
      for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        if (which == 0) k_dmma<<<blocks, threads>>>(out, iters); else k_dfma<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      double fma_total = which == 0 ? (double)blocks * threads / 32 * iters * 8 * 256 : (double)blocks * threads * iters * 16;
      printf("%s warps/SM %d: %.3f ms  %.2f TFLOP/s\n", which == 0 ? "DMMA.m8n8k4" : "DFMA", warps, best, 2 * fma_total / best / 1e9);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
