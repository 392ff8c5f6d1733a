// Probe: does DMMA.8x8x4 keep its 37 TFLOP/s when every instruction reads DIFFERENT A / B registers (as a real kernel does),
// or only when the operands repeat (scripts/probe_dmma.cu reuses one a / b pair)?  16 warps / SM, 8 independent accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/probe_dmma_operands.cu -o /tmp/probe_dmma_ops && /tmp/probe_dmma_ops
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NA, int NB>
__global__ void __launch_bounds__(256) k(double* out, int iters) {
  double c[8][2], a[8], b[8];
  for (int i = 0; i < 8; ++i) { c[i][0] = c[i][1] = 0.0; a[i] = threadIdx.x * 1e-3 + i; b[i] = threadIdx.x * 2e-3 + 1.0 + 0.5 * i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a[i % NA], b[i % NB]);
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NA, int NB> void run(double* out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = 148 * 2, iters = 20000;
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); k<NA, NB><<<blocks, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  printf("distinct A regs %d, distinct B regs %d: %.3f ms  %.2f TFLOP/s\n", NA, NB, best, 2.0 * blocks * 8 * iters * 8 * 256 / best / 1e9);
}
int main() {
  double* out; cudaMalloc(&out, 148 * 2 * 256 * 8);
  run<1, 1>(out); run<1, 8>(out); run<8, 1>(out); run<8, 8>(out); run<4, 2>(out); run<2, 4>(out);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
