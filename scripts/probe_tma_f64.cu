// Probe: which cp.async.bulk.tensor (tile mode) configurations the B200 accepts for 8-byte elements.  Build and run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 scripts/probe_tma_f64.cu -o /tmp/probe_tma -lcuda && /tmp/probe_tma
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k2d(const __grid_constant__ CUtensorMap tm, int c0, int c1, int nbytes, double* out, int nout) {
  extern __shared__ __align__(128) double buf[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < nout; i += blockDim.x) buf[i] = -1.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(nbytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(buf)),
                 "l"(&tm), "r"(c0), "r"(c1), "r"(s32(&bar))
                 : "memory");
  }
  asm volatile(
      "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(s32(&bar))
      : "memory");
  for (int i = threadIdx.x; i < nout; i += blockDim.x) out[i] = buf[i];
}

int run(const char* name, CUtensorMapDataType dt, uint64_t d0, uint64_t d1, uint64_t stride_bytes, uint32_t b0, uint32_t b1, int c0, int c1,
        size_t base_off_doubles) {
  std::vector<double> h(d0 * d1 + 4096);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i;
  double *d, *out;
  cudaMalloc(&d, h.size() * 8);
  cudaMalloc(&out, 65536);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t dims[2] = {d0, d1}, strides[1] = {stride_bytes};
  cuuint32_t box[2] = {b0, b1}, es[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&tm, dt, 2, d + base_off_doubles, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("%-34s encode failed %d\n", name, (int)r); return 1; }
  const int esz = dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : 8;
  int nout = b0 * b1 * esz / 8;
  k2d<<<1, 128, nout * 8>>>(tm, c0, c1, nout * 8, out, nout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-34s KERNEL ERROR: %s\n", name, cudaGetErrorString(e)); return 2; }
  std::vector<double> o(nout);
  cudaMemcpy(o.data(), out, nout * 8, cudaMemcpyDeviceToHost);
  printf("%-40s ok: first %.0f [1] %.0f [4] %.0f second-row-first %.0f last %.0f\n", name, o[0], o[1], o[4], o[b0 * esz / 8], o[nout - 1]);
  cudaFree(d); cudaFree(out);
  return 0;
}

int main(int argc, char** argv) {
  cuInit(0);
  cudaFree(0);
  const int which = argc > 1 ? atoi(argv[1]) : 0;  // one case per process: a trapped kernel poisons the context
  switch (which) {
    case 0: return run("f64 1024x64 box102x5 (0,0)", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 1024, 64, 8192, 102, 5, 0, 0, 0);
    case 1: return run("f64 1024x64 box102x5 (5,1) odd coord", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 1024, 64, 8192, 102, 5, 5, 1, 0);
    case 2: return run("f64 1024x64 box102x5 (-4,0) neg even", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 1024, 64, 8192, 102, 5, -4, 0, 0);
    case 3: return run("f64 1024x64 box102x5 (-3,0) neg odd", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 1024, 64, 8192, 102, 5, -3, 0, 0);
    case 4: return run("f64 390x50 stride3120 box102x5 (194,3)", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 390, 50, 3120, 102, 5, 194, 3, 0);
    case 5: return run("f64 390x50 stride3120 box102x5 (195,3)", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 390, 50, 3120, 102, 5, 195, 3, 0);
    case 6: return run("f64 1024x64 box102x5 (1000,62) oob", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 1024, 64, 8192, 102, 5, 1000, 62, 0);
    case 7: return run("f32x2 view 2048x64 box204x5 (-6,0)", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2048, 64, 8192, 204, 5, -6, 0, 0);
    case 8: return run("f32x2 view 2048x64 box204x5 (10,1)", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2048, 64, 8192, 204, 5, 10, 1, 0);
  }
  return 0;
}
