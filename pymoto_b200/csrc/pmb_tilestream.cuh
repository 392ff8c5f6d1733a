// Shared-memory tile streaming for the stencil-CSR layout: row-aligned tile geometry + TMA bulk-copy / mbarrier
// primitives used by the operator kernel (pmb_spmv.cu) and the Galerkin column-collapse pass (pmb_multigrid.cu).
#pragma once
#include <cstdint>
#include "pmb_common.cuh"

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- mbarrier / TMA bulk-copy primitives (PTX ISA 8.x, sm_90+; SASS: UBLKCP / SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, uint64_t* bar, bool stream_hint) {
  if (stream_hint) {
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
  } else {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
  }
}

// Tiles never straddle an x-row of nodes: tile = (row (j,k), i0 .. i0+T), so (cy, cz) are uniform per tile and a
// node's offset inside the tile is closed-form in i alone.
struct TileGeom {
  int i0, ni, j, k, cy, cz, jlo, klo;
  long long e0;  // entry offset (slab-relative) of the tile's first entry
  long long e1;  // one past its last entry
};

template <int NDOF, int T>
__device__ __forceinline__ TileGeom tile_geom(const Geo& g, int tile, int tiles_per_row) {
  TileGeom t;
  const int row = tile / tiles_per_row;
  const int ti = tile - row * tiles_per_row;
  const int kl = row / g.NY;
  t.j = row - kl * g.NY;
  t.k = kl + g.kz0;
  t.i0 = ti * T;
  t.ni = min(T, g.NX - t.i0);
  t.cy = cnt1(t.j, g.NY);
  t.cz = cnt1(t.k, g.NZ);
  t.jlo = max(t.j - 1, 0);
  t.klo = max(t.k - 1, 0);
  const long long rowbase = pre1(t.k, g.NZ) * g.Sy * g.Sx + (long long)t.cz * (pre1(t.j, g.NY) * g.Sx) - g.bo0;
  const long long per = (long long)(NDOF * NDOF) * t.cy * t.cz;
  t.e0 = (long long)(NDOF * NDOF) * rowbase + per * pre1(t.i0, g.NX);
  t.e1 = (long long)(NDOF * NDOF) * rowbase + per * pre1(t.i0 + t.ni, g.NX);
  return t;
}


// final deterministic reduction of the per-CTA partial triples
static __global__ void __launch_bounds__(1024) reduce_triples_kernel(const double* __restrict__ partials, long long nblocks,
                                                               double* __restrict__ out) {
  __shared__ double s[3][32];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (long long v = threadIdx.x; v < nblocks; v += 1024)
    a0 += partials[3 * v], a1 += partials[3 * v + 1], a2 += partials[3 * v + 2];
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  a2 = warp_sum(a2);
  if ((threadIdx.x & 31) == 0) s[0][threadIdx.x >> 5] = a0, s[1][threadIdx.x >> 5] = a1, s[2][threadIdx.x >> 5] = a2;
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int v = 0; v < 32; ++v) t += s[threadIdx.x][v];
    out[threadIdx.x] = t;
  }
}

