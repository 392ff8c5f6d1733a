// libpmb: density filter stencil (K10).
//
// Replaces the sparse products with H of pymoto/modules/filter.py:266-270 (H built by the Python loop of
// :307-378).  H is never formed: the cone weights max(0, r - dist) live in a (2d+1)^dim table and the window is
// clipped to the domain exactly like the reference's (no padding).  Each output accumulates w*x over its window
// in ascending element number with a separate multiply and add, which is what scipy's csc_matvec does for a row
// of H, so forward values, row sums Hs and the backward pass are bit-identical to the reference.
//
// A CTA produces a TX x TY x TZ brick of outputs from a shared-memory tile with a d-wide apron.
#include "pmb_common.cuh"

static constexpr int FTX = 32, FTY = 4, FTZ = 4;

__global__ void __launch_bounds__(FTX* FTY* FTZ) filter_kernel(int nx, int ny, int nzE, int ez0, int nezl, int d,
                                                                const double* __restrict__ wtab, const double* __restrict__ in,
                                                                const double* __restrict__ hs, double* __restrict__ out,
                                                                int has_z) {
  extern __shared__ double tile[];
  const int dz = has_z ? d : 0;
  const int SX = FTX + 2 * d, SY = FTY + 2 * d, SZ = FTZ + 2 * dz;
  const int W = 2 * d + 1;
  double* swt = tile + (size_t)SX * SY * SZ;  // weight table copy
  const int nw = W * W * (has_z ? W : 1);

  const int bx = blockIdx.x * FTX, by = blockIdx.y * FTY, bz = blockIdx.z * FTZ;  // bz relative to ez0
  const int tid = (threadIdx.z * FTY + threadIdx.y) * FTX + threadIdx.x;
  const int NT = FTX * FTY * FTZ;

  for (int q = tid; q < nw; q += NT) swt[q] = wtab[q];
  // stage the input brick + apron (zeros outside the domain are never used: the loops below clip)
  const int tot = SX * SY * SZ;
  for (int q = tid; q < tot; q += NT) {
    int sx = q % SX, sy = (q / SX) % SY, sz = q / (SX * SY);
    int gx = bx + sx - d, gy = by + sy - d, gz = ez0 + bz + sz - dz;  // global element indices
    double v = 0.0;
    if (gx >= 0 && gx < nx && gy >= 0 && gy < ny && gz >= 0 && gz < nzE)
      v = in ? __ldg(in + ((long long)(gz - ez0) * ny + gy) * nx + gx) : 1.0;
    tile[q] = v;
  }
  __syncthreads();

  const int ex = bx + threadIdx.x, ey = by + threadIdx.y, ezl = bz + threadIdx.z;
  if (ex >= nx || ey >= ny || ezl >= nezl) return;
  const int ez = ez0 + ezl;
  const int x0 = max(ex - d, 0), x1 = min(ex + d, nx - 1);
  const int y0 = max(ey - d, 0), y1 = min(ey + d, ny - 1);
  const int z0 = has_z ? max(ez - d, 0) : 0, z1 = has_z ? min(ez + d, nzE - 1) : 0;
  double acc = 0.0;
  for (int z = z0; z <= z1; ++z) {
    const int sz = z - (ez0 + bz) + dz;
    const int wz = has_z ? (z - ez + d) : 0;
    for (int y = y0; y <= y1; ++y) {
      const int sy = y - by + d;
      const int wy = y - ey + d;
      const double* trow = tile + ((size_t)sz * SY + sy) * SX + (x0 - bx + d);
      const double* wrow = swt + ((size_t)wz * W + wy) * W + (x0 - ex + d);
      for (int x = 0; x <= x1 - x0; ++x) acc = __dadd_rn(acc, __dmul_rn(wrow[x], trow[x]));
    }
  }
  const long long e = ((long long)ezl * ny + ey) * nx + ex;
  out[e] = hs ? acc / hs[e] : acc;
}

extern "C" int pmb_filter_apply(const pmb_grid* p, int ez0, int nezl, int d, const double* wtab, const double* in,
                                const double* hs, double* out, void* stream) {
  if (validate_grid(p, "pmb_filter_apply")) return 1;
  PMB_REQUIRE(wtab && out, "pmb_filter_apply: NULL pointer argument");
  PMB_REQUIRE(d >= 0 && d <= 12, "pmb_filter_apply: window half-width %d not in 0..12", d);
  const int has_z = p->nz > 0;
  const int nzE = has_z ? p->nz : 1;
  PMB_REQUIRE(ez0 >= 0 && nezl >= 1 && ez0 + nezl <= nzE, "pmb_filter_apply: element layers [%d, %d) outside [0, %d)", ez0,
              ez0 + nezl, nzE);
  const int dz = has_z ? d : 0;
  const int W = 2 * d + 1;
  size_t smem = sizeof(double) * ((size_t)(FTX + 2 * d) * (FTY + 2 * d) * (FTZ + 2 * dz) + (size_t)W * W * (has_z ? W : 1));
  PMB_REQUIRE(smem <= 200 * 1024, "pmb_filter_apply: radius too large for the shared-memory tile (%zu bytes)", smem);
  cudaError_t e = cudaFuncSetAttribute(filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return pmb_set_error("pmb_filter_apply: %s", cudaGetErrorString(e));
  dim3 block(FTX, FTY, FTZ);
  dim3 grid((p->nx + FTX - 1) / FTX, (p->ny + FTY - 1) / FTY, (nezl + FTZ - 1) / FTZ);
  filter_kernel<<<grid, block, smem, (cudaStream_t)stream>>>(p->nx, p->ny, nzE, ez0, nezl, d, wtab, in, hs, out, has_z);
  PMB_CHECK_LAUNCH("pmb_filter_apply");
  return 0;
}
