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
    // (layers further than dz from the produced range are never used and, on a slab, may not be mapped)
    if (gx >= 0 && gx < nx && gy >= 0 && gy < ny && gz >= 0 && gz < nzE && gz >= ez0 - dz && gz < ez0 + nezl + dz)
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
      // zero weights (98 of the 125 window slots at r = 2) are skipped: adding w * x = +-0 never changes the sum, so the
      // result keeps the bits of the reference's csc_matvec; the test is warp-uniform away from the domain faces
      for (int x = 0; x <= x1 - x0; ++x) {
        const double wv = wrow[x];
        if (wv != 0.0) acc = __dadd_rn(acc, __dmul_rn(wv, trow[x]));
      }
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

// ------------------------------------------------------------------------------------------------- FilterConv
// Padded convolution filter of pymoto/modules/filter.py:8-220.  The reference materialises an index array el3d_pad
// ((n+2p)^3 int64) and calls scipy.signal.convolve / correlate; here the padding is a separable per-axis index map
// (symmetric / edge / wrap / constant, built on the host with the reference's own np.pad sequence), applied by a
// gather kernel, and the convolution is a direct stencil.

// xpad[pz][py][px] = constant of the outermost constant-padded axis (z wins over y over x, the order in which the
// reference applies its overrides, filter.py:99-160,183-187), else x[(mz*ny + my)*nx + mx]
__global__ void __launch_bounds__(256) pad_gather_kernel(int nx, int ny, int px, int py, int pz, const int* __restrict__ mapx,
                                                          const int* __restrict__ mapy, const int* __restrict__ mapz,
                                                          const double* __restrict__ cvx, const double* __restrict__ cvy,
                                                          const double* __restrict__ cvz, const double* __restrict__ x,
                                                          double* __restrict__ xpad) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long n = (long long)px * py * pz;
  if (t >= n) return;
  int ix = (int)(t % px), iy = (int)((t / px) % py), iz = (int)(t / ((long long)px * py));
  int mx = mapx[ix], my = mapy[iy], mz = mapz[iz];
  double v;
  if (mz < 0) v = cvz[iz];
  else if (my < 0) v = cvy[iy];
  else if (mx < 0) v = cvx[ix];
  else v = x[((long long)mz * ny + my) * nx + mx];
  xpad[t] = v;
}

extern "C" int pmb_pad_gather(int nx, int ny, int nz, int px, int py, int pz, const int* mapx, const int* mapy, const int* mapz,
                              const double* cvx, const double* cvy, const double* cvz, const double* x, double* xpad,
                              void* stream) {
  PMB_REQUIRE(nx > 0 && ny > 0 && nz > 0 && px > 0 && py > 0 && pz > 0, "pmb_pad_gather: invalid sizes");
  PMB_REQUIRE(mapx && mapy && mapz && cvx && cvy && cvz && x && xpad, "pmb_pad_gather: NULL pointer argument");
  long long n = (long long)px * py * pz;
  pad_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nx, ny, px, py, pz, mapx, mapy, mapz, cvx, cvy,
                                                                                     cvz, x, xpad);
  PMB_CHECK_LAUNCH("pmb_pad_gather");
  return 0;
}

// dx[s] = sum over the padded positions that map to s (per-axis inverse lists in CSR form) of dxpad
__global__ void __launch_bounds__(256) pad_scatter_kernel(int nx, int ny, int nz, int px, int py, const int* __restrict__ ptrx,
                                                           const int* __restrict__ lstx, const int* __restrict__ ptry,
                                                           const int* __restrict__ lsty, const int* __restrict__ ptrz,
                                                           const int* __restrict__ lstz, const double* __restrict__ dxpad,
                                                           double* __restrict__ dx) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long n = (long long)nx * ny * nz;
  if (t >= n) return;
  int ix = (int)(t % nx), iy = (int)((t / nx) % ny), iz = (int)(t / ((long long)nx * ny));
  double acc = 0.0;
  for (int a = ptrz[iz]; a < ptrz[iz + 1]; ++a)
    for (int b = ptry[iy]; b < ptry[iy + 1]; ++b)
      for (int c = ptrx[ix]; c < ptrx[ix + 1]; ++c) acc += dxpad[((long long)lstz[a] * py + lsty[b]) * px + lstx[c]];
  dx[t] = acc;
}

extern "C" int pmb_pad_scatter(int nx, int ny, int nz, int px, int py, int pz, const int* ptrx, const int* lstx, const int* ptry,
                               const int* lsty, const int* ptrz, const int* lstz, const double* dxpad, double* dx, void* stream) {
  PMB_REQUIRE(nx > 0 && ny > 0 && nz > 0 && px > 0 && py > 0 && pz > 0, "pmb_pad_scatter: invalid sizes");
  PMB_REQUIRE(ptrx && lstx && ptry && lsty && ptrz && lstz && dxpad && dx, "pmb_pad_scatter: NULL pointer argument");
  long long n = (long long)nx * ny * nz;
  pad_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nx, ny, nz, px, py, ptrx, lstx, ptry, lsty, ptrz,
                                                                                     lstz, dxpad, dx);
  PMB_CHECK_LAUNCH("pmb_pad_scatter");
  return 0;
}

// out[o] = sum_q w[q] in[o + q - off]  (per axis), `in` taken as zero outside its extent.  Forward filter: in = padded
// field, off = 0, w = flipped weights ("valid" convolution); backward: in = dy, off = K-1, w = weights ("full" correlation).
__global__ void __launch_bounds__(256) stencil_corr_kernel(int inx, int iny, int inz, const double* __restrict__ in, int ox, int oy,
                                                            int oz, double* __restrict__ out, int kx, int ky, int kz,
                                                            const double* __restrict__ w, int offx, int offy, int offz) {
  extern __shared__ double sw[];
  for (int q = threadIdx.x; q < kx * ky * kz; q += blockDim.x) sw[q] = w[q];
  __syncthreads();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long n = (long long)ox * oy * oz;
  if (t >= n) return;
  int x = (int)(t % ox), y = (int)((t / ox) % oy), z = (int)(t / ((long long)ox * oy));
  double acc = 0.0;
  for (int c = 0; c < kz; ++c) {
    int zz = z + c - offz;
    if (zz < 0 || zz >= inz) continue;
    for (int b = 0; b < ky; ++b) {
      int yy = y + b - offy;
      if (yy < 0 || yy >= iny) continue;
      const double* row = in + ((long long)zz * iny + yy) * inx;
      const double* wr = sw + (c * ky + b) * kx;
      for (int a = 0; a < kx; ++a) {
        int xx = x + a - offx;
        if (xx >= 0 && xx < inx) acc = fma(wr[a], __ldg(row + xx), acc);
      }
    }
  }
  out[t] = acc;
}

extern "C" int pmb_stencil_corr(int inx, int iny, int inz, const double* in, int ox, int oy, int oz, double* out, int kx, int ky,
                                int kz, const double* w, int offx, int offy, int offz, void* stream) {
  PMB_REQUIRE(inx > 0 && iny > 0 && inz > 0 && ox > 0 && oy > 0 && oz > 0 && kx > 0 && ky > 0 && kz > 0, "pmb_stencil_corr: invalid sizes");
  PMB_REQUIRE(in && out && w, "pmb_stencil_corr: NULL pointer argument");
  size_t smem = sizeof(double) * kx * ky * kz;
  PMB_REQUIRE(smem <= 48 * 1024, "pmb_stencil_corr: kernel of %d x %d x %d weights too large", kx, ky, kz);
  long long n = (long long)ox * oy * oz;
  stencil_corr_kernel<<<(unsigned)((n + 255) / 256), 256, smem, (cudaStream_t)stream>>>(inx, iny, inz, in, ox, oy, oz, out, kx, ky, kz, w,
                                                                                         offx, offy, offz);
  PMB_CHECK_LAUNCH("pmb_stencil_corr");
  return 0;
}
