// libpmb: stencil-CSR operator application (K2 SpMV, residual, K3 fused damped-Jacobi sweep) and row statistics.
//
// Replaces scipy's csr_matvec as called from pymoto/solvers/iterative.py:236-255 (smoother and residual inside
// GeometricMultigrid.solve), :359,375,382 (CG) and pymoto/solvers/solvers.py:84,237 (LDAS residual / database),
// plus DampedJacobi (iterative.py:38-47) and get_diagonal_indices (solvers.py:88-96).
//
// Layout: the matrix is the reference's own CSR `data` array.  On a structured grid all rows of T consecutive
// nodes are ONE contiguous run of doubles, so a CTA streams that run with 128-bit loads into shared memory
// (fully coalesced, no index traffic: indptr/indices are closed-form), then PARTS threads per node multiply
// their part of the node's NDOF x (27*NDOF) block against x gathered through L1.
#include "pmb_common.cuh"

enum { MODE_SPMV = PMB_SPMV, MODE_RESID = PMB_RESIDUAL, MODE_JACOBI = PMB_JACOBI, MODE_ROWSTATS = 3 };

// T nodes per CTA, PARTS threads per node; NT = threads per CTA rounded up to whole warps (the padding threads
// only help with the staging loads); SMEM_DOUBLES = largest staged run (+2 for the 16-byte alignment slack)
template <int NDOF, int T_, int PARTS_>
struct TileCfgBase {
  static constexpr int T = T_, PARTS = PARTS_;
  static constexpr int NT = (T_ * PARTS_ + 31) / 32 * 32;
  static constexpr int SMEM_DOUBLES = T_ * NDOF * NDOF * 27 + 2;
};
template <int NDOF>
struct TileCfg;
template <>
struct TileCfg<3> : TileCfgBase<3, 16, 9> {};
template <>
struct TileCfg<2> : TileCfgBase<2, 48, 3> {};
template <>
struct TileCfg<1> : TileCfgBase<1, 96, 3> {};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NDOF, int MODE>
__global__ void __launch_bounds__(TileCfg<NDOF>::NT)
    tile_kernel(Geo g, const double* __restrict__ A, const double* __restrict__ x, const double* __restrict__ b,
                const double* __restrict__ diag, double w, double* __restrict__ y, const double* __restrict__ dotv,
                double* __restrict__ partials, int* __restrict__ nnz_out) {
  constexpr int T = TileCfg<NDOF>::T, PARTS = TileCfg<NDOF>::PARTS, NT = TileCfg<NDOF>::NT;
  extern __shared__ double2 smem2[];
  double* sA = reinterpret_cast<double*>(smem2);
  __shared__ double red[PARTS][T * NDOF];
  __shared__ double red2[MODE == MODE_ROWSTATS ? PARTS : 1][MODE == MODE_ROWSTATS ? T * NDOF : 1];
  __shared__ double wred[3][NT / 32];

  const int tid = threadIdx.x;
  const long long n0 = (long long)blockIdx.x * T;
  const long long nEnd = min(n0 + T, g.nOwned);
  const long long e0 = node_entry_offset(g, n0);
  const long long e1 = node_entry_offset(g, nEnd);
  const long long lo = e0 & ~1LL;
  const long long hi = (e1 + 1) & ~1LL;

  // ---- stage the contiguous run of matrix values (streaming, 128-bit, evict-first)
  {
    const double2* g2 = reinterpret_cast<const double2*>(A + lo);
    const int nvec = (int)((hi - lo) >> 1);
    for (int v = tid; v < nvec; v += NT) smem2[v] = __ldcs(g2 + v);
  }
  __syncthreads();

  const int gI = tid % T, q = tid / T;
  const long long ln = n0 + gI;
  double acc[NDOF], acc2[NDOF];
#pragma unroll
  for (int d = 0; d < NDOF; ++d) acc[d] = 0.0, acc2[d] = 0.0;

  if (ln < nEnd && q < PARTS) {
    int i, j, k;
    node_ijk(g, ln, i, j, k);
    const int cx = cnt1(i, g.NX), cy = cnt1(j, g.NY), cz = cnt1(k, g.NZ);
    const int ilo = max(i - 1, 0), jlo = max(j - 1, 0), klo = max(k - 1, 0);
    const int L = cx * cy * cz * NDOF;
    const double* rowp = sA + ((long long)(NDOF * NDOF) * (block_offset(g, i, j, k) - g.bo0) - lo);
    int kk0, kk1, jj0, jj1;
    if (PARTS == 9) {
      kk0 = q / 3; kk1 = kk0 + 1; jj0 = q % 3; jj1 = jj0 + 1;
    } else {
      kk0 = q; kk1 = q + 1; jj0 = 0; jj1 = 3;
    }
    kk1 = min(kk1, cz);
    jj1 = min(jj1, cy);
    for (int kkI = kk0; kkI < kk1; ++kkI) {
      for (int jjI = jj0; jjI < jj1; ++jjI) {
        const int nbr0 = (kkI * cy + jjI) * cx;
        const long long c0 = ((long long)(klo + kkI - g.kz0) * g.NY + (jlo + jjI)) * g.NX + ilo;
        for (int iiI = 0; iiI < cx; ++iiI) {
          const double* ap = rowp + (nbr0 + iiI) * NDOF;
          if (MODE == MODE_ROWSTATS) {
            const bool self = (c0 + iiI) == ln;
#pragma unroll
            for (int d = 0; d < NDOF; ++d)
#pragma unroll
              for (int cd = 0; cd < NDOF; ++cd) {
                double a = ap[d * L + cd];
                if (self && cd == d) acc[d] = a;
                else acc2[d] += (a != 0.0) ? 1.0 : 0.0;
              }
          } else {
            double xv[NDOF];
            const double* xp = x + (c0 + iiI) * NDOF;
#pragma unroll
            for (int cd = 0; cd < NDOF; ++cd) xv[cd] = __ldg(xp + cd);
#pragma unroll
            for (int d = 0; d < NDOF; ++d)
#pragma unroll
              for (int cd = 0; cd < NDOF; ++cd) acc[d] = fma(ap[d * L + cd], xv[cd], acc[d]);
          }
        }
      }
    }
  }
  if (q < PARTS) {
#pragma unroll
    for (int d = 0; d < NDOF; ++d) {
      red[q][gI * NDOF + d] = acc[d];
      if (MODE == MODE_ROWSTATS) red2[q][gI * NDOF + d] = acc2[d];
    }
  }
  __syncthreads();

  // ---- combine the PARTS partial sums of each row in fixed order, then the per-row epilogue
  double d0 = 0.0, d1 = 0.0, d2 = 0.0;
  if (tid < T * NDOF) {
    const long long node = n0 + tid / NDOF;
    if (node < nEnd) {
      const long long r = n0 * NDOF + tid;
      double ax = 0.0;
#pragma unroll
      for (int p = 0; p < PARTS; ++p) ax += red[p][tid];
      if (MODE == MODE_ROWSTATS) {
        double c = 0.0;
#pragma unroll
        for (int p = 0; p < PARTS; ++p) c += red2[p][tid];
        if (y) y[r] = ax;
        if (nnz_out) nnz_out[r] = (int)c;
      } else {
        double out;
        if (MODE == MODE_SPMV) out = ax;
        else if (MODE == MODE_RESID) out = b[r] - ax;
        else out = x[r] + w * ((b[r] - ax) / diag[r]);
        y[r] = out;
        if (partials) {
          d0 = out * x[r];
          if (dotv) d1 = x[r] * dotv[r], d2 = out * dotv[r];
        }
      }
    }
  }
  if (MODE != MODE_ROWSTATS && partials) {
    d0 = warp_sum(d0);
    d1 = warp_sum(d1);
    d2 = warp_sum(d2);
    if ((tid & 31) == 0) wred[0][tid >> 5] = d0, wred[1][tid >> 5] = d1, wred[2][tid >> 5] = d2;
    __syncthreads();
    if (tid == 0) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      for (int v = 0; v < NT / 32; ++v) s0 += wred[0][v], s1 += wred[1][v], s2 += wred[2][v];
      partials[3 * (long long)blockIdx.x] = s0;
      partials[3 * (long long)blockIdx.x + 1] = s1;
      partials[3 * (long long)blockIdx.x + 2] = s2;
    }
  }
}

// final deterministic reduction of the per-CTA partial triples
__global__ void __launch_bounds__(1024) reduce_triples_kernel(const double* __restrict__ partials, long long nblocks,
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

template <int NDOF>
static long long tile_blocks(const Geo& g) { return (g.nOwned + TileCfg<NDOF>::T - 1) / TileCfg<NDOF>::T; }

extern "C" long long pmb_spmv_ws_doubles(const pmb_grid* p) {
  if (validate_grid(p, "pmb_spmv_ws_doubles")) return -1;
  Geo g = make_geo(p);
  long long nb = g.ndof == 3 ? tile_blocks<3>(g) : g.ndof == 2 ? tile_blocks<2>(g) : tile_blocks<1>(g);
  return 3 * nb;
}

template <int NDOF, int MODE>
static int launch_tile(const Geo& g, const double* A, const double* x, const double* b, const double* diag, double w,
                       double* y, const double* dotv, double* dot_out, double* ws, int* nnz_out, cudaStream_t st) {
  constexpr int NT = TileCfg<NDOF>::NT;
  const size_t smem = sizeof(double) * TileCfg<NDOF>::SMEM_DOUBLES;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tile_kernel<NDOF, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return pmb_set_error("tile_kernel attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  long long nb = tile_blocks<NDOF>(g);
  tile_kernel<NDOF, MODE><<<(unsigned)nb, NT, smem, st>>>(g, A, x, b, diag, w, y, dotv, dot_out ? ws : nullptr, nnz_out);
  PMB_CHECK_LAUNCH("pmb_spmv");
  if (dot_out) {
    reduce_triples_kernel<<<1, 1024, 0, st>>>(ws, nb, dot_out);
    PMB_CHECK_LAUNCH("pmb_spmv(reduce)");
  }
  return 0;
}

template <int NDOF>
static int dispatch_mode(int mode, const Geo& g, const double* A, const double* x, const double* b, const double* diag,
                         double w, double* y, const double* dotv, double* dot_out, double* ws, cudaStream_t st) {
  switch (mode) {
    case MODE_SPMV: return launch_tile<NDOF, MODE_SPMV>(g, A, x, b, diag, w, y, dotv, dot_out, ws, nullptr, st);
    case MODE_RESID: return launch_tile<NDOF, MODE_RESID>(g, A, x, b, diag, w, y, dotv, dot_out, ws, nullptr, st);
    case MODE_JACOBI: return launch_tile<NDOF, MODE_JACOBI>(g, A, x, b, diag, w, y, dotv, dot_out, ws, nullptr, st);
  }
  return pmb_set_error("pmb_spmv: unknown mode %d", mode);
}

extern "C" int pmb_spmv(const pmb_grid* p, int mode, const double* data, const double* x, const double* b,
                        const double* diag, double w, double* y, const double* dotv, double* dot_out, double* ws,
                        void* stream) {
  if (validate_grid(p, "pmb_spmv")) return 1;
  PMB_REQUIRE(data && x && y, "pmb_spmv: NULL pointer argument");
  PMB_REQUIRE(x != y, "pmb_spmv: y must not alias x");
  PMB_REQUIRE(mode == MODE_SPMV || b, "pmb_spmv: b required for residual / Jacobi");
  PMB_REQUIRE(mode != MODE_JACOBI || diag, "pmb_spmv: diag required for Jacobi");
  PMB_REQUIRE(!dot_out || ws, "pmb_spmv: workspace required for the fused dot products");
  PMB_REQUIRE((reinterpret_cast<size_t>(data) & 15) == 0, "pmb_spmv: data must be 16-byte aligned");
  Geo g = make_geo(p);
  cudaStream_t st = (cudaStream_t)stream;
  switch (g.ndof) {
    case 1: return dispatch_mode<1>(mode, g, data, x, b, diag, w, y, dotv, dot_out, ws, st);
    case 2: return dispatch_mode<2>(mode, g, data, x, b, diag, w, y, dotv, dot_out, ws, st);
    case 3: return dispatch_mode<3>(mode, g, data, x, b, diag, w, y, dotv, dot_out, ws, st);
  }
  return 1;
}

extern "C" int pmb_rowstats(const pmb_grid* p, const double* data, double* diag, int* nnz_offdiag, void* stream) {
  if (validate_grid(p, "pmb_rowstats")) return 1;
  PMB_REQUIRE(data, "pmb_rowstats: NULL data");
  PMB_REQUIRE((reinterpret_cast<size_t>(data) & 15) == 0, "pmb_rowstats: data must be 16-byte aligned");
  Geo g = make_geo(p);
  cudaStream_t st = (cudaStream_t)stream;
  switch (g.ndof) {
    case 1: return launch_tile<1, MODE_ROWSTATS>(g, data, nullptr, nullptr, nullptr, 0.0, diag, nullptr, nullptr, nullptr, nnz_offdiag, st);
    case 2: return launch_tile<2, MODE_ROWSTATS>(g, data, nullptr, nullptr, nullptr, 0.0, diag, nullptr, nullptr, nullptr, nnz_offdiag, st);
    case 3: return launch_tile<3, MODE_ROWSTATS>(g, data, nullptr, nullptr, nullptr, 0.0, diag, nullptr, nullptr, nullptr, nnz_offdiag, st);
  }
  return 1;
}

// u = w * (r / diag)   (first pre-smoothing step from a zero initial guess, iterative.py:43,233-234)
__global__ void __launch_bounds__(256) smooth0_kernel(long long n, double w, const double* __restrict__ r,
                                                       const double* __restrict__ diag, double* __restrict__ u) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) u[t] = w * (r[t] / diag[t]);
}

extern "C" int pmb_smooth0(long long n, double w, const double* r, const double* diag, double* u, void* stream) {
  PMB_REQUIRE(r && diag && u, "pmb_smooth0: NULL pointer argument");
  if (n <= 0) return 0;
  smooth0_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, w, r, diag, u);
  PMB_CHECK_LAUNCH("pmb_smooth0");
  return 0;
}
