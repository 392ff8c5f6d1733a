// libpmb: stencil-CSR operator application (K2 SpMV, residual, K3 fused damped-Jacobi sweep) and row statistics.
//
// Replaces scipy's csr_matvec as called from pymoto/solvers/iterative.py:236-255 (smoother and residual inside
// GeometricMultigrid.solve), :359,375,382 (CG) and pymoto/solvers/solvers.py:84,237 (LDAS residual / database),
// plus DampedJacobi (iterative.py:38-47) and get_diagonal_indices (solvers.py:88-96).
//
// Layout: the matrix is the reference's own CSR `data` array.  On a structured grid all rows of T consecutive
// nodes are ONE contiguous run of doubles and indptr/indices are closed-form, so no index is ever read.
//
// Kernel: persistent CTAs (2 per SM) walk the tiles round-robin.  One elected thread streams each tile's run into
// a 3-stage shared-memory ring with TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx::bytes, L2
// evict-first for matrices larger than L2), so loads of tiles t+1, t+2 are in flight while tile t is multiplied.
// PARTS threads per node each take one line of the 27-point stencil (<= 3 neighbour nodes x NDOF columns for all
// NDOF rows of the node), x is gathered through L1, the PARTS partial sums of a row are combined in fixed order
// through shared memory and the epilogue (residual / Jacobi update / fused dot products) is applied per row.
#include "pmb_tilestream.cuh"

enum { MODE_SPMV = PMB_SPMV, MODE_RESID = PMB_RESIDUAL, MODE_JACOBI = PMB_JACOBI, MODE_ROWSTATS = 3 };

// T nodes per tile, PARTS threads per node, STAGES ring slots; NT = threads per CTA rounded up to whole warps;
// TILE_DOUBLES = largest staged run (+2 for the 16-byte alignment slack), itself kept a multiple of 2.
template <int NDOF, int T_, int PARTS_, int STAGES_, int CTAS_>
struct TileCfgBase {
  static constexpr int T = T_, PARTS = PARTS_, STAGES = STAGES_, CTAS = CTAS_;  // CTAS = resident CTAs per SM
  static constexpr int NT = (T_ * PARTS_ + 31) / 32 * 32;
  static constexpr int TILE_DOUBLES = (T_ * NDOF * NDOF * 27 + 2 + 1) / 2 * 2;
  static constexpr size_t SMEM_BYTES = sizeof(double) * TILE_DOUBLES * STAGES_;
};
template <int NDOF>
struct TileCfg;
template <>
struct TileCfg<3> : TileCfgBase<3, 16, 9, 3, 2> {};
template <>
struct TileCfg<2> : TileCfgBase<2, 48, 3, 2, 2> {};
template <>
struct TileCfg<1> : TileCfgBase<1, 96, 3, 2, 4> {};  // small tiles: more CTAs in flight hide the per-tile latency


template <int NDOF, int MODE>
__global__ void __launch_bounds__(TileCfg<NDOF>::NT, TileCfg<NDOF>::CTAS)
    tile_kernel(Geo g, int ntiles, int tiles_per_row, int stream_hint, const double* __restrict__ A,
                const double* __restrict__ x, const double* __restrict__ b, const double* __restrict__ diag, double w,
                double* __restrict__ y, const double* __restrict__ dotv, double* __restrict__ partials,
                int* __restrict__ nnz_out) {
  using Cfg = TileCfg<NDOF>;
  constexpr int T = Cfg::T, PARTS = Cfg::PARTS, NT = Cfg::NT, STAGES = Cfg::STAGES, TD = Cfg::TILE_DOUBLES;
  constexpr int LINES = 9 / PARTS;  // stencil lines (jj, kk) per thread: 1 (PARTS = 9) or 3 (PARTS = 3: one kk plane)
  extern __shared__ __align__(128) double sTiles[];  // STAGES x TD doubles
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ double red[2][PARTS][T * NDOF];
  __shared__ double red2[MODE == MODE_ROWSTATS ? 2 : 1][MODE == MODE_ROWSTATS ? PARTS : 1][MODE == MODE_ROWSTATS ? T * NDOF : 1];
  __shared__ double wred[3][NT / 32];

  const int tid = threadIdx.x;
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles blockIdx.x + i*gridDim.x
  const long long amis = (long long)((reinterpret_cast<uintptr_t>(A) >> 3) & 1);

  // producer: issue the TMA bulk copy of this CTA's i-th tile into ring slot i % STAGES
  auto issue = [&](int i) {
    const int s = i % STAGES;
    const TileGeom t = tile_geom<NDOF, T>(g, (int)blockIdx.x + i * (int)gridDim.x, tiles_per_row);
    // 16-byte aligned window [lo, hi) around the tile's entries; `amis` = 1 when A itself sits on an odd double
    // (a sub-slab view of a larger matrix), in which case the window may start one double before A
    const long long lo = ((t.e0 + amis) & ~1LL) - amis;
    const long long hi = ((t.e1 + amis + 1) & ~1LL) - amis;
    const unsigned bytes = (unsigned)((hi - lo) * sizeof(double));
    mbar_expect_tx(&full_bar[s], bytes);
    tma_load_1d(sTiles + (size_t)s * TD, A + lo, bytes, &full_bar[s], stream_hint != 0);
  };

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < STAGES && i < my_tiles; ++i) issue(i);
  }
  __syncthreads();

  const int gI = tid % T, q = tid / T;
  double d0 = 0.0, d1 = 0.0, d2 = 0.0;  // fused dot-product partials, accumulated over this CTA's tiles

  for (int it = 0; it < my_tiles; ++it) {
    const int s = it % STAGES;
    const TileGeom t = tile_geom<NDOF, T>(g, (int)blockIdx.x + it * (int)gridDim.x, tiles_per_row);
    const long long lnode0 = ((long long)(t.k - g.kz0) * g.NY + t.j) * g.NX + t.i0;  // slab-local index of the tile's first node
    const int i = t.i0 + gI;
    const bool active = (gI < t.ni) && (q < PARTS);

    // ---- everything that does not depend on the matrix values is issued before waiting for the tile:
    //      the x gathers of this thread's stencil line(s) and the epilogue operands of this thread's row
    const int cx = cnt1(i, g.NX), ilo = max(i - 1, 0);
    double xv[LINES][3][NDOF];
    bool lvalid[LINES];
#pragma unroll
    for (int l = 0; l < LINES; ++l) {
      const int kkI = (PARTS == 9) ? q / 3 : q;
      const int jjI = (PARTS == 9) ? q % 3 : l;
      lvalid[l] = active && kkI < t.cz && jjI < t.cy;
      const long long c0 = ((long long)(t.klo + kkI - g.kz0) * g.NY + (t.jlo + jjI)) * g.NX + ilo;
#pragma unroll
      for (int n = 0; n < 3; ++n)
#pragma unroll
        for (int cd = 0; cd < NDOF; ++cd)
          xv[l][n][cd] = (MODE != MODE_ROWSTATS && lvalid[l] && n < cx) ? __ldg(x + (c0 + n) * NDOF + cd) : 0.0;
    }
    const bool erow = (tid < T * NDOF) && (tid / NDOF < t.ni);
    const long long r = (lnode0 * NDOF) + tid;  // global (slab-local) row of the epilogue thread
    double xr = 0.0, br = 0.0, dr = 1.0, dvr = 0.0;
    if (MODE != MODE_ROWSTATS && erow) {
      if (MODE == MODE_JACOBI || partials) xr = x[r];
      if (MODE != MODE_SPMV) br = b[r];
      if (MODE == MODE_JACOBI) dr = diag[r];
      if (partials && dotv) dvr = dotv[r];
    }

    mbar_wait(&full_bar[s], (unsigned)((it / STAGES) & 1));

    double acc[NDOF], acc2[NDOF];
#pragma unroll
    for (int d = 0; d < NDOF; ++d) acc[d] = 0.0, acc2[d] = 0.0;
    {
      const int L = cx * t.cy * t.cz * NDOF;
      const long long per = (long long)(NDOF * NDOF) * t.cy * t.cz;
      const double* rowp = sTiles + (size_t)s * TD + ((t.e0 - (((t.e0 + amis) & ~1LL) - amis)) + per * (pre1(i, g.NX) - pre1(t.i0, g.NX)));
#pragma unroll
      for (int l = 0; l < LINES; ++l) {
        if (!lvalid[l]) continue;
        const int kkI = (PARTS == 9) ? q / 3 : q;
        const int jjI = (PARTS == 9) ? q % 3 : l;
        const double* lp = rowp + (kkI * t.cy + jjI) * cx * NDOF;
#pragma unroll
        for (int n = 0; n < 3; ++n) {
          if (n < cx) {
            const double* ap = lp + n * NDOF;
            if (MODE == MODE_ROWSTATS) {
              const bool self = (t.klo + kkI == t.k) && (t.jlo + jjI == t.j) && (ilo + n == i);
#pragma unroll
              for (int d = 0; d < NDOF; ++d)
#pragma unroll
                for (int cd = 0; cd < NDOF; ++cd) {
                  const double a = ap[d * L + cd];
                  if (self && cd == d) acc[d] = a;
                  else acc2[d] += (a != 0.0) ? 1.0 : 0.0;
                }
            } else {
#pragma unroll
              for (int d = 0; d < NDOF; ++d)
#pragma unroll
                for (int cd = 0; cd < NDOF; ++cd) acc[d] = fma(ap[d * L + cd], xv[l][n][cd], acc[d]);
            }
          }
        }
      }
    }
    const int rb = it & 1;  // double-buffered combine area: the next tile's writes cannot race this tile's reads
    if (q < PARTS) {
#pragma unroll
      for (int d = 0; d < NDOF; ++d) {
        red[rb][q][gI * NDOF + d] = acc[d];
        if (MODE == MODE_ROWSTATS) red2[rb][q][gI * NDOF + d] = acc2[d];
      }
    }
    __syncthreads();  // every thread is done reading ring slot s

    if (tid == 0 && it + STAGES < my_tiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads before the async-proxy refill
      issue(it + STAGES);
    }

    // ---- combine the PARTS partial sums of each row in fixed order, then the per-row epilogue
    if (erow) {
      double ax = 0.0;
#pragma unroll
      for (int p = 0; p < PARTS; ++p) ax += red[rb][p][tid];
      if (MODE == MODE_ROWSTATS) {
        double c = 0.0;
#pragma unroll
        for (int p = 0; p < PARTS; ++p) c += red2[rb][p][tid];
        if (y) y[r] = ax;
        if (nnz_out) nnz_out[r] = (int)c;
      } else {
        double out;
        if (MODE == MODE_SPMV) out = ax;
        else if (MODE == MODE_RESID) out = br - ax;
        else out = xr + w * ((br - ax) / dr);
        y[r] = out;
        if (partials) {
          d0 = fma(out, xr, d0);
          d1 = fma(xr, dvr, d1);
          d2 = fma(out, dvr, d2);
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

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

template <int NDOF>
static int tiles_per_row(const Geo& g) { return (g.NX + TileCfg<NDOF>::T - 1) / TileCfg<NDOF>::T; }
template <int NDOF>
static long long tile_count(const Geo& g) { return (long long)tiles_per_row<NDOF>(g) * g.NY * g.nzl; }

static int grid_for(long long ntiles, int ctas_per_sm) {
  long long cap = (long long)sm_count() * ctas_per_sm;
  return (int)(ntiles < cap ? ntiles : cap);
}

extern "C" long long pmb_spmv_ws_doubles(const pmb_grid* p) {
  if (validate_grid(p, "pmb_spmv_ws_doubles")) return -1;
  return 3LL * sm_count() * 4;  // largest resident-CTA count of any configuration
}

template <int NDOF, int MODE>
static int launch_tile(const Geo& g, const double* A, const double* x, const double* b, const double* diag, double w,
                       double* y, const double* dotv, double* dot_out, double* ws, int* nnz_out, cudaStream_t st) {
  using Cfg = TileCfg<NDOF>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tile_kernel<NDOF, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return pmb_set_error("tile_kernel attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const long long ntiles = tile_count<NDOF>(g);
  PMB_REQUIRE(ntiles < 2147483647LL, "pmb_spmv: too many tiles");
  const int grid = grid_for(ntiles, Cfg::CTAS);
  // matrices that do not fit in L2 anyway are streamed evict-first so the x / b / y vectors keep their lines
  const long long nnz_bytes = 8LL * NDOF * NDOF * (pre1(g.kz0 + g.nzl, g.NZ) * g.Sy * g.Sx - g.bo0);
  const int stream_hint = nnz_bytes > (96LL << 20);
  tile_kernel<NDOF, MODE><<<grid, Cfg::NT, Cfg::SMEM_BYTES, st>>>(g, (int)ntiles, tiles_per_row<NDOF>(g), stream_hint, A, x, b, diag, w, y, dotv,
                                                                  dot_out ? ws : nullptr, nnz_out);
  PMB_CHECK_LAUNCH("pmb_spmv");
  if (dot_out) {
    reduce_triples_kernel<<<1, 1024, 0, st>>>(ws, grid, dot_out);
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
  PMB_REQUIRE((reinterpret_cast<size_t>(data) & 7) == 0, "pmb_spmv: data must be 8-byte aligned");
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
