// libpmb: symmetric half-stencil storage of a coarse-level operator and its application (north star: "vectorised loads of
// a symmetric-block CSR").
//
// The stencil-CSR layout of pmb_spmv.cu streams all 27 neighbour blocks of every node: 8 B per non-zero, at 0.91 of the HBM
// peak -- only fewer bytes make that sweep faster.  A Galerkin operator R^T A R is symmetric (to rounding), so the block
// that couples node i to its neighbour j is the transpose of the one that couples j to i.  Layout kept here:
//     S[slot][d][c][node],  slot 0 = the node's diagonal block, slots 1..13 = its 13 UPPER neighbours
//     (dk, dj, di) > (0, 0, 0) in lexicographic order, slot = (dk*3 + dj+1)*3 + di+1 - 4,
// i.e. direction-major, node-minor ("structure of arrays"): 14 ndof^2 doubles per node instead of 27 ndof^2 (0.52 of the
// bytes in HBM).  A thread owns one node and accumulates its ndof rows from
//     its own 14 blocks                              y_i += S[s][i]   x_(i + dir_s)     and
//     the 13 blocks stored at its LOWER neighbours   y_i += S[s][i - dir_s]^T x_(i - dir_s),
// every access a fully coalesced 8-byte load across the threads of a warp (consecutive nodes are consecutive in memory for
// every slot, shifted by the direction's node offset).  The transposed products are formed by the thread that loaded the
// block and handed to the neighbour's thread through shared memory (see sym_kernel); no atomics, deterministic.
// Results differ from the stencil-CSR kernel by the asymmetry of the assembled values (rounding of R^T A R, ~1e-16
// relative); pmb_sym_pack measures that asymmetry so that the host can refuse the layout for a non-symmetric operator.
#include "pmb_tilestream.cuh"

enum { SMODE_SPMV = PMB_SPMV, SMODE_RESID = PMB_RESIDUAL, SMODE_JACOBI = PMB_JACOBI };
constexpr int SYM_SLOTS = 14;
// Each block is fetched from HBM ONCE and used twice on chip.  Two other designs were built and measured first
// (256x128x128 bench, level 1, 0.187 ms for the stencil-CSR kernel):
//   * a pure gather -- every node also loading the 13 blocks stored with its lower neighbours -- moves the same 27 blocks
//     per node over the L2 -> SM fabric as the full layout does from HBM: 0.21 ms, the fabric (~7 TB/s) is the limit;
//   * node columns marching in z with the transposed products handed upward through shared memory read only 0.66 GB
//     from DRAM but exposes one memory round trip per plane to ~1000 resident warps: 0.35 ms, latency bound.
// Here a CTA owns a BRICK of SYM_BX x SYM_BY x SYM_BZ nodes, one thread per node, all bricks independent.  A thread
//   * accumulates its own rows from its diagonal block and its 13 upper blocks (coalesced 8-byte loads: consecutive lanes
//     = consecutive nodes of an x-row; the next block is already in flight while the current one is used);
//   * forms the TRANSPOSED product B^T x_i of each upper block and adds it to the shared-memory accumulator of the
//     neighbour it belongs to when that neighbour is a node of the brick -- one direction per phase, so every accumulator
//     has exactly one writer per phase (a block barrier between phases, no atomics, fixed order: deterministic);
//   * for the lower neighbours OUTSIDE the brick reads the block stored with the neighbour, as the pure gather would.
// Measured (same bench): 0.206 ms with 32 x 4 x 2 bricks (0.215 with 32 x 4 x 4, 0.24 with three CTAs of 80 registers) -- the
// best of the three, still behind the stencil-CSR kernel, whose TMA-streamed contiguous runs reach 0.91 of the HBM peak
// while 126 interleaved per-thread load streams do not.  The layout therefore stays OFF by default
// (PMB_SYMMETRIC_STORAGE=1 / DeviceCSR.symmetric_storage enables it); see DESIGN.md section 6c.
#ifndef PMB_SYM_BY
#define PMB_SYM_BY 4
#define PMB_SYM_BZ 2
#define PMB_SYM_MINB 2
#endif
constexpr int SYM_BX = 32, SYM_BY = PMB_SYM_BY, SYM_BZ = PMB_SYM_BZ, SYM_NT = SYM_BX * SYM_BY * SYM_BZ;

__device__ __forceinline__ void sym_dir(int slot, int& di, int& dj, int& dk) {
  const int o = slot + 4;
  dk = o / 9;
  dj = (o / 3) % 3 - 1;
  di = o % 3 - 1;
}

__device__ __forceinline__ unsigned long long dbits(double v) { return (unsigned long long)__double_as_longlong(fabs(v)); }

// One thread per (slot, node): copy the block out of the node's stencil-CSR row group (zeros where the neighbour is outside
// the grid) and compare it with the transposed block stored with the neighbour.  stats[0] = max |A_ij - A_ji^T|,
// stats[1] = max |A_ij| as bit patterns of non-negative doubles (order-preserving): atomicMax.
template <int NDOF>
__global__ void __launch_bounds__(256) sym_pack_kernel(Geo g, const double* __restrict__ A, double* __restrict__ S,
                                                       unsigned long long* __restrict__ stats) {
  const long long N = g.nOwned;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double mdiff = 0.0, mabs = 0.0;
  if (gid < N * SYM_SLOTS) {
    const int slot = (int)(gid / N);
    const long long ln = gid - (long long)slot * N;
    int i, j, k, di, dj, dk;
    node_ijk(g, ln, i, j, k);
    sym_dir(slot, di, dj, dk);
    const int i2 = i + di, j2 = j + dj, k2 = k + dk;
    const bool in = i2 >= 0 && i2 < g.NX && j2 >= 0 && j2 < g.NY && k2 >= 0 && k2 < g.NZ;
    double blk[NDOF][NDOF];
#pragma unroll
    for (int d = 0; d < NDOF; ++d)
#pragma unroll
      for (int c = 0; c < NDOF; ++c) blk[d][c] = 0.0;
    if (in) {
      const int cx = cnt1(i, g.NX), cy = cnt1(j, g.NY), cz = cnt1(k, g.NZ);
      const int L = cx * cy * cz * NDOF;
      const double* rp = A + node_entry_offset(g, ln) + (((k2 - max(k - 1, 0)) * cy + (j2 - max(j - 1, 0))) * cx + (i2 - max(i - 1, 0))) * NDOF;
#pragma unroll
      for (int d = 0; d < NDOF; ++d)
#pragma unroll
        for (int c = 0; c < NDOF; ++c) blk[d][c] = rp[d * L + c];
      // the transposed partner: block (-di, -dj, -dk) of node (i2, j2, k2), if that node's rows live in this slab
      const int kl2 = k2 - g.kz0;
      if (kl2 >= 0 && kl2 < g.nzl) {
        const long long ln2 = ((long long)kl2 * g.NY + j2) * g.NX + i2;
        const int cx2 = cnt1(i2, g.NX), cy2 = cnt1(j2, g.NY), cz2 = cnt1(k2, g.NZ);
        const int L2 = cx2 * cy2 * cz2 * NDOF;
        const double* rq = A + node_entry_offset(g, ln2) + (((k - max(k2 - 1, 0)) * cy2 + (j - max(j2 - 1, 0))) * cx2 + (i - max(i2 - 1, 0))) * NDOF;
#pragma unroll
        for (int d = 0; d < NDOF; ++d)
#pragma unroll
          for (int c = 0; c < NDOF; ++c) {
            mdiff = fmax(mdiff, fabs(blk[d][c] - rq[c * L2 + d]));
            mabs = fmax(mabs, fabs(blk[d][c]));
          }
      }
    }
#pragma unroll
    for (int d = 0; d < NDOF; ++d)
#pragma unroll
      for (int c = 0; c < NDOF; ++c) S[((long long)(slot * NDOF + d) * NDOF + c) * N + ln] = blk[d][c];
  }
  // block maxima -> two atomics per CTA
  __shared__ double smax[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mdiff = fmax(mdiff, __shfl_xor_sync(0xffffffffu, mdiff, o));
    mabs = fmax(mabs, __shfl_xor_sync(0xffffffffu, mabs, o));
  }
  if ((threadIdx.x & 31) == 0) smax[0][threadIdx.x >> 5] = mdiff, smax[1][threadIdx.x >> 5] = mabs;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int v = 1; v < 8; ++v) mdiff = fmax(mdiff, smax[0][v]), mabs = fmax(mabs, smax[1][v]);
    mdiff = fmax(mdiff, smax[0][0]);
    mabs = fmax(mabs, smax[1][0]);
    if (mdiff > 0.0) atomicMax(stats, dbits(mdiff));
    if (mabs > 0.0) atomicMax(stats + 1, dbits(mabs));
  }
}

template <int NDOF, int MODE>
__global__ void __launch_bounds__(SYM_NT, PMB_SYM_MINB) sym_kernel(Geo g, const double* __restrict__ S, const double* __restrict__ x,
                                                        const double* __restrict__ b, const double* __restrict__ diag, double w,
                                                        double* __restrict__ y, const double* __restrict__ dotv,
                                                        double* __restrict__ partials) {
  const long long N = g.nOwned;
  __shared__ double yacc[NDOF][SYM_NT];   // transposed contributions received by the nodes of the brick
  __shared__ double wred[3][SYM_NT / 32];
  const int tid = threadIdx.x, tx = tid % SYM_BX, ty = (tid / SYM_BX) % SYM_BY, tz = tid / (SYM_BX * SYM_BY);
  const int nbx = (g.NX + SYM_BX - 1) / SYM_BX, nby = (g.NY + SYM_BY - 1) / SYM_BY, nbz = (g.NZ + SYM_BZ - 1) / SYM_BZ;
  const long long nbricks = (long long)nbx * nby * nbz;
  double d0 = 0.0, d1 = 0.0, d2 = 0.0;

  for (long long brick = blockIdx.x; brick < nbricks; brick += gridDim.x) {
    const int bx = (int)(brick % nbx), by = (int)((brick / nbx) % nby), bz = (int)(brick / ((long long)nbx * nby));
    const int i = bx * SYM_BX + tx, j = by * SYM_BY + ty, k = bz * SYM_BZ + tz;
    const bool in = i < g.NX && j < g.NY && k < g.NZ;
    const long long ln = in ? ((long long)k * g.NY + j) * g.NX + i : 0;
    double acc[NDOF], xc[NDOF];
#pragma unroll
    for (int d = 0; d < NDOF; ++d) acc[d] = 0.0, xc[d] = in ? __ldg(x + ln * NDOF + d) : 0.0, yacc[d][tid] = 0.0;

    // block of slot s of this node (zeros are stored for neighbours outside the grid)
    auto load_block = [&](int slot, double (&blk)[NDOF * NDOF]) {
      const double* sp = S + (long long)slot * NDOF * NDOF * N + ln;
#pragma unroll
      for (int e = 0; e < NDOF * NDOF; ++e) blk[e] = in ? __ldg(sp + (long long)e * N) : 0.0;
    };
    double cur[NDOF * NDOF], nxt[NDOF * NDOF];
    load_block(0, cur);
    load_block(1, nxt);
    // ---- lower neighbours outside the brick: the block stored with the neighbour, transposed (issued early: independent)
    if (in) {
#pragma unroll
      for (int slot = 1; slot < SYM_SLOTS; ++slot) {
        const int o = slot + 4, dk = o / 9, dj = (o / 3) % 3 - 1, di = o % 3 - 1;
        const int i2 = i - di, j2 = j - dj, k2 = k - dk;
        const int tx2 = tx - di, ty2 = ty - dj, tz2 = tz - dk;
        const bool inbrick = tx2 >= 0 && tx2 < SYM_BX && ty2 >= 0 && ty2 < SYM_BY && tz2 >= 0;   // (tz2 < SYM_BZ: dk >= 0)
        if (!inbrick && i2 >= 0 && i2 < g.NX && j2 >= 0 && j2 < g.NY && k2 >= 0) {
          const long long nb = ln - di - (long long)dj * g.NX - (long long)dk * g.plane;
          const double* sp = S + (long long)slot * NDOF * NDOF * N + nb;
          double xv[NDOF];
#pragma unroll
          for (int c = 0; c < NDOF; ++c) xv[c] = __ldg(x + nb * NDOF + c);
#pragma unroll
          for (int c = 0; c < NDOF; ++c)
#pragma unroll
            for (int d = 0; d < NDOF; ++d) acc[d] = fma(__ldg(sp + (long long)(c * NDOF + d) * N), xv[c], acc[d]);
        }
      }
      // diagonal block
#pragma unroll
      for (int d = 0; d < NDOF; ++d)
#pragma unroll
        for (int c = 0; c < NDOF; ++c) acc[d] = fma(cur[d * NDOF + c], xc[c], acc[d]);
    }
    __syncthreads();   // every accumulator of the brick is zeroed
    // ---- the 13 upper directions, one phase each
#pragma unroll
    for (int slot = 1; slot < SYM_SLOTS; ++slot) {
      const int o = slot + 4, dk = o / 9, dj = (o / 3) % 3 - 1, di = o % 3 - 1;
#pragma unroll
      for (int e = 0; e < NDOF * NDOF; ++e) cur[e] = nxt[e];
      if (slot + 1 < SYM_SLOTS) load_block(slot + 1, nxt);
      const int i2 = i + di, j2 = j + dj, k2 = k + dk;
      if (in && i2 >= 0 && i2 < g.NX && j2 < g.NY && k2 < g.NZ) {   // (j2 >= 0: an upper direction with dj = -1 has dk = 1 ... checked below)
        if (j2 >= 0) {
          const long long nb = ln + di + (long long)dj * g.NX + (long long)dk * g.plane;
#pragma unroll
          for (int d = 0; d < NDOF; ++d)
#pragma unroll
            for (int c = 0; c < NDOF; ++c) acc[d] = fma(cur[d * NDOF + c], __ldg(x + nb * NDOF + c), acc[d]);
          const int tx2 = tx + di, ty2 = ty + dj, tz2 = tz + dk;
          if (tx2 >= 0 && tx2 < SYM_BX && ty2 >= 0 && ty2 < SYM_BY && tz2 < SYM_BZ) {
            const int tgt = (tz2 * SYM_BY + ty2) * SYM_BX + tx2;
#pragma unroll
            for (int c = 0; c < NDOF; ++c) {
              double tv = 0.0;
#pragma unroll
              for (int d = 0; d < NDOF; ++d) tv = fma(cur[d * NDOF + c], xc[d], tv);
              yacc[c][tgt] += tv;
            }
          }
        }
      }
      __syncthreads();   // one writer per accumulator and phase
    }
    if (in) {
#pragma unroll
      for (int d = 0; d < NDOF; ++d) {
        const long long r = ln * NDOF + d;
        const double ax = acc[d] + yacc[d][tid];
        double out;
        if (MODE == SMODE_SPMV) out = ax;
        else if (MODE == SMODE_RESID) out = b[r] - ax;
        else out = xc[d] + w * ((b[r] - ax) / diag[r]);
        y[r] = out;
        if (partials) {
          const double dvv = dotv ? dotv[r] : 0.0;
          d0 = fma(out, xc[d], d0);
          d1 = fma(xc[d], dvv, d1);
          d2 = fma(out, dvv, d2);
        }
      }
    }
    __syncthreads();   // the accumulators are read before the next brick zeroes them
  }
  if (partials) {
    d0 = warp_sum(d0);
    d1 = warp_sum(d1);
    d2 = warp_sum(d2);
    if ((tid & 31) == 0) wred[0][tid >> 5] = d0, wred[1][tid >> 5] = d1, wred[2][tid >> 5] = d2;
    __syncthreads();
    if (tid == 0) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      for (int v = 0; v < SYM_NT / 32; ++v) s0 += wred[0][v], s1 += wred[1][v], s2 += wred[2][v];
      partials[3 * (long long)blockIdx.x] = s0;
      partials[3 * (long long)blockIdx.x + 1] = s1;
      partials[3 * (long long)blockIdx.x + 2] = s2;
    }
  }
}

static int sym_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

extern "C" long long pmb_sym_doubles(const pmb_grid* p) {
  if (validate_grid(p, "pmb_sym_doubles")) return -1;
  Geo g = make_geo(p);
  return (long long)SYM_SLOTS * g.ndof * g.ndof * g.nOwned;
}

extern "C" int pmb_sym_pack(const pmb_grid* p, const double* data, double* S, unsigned long long* stats, void* stream) {
  if (validate_grid(p, "pmb_sym_pack")) return 1;
  PMB_REQUIRE(data && S && stats, "pmb_sym_pack: NULL pointer argument");
  PMB_REQUIRE(p->nz > 0, "pmb_sym_pack: 3-D grids only");
  PMB_REQUIRE(p->kz0 == 0 && p->nzl == p->nz + 1, "pmb_sym_pack: whole grids only (the lower neighbours of a slab's first plane live on another rank)");
  Geo g = make_geo(p);
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(stats, 0, 2 * sizeof(unsigned long long), st) != cudaSuccess) return pmb_set_error("pmb_sym_pack: cudaMemsetAsync failed");
  const long long total = g.nOwned * SYM_SLOTS;
  const long long nblk = (total + 255) / 256;
  PMB_REQUIRE(nblk < 2147483647LL, "pmb_sym_pack: grid too large");
  switch (g.ndof) {
    case 1: sym_pack_kernel<1><<<(unsigned)nblk, 256, 0, st>>>(g, data, S, stats); break;
    case 2: sym_pack_kernel<2><<<(unsigned)nblk, 256, 0, st>>>(g, data, S, stats); break;
    case 3: sym_pack_kernel<3><<<(unsigned)nblk, 256, 0, st>>>(g, data, S, stats); break;
  }
  PMB_CHECK_LAUNCH("pmb_sym_pack");
  return 0;
}

template <int NDOF, int MODE>
static int launch_sym(const Geo& g, const double* S, const double* x, const double* b, const double* diag, double w, double* y,
                      const double* dotv, double* dot_out, double* ws, cudaStream_t st) {
  const long long nbricks = (long long)((g.NX + SYM_BX - 1) / SYM_BX) * ((g.NY + SYM_BY - 1) / SYM_BY) * ((g.NZ + SYM_BZ - 1) / SYM_BZ);
  const long long cap = 4LL * sym_sm_count();   // == pmb_spmv_ws_doubles() / 3 partial triples
  const int grid = (int)(dot_out ? (nbricks < cap ? nbricks : cap) : (nbricks < 2147483647LL ? nbricks : 2147483647LL));
  sym_kernel<NDOF, MODE><<<grid, SYM_NT, 0, st>>>(g, S, x, b, diag, w, y, dotv, dot_out ? ws : nullptr);
  PMB_CHECK_LAUNCH("pmb_sym_spmv");
  if (dot_out) {
    reduce_triples_kernel<<<1, 1024, 0, st>>>(ws, grid, dot_out);
    PMB_CHECK_LAUNCH("pmb_sym_spmv(reduce)");
  }
  return 0;
}

template <int NDOF>
static int dispatch_sym(int mode, const Geo& g, const double* S, const double* x, const double* b, const double* diag, double w,
                        double* y, const double* dotv, double* dot_out, double* ws, cudaStream_t st) {
  switch (mode) {
    case SMODE_SPMV: return launch_sym<NDOF, SMODE_SPMV>(g, S, x, b, diag, w, y, dotv, dot_out, ws, st);
    case SMODE_RESID: return launch_sym<NDOF, SMODE_RESID>(g, S, x, b, diag, w, y, dotv, dot_out, ws, st);
    case SMODE_JACOBI: return launch_sym<NDOF, SMODE_JACOBI>(g, S, x, b, diag, w, y, dotv, dot_out, ws, st);
  }
  return pmb_set_error("pmb_sym_spmv: unknown mode %d", mode);
}

extern "C" int pmb_sym_spmv(const pmb_grid* p, int mode, const double* S, const double* x, const double* b, const double* diag,
                            double w, double* y, const double* dotv, double* dot_out, double* ws, void* stream) {
  if (validate_grid(p, "pmb_sym_spmv")) return 1;
  PMB_REQUIRE(S && x && y, "pmb_sym_spmv: NULL pointer argument");
  PMB_REQUIRE(x != y, "pmb_sym_spmv: y must not alias x");
  PMB_REQUIRE(p->nz > 0 && p->kz0 == 0 && p->nzl == p->nz + 1, "pmb_sym_spmv: whole 3-D grids only");
  PMB_REQUIRE(mode == SMODE_SPMV || b, "pmb_sym_spmv: b required for residual / Jacobi");
  PMB_REQUIRE(mode != SMODE_JACOBI || diag, "pmb_sym_spmv: diag required for Jacobi");
  PMB_REQUIRE(!dot_out || ws, "pmb_sym_spmv: workspace required for the fused dot products");
  Geo g = make_geo(p);
  cudaStream_t st = (cudaStream_t)stream;
  switch (g.ndof) {
    case 1: return dispatch_sym<1>(mode, g, S, x, b, diag, w, y, dotv, dot_out, ws, st);
    case 2: return dispatch_sym<2>(mode, g, S, x, b, diag, w, y, dotv, dot_out, ws, st);
    case 3: return dispatch_sym<3>(mode, g, S, x, b, diag, w, y, dotv, dot_out, ws, st);
  }
  return 1;
}
