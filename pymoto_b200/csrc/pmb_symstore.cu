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
// Each block is fetched from HBM ONCE and used twice on chip.  (A pure gather -- every node also loading the 13 blocks stored
// with its lower neighbours -- moves the same 27 blocks per node over the L2 -> SM fabric as the full layout does from HBM and
// was measured at 0.21 ms against 0.187 ms for the stencil-CSR kernel: the fabric, ~7 TB/s, is the limit, not the DRAM.)
// A CTA owns SYM_TX x SYM_TY node columns and marches through its chunk of node planes, one thread per node of the
// current plane.  Per plane a thread
//   A. reads its 14 blocks (coalesced 8-byte loads: consecutive lanes = consecutive nodes of an x-row) and the x of its 27
//      neighbours (L1); accumulates its own rows from the diagonal and upper blocks; forms the TRANSPOSED products
//      B^T x_i for its 13 upper neighbours and leaves each in the shared-memory slot of its target node (4 in-plane
//      slots, 9 slots for the plane above; one writer per slot and target -> no conflicts, no atomics);
//      picks up the 9 contributions that the plane below left for it in the previous step;
//   B. after ONE barrier picks up the 4 in-plane contributions, applies the epilogue and stores.
// Links that leave the CTA's tile sideways, or come from the plane below the chunk's first, cannot be served on chip: for
// those (19 % of the lower links of a 32 x 4 tile) the node reads the block stored with its lower neighbour, as in the
// pure gather.  Fixed summation order: deterministic.
constexpr int SYM_TX = 32, SYM_TY = 4, SYM_NT = SYM_TX * SYM_TY;
template <int NDOF>
struct SymCfg {
  static constexpr int TIN = 4 * NDOF * SYM_NT, TUP = 9 * NDOF * SYM_NT;   // doubles of one in-plane / upward buffer
  static constexpr size_t SMEM = sizeof(double) * 2 * (TIN + TUP);
};

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

// slot (1..13) -> index among the in-plane upper links (0..3: (di,dj) = (1,0), (-1,1), (0,1), (1,1)) or the upward ones (0..8)
__host__ __device__ constexpr int sym_sub(int slot) { return slot <= 4 ? slot - 1 : slot - 5; }

template <int NDOF, int MODE>
__global__ void __launch_bounds__(SYM_NT, 2) sym_kernel(Geo g, int zl, const double* __restrict__ S, const double* __restrict__ x,
                                                        const double* __restrict__ b, const double* __restrict__ diag, double w,
                                                        double* __restrict__ y, const double* __restrict__ dotv,
                                                        double* __restrict__ partials) {
  using C = SymCfg<NDOF>;
  const long long N = g.nOwned;
  extern __shared__ __align__(16) double sym_smem[];
  double* tin = sym_smem;                 // [2][4][NDOF][NT]  in-plane contributions of the current step (by parity of the step)
  double* tup = sym_smem + 2 * C::TIN;    // [2][9][NDOF][NT]  contributions to the plane above
  __shared__ double wred[3][SYM_NT / 32];
  const int tid = threadIdx.x, tx = tid % SYM_TX, ty = tid / SYM_TX;
  const int i = blockIdx.x * SYM_TX + tx, j = blockIdx.y * SYM_TY + ty;
  const int kA = blockIdx.z * zl, kB = min(kA + zl, g.NZ);
  const bool col = i < g.NX && j < g.NY;
  double d0 = 0.0, d1 = 0.0, d2 = 0.0;

  for (int k = kA, t = 0; k < kB; ++k, ++t) {
    const long long ln = ((long long)k * g.NY + j) * g.NX + i;
    double* tin_w = tin + (t & 1) * C::TIN;
    double* tup_w = tup + (t & 1) * C::TUP;
    const double* tup_r = tup + ((t + 1) & 1) * C::TUP;   // written in the previous step
    double acc[NDOF], xc[NDOF];
#pragma unroll
    for (int d = 0; d < NDOF; ++d) acc[d] = 0.0, xc[d] = 0.0;
    if (col) {
#pragma unroll
      for (int c = 0; c < NDOF; ++c) xc[c] = __ldg(x + ln * NDOF + c);
      // ---- A. own blocks: diagonal, 13 upper neighbours (+ the transposed products for them), in two batches of 7 slots:
      //      all loads of a batch are issued before the first is used (two memory round trips per plane instead of 14)
#pragma unroll
      for (int bt = 0; bt < 2; ++bt) {
        double blk[7][NDOF * NDOF], xv[7][NDOF];
        bool ok[7];
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          const int slot = 7 * bt + q, o = slot + 4, dk = o / 9, dj = (o / 3) % 3 - 1, di = o % 3 - 1;
          const int i2 = i + di, j2 = j + dj, k2 = k + dk;
          ok[q] = slot == 0 || (i2 >= 0 && i2 < g.NX && j2 >= 0 && j2 < g.NY && k2 < g.NZ);
          const long long nb = ok[q] ? ln + di + (long long)dj * g.NX + (long long)dk * g.plane : ln;
          const double* sp = S + (long long)slot * NDOF * NDOF * N + ln;   // (blocks of missing neighbours are stored as zeros)
#pragma unroll
          for (int e = 0; e < NDOF * NDOF; ++e) blk[q][e] = __ldg(sp + (long long)e * N);
#pragma unroll
          for (int c = 0; c < NDOF; ++c) xv[q][c] = __ldg(x + nb * NDOF + c);
        }
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          const int slot = 7 * bt + q, o = slot + 4, dk = o / 9, dj = (o / 3) % 3 - 1, di = o % 3 - 1;
#pragma unroll
          for (int d = 0; d < NDOF; ++d)
#pragma unroll
            for (int c = 0; c < NDOF; ++c) acc[d] = fma(blk[q][d * NDOF + c], xv[q][c], acc[d]);
          if (slot > 0 && ok[q]) {
            // the neighbour's rows take B^T x_i: on chip when the neighbour is a node of this tile (and, upward, of this chunk)
            const int tx2 = tx + di, ty2 = ty + dj;
            if (tx2 >= 0 && tx2 < SYM_TX && ty2 >= 0 && ty2 < SYM_TY && (dk == 0 || k + dk < kB)) {
              double* dst = (dk == 0 ? tin_w : tup_w) + sym_sub(slot) * NDOF * SYM_NT + ty2 * SYM_TX + tx2;
#pragma unroll
              for (int c = 0; c < NDOF; ++c) {
                double tv = 0.0;
#pragma unroll
                for (int d = 0; d < NDOF; ++d) tv = fma(blk[q][d * NDOF + c], xc[d], tv);
                dst[c * SYM_NT] = tv;
              }
            }
          }
        }
      }
      // ---- lower neighbours: on-chip contributions of the plane below (left in the previous step), or -- when the link
      //      leaves the tile / the chunk -- the block stored with the neighbour, transposed
#pragma unroll
      for (int slot = 1; slot < SYM_SLOTS; ++slot) {
        const int o = slot + 4, dk = o / 9, dj = (o / 3) % 3 - 1, di = o % 3 - 1;
        const int i2 = i - di, j2 = j - dj, k2 = k - dk;
        if (i2 >= 0 && i2 < g.NX && j2 >= 0 && j2 < g.NY && k2 >= 0) {
          const int tx2 = tx - di, ty2 = ty - dj;
          const bool onchip = tx2 >= 0 && tx2 < SYM_TX && ty2 >= 0 && ty2 < SYM_TY && (dk == 0 || k2 >= kA);
          if (!onchip) {
            const long long nb = ln - di - (long long)dj * g.NX - (long long)dk * g.plane;
            const double* sp = S + (long long)slot * NDOF * NDOF * N + nb;
            double xv[NDOF];
#pragma unroll
            for (int c = 0; c < NDOF; ++c) xv[c] = __ldg(x + nb * NDOF + c);
#pragma unroll
            for (int c = 0; c < NDOF; ++c)
#pragma unroll
              for (int d = 0; d < NDOF; ++d) acc[d] = fma(__ldg(sp + (long long)(c * NDOF + d) * N), xv[c], acc[d]);
          } else if (dk == 1) {
            const double* src = tup_r + sym_sub(slot) * NDOF * SYM_NT + tid;
#pragma unroll
            for (int d = 0; d < NDOF; ++d) acc[d] += src[d * SYM_NT];
          }
        }
      }
    }
    __syncthreads();
    // ---- B. in-plane contributions of this step, epilogue
    if (col) {
#pragma unroll
      for (int slot = 1; slot <= 4; ++slot) {
        const int o = slot + 4, dj = (o / 3) % 3 - 1, di = o % 3 - 1;
        const int tx2 = tx - di, ty2 = ty - dj;
        if (tx2 >= 0 && tx2 < SYM_TX && ty2 >= 0 && ty2 < SYM_TY && i - di < g.NX && j - dj < g.NY) {   // (i - di, j - dj >= 0: inside the tile)
          const double* src = tin_w + sym_sub(slot) * NDOF * SYM_NT + tid;
#pragma unroll
          for (int d = 0; d < NDOF; ++d) acc[d] += src[d * SYM_NT];
        }
      }
#pragma unroll
      for (int d = 0; d < NDOF; ++d) {
        const long long r = ln * NDOF + d;
        double out;
        if (MODE == SMODE_SPMV) out = acc[d];
        else if (MODE == SMODE_RESID) out = b[r] - acc[d];
        else out = xc[d] + w * ((b[r] - acc[d]) / diag[r]);
        y[r] = out;
        if (partials) {
          const double dvv = dotv ? dotv[r] : 0.0;
          d0 = fma(out, xc[d], d0);
          d1 = fma(xc[d], dvv, d1);
          d2 = fma(out, dvv, d2);
        }
      }
    }
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
      const long long bid = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      partials[3 * bid] = s0;
      partials[3 * bid + 1] = s1;
      partials[3 * bid + 2] = s2;
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

// planes per CTA: at most 4 CTAs per SM worth of partial triples (pmb_spmv_ws_doubles), few CTAs lost to the last wave, short
// chunks cost one gathered plane each
static int sym_zl(const Geo& g, long long tiles, int sms) {
  const long long slots = 2LL * sms, cap = 4LL * sms;
  int best = g.NZ;
  double best_cost = 1e300;
  for (int chunks = 1; chunks <= g.NZ; ++chunks) {
    const int zl = (g.NZ + chunks - 1) / chunks;
    const long long ctas = tiles * ((g.NZ + zl - 1) / zl);
    if (ctas > cap && chunks > 1) break;
    const long long waves = (ctas + slots - 1) / slots;
    const double cost = (double)waves * (zl + 0.35);   // the first plane of a chunk gathers 9 of its 27 blocks a second time
    if (cost < best_cost - 1e-12) best_cost = cost, best = zl;
  }
  return best;
}

template <int NDOF, int MODE>
static int launch_sym(const Geo& g, const double* S, const double* x, const double* b, const double* diag, double w, double* y,
                      const double* dotv, double* dot_out, double* ws, cudaStream_t st) {
  using C = SymCfg<NDOF>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(sym_kernel<NDOF, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return pmb_set_error("sym_kernel attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int nbx = (g.NX + SYM_TX - 1) / SYM_TX, nby = (g.NY + SYM_TY - 1) / SYM_TY;
  const int zl = sym_zl(g, (long long)nbx * nby, sym_sm_count());
  const int nbz = (g.NZ + zl - 1) / zl;
  PMB_REQUIRE(nby <= 65535 && nbz <= 65535, "pmb_sym_spmv: grid too large");
  PMB_REQUIRE(!dot_out || (long long)nbx * nby * nbz <= 4LL * sym_sm_count(), "pmb_sym_spmv: grid too large for the fused dot products");
  sym_kernel<NDOF, MODE><<<dim3(nbx, nby, nbz), SYM_NT, C::SMEM, st>>>(g, zl, S, x, b, diag, w, y, dotv, dot_out ? ws : nullptr);
  PMB_CHECK_LAUNCH("pmb_sym_spmv");
  if (dot_out) {
    reduce_triples_kernel<<<1, 1024, 0, st>>>(ws, (long long)nbx * nby * nbz, dot_out);
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
