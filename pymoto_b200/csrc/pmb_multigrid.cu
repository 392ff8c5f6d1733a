// libpmb: geometric-multigrid transfer operators (K4, K5), Galerkin coarse operator (K6), coarsest-level dense
// inverse (K7).
//
// Replaces pymoto/solvers/iterative.py:178-220 (the prolongation matrix R is never formed: its trilinear
// weights 1, 1/2, 1/4, 1/8 are closed-form), :244 (R^T r via csc_matvec), :250 (u += R u_c via csr_matvec),
// :173 (R^T A R via two csr_matmat) and the coarsest-level splu of pymoto/solvers/sparse.py:533-550.
#include <cooperative_groups.h>
#include <cstdlib>
#include "pmb_tilestream.cuh"

// ------------------------------------------------------------------------------------------------- K4
// rc[C] = sum over the <= 27 fine nodes 2C+d of w(d) rf[fine]; ascending fine node number, separate multiply and
// add: the order and rounding of scipy's csc_matvec for R^T (bit-identical to the reference).
// Launch geometry of both transfer kernels: x = the dofs of one node row (coalesced), y = node row j, z = local plane --
// no per-thread 64-bit division (the flat-index version spent its time there: 171 us for 206 MB at 256x128x128).
template <int NDOF>
__global__ void __launch_bounds__(128) restrict_kernel(Geo gf, Geo gc, const double* __restrict__ rf, double* __restrict__ rc) {
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= gc.NX * NDOF) return;
  const int I = t / NDOF, d = t - I * NDOF, J = blockIdx.y, K = (int)blockIdx.z + gc.kz0;
  double acc = 0.0;
  for (int dk = -1; dk <= 1; ++dk) {
    int fk = 2 * K + dk;
    if (fk < 0 || fk >= gf.NZ) continue;
    for (int dj = -1; dj <= 1; ++dj) {
      int fj = 2 * J + dj;
      if (fj < 0 || fj >= gf.NY) continue;
      const double* row = rf + ((long long)(fk - gf.kz0) * gf.NY + fj) * gf.NX * NDOF + d;
      for (int di = -1; di <= 1; ++di) {
        int fi = 2 * I + di;
        if (fi < 0 || fi >= gf.NX) continue;
        double w = (dk ? 0.5 : 1.0) * (dj ? 0.5 : 1.0) * (di ? 0.5 : 1.0);
        acc = __dadd_rn(acc, __dmul_rn(w, __ldg(row + fi * NDOF)));
      }
    }
  }
  rc[((long long)blockIdx.z * gc.NY + J) * gc.NX * NDOF + t] = acc;
}

extern "C" int pmb_restrict(const pmb_grid* pf, const pmb_grid* pc, const double* rf, double* rc, void* stream) {
  if (validate_grid(pf, "pmb_restrict(fine)") || validate_grid(pc, "pmb_restrict(coarse)")) return 1;
  PMB_REQUIRE(pf->nx == 2 * pc->nx && pf->ny == 2 * pc->ny && pf->nz == 2 * pc->nz && pf->ndof == pc->ndof,
              "pmb_restrict: coarse grid is not the 2:1 coarsening of the fine grid");
  PMB_REQUIRE(rf && rc, "pmb_restrict: NULL pointer argument");
  Geo gf = make_geo(pf), gc = make_geo(pc);
  PMB_REQUIRE(gc.NY <= 65535 && gc.nzl <= 65535, "pmb_restrict: grid too large");
  const dim3 grid((gc.NX * gc.ndof + 127) / 128, gc.NY, gc.nzl);
  cudaStream_t st = (cudaStream_t)stream;
  switch (gc.ndof) {
    case 1: restrict_kernel<1><<<grid, 128, 0, st>>>(gf, gc, rf, rc); break;
    case 2: restrict_kernel<2><<<grid, 128, 0, st>>>(gf, gc, rf, rc); break;
    case 3: restrict_kernel<3><<<grid, 128, 0, st>>>(gf, gc, rf, rc); break;
  }
  PMB_CHECK_LAUNCH("pmb_restrict");
  return 0;
}

// ------------------------------------------------------------------------------------------------- K5
// uf[f] += sum over the <= 8 coarse parents (ascending coarse node number) of w uc[parent]
template <int NDOF>
__global__ void __launch_bounds__(128) prolong_add_kernel(Geo gf, Geo gc, const double* __restrict__ uc, double* __restrict__ uf) {
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= gf.NX * NDOF) return;
  const int i = t / NDOF, d = t - i * NDOF, j = blockIdx.y, k = (int)blockIdx.z + gf.kz0;
  const int ni = (i & 1) + 1, nj = (j & 1) + 1, nk = (k & 1) + 1;
  const int I0 = i >> 1, J0 = j >> 1, K0 = k >> 1;
  const double w = (nk == 2 ? 0.5 : 1.0) * (nj == 2 ? 0.5 : 1.0) * (ni == 2 ? 0.5 : 1.0);
  double acc = 0.0;
  for (int a = 0; a < nk; ++a)
    for (int b = 0; b < nj; ++b) {
      const double* row = uc + (((long long)(K0 + a - gc.kz0) * gc.NY + (J0 + b)) * gc.NX + I0) * NDOF + d;
      for (int c = 0; c < ni; ++c) acc = __dadd_rn(acc, __dmul_rn(w, __ldg(row + c * NDOF)));
    }
  double* dst = uf + ((long long)blockIdx.z * gf.NY + j) * gf.NX * NDOF + t;
  *dst = __dadd_rn(*dst, acc);
}

extern "C" int pmb_prolong_add(const pmb_grid* pf, const pmb_grid* pc, const double* uc, double* uf, void* stream) {
  if (validate_grid(pf, "pmb_prolong_add(fine)") || validate_grid(pc, "pmb_prolong_add(coarse)")) return 1;
  PMB_REQUIRE(pf->nx == 2 * pc->nx && pf->ny == 2 * pc->ny && pf->nz == 2 * pc->nz && pf->ndof == pc->ndof,
              "pmb_prolong_add: coarse grid is not the 2:1 coarsening of the fine grid");
  PMB_REQUIRE(uc && uf, "pmb_prolong_add: NULL pointer argument");
  Geo gf = make_geo(pf), gc = make_geo(pc);
  PMB_REQUIRE(gf.NY <= 65535 && gf.nzl <= 65535, "pmb_prolong_add: grid too large");
  const dim3 grid((gf.NX * gf.ndof + 127) / 128, gf.NY, gf.nzl);
  cudaStream_t st = (cudaStream_t)stream;
  switch (gf.ndof) {
    case 1: prolong_add_kernel<1><<<grid, 128, 0, st>>>(gf, gc, uc, uf); break;
    case 2: prolong_add_kernel<2><<<grid, 128, 0, st>>>(gf, gc, uc, uf); break;
    case 3: prolong_add_kernel<3><<<grid, 128, 0, st>>>(gf, gc, uc, uf); break;
  }
  PMB_CHECK_LAUNCH("pmb_prolong_add");
  return 0;
}

// ------------------------------------------------------------------------------------------------- K6
// Ac = R^T A R as two streaming gather passes (no atomics, no SpGEMM, R never formed):
//   pass 1 (columns): B[i, C] = sum_{j nbr of i, j in supp(C)} w_C(j) A[i, j]   for fine row-node i and the <= 27 coarse
//                     nodes C around it; the fine matrix is streamed once through the same TMA tile ring as the
//                     operator kernel and each (i, C) block is gathered from shared memory;
//   pass 2 (rows):    Ac[I, C] = sum_{i in supp(I)} w_I(i) B[i, C], written straight into the coarse stencil-CSR layout.
// B is stored padded as [fine node][NDOF^2][27 coarse slots] (slot = per-dimension offset C - (f >> 1) + 1), so that
// the threads of a warp (consecutive slots) store and load consecutive doubles.
template <int NDOF>
struct GalCfg;
template <>
struct GalCfg<3> { static constexpr int T = 16, STAGES = 2; };
template <>
struct GalCfg<2> { static constexpr int T = 32, STAGES = 2; };
template <>
struct GalCfg<1> { static constexpr int T = 96, STAGES = 2; };
static constexpr int GAL_NT = 256;

// trilinear weight of fine index p in the support of coarse index C (0 outside): 1 at p = 2C, 1/2 at 2C +- 1
__device__ __forceinline__ double prolong_w(int p, int C) {
  const int t = p - 2 * C;
  return t == 0 ? 1.0 : ((t == 1 || t == -1) ? 0.5 : 0.0);
}

// Pass 1 is separable: the 27 fine neighbour blocks of a node are collapsed onto its 27 coarse slots one dimension
// at a time (x, then y, then z), each step a uniform 3-term combination -> no divergent trip counts.
template <int NDOF>
__global__ void __launch_bounds__(GAL_NT, 2) galerkin_cols_kernel(Geo gf, Geo gc, int ntiles, int tiles_per_row, int stream_hint,
                                                                  const double* __restrict__ Af, double* __restrict__ B) {
  constexpr int T = GalCfg<NDOF>::T, STAGES = GalCfg<NDOF>::STAGES;
  constexpr int TD = (T * NDOF * NDOF * 27 + 2 + 1) / 2 * 2;
  constexpr int NB = NDOF * NDOF;      // doubles per block
  constexpr int PB = 27 * NB;          // padded doubles per node in the intermediates
  extern __shared__ __align__(128) double sTiles[];   // STAGES x TD ring, then one T x PB intermediate
  double* s1 = sTiles + (size_t)STAGES * TD;
  static_assert(TD >= T * PB, "the ring slot doubles as the second intermediate");
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  const int tid = threadIdx.x;
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  auto issue = [&](int i) {
    const int s = i % STAGES;
    const TileGeom t = tile_geom<NDOF, T>(gf, (int)blockIdx.x + i * (int)gridDim.x, tiles_per_row);
    const long long lo = t.e0 & ~1LL;
    const long long hi = (t.e1 + 1) & ~1LL;
    const unsigned bytes = (unsigned)((hi - lo) * sizeof(double));
    mbar_expect_tx(&full_bar[s], bytes);
    tma_load_1d(sTiles + (size_t)s * TD, Af + lo, bytes, &full_bar[s], stream_hint != 0);
  };
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < STAGES && i < my_tiles; ++i) issue(i);
  }
  __syncthreads();

  for (int it = 0; it < my_tiles; ++it) {
    const int s = it % STAGES;
    const TileGeom t = tile_geom<NDOF, T>(gf, (int)blockIdx.x + it * (int)gridDim.x, tiles_per_row);
    const long long lnode0 = ((long long)(t.k - gf.kz0) * gf.NY + t.j) * gf.NX + t.i0;
    const double* tile = sTiles + (size_t)s * TD + (t.e0 - (t.e0 & ~1LL));
    const long long per = (long long)NB * t.cy * t.cz;
    const int fj = t.j, fk = t.k;
    double* s2 = sTiles + (size_t)s * TD;  // overwrites the staged values once the x-pass has consumed them
    mbar_wait(&full_bar[s], (unsigned)((it / STAGES) & 1));

    // ---- x: s1[g][ez][ey][sx] = sum_ex w_x A[g][(ez,ey,ex)]
    for (int p = tid; p < t.ni * 27; p += GAL_NT) {
      const int g = p / 27, q = p - g * 27;
      const int ez = q / 9, ey = (q / 3) % 3, sx = q % 3;
      const int fi = t.i0 + g;
      const int pk = fk + ez - 1, pj = fj + ey - 1;
      double acc[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) acc[c] = 0.0;
      if (pk >= 0 && pk < gf.NZ && pj >= 0 && pj < gf.NY) {
        const int cx = cnt1(fi, gf.NX), ilo = max(fi - 1, 0);
        const int L = cx * t.cy * t.cz * NDOF;
        const double* nodep = tile + per * (pre1(fi, gf.NX) - pre1(t.i0, gf.NX));
        const int Cx = (fi >> 1) + sx - 1;
        const int rowb = ((pk - t.klo) * t.cy + (pj - t.jlo)) * cx;
#pragma unroll
        for (int ex = -1; ex <= 1; ++ex) {
          const int pi = fi + ex;
          const double w = (pi >= 0 && pi < gf.NX && Cx >= 0 && Cx < gc.NX) ? prolong_w(pi, Cx) : 0.0;
          if (w != 0.0) {
            const double* ap = nodep + (rowb + (pi - ilo)) * NDOF;
#pragma unroll
            for (int a = 0; a < NDOF; ++a)
#pragma unroll
              for (int c = 0; c < NDOF; ++c) acc[a * NDOF + c] = fma(w, ap[a * L + c], acc[a * NDOF + c]);
          }
        }
      }
      double* o = s1 + (size_t)g * PB + q * NB;
#pragma unroll
      for (int c = 0; c < NB; ++c) o[c] = acc[c];
    }
    __syncthreads();
    // ---- y: s2[g][ez][sy][sx] = sum_ey w_y s1[g][ez][ey][sx]       (weights uniform over the tile)
    for (int p = tid; p < t.ni * 27; p += GAL_NT) {
      const int g = p / 27, q = p - g * 27;
      const int ez = q / 9, sy = (q / 3) % 3, sx = q % 3;
      const int Cy = (fj >> 1) + sy - 1;
      double acc[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) acc[c] = 0.0;
#pragma unroll
      for (int ey = 0; ey < 3; ++ey) {
        const int pj = fj + ey - 1;
        const double w = (pj >= 0 && pj < gf.NY && Cy >= 0 && Cy < gc.NY) ? prolong_w(pj, Cy) : 0.0;
        const double* ip = s1 + (size_t)g * PB + ((ez * 3 + ey) * 3 + sx) * NB;
#pragma unroll
        for (int c = 0; c < NB; ++c) acc[c] = fma(w, ip[c], acc[c]);
      }
      double* o = s2 + (size_t)g * PB + q * NB;
#pragma unroll
      for (int c = 0; c < NB; ++c) o[c] = acc[c];
    }
    __syncthreads();
    // ---- z: B[g][sz][sy][sx] = sum_ez w_z s2[g][ez][sy][sx]  -> global, layout [fine node][NB][27 slots]
    for (int p = tid; p < t.ni * 27; p += GAL_NT) {
      const int g = p / 27, q = p - g * 27;
      const int sz = q / 9, sy = (q / 3) % 3, sx = q % 3;
      const int fi = t.i0 + g;
      const int Cx = (fi >> 1) + sx - 1, Cy = (fj >> 1) + sy - 1, Cz = (fk >> 1) + sz - 1;
      const bool valid = ((fi & 1) == 0 || sx >= 1) && ((fj & 1) == 0 || sy >= 1) && ((fk & 1) == 0 || sz >= 1) && Cx >= 0 &&
                         Cx < gc.NX && Cy >= 0 && Cy < gc.NY && Cz >= 0 && Cz < gc.NZ;
      if (!valid) continue;
      double acc[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) acc[c] = 0.0;
#pragma unroll
      for (int ez = 0; ez < 3; ++ez) {
        const int pk = fk + ez - 1;
        const double w = (pk >= 0 && pk < gf.NZ) ? prolong_w(pk, Cz) : 0.0;
        const double* ip = s2 + (size_t)g * PB + ((ez * 3 + sy) * 3 + sx) * NB;
#pragma unroll
        for (int c = 0; c < NB; ++c) acc[c] = fma(w, ip[c], acc[c]);
      }
      double* bp = B + (lnode0 + g) * (27 * NB) + q;
#pragma unroll
      for (int c = 0; c < NB; ++c) bp[c * 27] = acc[c];
    }
    __syncthreads();
    if (tid == 0 && it + STAGES < my_tiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(it + STAGES);
    }
  }
}

// pass 2: one thread per coarse (node, neighbour slot) = one NDOF x NDOF block
template <int NDOF>
__global__ void __launch_bounds__(128) galerkin_rows_kernel(Geo gf, Geo gc, const double* __restrict__ B, double* __restrict__ Ac) {
  long long tg = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long nslots = gc.nOwned * 27;
  if (tg >= nslots) return;
  int s = (int)(tg % 27);
  long long lc = tg / 27;
  int I, J, K;
  node_ijk(gc, lc, I, J, K);
  const int Dk = s / 9 - 1, Dj = (s / 3) % 3 - 1, Di = s % 3 - 1;
  const int Ci = I + Di, Cj = J + Dj, Ck = K + Dk;
  if (Ci < 0 || Ci >= gc.NX || Cj < 0 || Cj >= gc.NY || Ck < 0 || Ck >= gc.NZ) return;
  double acc[NDOF][NDOF];
#pragma unroll
  for (int a = 0; a < NDOF; ++a)
#pragma unroll
    for (int c = 0; c < NDOF; ++c) acc[a][c] = 0.0;
  for (int dz = -1; dz <= 1; ++dz) {
    const int fk = 2 * K + dz, sk = Ck - (fk >> 1) + 1;
    if (fk < 0 || fk >= gf.NZ || sk < ((fk & 1) ? 1 : 0) || sk > 2) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      const int fj = 2 * J + dy, sj = Cj - (fj >> 1) + 1;
      if (fj < 0 || fj >= gf.NY || sj < ((fj & 1) ? 1 : 0) || sj > 2) continue;
      const double wzy = (dz ? 0.5 : 1.0) * (dy ? 0.5 : 1.0);
      for (int dx = -1; dx <= 1; ++dx) {
        const int fi = 2 * I + dx, si = Ci - (fi >> 1) + 1;
        if (fi < 0 || fi >= gf.NX || si < ((fi & 1) ? 1 : 0) || si > 2) continue;
        const double w = wzy * (dx ? 0.5 : 1.0);
        const long long lf = ((long long)(fk - gf.kz0) * gf.NY + fj) * gf.NX + fi;
        const double* bp = B + lf * (27 * NDOF * NDOF) + (sk * 9 + sj * 3 + si);
#pragma unroll
        for (int a = 0; a < NDOF; ++a)
#pragma unroll
          for (int c = 0; c < NDOF; ++c) acc[a][c] = fma(w, __ldg(bp + (a * NDOF + c) * 27), acc[a][c]);
      }
    }
  }
  const int cxc = cnt1(I, gc.NX), cyc = cnt1(J, gc.NY), czc = cnt1(K, gc.NZ);
  const int Ilo = max(I - 1, 0), Jlo = max(J - 1, 0), Klo = max(K - 1, 0);
  const long long Lc = (long long)cxc * cyc * czc * NDOF;
  const int nbrc = ((Ck - Klo) * cyc + (Cj - Jlo)) * cxc + (Ci - Ilo);
  double* op = Ac + (long long)(NDOF * NDOF) * (block_offset(gc, I, J, K) - gc.bo0) + (long long)nbrc * NDOF;
#pragma unroll
  for (int a = 0; a < NDOF; ++a)
#pragma unroll
    for (int c = 0; c < NDOF; ++c) op[a * Lc + c] = acc[a][c];
}

extern "C" long long pmb_galerkin_ws_doubles(const pmb_grid* pf) {
  if (validate_grid(pf, "pmb_galerkin_ws_doubles")) return -1;
  Geo gf = make_geo(pf);
  return gf.nOwned * 27 * gf.ndof * gf.ndof;
}

template <int NDOF>
static int launch_galerkin_cols(const Geo& gf, const Geo& gc, const double* Af, double* B, cudaStream_t st) {
  constexpr int T = GalCfg<NDOF>::T, STAGES = GalCfg<NDOF>::STAGES;
  constexpr int TD = (T * NDOF * NDOF * 27 + 2 + 1) / 2 * 2;
  const size_t smem = sizeof(double) * ((size_t)TD * STAGES + (size_t)T * 27 * NDOF * NDOF);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(galerkin_cols_kernel<NDOF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return pmb_set_error("galerkin_cols_kernel attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int tpr = (gf.NX + T - 1) / T;
  const long long ntiles = (long long)tpr * gf.NY * gf.nzl;
  PMB_REQUIRE(ntiles < 2147483647LL, "pmb_galerkin: too many tiles");
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(ntiles < 2LL * sms ? ntiles : 2LL * sms);
  const long long nnz_bytes = 8LL * NDOF * NDOF * (pre1(gf.kz0 + gf.nzl, gf.NZ) * gf.Sy * gf.Sx - gf.bo0);
  galerkin_cols_kernel<NDOF><<<grid, GAL_NT, smem, st>>>(gf, gc, (int)ntiles, tpr, nnz_bytes > (96LL << 20), Af, B);
  PMB_CHECK_LAUNCH("pmb_galerkin_cols");
  return 0;
}

template <int NDOF>
static int launch_galerkin_rows(const Geo& gf, const Geo& gc, const double* B, double* Ac, cudaStream_t st) {
  const long long nslots = gc.nOwned * 27;
  galerkin_rows_kernel<NDOF><<<(unsigned)((nslots + 127) / 128), 128, 0, st>>>(gf, gc, B, Ac);
  PMB_CHECK_LAUNCH("pmb_galerkin_rows");
  return 0;
}

static int check_galerkin_grids(const pmb_grid* pf, const pmb_grid* pc, const char* who) {
  if (validate_grid(pf, who) || validate_grid(pc, who)) return 1;
  PMB_REQUIRE(pf->nx == 2 * pc->nx && pf->ny == 2 * pc->ny && pf->nz == 2 * pc->nz && pf->ndof == pc->ndof,
              "%s: coarse grid is not the 2:1 coarsening of the fine grid", who);
  // coarse plane K is produced from fine planes 2K-1 .. 2K+1: 2K and 2K+1 must be owned (2K-1 may be the lower halo)
  PMB_REQUIRE(2 * pc->kz0 >= pf->kz0 && 2 * (pc->kz0 + pc->nzl - 1) < pf->kz0 + pf->nzl,
              "%s: coarse slab [%d,%d) does not sit inside fine slab [%d,%d)", who, pc->kz0, pc->kz0 + pc->nzl, pf->kz0,
              pf->kz0 + pf->nzl);
  return 0;
}

extern "C" int pmb_galerkin_cols(const pmb_grid* pf, const pmb_grid* pc, const double* Af, double* work, void* stream) {
  if (check_galerkin_grids(pf, pc, "pmb_galerkin_cols")) return 1;
  PMB_REQUIRE(Af && work, "pmb_galerkin_cols: NULL pointer argument");
  PMB_REQUIRE((reinterpret_cast<size_t>(Af) & 15) == 0, "pmb_galerkin_cols: fine data must be 16-byte aligned");
  Geo gf = make_geo(pf), gc = make_geo(pc);
  cudaStream_t st = (cudaStream_t)stream;
  switch (gc.ndof) {
    case 1: return launch_galerkin_cols<1>(gf, gc, Af, work, st);
    case 2: return launch_galerkin_cols<2>(gf, gc, Af, work, st);
    case 3: return launch_galerkin_cols<3>(gf, gc, Af, work, st);
  }
  return 1;
}

extern "C" int pmb_galerkin_rows(const pmb_grid* pf, const pmb_grid* pc, const double* work, double* Ac, void* stream) {
  if (check_galerkin_grids(pf, pc, "pmb_galerkin_rows")) return 1;
  PMB_REQUIRE(work && Ac, "pmb_galerkin_rows: NULL pointer argument");
  Geo gf = make_geo(pf), gc = make_geo(pc);
  cudaStream_t st = (cudaStream_t)stream;
  switch (gc.ndof) {
    case 1: return launch_galerkin_rows<1>(gf, gc, work, Ac, st);
    case 2: return launch_galerkin_rows<2>(gf, gc, work, Ac, st);
    case 3: return launch_galerkin_rows<3>(gf, gc, work, Ac, st);
  }
  return 1;
}

// ------------------------------------------------------------------------------------------------- K6 (direct)
// Level-1 operator straight from the element densities (host set-up and derivation: pymoto_b200/coarse.py):
//     Ac = sum_E sum_{p<8} s_child(E,p) G_id(E,p)  [+ the Dirichlet diagonal term, added by pmb_scatter_add]
// One warp = 32 consecutive coarse nodes of an x-row x ONE (dk, dj) line of neighbour slots (9 warps per CTA, the three di
// slots in turn; 2 CTAs per SM so that one CTA computes while another writes out): the element / local-node structure is
// warp-uniform, so the table entry G[id][a][b] is a shared-memory broadcast for unmasked children (id = p)
// and only coarse elements with a Dirichlet child (cidx >= 0) read per-child ids and the table in global memory.
// The CTA builds the contiguous run of its 32 nodes in shared memory (same layout as assemble_kernel) and writes it
// with coalesced stores.  Densities of the 66 x 4 x 4 fine elements around the run are staged once (0.0 outside the grid).
template <int NDOF>
__global__ void __launch_bounds__(9 * 32, 2)
    galerkin_direct_kernel(Geo gf, Geo gc, const double* __restrict__ Gtab, const int* __restrict__ cidx,
                           const unsigned short* __restrict__ child_ids, const double* __restrict__ s, double* __restrict__ Ac) {
  constexpr int T = 32, ND2 = NDOF * NDOF, NT = 9 * 32, SW = 2 * T + 2;
  extern __shared__ double dyn[];
  double* tile = dyn;                    // T * ND2 * 27 doubles: the output run
  double* sG = tile + T * ND2 * 27;      // 8 * 64 * ND2: unmasked child tables
  double* sS = sG + 8 * 64 * ND2;        // 4 x 4 x SW staged densities
  const int tid = threadIdx.x, gI = tid & 31, line = tid >> 5;
  const int I0 = blockIdx.x * T, J = blockIdx.y, kl = blockIdx.z, K = gc.kz0 + kl;
  const int ni = min(T, gc.NX - I0);
  for (int p = tid; p < 8 * 64 * ND2; p += NT) sG[p] = Gtab[p];
  for (int p = tid; p < 16 * SW; p += NT) {
    const int tx = p % SW, ty = (p / SW) & 3, tz = p / (4 * SW);
    const int ei = 2 * I0 - 2 + tx, ej = 2 * J - 2 + ty, ek = 2 * K - 2 + tz;
    // fine layers below the fine slab's two halo layers / above its last owned node plane are never needed
    const bool in = ei >= 0 && ei < gf.nx && ej >= 0 && ej < gf.ny && ek >= 0 && ek < gf.nzE && ek - gf.kz0 >= -2 &&
                    ek - gf.kz0 < gf.nzl;
    sS[p] = in ? __ldg(s + ((long long)(ek - gf.kz0) * gf.ny + ej) * gf.nx + ei) : 0.0;
  }
  __syncthreads();

  const int cy = cnt1(J, gc.NY), cz = cnt1(K, gc.NZ);
  const int Jlo = max(J - 1, 0), Klo = max(K - 1, 0);
  const long long per = (long long)ND2 * cy * cz;
  const long long rowbase = pre1(K, gc.NZ) * gc.Sy * gc.Sx + (long long)cz * (pre1(J, gc.NY) * gc.Sx) - gc.bo0;
  const long long e0 = (long long)ND2 * rowbase + per * pre1(I0, gc.NX);
  const int nelem = (int)(per * (pre1(I0 + ni, gc.NX) - pre1(I0, gc.NX)));

  const int dk = line / 3 - 1, dj = line % 3 - 1;   // warp-uniform: one (dk, dj) line of neighbour slots per warp
  const int I = I0 + gI, CJ = J + dj, CK = K + dk;
  if (gI < ni && CJ >= 0 && CJ < gc.NY && CK >= 0 && CK < gc.NZ) {
    const int cx = cnt1(I, gc.NX), Ilo = max(I - 1, 0);
    const int L = cx * cy * cz * NDOF;
#pragma unroll
    for (int di = -1; di <= 1; ++di) {
      const int CI = I + di;
      if (CI < 0 || CI >= gc.NX) continue;
      double acc[ND2];
#pragma unroll
      for (int q = 0; q < ND2; ++q) acc[q] = 0.0;
      for (int oz = 0; oz < 2; ++oz) {
        const int az = 1 - oz, bz = az + dk, EK = K - 1 + oz;
        if (bz < 0 || bz > 1 || EK < 0 || EK >= gc.nzE) continue;
        for (int oy = 0; oy < 2; ++oy) {
          const int ay = 1 - oy, by = ay + dj, EJ = J - 1 + oy;
          if (by < 0 || by > 1 || EJ < 0 || EJ >= gc.ny) continue;
#pragma unroll
          for (int ox = 0; ox < 2; ++ox) {
            const int ax = 1 - ox, bx = ax + di, EI = I - 1 + ox;
            if (bx < 0 || bx > 1) continue;        // compile-time
            if (EI < 0 || EI >= gc.nx) continue;   // per lane (first / last node of the row)
            const int a = ax + 2 * ay + 4 * az, bn = bx + 2 * by + 4 * bz;
            const int m = cidx ? __ldg(cidx + ((long long)EK * gc.ny + EJ) * gc.nx + EI) : -1;
            const double* sp = sS + ((2 * oz) * 4 + 2 * oy) * SW + 2 * (gI + ox);
            if (m < 0) {
              const double* gp = sG + (a * 8 + bn) * ND2;
#pragma unroll
              for (int p = 0; p < 8; ++p) {
                const double sv = sp[((p >> 2) * 4 + ((p >> 1) & 1)) * SW + (p & 1)];
#pragma unroll
                for (int q = 0; q < ND2; ++q) acc[q] = fma(sv, gp[p * 64 * ND2 + q], acc[q]);
              }
            } else {
              const unsigned short* ids = child_ids + 8 * (long long)m;
#pragma unroll
              for (int p = 0; p < 8; ++p) {
                const double sv = sp[((p >> 2) * 4 + ((p >> 1) & 1)) * SW + (p & 1)];
                const double* gp = Gtab + ((long long)__ldg(ids + p) * 64 + a * 8 + bn) * ND2;
#pragma unroll
                for (int q = 0; q < ND2; ++q) acc[q] = fma(sv, __ldg(gp + q), acc[q]);
              }
            }
          }
        }
      }
      const int nbr = ((CK - Klo) * cy + (CJ - Jlo)) * cx + (CI - Ilo);
      double* nodep = tile + per * (pre1(I, gc.NX) - pre1(I0, gc.NX)) + nbr * NDOF;
#pragma unroll
      for (int d = 0; d < NDOF; ++d)
#pragma unroll
        for (int c = 0; c < NDOF; ++c) nodep[d * L + c] = acc[d * NDOF + c];
    }
  }
  __syncthreads();
  for (int q = tid; q < nelem; q += NT) Ac[e0 + q] = tile[q];
}

template <int NDOF>
static int launch_galerkin_direct(const Geo& gf, const Geo& gc, const double* Gtab, const int* cidx, const unsigned short* child_ids,
                                  const double* s, double* Ac, cudaStream_t st) {
  constexpr int T = 32, ND2 = NDOF * NDOF;
  const size_t smem = sizeof(double) * ((size_t)T * ND2 * 27 + 8 * 64 * ND2 + 16 * (2 * T + 2));
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(galerkin_direct_kernel<NDOF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return pmb_set_error("galerkin_direct_kernel attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  dim3 blocks((gc.NX + T - 1) / T, gc.NY, gc.nzl);
  galerkin_direct_kernel<NDOF><<<blocks, 9 * 32, smem, st>>>(gf, gc, Gtab, cidx, child_ids, s, Ac);
  PMB_CHECK_LAUNCH("pmb_galerkin_direct");
  return 0;
}

extern "C" int pmb_galerkin_direct(const pmb_grid* pf, const pmb_grid* pc, const double* Gtab, const int* cidx,
                                   const unsigned short* child_ids, const double* s, double* Ac, void* stream) {
  if (check_galerkin_grids(pf, pc, "pmb_galerkin_direct")) return 1;
  PMB_REQUIRE(pf->nz > 0, "pmb_galerkin_direct: 3-D grids only");
  PMB_REQUIRE(Gtab && s && Ac, "pmb_galerkin_direct: NULL pointer argument");
  PMB_REQUIRE(!cidx || child_ids, "pmb_galerkin_direct: cidx without child_ids");
  Geo gf = make_geo(pf), gc = make_geo(pc);
  PMB_REQUIRE(gc.NY <= 65535 && gc.nzl <= 65535, "pmb_galerkin_direct: grid too large for the 3-D launch");
  cudaStream_t st = (cudaStream_t)stream;
  switch (gc.ndof) {
    case 1: return launch_galerkin_direct<1>(gf, gc, Gtab, cidx, child_ids, s, Ac, st);
    case 2: return launch_galerkin_direct<2>(gf, gc, Gtab, cidx, child_ids, s, Ac, st);
    case 3: return launch_galerkin_direct<3>(gf, gc, Gtab, cidx, child_ids, s, Ac, st);
  }
  return 1;
}

// data[idx[i]] += val[i] for unique idx (the constant Dirichlet term of the direct coarse operator)
__global__ void __launch_bounds__(256) scatter_add_kernel(long long n, const long long* __restrict__ idx, const double* __restrict__ val,
                                                           double* __restrict__ data) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) data[idx[t]] += val[t];
}

extern "C" int pmb_scatter_add(long long n, const long long* idx, const double* val, double* data, void* stream) {
  if (n <= 0) return 0;
  PMB_REQUIRE(idx && val && data, "pmb_scatter_add: NULL pointer argument");
  scatter_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, idx, val, data);
  PMB_CHECK_LAUNCH("pmb_scatter_add");
  return 0;
}

extern "C" int pmb_galerkin(const pmb_grid* pf, const pmb_grid* pc, const double* Af, double* Ac, double* work, void* stream) {
  if (pmb_galerkin_cols(pf, pc, Af, work, stream)) return 1;
  return pmb_galerkin_rows(pf, pc, work, Ac, stream);
}

// ------------------------------------------------------------------------------------------------- K7
__global__ void __launch_bounds__(256) densify_kernel(Geo g, const double* __restrict__ data, double* __restrict__ dense, long long n) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long nslots = g.nOwned * g.ndof * 27;
  if (t >= nslots) return;
  int s = (int)(t % 27);
  long long r = t / 27;
  long long ln = r / g.ndof;
  int d = (int)(r - ln * g.ndof);
  int i, j, k;
  node_ijk(g, ln, i, j, k);
  int dk = s / 9 - 1, dj = (s / 3) % 3 - 1, di = s % 3 - 1;
  int ci = i + di, cj = j + dj, ck = k + dk;
  if (ci < 0 || ci >= g.NX || cj < 0 || cj >= g.NY || ck < 0 || ck >= g.NZ) return;
  int cx = cnt1(i, g.NX), cy = cnt1(j, g.NY), cz = cnt1(k, g.NZ);
  int ilo = max(i - 1, 0), jlo = max(j - 1, 0), klo = max(k - 1, 0);
  long long L = (long long)cx * cy * cz * g.ndof;
  int nbr = ((ck - klo) * cy + (cj - jlo)) * cx + (ci - ilo);
  long long off = (long long)(g.ndof * g.ndof) * (block_offset(g, i, j, k) - g.bo0) + d * L + (long long)nbr * g.ndof;
  long long c = ((long long)ck * g.NY + cj) * g.NX + ci;
  for (int cd = 0; cd < g.ndof; ++cd) dense[r * n + c * g.ndof + cd] = data[off + cd];
}

extern "C" int pmb_densify(const pmb_grid* p, const double* data, double* dense, void* stream) {
  if (validate_grid(p, "pmb_densify")) return 1;
  PMB_REQUIRE(p->kz0 == 0 && p->nzl == p->nz + 1, "pmb_densify: needs the whole grid on one rank");
  PMB_REQUIRE(data && dense, "pmb_densify: NULL pointer argument");
  Geo g = make_geo(p);
  long long n = g.nOwned * g.ndof;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(dense, 0, sizeof(double) * n * n, st);
  if (e != cudaSuccess) return pmb_set_error("pmb_densify: %s", cudaGetErrorString(e));
  long long nslots = n * 27;
  densify_kernel<<<(unsigned)((nslots + 255) / 256), 256, 0, st>>>(g, data, dense, n);
  PMB_CHECK_LAUNCH("pmb_densify");
  return 0;
}

// In-place BLOCKED Gauss-Jordan inverse without pivoting (SPD input => every pivot block is SPD).  Per block of 32
// pivots: (1) one CTA inverts the 32x32 pivot block P = inv(A_kk) in shared memory, (2) the pivot row panel becomes
// R = P A_k,: (R_k,k = P) while the old pivot column panel C = A_:,k is saved, (3) one rank-32 update of the whole
// matrix: rows of the block <- R, every other row i: A_i,: <- (A_i,: with the block columns zeroed) - C_i R.
// 3 launches per 32 pivots instead of 2 per pivot; the matrix (<= a few 10 MB) stays in L2.
static constexpr int GJB = 32;

__global__ void __launch_bounds__(GJB* GJB) gj_pivot_block_kernel(int n, int k0, int bs, const double* __restrict__ A,
                                                                   double* __restrict__ P, int* __restrict__ info) {
  __shared__ double M[GJB][GJB + 1];
  const int tx = threadIdx.x % GJB, ty = threadIdx.x / GJB;
  const bool act = tx < bs && ty < bs;
  M[ty][tx] = act ? A[(long long)(k0 + ty) * n + (k0 + tx)] : (tx == ty ? 1.0 : 0.0);
  for (int k = 0; k < bs; ++k) {
    __syncthreads();
    const double piv = M[k][k];
    const double rk = (tx == k ? 1.0 : M[k][tx]) / piv;
    const double ck = M[ty][k];
    const double cur = (tx == k) ? 0.0 : M[ty][tx];
    if (threadIdx.x == 0 && !(piv > 0.0) && *info == 0) *info = k0 + k + 1;
    __syncthreads();
    M[ty][tx] = (ty == k) ? rk : cur - ck * rk;
  }
  __syncthreads();
  if (act) P[ty * GJB + tx] = M[ty][tx];
}

// R[t][j] = (P A_k,:)[t][j] outside the block columns, P inside; C[i][t] = A[i][k0+t]
__global__ void __launch_bounds__(256) gj_panels_kernel(int n, int k0, int bs, const double* __restrict__ A,
                                                         const double* __restrict__ P, double* __restrict__ R,
                                                         double* __restrict__ C) {
  __shared__ double sP[GJB * GJB];
  for (int q = threadIdx.x; q < GJB * GJB; q += 256) sP[q] = P[q];
  __syncthreads();
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  double col[GJB];
#pragma unroll 8
  for (int t = 0; t < bs; ++t) col[t] = A[(long long)(k0 + t) * n + j];
  const bool inblk = j >= k0 && j < k0 + bs;
  for (int t = 0; t < bs; ++t) {
    double acc = 0.0;
    if (inblk) acc = sP[t * GJB + (j - k0)];
    else
      for (int q = 0; q < bs; ++q) acc = fma(sP[t * GJB + q], col[q], acc);
    R[(long long)t * n + j] = acc;
  }
  // old pivot column panel (thread j doubles as row index i = j)
  for (int t = 0; t < bs; ++t) C[(long long)j * GJB + t] = A[(long long)j * n + (k0 + t)];
}

__global__ void __launch_bounds__(256) gj_rank_update_kernel(int n, int k0, int bs, double* __restrict__ A,
                                                              const double* __restrict__ R, const double* __restrict__ C) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= n) return;
  if (i >= k0 && i < k0 + bs) {
    A[(long long)i * n + j] = R[(long long)(i - k0) * n + j];
    return;
  }
  const bool inblk = j >= k0 && j < k0 + bs;
  double acc = inblk ? 0.0 : A[(long long)i * n + j];
  const double* ci = C + (long long)i * GJB;
  for (int t = 0; t < bs; ++t) acc = fma(-ci[t], R[(long long)t * n + j], acc);
  A[(long long)i * n + j] = acc;
}

extern "C" long long pmb_dense_invert_ws_doubles(int n) { return (long long)GJB * GJB + 2LL * GJB * (n > 0 ? n : 0); }

// The same elimination as ONE cooperative kernel (one CTA per SM, grid-wide barriers instead of 3 launches per 32 pivots:
// the 66 dependent launches of a 675 x 675 inverse cost 1.2 ms, a quarter of a whole design iteration at 64x32x32).
//   phase 1  every CTA inverts the 32 x 32 pivot block itself (no barrier needed to publish it): ONE warp, lane = column,
//            the column in registers, pivot column broadcast by shuffles -- ~2 us instead of ~10 us for the 1024-thread
//            shared-memory version with two block barriers per pivot;
//   phase 2  R = P A_k,: and the old pivot column panel C, one (t, j) / (i, t) item per thread;       grid barrier
//   phase 3  rank-32 update of the whole matrix, one (i, j) item per thread;                          grid barrier
__global__ void __launch_bounds__(256) gj_coop_kernel(int n, double* __restrict__ A, double* __restrict__ R, double* __restrict__ C,
                                                      int* __restrict__ info) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  __shared__ double sP[GJB][GJB + 1];
  __shared__ double sC[64][GJB + 1], sR[GJB][64];
  const int tid = threadIdx.x, lane = tid & 31;
  const long long gtid = (long long)blockIdx.x * blockDim.x + tid, gsize = (long long)gridDim.x * blockDim.x;
  for (int k0 = 0; k0 < n; k0 += GJB) {
    const int bs = n - k0 < GJB ? n - k0 : GJB;
    // ---- phase 1 (warp 0 of every CTA): lane tx holds column tx of the pivot block (identity beyond bs)
    if (tid < 32) {
      double m[GJB];
#pragma unroll
      for (int ty = 0; ty < GJB; ++ty)
        m[ty] = (lane < bs && ty < bs) ? A[(long long)(k0 + ty) * n + (k0 + lane)] : (lane == ty ? 1.0 : 0.0);
#pragma unroll
      for (int k = 0; k < GJB; ++k) {
        const double piv = __shfl_sync(0xffffffffu, m[k], k);
        if (blockIdx.x == 0 && lane == 0 && k < bs && !(piv > 0.0) && *info == 0) *info = k0 + k + 1;
        const double rk = (lane == k ? 1.0 : m[k]) / piv;   // pivot row, my column
#pragma unroll
        for (int ty = 0; ty < GJB; ++ty) {
          const double ck = __shfl_sync(0xffffffffu, m[ty], k);   // pivot column, row ty
          const double cur = (lane == k) ? 0.0 : m[ty];
          m[ty] = (ty == k) ? rk : cur - ck * rk;
        }
      }
#pragma unroll
      for (int ty = 0; ty < GJB; ++ty) sP[ty][lane] = m[ty];
    }
    __syncthreads();
    // ---- phase 2: R[t][j] = (P A_k,:)[t][j] outside the block columns, P inside; C[i][t] = A[i][k0 + t]
    for (long long it = gtid; it < (long long)bs * n; it += gsize) {
      const int t = (int)(it / n), j = (int)(it - (long long)t * n);
      double acc = 0.0;
      if (j >= k0 && j < k0 + bs) acc = sP[t][j - k0];
      else
        for (int q = 0; q < bs; ++q) acc = fma(sP[t][q], A[(long long)(k0 + q) * n + j], acc);
      R[(long long)t * n + j] = acc;
    }
    for (long long it = gtid; it < (long long)bs * n; it += gsize) {
      const int i = (int)(it / bs), t = (int)(it - (long long)i * bs);
      C[(long long)i * GJB + t] = A[(long long)i * n + (k0 + t)];
    }
    grid.sync();
    // ---- phase 3: rows of the block <- R, every other row i: A_i,: <- (A_i,: with the block columns zeroed) - C_i R,
    //      as 64 x 64 tiles with the two panels staged in shared memory (each thread a 4 x 4 sub-tile)
    {
      const int ntile = (n + 63) / 64;
      const int tx = tid & 15, ty = tid >> 4;
      for (int tile = blockIdx.x; tile < ntile * ntile; tile += gridDim.x) {
        const int i0 = (tile / ntile) * 64, j0 = (tile % ntile) * 64;
        __syncthreads();  // the previous tile's panels are no longer read
        for (int q = tid; q < 64 * GJB; q += 256) {
          const int r = q / GJB, t = q - r * GJB;   // C panel: rows i0 .. i0+63
          sC[r][t] = (i0 + r < n && t < bs) ? C[(long long)(i0 + r) * GJB + t] : 0.0;
          const int t2 = q / 64, c = q - t2 * 64;   // R panel: columns j0 .. j0+63
          sR[t2][c] = (j0 + c < n && t2 < bs) ? R[(long long)t2 * n + j0 + c] : 0.0;
        }
        __syncthreads();
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b2 = 0; b2 < 4; ++b2) {
            const int i = i0 + 4 * ty + a, j = j0 + 4 * tx + b2;
            acc[a][b2] = (i < n && j < n && !(j >= k0 && j < k0 + bs)) ? A[(long long)i * n + j] : 0.0;
          }
#pragma unroll 8
        for (int t = 0; t < GJB; ++t) {
          double cv[4], rv[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) cv[a] = sC[4 * ty + a][t], rv[a] = sR[t][4 * tx + a];
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b2 = 0; b2 < 4; ++b2) acc[a][b2] = fma(-cv[a], rv[b2], acc[a][b2]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b2 = 0; b2 < 4; ++b2) {
            const int i = i0 + 4 * ty + a, j = j0 + 4 * tx + b2;
            if (i < n && j < n) A[(long long)i * n + j] = (i >= k0 && i < k0 + bs) ? sR[i - k0][4 * tx + b2] : acc[a][b2];
          }
      }
    }
    grid.sync();
  }
}

extern "C" int pmb_dense_invert(int n, double* dense, double* scratch, int* info, void* stream) {
  PMB_REQUIRE(n > 0 && dense && scratch && info, "pmb_dense_invert: invalid argument");
  PMB_REQUIRE(n <= 8192, "pmb_dense_invert: n=%d too large for the dense coarsest-level solve", n);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
  if (e != cudaSuccess) return pmb_set_error("pmb_dense_invert: %s", cudaGetErrorString(e));
  double* P = scratch;
  double* R = scratch + GJB * GJB;
  double* C = R + (long long)GJB * n;
  // one cooperative kernel when the device supports it (PMB_DENSE_COOP=0 forces the three-kernel version)
  static int coop = -1, sms = 0;
  if (coop < 0) {
    int dev = 0, v = 0;
    const char* env = getenv("PMB_DENSE_COOP");
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
      coop = (v != 0 && !(env && env[0] == '0')) ? 1 : 0;
    else
      coop = 0;
  }
  if (coop == 1) {
    void* args[] = {&n, &dense, &R, &C, &info};
    e = cudaLaunchCooperativeKernel((const void*)gj_coop_kernel, dim3(sms), dim3(256), args, 0, st);
    if (e == cudaSuccess) return 0;
    if (getenv("PMB_DEBUG")) fprintf(stderr, "pmb_dense_invert: cooperative launch refused (%s), using the three-kernel version\n", cudaGetErrorString(e));
    cudaGetLastError();  // e.g. inside a stream capture that does not take cooperative launches: fall through
  }
  dim3 grid((n + 255) / 256, n);
  for (int k0 = 0; k0 < n; k0 += GJB) {
    const int bs = n - k0 < GJB ? n - k0 : GJB;
    gj_pivot_block_kernel<<<1, GJB * GJB, 0, st>>>(n, k0, bs, dense, P, info);
    gj_panels_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, k0, bs, dense, P, R, C);
    gj_rank_update_kernel<<<grid, 256, 0, st>>>(n, k0, bs, dense, R, C);
  }
  PMB_CHECK_LAUNCH("pmb_dense_invert");
  return 0;
}

// y = M x, one warp per row
__global__ void __launch_bounds__(256) dense_gemv_kernel(int n, const double* __restrict__ M, const double* __restrict__ x,
                                                          double* __restrict__ y) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* mp = M + (long long)row * n;
  double acc = 0.0;
  for (int c = lane; c < n; c += 32) acc = fma(mp[c], __ldg(x + c), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[row] = acc;
}

extern "C" int pmb_dense_gemv(int n, const double* M, const double* x, double* y, void* stream) {
  PMB_REQUIRE(n > 0 && M && x && y, "pmb_dense_gemv: invalid argument");
  long long threads = (long long)n * 32;
  dense_gemv_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, M, x, y);
  PMB_CHECK_LAUNCH("pmb_dense_gemv");
  return 0;
}
