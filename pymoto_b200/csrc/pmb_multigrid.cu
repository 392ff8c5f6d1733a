// libpmb: geometric-multigrid transfer operators (K4, K5), Galerkin coarse operator (K6), coarsest-level dense
// inverse (K7).
//
// Replaces pymoto/solvers/iterative.py:178-220 (the prolongation matrix R is never formed: its trilinear
// weights 1, 1/2, 1/4, 1/8 are closed-form), :244 (R^T r via csc_matvec), :250 (u += R u_c via csr_matvec),
// :173 (R^T A R via two csr_matmat) and the coarsest-level splu of pymoto/solvers/sparse.py:533-550.
#include "pmb_common.cuh"

// ------------------------------------------------------------------------------------------------- K4
// rc[C] = sum over the <= 27 fine nodes 2C+d of w(d) rf[fine]; ascending fine node number, separate multiply and
// add: the order and rounding of scipy's csc_matvec for R^T (bit-identical to the reference).
__global__ void __launch_bounds__(256) restrict_kernel(Geo gf, Geo gc, const double* __restrict__ rf, double* __restrict__ rc) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long n = gc.nOwned * gc.ndof;
  if (t >= n) return;
  long long lc = t / gc.ndof;
  int d = (int)(t - lc * gc.ndof);
  int I, J, K;
  node_ijk(gc, lc, I, J, K);
  double acc = 0.0;
  for (int dk = -1; dk <= 1; ++dk) {
    int fk = 2 * K + dk;
    if (fk < 0 || fk >= gf.NZ) continue;
    for (int dj = -1; dj <= 1; ++dj) {
      int fj = 2 * J + dj;
      if (fj < 0 || fj >= gf.NY) continue;
      for (int di = -1; di <= 1; ++di) {
        int fi = 2 * I + di;
        if (fi < 0 || fi >= gf.NX) continue;
        double w = (dk ? 0.5 : 1.0) * (dj ? 0.5 : 1.0) * (di ? 0.5 : 1.0);
        long long lf = ((long long)(fk - gf.kz0) * gf.NY + fj) * gf.NX + fi;
        acc = __dadd_rn(acc, __dmul_rn(w, __ldg(rf + lf * gf.ndof + d)));
      }
    }
  }
  rc[t] = acc;
}

extern "C" int pmb_restrict(const pmb_grid* pf, const pmb_grid* pc, const double* rf, double* rc, void* stream) {
  if (validate_grid(pf, "pmb_restrict(fine)") || validate_grid(pc, "pmb_restrict(coarse)")) return 1;
  PMB_REQUIRE(pf->nx == 2 * pc->nx && pf->ny == 2 * pc->ny && pf->nz == 2 * pc->nz && pf->ndof == pc->ndof,
              "pmb_restrict: coarse grid is not the 2:1 coarsening of the fine grid");
  PMB_REQUIRE(rf && rc, "pmb_restrict: NULL pointer argument");
  Geo gf = make_geo(pf), gc = make_geo(pc);
  long long n = gc.nOwned * gc.ndof;
  restrict_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gf, gc, rf, rc);
  PMB_CHECK_LAUNCH("pmb_restrict");
  return 0;
}

// ------------------------------------------------------------------------------------------------- K5
// uf[f] += sum over the <= 8 coarse parents (ascending coarse node number) of w uc[parent]
__global__ void __launch_bounds__(256) prolong_add_kernel(Geo gf, Geo gc, const double* __restrict__ uc, double* __restrict__ uf) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long n = gf.nOwned * gf.ndof;
  if (t >= n) return;
  long long lf = t / gf.ndof;
  int d = (int)(t - lf * gf.ndof);
  int i, j, k;
  node_ijk(gf, lf, i, j, k);
  const int ni = (i & 1) + 1, nj = (j & 1) + 1, nk = (k & 1) + 1;
  const int I0 = i >> 1, J0 = j >> 1, K0 = k >> 1;
  double acc = 0.0;
  for (int a = 0; a < nk; ++a)
    for (int b = 0; b < nj; ++b)
      for (int c = 0; c < ni; ++c) {
        double w = (nk == 2 ? 0.5 : 1.0) * (nj == 2 ? 0.5 : 1.0) * (ni == 2 ? 0.5 : 1.0);
        long long lc = ((long long)(K0 + a - gc.kz0) * gc.NY + (J0 + b)) * gc.NX + (I0 + c);
        acc = __dadd_rn(acc, __dmul_rn(w, __ldg(uc + lc * gc.ndof + d)));
      }
  uf[t] = __dadd_rn(uf[t], acc);
}

extern "C" int pmb_prolong_add(const pmb_grid* pf, const pmb_grid* pc, const double* uc, double* uf, void* stream) {
  if (validate_grid(pf, "pmb_prolong_add(fine)") || validate_grid(pc, "pmb_prolong_add(coarse)")) return 1;
  PMB_REQUIRE(pf->nx == 2 * pc->nx && pf->ny == 2 * pc->ny && pf->nz == 2 * pc->nz && pf->ndof == pc->ndof,
              "pmb_prolong_add: coarse grid is not the 2:1 coarsening of the fine grid");
  PMB_REQUIRE(uc && uf, "pmb_prolong_add: NULL pointer argument");
  Geo gf = make_geo(pf), gc = make_geo(pc);
  long long n = gf.nOwned * gf.ndof;
  prolong_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gf, gc, uc, uf);
  PMB_CHECK_LAUNCH("pmb_prolong_add");
  return 0;
}

// ------------------------------------------------------------------------------------------------- K6
// Ac[(C,dI),(C+D,dJ)] = sum_{i in supp(C), j in supp(C+D), j neighbour of i} w_C(i) w_{C+D}(j) Af[(i,dI),(j,dJ)]
// One thread per coarse (node, neighbour slot) = one NDOF x NDOF block, gathered straight from the fine
// stencil-CSR values; the result lands in the coarse grid's own stencil-CSR layout.
template <int NDOF>
__global__ void __launch_bounds__(128) galerkin_kernel(Geo gf, Geo gc, const double* __restrict__ Af, double* __restrict__ Ac) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long nslots = gc.nOwned * 27;
  if (t >= nslots) return;
  int s = (int)(t % 27);
  long long lc = t / 27;
  int I, J, K;
  node_ijk(gc, lc, I, J, K);
  const int Dk = s / 9 - 1, Dj = (s / 3) % 3 - 1, Di = s % 3 - 1;
  const int Ci = I + Di, Cj = J + Dj, Ck = K + Dk;
  if (Ci < 0 || Ci >= gc.NX || Cj < 0 || Cj >= gc.NY || Ck < 0 || Ck >= gc.NZ) return;

  double acc[NDOF][NDOF];
#pragma unroll
  for (int a = 0; a < NDOF; ++a)
#pragma unroll
    for (int b = 0; b < NDOF; ++b) acc[a][b] = 0.0;

  for (int dz = -1; dz <= 1; ++dz) {
    const int fk = 2 * K + dz;
    if (fk < 0 || fk >= gf.NZ) continue;
    const int czf = cnt1(fk, gf.NZ), klo = max(fk - 1, 0);
    for (int ez = -1; ez <= 1; ++ez) {
      const int pk = fk + ez, tz = pk - 2 * Ck;
      if (pk < 0 || pk >= gf.NZ || tz < -1 || tz > 1) continue;
      const double wz = (dz ? 0.5 : 1.0) * (tz ? 0.5 : 1.0);
      for (int dy = -1; dy <= 1; ++dy) {
        const int fj = 2 * J + dy;
        if (fj < 0 || fj >= gf.NY) continue;
        const int cyf = cnt1(fj, gf.NY), jlo = max(fj - 1, 0);
        for (int ey = -1; ey <= 1; ++ey) {
          const int pj = fj + ey, ty = pj - 2 * Cj;
          if (pj < 0 || pj >= gf.NY || ty < -1 || ty > 1) continue;
          const double wzy = wz * (dy ? 0.5 : 1.0) * (ty ? 0.5 : 1.0);
          for (int dx = -1; dx <= 1; ++dx) {
            const int fi = 2 * I + dx;
            if (fi < 0 || fi >= gf.NX) continue;
            const int cxf = cnt1(fi, gf.NX), ilo = max(fi - 1, 0);
            const long long L = (long long)cxf * cyf * czf * NDOF;
            const long long rowbase = (long long)(NDOF * NDOF) * (block_offset(gf, fi, fj, fk) - gf.bo0);
            for (int ex = -1; ex <= 1; ++ex) {
              const int pi = fi + ex, tx = pi - 2 * Ci;
              if (pi < 0 || pi >= gf.NX || tx < -1 || tx > 1) continue;
              const double w = wzy * (dx ? 0.5 : 1.0) * (tx ? 0.5 : 1.0);
              const int nbr = ((pk - klo) * cyf + (pj - jlo)) * cxf + (pi - ilo);
              const double* ap = Af + rowbase + (long long)nbr * NDOF;
#pragma unroll
              for (int a = 0; a < NDOF; ++a)
#pragma unroll
                for (int b = 0; b < NDOF; ++b) acc[a][b] = fma(w, __ldg(ap + a * L + b), acc[a][b]);
            }
          }
        }
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
    for (int b = 0; b < NDOF; ++b) op[a * Lc + b] = acc[a][b];
}

extern "C" int pmb_galerkin(const pmb_grid* pf, const pmb_grid* pc, const double* Af, double* Ac, void* stream) {
  if (validate_grid(pf, "pmb_galerkin(fine)") || validate_grid(pc, "pmb_galerkin(coarse)")) return 1;
  PMB_REQUIRE(pf->nx == 2 * pc->nx && pf->ny == 2 * pc->ny && pf->nz == 2 * pc->nz && pf->ndof == pc->ndof,
              "pmb_galerkin: coarse grid is not the 2:1 coarsening of the fine grid");
  PMB_REQUIRE(Af && Ac, "pmb_galerkin: NULL pointer argument");
  Geo gf = make_geo(pf), gc = make_geo(pc);
  long long nslots = gc.nOwned * 27;
  unsigned blocks = (unsigned)((nslots + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  switch (gc.ndof) {
    case 1: galerkin_kernel<1><<<blocks, 128, 0, st>>>(gf, gc, Af, Ac); break;
    case 2: galerkin_kernel<2><<<blocks, 128, 0, st>>>(gf, gc, Af, Ac); break;
    case 3: galerkin_kernel<3><<<blocks, 128, 0, st>>>(gf, gc, Af, Ac); break;
  }
  PMB_CHECK_LAUNCH("pmb_galerkin");
  return 0;
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

// In-place Gauss-Jordan inverse without pivoting (SPD input => positive pivots), one elimination step per
// launch pair so the whole GPU works on every step (the matrix, <= ~30 MB, stays in L2).
__global__ void gj_copy_kernel(int n, int k, const double* __restrict__ A, double* __restrict__ rowk, double* __restrict__ colk,
                               int* __restrict__ info) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    rowk[t] = A[(long long)k * n + t];
    colk[t] = A[(long long)t * n + k];
  }
  if (t == 0) {
    double p = A[(long long)k * n + k];
    if (!(p > 0.0) && *info == 0) *info = k + 1;
  }
}

__global__ void __launch_bounds__(256) gj_update_kernel(int n, int k, double* __restrict__ A, const double* __restrict__ rowk,
                                                         const double* __restrict__ colk) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= n) return;
  const double p = rowk[k];
  const double rk = (j == k ? 1.0 : rowk[j]) / p;
  if (i == k) {
    A[(long long)i * n + j] = rk;
  } else {
    const double aij = (j == k) ? 0.0 : A[(long long)i * n + j];
    A[(long long)i * n + j] = aij - colk[i] * rk;
  }
}

extern "C" int pmb_dense_invert(int n, double* dense, double* scratch, int* info, void* stream) {
  PMB_REQUIRE(n > 0 && dense && scratch && info, "pmb_dense_invert: invalid argument");
  PMB_REQUIRE(n <= 8192, "pmb_dense_invert: n=%d too large for the dense coarsest-level solve", n);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
  if (e != cudaSuccess) return pmb_set_error("pmb_dense_invert: %s", cudaGetErrorString(e));
  dim3 grid((n + 255) / 256, n);
  for (int k = 0; k < n; ++k) {
    gj_copy_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, k, dense, scratch, scratch + n, info);
    gj_update_kernel<<<grid, 256, 0, st>>>(n, k, dense, scratch, scratch + n);
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
