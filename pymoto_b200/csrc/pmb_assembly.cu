// libpmb: CSR pattern (K0), SIMP-scaled assembly (K1), element sensitivity (K11).
//
// Replaces pymoto/modules/assembly.py:130-206 (pattern), :255-275 (np.add.at scatter + Dirichlet handling) and
// :298-315 -> pymoto/common/dyadcarrier.py:408-412 (einsum "Ai,ij,Aj->A").
// The scatter of the reference is turned into a gather: one thread owns one (row, neighbour-node) slot and
// sums the <= 8 (4 in 2-D) element contributions in ascending element number with a separate multiply and add
// (no FMA), which is the order and rounding of the sequential np.add.at loop -> values are bit-identical.
#include <cstdlib>
#include "pmb_common.cuh"

// ------------------------------------------------------------------------------------------------- K0
template <typename IDX>
__global__ void __launch_bounds__(256) csr_pattern_kernel(Geo g, IDX* __restrict__ indptr, IDX* __restrict__ indices) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long nslots = g.nOwned * g.ndof * 27;
  if (t >= nslots) return;
  int s = (int)(t % 27);
  long long r = t / 27;  // local row
  long long ln = r / g.ndof;
  int d = (int)(r - ln * g.ndof);
  int i, j, k;
  node_ijk(g, ln, i, j, k);
  int cx = cnt1(i, g.NX), cy = cnt1(j, g.NY), cz = cnt1(k, g.NZ);
  int ilo = max(i - 1, 0), jlo = max(j - 1, 0), klo = max(k - 1, 0);
  long long L = (long long)cx * cy * cz * g.ndof;
  long long rowoff = (long long)(g.ndof * g.ndof) * (block_offset(g, i, j, k) - g.bo0) + d * L;
  if (s == 0) {
    indptr[r] = (IDX)rowoff;
    if (r == g.nOwned * g.ndof - 1) indptr[r + 1] = (IDX)(rowoff + L);
  }
  int dk = s / 9 - 1, dj = (s / 3) % 3 - 1, di = s % 3 - 1;
  int ci = i + di, cj = j + dj, ck = k + dk;
  if (ci < 0 || ci >= g.NX || cj < 0 || cj >= g.NY || ck < 0 || ck >= g.NZ) return;
  int nbr = ((ck - klo) * cy + (cj - jlo)) * cx + (ci - ilo);
  long long c = ((long long)ck * g.NY + cj) * g.NX + ci;  // global column node
  for (int cd = 0; cd < g.ndof; ++cd) indices[rowoff + (long long)nbr * g.ndof + cd] = (IDX)(c * g.ndof + cd);
}

extern "C" int pmb_csr_pattern(const pmb_grid* p, void* indptr, void* indices, int index_bits, void* stream) {
  if (validate_grid(p, "pmb_csr_pattern")) return 1;
  PMB_REQUIRE(indptr && indices, "pmb_csr_pattern: NULL output");
  PMB_REQUIRE(index_bits == 32 || index_bits == 64, "pmb_csr_pattern: index_bits must be 32 or 64");
  Geo g = make_geo(p);
  long long nnz = pmb_nnz(p);
  PMB_REQUIRE(index_bits == 64 || nnz < 2147483647LL, "pmb_csr_pattern: nnz=%lld needs 64-bit indices", nnz);
  long long nslots = g.nOwned * g.ndof * 27;
  unsigned blocks = (unsigned)((nslots + 255) / 256);
  if (index_bits == 32)
    csr_pattern_kernel<int><<<blocks, 256, 0, (cudaStream_t)stream>>>(g, (int*)indptr, (int*)indices);
  else
    csr_pattern_kernel<long long><<<blocks, 256, 0, (cudaStream_t)stream>>>(g, (long long*)indptr, (long long*)indices);
  PMB_CHECK_LAUNCH("pmb_csr_pattern");
  return 0;
}

// ------------------------------------------------------------------------------------------------- K1
// A CTA builds the CSR values of T consecutive nodes of one x-row -- ONE contiguous run of the data array -- in shared
// memory and writes it out with fully coalesced stores.  One thread per (node, z-plane of neighbours): all element /
// local-node / dof indices are compile-time after unrolling, so the element matrix is read from the kernel-parameter
// constant bank and the only run-time quantities are the <= 8 element scalings, the Dirichlet mask and the slot
// offsets of boundary nodes.  Every entry is the sum of its element contributions in ascending element number with a
// separate multiply and add (no FMA): the order and rounding of np.add.at (assembly.py:267-268) -> bit-identical.
// (Elements outside the grid enter with scaling +0.0: x + (+-0.0) == x for every x the sum can hold, so the result
// does not change.)
template <int NDOF, bool DIM3>
struct AsmKe {
  static constexpr int NN = DIM3 ? 8 : 4;
  static constexpr int LD = NN * NDOF;
  double v[LD * LD];
};

// the part of one node's block row that couples to the neighbour plane k + DK (DK compile-time)
template <int NDOF, bool DIM3, int DK>
__device__ __forceinline__ void assemble_node_plane(const Geo& g, const AsmKe<NDOF, DIM3>& ke, const double* __restrict__ x,
                                                    const unsigned char* __restrict__ bcmask, double bcdiagval, double* nodep,
                                                    int i, int j, int k, int kl, int cy, int cz, int jlo, int klo) {
  constexpr int LD = AsmKe<NDOF, DIM3>::LD;
  const int ck = k + DK;
  if (ck < 0 || ck >= g.NZ) return;
  const int cx = cnt1(i, g.NX), ilo = max(i - 1, 0);
  const int L = cx * cy * cz * NDOF;
  const long long ln = ((long long)kl * g.NY + j) * g.NX + i;
  // element scalings of the (up to) 8 elements around the node; absent elements count as +0.0
  double xe[2][2][2];
#pragma unroll
  for (int oz = 0; oz < (DIM3 ? 2 : 1); ++oz)
#pragma unroll
    for (int oy = 0; oy < 2; ++oy)
#pragma unroll
      for (int ox = 0; ox < 2; ++ox) {
        const int ei = i - 1 + ox, ej = j - 1 + oy, ek = DIM3 ? k - 1 + oz : 0;
        const bool in = ei >= 0 && ei < g.nx && ej >= 0 && ej < g.ny && ek >= 0 && ek < g.nzE;
        // element layer index relative to the slab's first owned layer (layer kz0); kz0-1 is the halo
        xe[oz][oy][ox] = in ? __ldg(x + ((long long)(ek - (DIM3 ? g.kz0 : 0)) * g.ny + ej) * g.nx + ei) : 0.0;
      }
  bool rowbc[NDOF];
#pragma unroll
  for (int d = 0; d < NDOF; ++d) rowbc[d] = bcmask && bcmask[ln * NDOF + d];

#pragma unroll
  for (int dj = -1; dj <= 1; ++dj) {
    const int cj = j + dj;
#pragma unroll
    for (int di = -1; di <= 1; ++di) {
      const int ci = i + di;
      if (cj < 0 || cj >= g.NY || ci < 0 || ci >= g.NX) continue;
      const int nbr = ((ck - klo) * cy + (cj - jlo)) * cx + (ci - ilo);
      const long long lc = ((long long)(ck - g.kz0) * g.NY + cj) * g.NX + ci;  // may lie in a halo plane
      bool colbc[NDOF];
#pragma unroll
      for (int c = 0; c < NDOF; ++c) colbc[c] = bcmask && bcmask[lc * NDOF + c];
      constexpr bool zself = (DK == 0);
      const bool self = zself && di == 0 && dj == 0;
#pragma unroll
      for (int d = 0; d < NDOF; ++d)
#pragma unroll
        for (int c = 0; c < NDOF; ++c) {
          double acc = 0.0;
          // elements in ascending number (oz, oy, ox); every Ke index below is a compile-time constant
#pragma unroll
          for (int oz = 0; oz < (DIM3 ? 2 : 1); ++oz) {
            const int az = DIM3 ? 1 - oz : 0, bz = az + DK;
            if (bz < 0 || bz > (DIM3 ? 1 : 0)) continue;
#pragma unroll
            for (int oy = 0; oy < 2; ++oy) {
              const int ay = 1 - oy, by = ay + dj;
              if (by < 0 || by > 1) continue;
#pragma unroll
              for (int ox = 0; ox < 2; ++ox) {
                const int ax = 1 - ox, bx = ax + di;
                if (bx < 0 || bx > 1) continue;
                const int a = ax + 2 * ay + 4 * az, bn = bx + 2 * by + 4 * bz;
                acc = __dadd_rn(acc, __dmul_rn(ke.v[(a * NDOF + d) * LD + bn * NDOF + c], xe[oz][oy][ox]));
              }
            }
          }
          double v = acc;
          if (rowbc[d] || colbc[c]) v = (self && c == d) ? bcdiagval : 0.0;
          nodep[d * L + nbr * NDOF + c] = v;
        }
    }
  }
}

template <int NDOF, bool DIM3>
__global__ void __launch_bounds__(DIM3 ? 96 : 32) assemble_kernel(Geo g, const __grid_constant__ AsmKe<NDOF, DIM3> ke,
                                                                  const double* __restrict__ x,
                                                                  const unsigned char* __restrict__ bcmask, double bcdiagval,
                                                                  double* __restrict__ data) {
  constexpr int T = 32, PARTS = DIM3 ? 3 : 1, NT = T * PARTS;
  extern __shared__ double tile[];  // T * NDOF * NDOF * 27 doubles
  const int tid = threadIdx.x;
  const int gI = tid % T, part = tid / T;  // part = neighbour z-plane (dk = part - 1) in 3-D: uniform per warp
  const int i0 = blockIdx.x * T, j = blockIdx.y, kl = blockIdx.z, k = g.kz0 + kl;
  const int ni = min(T, g.NX - i0);
  const int cy = cnt1(j, g.NY), cz = cnt1(k, g.NZ);
  const int jlo = max(j - 1, 0), klo = max(k - 1, 0);
  const long long per = (long long)(NDOF * NDOF) * cy * cz;
  const long long rowbase = pre1(k, g.NZ) * g.Sy * g.Sx + (long long)cz * (pre1(j, g.NY) * g.Sx) - g.bo0;
  const long long e0 = (long long)(NDOF * NDOF) * rowbase + per * pre1(i0, g.NX);
  const int nelem = (int)(per * (pre1(i0 + ni, g.NX) - pre1(i0, g.NX)));
  if (gI < ni) {
    const int i = i0 + gI;
    double* nodep = tile + per * (pre1(i, g.NX) - pre1(i0, g.NX));
if (!DIM3 || part == 1) assemble_node_plane<NDOF, DIM3, 0>(g, ke, x, bcmask, bcdiagval, nodep, i, j, k, kl, cy, cz, jlo, klo);
    else if (part == 0) assemble_node_plane<NDOF, DIM3, -1>(g, ke, x, bcmask, bcdiagval, nodep, i, j, k, kl, cy, cz, jlo, klo);
    else assemble_node_plane<NDOF, DIM3, 1>(g, ke, x, bcmask, bcdiagval, nodep, i, j, k, kl, cy, cz, jlo, klo);
  }
  __syncthreads();
  // coalesced write-out of the contiguous run
  for (int q = tid; q < nelem; q += NT) data[e0 + q] = tile[q];
}

// 3-D layout with one WARP per (32 nodes of an x-row, neighbour line): 9 warps per CTA, each taking the three slots
// (di = -1, 0, 1) of one (dk, dj) line in turn, 3 CTAs per SM so that the compute phase of one CTA overlaps the write-out of
// another.  assemble_kernel above keeps 3 threads per node and 62 KB of shared memory per 96-thread CTA (9 resident warps per
// SM: instruction-issue bound at 0.36 of the HBM write rate); here the line -- hence the element / local-node structure and
// every Ke index -- is warp-uniform and a thread owns one NDOF x NDOF block at a time.  Same arithmetic: per entry the
// element terms in ascending element number with separate multiply and add -> the same bits as np.add.at.
template <int NDOF>
__global__ void __launch_bounds__(9 * 32, 3) assemble_slot_kernel(Geo g, const __grid_constant__ AsmKe<NDOF, true> ke,
                                                                   const double* __restrict__ x,
                                                                   const unsigned char* __restrict__ bcmask, double bcdiagval,
                                                                   double* __restrict__ data, double* __restrict__ diag_out,
                                                                   int* __restrict__ nnz_out) {
  constexpr int T = 32, ND2 = NDOF * NDOF, LD = 8 * NDOF, NT = 9 * 32;
  extern __shared__ double tile[];  // T * ND2 * 27 doubles, then the element matrix
  double* sKe = tile + T * ND2 * 27;
  const int tid = threadIdx.x, gI = tid & 31, line = tid >> 5;
  const int i0 = blockIdx.x * T, j = blockIdx.y, kl = blockIdx.z, k = g.kz0 + kl;
  const int ni = min(T, g.NX - i0);
  for (int q = tid; q < LD * LD; q += NT) sKe[q] = ke.v[q];
  __syncthreads();
  const int cy = cnt1(j, g.NY), cz = cnt1(k, g.NZ);
  const int jlo = max(j - 1, 0), klo = max(k - 1, 0);
  const long long per = (long long)ND2 * cy * cz;
  const long long rowbase = pre1(k, g.NZ) * g.Sy * g.Sx + (long long)cz * (pre1(j, g.NY) * g.Sx) - g.bo0;
  const long long e0 = (long long)ND2 * rowbase + per * pre1(i0, g.NX);
  const int nelem = (int)(per * (pre1(i0 + ni, g.NX) - pre1(i0, g.NX)));
  const int dk = line / 3 - 1, dj = line % 3 - 1;   // warp-uniform
  const int i = i0 + gI, cj = j + dj, ck = k + dk;
  if (gI < ni && cj >= 0 && cj < g.NY && ck >= 0 && ck < g.NZ) {
    const long long ln = ((long long)kl * g.NY + j) * g.NX + i;
    const int cx = cnt1(i, g.NX), ilo = max(i - 1, 0);
    const int L = cx * cy * cz * NDOF;
    bool rowbc[NDOF];
#pragma unroll
    for (int d = 0; d < NDOF; ++d) rowbc[d] = bcmask && bcmask[ln * NDOF + d];
    // densities of the (up to) 4 elements of this line around the node: (i-1+ox, j-1+oy, k-1+oz) with oy / oz fixed by
    // dj / dk when they are non-zero
#pragma unroll
    for (int di = -1; di <= 1; ++di) {
      const int ci = i + di;
      if (ci < 0 || ci >= g.NX) continue;
      double acc[ND2];
#pragma unroll
      for (int q = 0; q < ND2; ++q) acc[q] = 0.0;
      for (int oz = 0; oz < 2; ++oz) {      // elements in ascending number
        const int az = 1 - oz, bz = az + dk, ek = k - 1 + oz;
        if (bz < 0 || bz > 1 || ek < 0 || ek >= g.nzE) continue;
        for (int oy = 0; oy < 2; ++oy) {
          const int ay = 1 - oy, by = ay + dj, ej = j - 1 + oy;
          if (by < 0 || by > 1 || ej < 0 || ej >= g.ny) continue;
#pragma unroll
          for (int ox = 0; ox < 2; ++ox) {
            const int ax = 1 - ox, bx = ax + di, ei = i - 1 + ox;
            if (bx < 0 || bx > 1) continue;       // compile-time
            if (ei < 0 || ei >= g.nx) continue;   // per lane
            const double xe = __ldg(x + ((long long)(ek - g.kz0) * g.ny + ej) * g.nx + ei);
            const double* kp = sKe + ((ax + 2 * ay + 4 * az) * NDOF) * LD + (bx + 2 * by + 4 * bz) * NDOF;
#pragma unroll
            for (int d = 0; d < NDOF; ++d)
#pragma unroll
              for (int c = 0; c < NDOF; ++c) acc[d * NDOF + c] = __dadd_rn(acc[d * NDOF + c], __dmul_rn(kp[d * LD + c], xe));
          }
        }
      }
      const long long lc = ((long long)(ck - g.kz0) * g.NY + cj) * g.NX + ci;   // may lie in a halo plane
      const bool self = dk == 0 && dj == 0 && di == 0;
      const int nbr = ((ck - klo) * cy + (cj - jlo)) * cx + (ci - ilo);
      double* nodep = tile + per * (pre1(i, g.NX) - pre1(i0, g.NX)) + nbr * NDOF;
#pragma unroll
      for (int d = 0; d < NDOF; ++d)
#pragma unroll
        for (int c = 0; c < NDOF; ++c) {
          const bool colbc = bcmask && bcmask[lc * NDOF + c];
          nodep[d * L + c] = (rowbc[d] || colbc) ? ((self && c == d) ? bcdiagval : 0.0) : acc[d * NDOF + c];
        }
    }
  }
  __syncthreads();
  for (int q = tid; q < nelem; q += NT) data[e0 + q] = tile[q];
  // row statistics while the rows are on chip (what pmb_rowstats would re-read 8 bytes per non-zero for): diagonal entry
  // and number of non-zero off-diagonals of every row of the run
  if ((diag_out || nnz_out) && tid < ni * NDOF) {
    const int gi = tid / NDOF, d = tid - gi * NDOF, ii = i0 + gi;
    const int cx = cnt1(ii, g.NX), ilo = max(ii - 1, 0);
    const int L = cx * cy * cz * NDOF;
    const double* row = tile + per * (pre1(ii, g.NX) - pre1(i0, g.NX)) + d * L;
    const int self = (((k - klo) * cy + (j - jlo)) * cx + (ii - ilo)) * NDOF + d;
    int cnt = 0;
    for (int q = 0; q < L; ++q) cnt += (q != self && row[q] != 0.0) ? 1 : 0;
    const long long r = (((long long)kl * g.NY + j) * g.NX + ii) * NDOF + d;
    if (diag_out) diag_out[r] = row[self];
    if (nnz_out) nnz_out[r] = cnt;
  }
}

template <int NDOF, bool DIM3>
static int launch_assemble(const Geo& g, const double* Ke_host, const double* x, const unsigned char* bcmask, double bcdiagval,
                           double* data, double* diag, int* nnz_offdiag, bool* stats_done, cudaStream_t st) {
  AsmKe<NDOF, DIM3> ke;
  memcpy(ke.v, Ke_host, sizeof(ke.v));
  if constexpr (DIM3) {
    constexpr int T = 32, LD = 8 * NDOF;
    const size_t smem = sizeof(double) * (T * NDOF * NDOF * 27 + LD * LD);
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(assemble_slot_kernel<NDOF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return pmb_set_error("assemble_slot_kernel attribute: %s", cudaGetErrorString(e));
      configured = true;
    }
    static const bool legacy = getenv("PMB_ASSEMBLE_LEGACY") != nullptr;  // the 3-threads-per-node kernel, for A/B timing
    if (!legacy) {
      dim3 blocks((g.NX + T - 1) / T, g.NY, g.nzl);
      assemble_slot_kernel<NDOF><<<blocks, 9 * 32, smem, st>>>(g, ke, x, bcmask, bcdiagval, data, diag, nnz_offdiag);
      PMB_CHECK_LAUNCH("pmb_assemble");
      *stats_done = true;
      return 0;
    }
  }
  constexpr int T = 32, NT = DIM3 ? 96 : 32;
  const size_t smem = sizeof(double) * T * NDOF * NDOF * 27;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(assemble_kernel<NDOF, DIM3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return pmb_set_error("assemble_kernel attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  dim3 blocks((g.NX + T - 1) / T, g.NY, g.nzl);
  assemble_kernel<NDOF, DIM3><<<blocks, NT, smem, st>>>(g, ke, x, bcmask, bcdiagval, data);
  PMB_CHECK_LAUNCH("pmb_assemble");
  return 0;
}

extern "C" int pmb_assemble(const pmb_grid* p, const double* Ke_host, const double* x, const unsigned char* bcmask,
                            double bcdiagval, double* data, double* diag, int* nnz_offdiag, void* stream) {
  if (validate_grid(p, "pmb_assemble")) return 1;
  PMB_REQUIRE(Ke_host && x && data, "pmb_assemble: NULL pointer argument");
  Geo g = make_geo(p);
  PMB_REQUIRE(g.NY <= 65535 && g.nzl <= 65535, "pmb_assemble: grid too large for the 3-D launch");
  cudaStream_t st = (cudaStream_t)stream;
  bool stats_done = false;
  int rc = 1;
  if (g.dim3) {
    switch (g.ndof) {
      case 1: rc = launch_assemble<1, true>(g, Ke_host, x, bcmask, bcdiagval, data, diag, nnz_offdiag, &stats_done, st); break;
      case 2: rc = launch_assemble<2, true>(g, Ke_host, x, bcmask, bcdiagval, data, diag, nnz_offdiag, &stats_done, st); break;
      case 3: rc = launch_assemble<3, true>(g, Ke_host, x, bcmask, bcdiagval, data, diag, nnz_offdiag, &stats_done, st); break;
    }
  } else {
    switch (g.ndof) {
      case 1: rc = launch_assemble<1, false>(g, Ke_host, x, bcmask, bcdiagval, data, diag, nnz_offdiag, &stats_done, st); break;
      case 2: rc = launch_assemble<2, false>(g, Ke_host, x, bcmask, bcdiagval, data, diag, nnz_offdiag, &stats_done, st); break;
      case 3: rc = launch_assemble<3, false>(g, Ke_host, x, bcmask, bcdiagval, data, diag, nnz_offdiag, &stats_done, st); break;
    }
  }
  if (rc) return rc;
  // layouts that do not produce the row statistics on the fly: one streaming pass over the values
  if ((diag || nnz_offdiag) && !stats_done) return pmb_rowstats(p, data, diag, nnz_offdiag, stream);
  return 0;
}

// ------------------------------------------------------------------------------------------------- K11
// One thread per element: dx_e = u_e^T Ke v_e with u, v zeroed at masked dofs.
template <int NDOF, int NN>
__global__ void __launch_bounds__(128) assemble_sens_kernel(Geo g, long long nel_owned, const double* __restrict__ Ke,
                                                             const double* __restrict__ u, const double* __restrict__ v,
                                                             const unsigned char* __restrict__ bcmask,
                                                             double* __restrict__ dx, int accumulate) {
  constexpr int M = NN * NDOF;
  __shared__ double sKe[M * M];
  for (int q = threadIdx.x; q < M * M; q += blockDim.x) sKe[q] = Ke[q];
  __syncthreads();
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nel_owned) return;
  long long lay = (long long)g.nx * g.ny;
  int ekl = (int)(e / lay);  // element layer relative to kz0
  int rem = (int)(e - (long long)ekl * lay);
  int ej = rem / g.nx, ei = rem - ej * g.nx;
  double ue[M], ve[M];
#pragma unroll
  for (int a = 0; a < NN; ++a) {
    long long ln = ((long long)(ekl + ((a >> 2) & 1)) * g.NY + (ej + ((a >> 1) & 1))) * g.NX + (ei + (a & 1));
#pragma unroll
    for (int d = 0; d < NDOF; ++d) {
      long long dof = ln * NDOF + d;
      bool m = bcmask && bcmask[dof];
      ue[a * NDOF + d] = m ? 0.0 : __ldg(u + dof);
      ve[a * NDOF + d] = m ? 0.0 : __ldg(v + dof);
    }
  }
  double acc = 0.0;
#pragma unroll 4
  for (int a = 0; a < M; ++a) {
    double tsum = 0.0;
#pragma unroll
    for (int b = 0; b < M; ++b) tsum = fma(sKe[a * M + b], ve[b], tsum);
    acc = fma(ue[a], tsum, acc);
  }
  dx[e] = accumulate ? dx[e] + acc : acc;
}

extern "C" int pmb_assemble_sens(const pmb_grid* p, const double* Ke, const double* u, const double* v,
                                 const unsigned char* bcmask, double* dx, int accumulate, void* stream) {
  if (validate_grid(p, "pmb_assemble_sens")) return 1;
  PMB_REQUIRE(Ke && u && v && dx, "pmb_assemble_sens: NULL pointer argument");
  Geo g = make_geo(p);
  int nlay = g.dim3 ? (min(g.kz0 + g.nzl, p->nz) - g.kz0) : 1;
  long long nel = (long long)g.nx * g.ny * nlay;
  if (nel <= 0) return 0;
  unsigned blocks = (unsigned)((nel + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (g.dim3) {
    switch (g.ndof) {
      case 1: assemble_sens_kernel<1, 8><<<blocks, 128, 0, st>>>(g, nel, Ke, u, v, bcmask, dx, accumulate); break;
      case 2: assemble_sens_kernel<2, 8><<<blocks, 128, 0, st>>>(g, nel, Ke, u, v, bcmask, dx, accumulate); break;
      case 3: assemble_sens_kernel<3, 8><<<blocks, 128, 0, st>>>(g, nel, Ke, u, v, bcmask, dx, accumulate); break;
    }
  } else {
    switch (g.ndof) {
      case 1: assemble_sens_kernel<1, 4><<<blocks, 128, 0, st>>>(g, nel, Ke, u, v, bcmask, dx, accumulate); break;
      case 2: assemble_sens_kernel<2, 4><<<blocks, 128, 0, st>>>(g, nel, Ke, u, v, bcmask, dx, accumulate); break;
      case 3: assemble_sens_kernel<3, 4><<<blocks, 128, 0, st>>>(g, nel, Ke, u, v, bcmask, dx, accumulate); break;
    }
  }
  PMB_CHECK_LAUNCH("pmb_assemble_sens");
  return 0;
}
