// libpmb: CSR pattern (K0), SIMP-scaled assembly (K1), element sensitivity (K11).
//
// Replaces pymoto/modules/assembly.py:130-206 (pattern), :255-275 (np.add.at scatter + Dirichlet handling) and
// :298-315 -> pymoto/common/dyadcarrier.py:408-412 (einsum "Ai,ij,Aj->A").
// The scatter of the reference is turned into a gather: one thread owns one (row, neighbour-node) slot and
// sums the <= 8 (4 in 2-D) element contributions in ascending element number with a separate multiply and add
// (no FMA), which is the order and rounding of the sequential np.add.at loop -> values are bit-identical.
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
// One thread per (node, neighbour slot): the NDOF x NDOF block A[n, c] (NDOF runs of NDOF consecutive CSR entries).
template <int NDOF>
__global__ void __launch_bounds__(256) assemble_kernel(Geo g, const double* __restrict__ Ke, const double* __restrict__ x,
                                                        const unsigned char* __restrict__ bcmask, double bcdiagval,
                                                        double* __restrict__ data) {
  __shared__ double sKe[8 * NDOF * 8 * NDOF];
  const int nn = g.dim3 ? 8 : 4;
  const int ke_ld = nn * NDOF;
  for (int q = threadIdx.x; q < ke_ld * ke_ld; q += blockDim.x) sKe[q] = Ke[q];
  __syncthreads();

  // 3-D launch: blockIdx.y = j, blockIdx.z = owned plane, blockIdx.x * 256 + threadIdx.x = i * 27 + slot.  All index
  // arithmetic below is 32-bit without runtime divisions (the kernel is instruction-bound, not bandwidth-bound, otherwise)
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= g.NX * 27) return;
  const int i = q / 27, s = q - i * 27;
  const int j = blockIdx.y, k = g.kz0 + blockIdx.z;
  const long long ln = ((long long)blockIdx.z * g.NY + j) * g.NX + i;
  int dk = s / 9 - 1, dj = (s / 3) % 3 - 1, di = s % 3 - 1;
  int ci = i + di, cj = j + dj, ck = k + dk;
  if (ci < 0 || ci >= g.NX || cj < 0 || cj >= g.NY || ck < 0 || ck >= g.NZ) return;
  int cx = cnt1(i, g.NX), cy = cnt1(j, g.NY), cz = cnt1(k, g.NZ);
  int ilo = max(i - 1, 0), jlo = max(j - 1, 0), klo = max(k - 1, 0);
  long long L = (long long)cx * cy * cz * NDOF;
  int nbr = ((ck - klo) * cy + (cj - jlo)) * cx + (ci - ilo);
  long long off = (long long)(NDOF * NDOF) * (block_offset(g, i, j, k) - g.bo0) + (long long)nbr * NDOF;

  // local (slab-relative) node numbers for the bc mask: row node ln, column node lc (may lie in a halo plane)
  long long lc = ((long long)(ck - g.kz0) * g.NY + cj) * g.NX + ci;

  double acc[NDOF][NDOF];
#pragma unroll
  for (int d = 0; d < NDOF; ++d)
#pragma unroll
    for (int cd = 0; cd < NDOF; ++cd) acc[d][cd] = 0.0;

  const int nzo = g.dim3 ? 2 : 1;
  for (int oz = 0; oz < nzo; ++oz) {
    int ek = g.dim3 ? (k - 1 + oz) : 0;
    int az = g.dim3 ? (1 - oz) : 0;
    int bz = az + dk;
    if (ek < 0 || ek >= g.nzE || bz < 0 || bz > (g.dim3 ? 1 : 0)) continue;
    for (int oy = 0; oy < 2; ++oy) {
      int ej = j - 1 + oy, ay = 1 - oy, by = ay + dj;
      if (ej < 0 || ej >= g.ny || by < 0 || by > 1) continue;
      for (int ox = 0; ox < 2; ++ox) {
        int ei = i - 1 + ox, ax = 1 - ox, bx = ax + di;
        if (ei < 0 || ei >= g.nx || bx < 0 || bx > 1) continue;
        // element layer index relative to the slab's first owned layer (layer kz0); kz0-1 is the halo
        long long e = ((long long)(ek - g.kz0) * g.ny + ej) * g.nx + ei;
        double xe = __ldg(x + e);
        int a = ax + 2 * ay + 4 * az, b = bx + 2 * by + 4 * bz;
        const double* kp = sKe + (a * NDOF) * ke_ld + b * NDOF;
        // ascending element number, separate multiply and add: the order and rounding of np.add.at (assembly.py:267-268)
#pragma unroll
        for (int d = 0; d < NDOF; ++d)
#pragma unroll
          for (int cd = 0; cd < NDOF; ++cd) acc[d][cd] = __dadd_rn(acc[d][cd], __dmul_rn(kp[d * ke_ld + cd], xe));
      }
    }
  }
  bool rowbc[NDOF], colbc[NDOF];
#pragma unroll
  for (int d = 0; d < NDOF; ++d) {
    rowbc[d] = bcmask && bcmask[ln * NDOF + d];
    colbc[d] = bcmask && bcmask[lc * NDOF + d];
  }
#pragma unroll
  for (int d = 0; d < NDOF; ++d)
#pragma unroll
    for (int cd = 0; cd < NDOF; ++cd) {
      double v = acc[d][cd];
      if (rowbc[d] || colbc[cd]) v = (lc == ln && cd == d) ? bcdiagval : 0.0;
      data[off + d * L + cd] = v;
    }
}

extern "C" int pmb_assemble(const pmb_grid* p, const double* Ke, const double* x, const unsigned char* bcmask,
                            double bcdiagval, double* data, void* stream) {
  if (validate_grid(p, "pmb_assemble")) return 1;
  PMB_REQUIRE(Ke && x && data, "pmb_assemble: NULL pointer argument");
  Geo g = make_geo(p);
  PMB_REQUIRE(g.NY <= 65535 && g.nzl <= 65535, "pmb_assemble: grid too large for the 3-D launch");
  dim3 blocks((g.NX * 27 + 255) / 256, g.NY, g.nzl);
  cudaStream_t st = (cudaStream_t)stream;
  switch (g.ndof) {
    case 1: assemble_kernel<1><<<blocks, 256, 0, st>>>(g, Ke, x, bcmask, bcdiagval, data); break;
    case 2: assemble_kernel<2><<<blocks, 256, 0, st>>>(g, Ke, x, bcmask, bcdiagval, data); break;
    case 3: assemble_kernel<3><<<blocks, 256, 0, st>>>(g, Ke, x, bcmask, bcdiagval, data); break;
  }
  PMB_CHECK_LAUNCH("pmb_assemble");
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
