// libpmb: matrix-free application of the finest-level operator.
//
// On the finest level the system matrix is a pure SIMP scaling of ONE element matrix,
//     K = P (sum_e s_e Ke) P + bcdiagval (I - P)          (P masks the Dirichlet dofs, pymoto/modules/assembly.py:208-272)
// so y = K x can be evaluated from the element densities s (8 B per element) instead of the assembled values
// (8 B per non-zero, 243 per node): 0.24 GB instead of 8.6 GB of HBM traffic per application at 256x128x128.  The
// kernel is FP64-pipe bound, not HBM bound.  Same modes / epilogues / fused dot products as the stencil-CSR kernel
// (pmb_spmv.cu); the assembled CSR matrix stays the source of truth for export, the diagonal, the Dirichlet
// detection and the Galerkin coarse operators.
//
// One thread per node: the masked x of the node's 27-neighbourhood and the densities of its 8 (4) elements are staged
// in a shared-memory brick, the element matrix lives in the kernel-parameter constant bank so every FMA takes its
// Ke operand straight from c[0x0][...] (compile-time indices after full unrolling).
#include <cstdlib>
#include "pmb_tilestream.cuh"

enum { EMODE_SPMV = PMB_SPMV, EMODE_RESID = PMB_RESIDUAL, EMODE_JACOBI = PMB_JACOBI };

template <int NDOF, bool DIM3>
struct KeParam {
  static constexpr int NN = DIM3 ? 8 : 4;
  static constexpr int LD = NN * NDOF;
  double v[LD * LD];
};

template <int NDOF, bool DIM3, int MODE>
__global__ void __launch_bounds__(256, 3) elem_kernel(Geo g, const __grid_constant__ KeParam<NDOF, DIM3> ke, const double* __restrict__ s,
                                                    const unsigned char* __restrict__ mask, double bcdiag,
                                                    const double* __restrict__ x, const double* __restrict__ b,
                                                    const double* __restrict__ diag, double w, double* __restrict__ y,
                                                    const double* __restrict__ dotv, double* __restrict__ partials) {
  constexpr int BX = 32, BY = DIM3 ? 4 : 8, BZ = DIM3 ? 2 : 1;
  constexpr int TX = BX + 2, TY = BY + 2, TZ = DIM3 ? BZ + 2 : 1;
  constexpr int SZ = DIM3 ? BZ + 1 : 1;
  constexpr int LD = KeParam<NDOF, DIM3>::LD;
  __shared__ double su[TZ][TY][TX * NDOF];
  __shared__ double ss[SZ][BY + 1][BX + 1];
  __shared__ double wred[3][8];

  const int tid = threadIdx.x;
  const int i0 = blockIdx.x * BX, j0 = blockIdx.y * BY, kl0 = blockIdx.z * BZ;  // kl0: plane index relative to kz0

  // ---- stage the masked input vector of the brick + 1-node apron (zero outside the grid).  A (tz, ty) row of the
  //      brick is TX*NDOF contiguous doubles of x: two rows per pass, 128 threads each, no per-element index division
  constexpr int ROWLEN = TX * NDOF;
  static_assert(ROWLEN <= 128, "row of the staged brick must fit 128 threads");
  // (loads are branch-free: out-of-grid slots read a safe in-range address and are zeroed by a select, so all of a
  //  thread's loads are in flight together)
  const long long safe = (((long long)kl0 * g.NY + j0) * g.NX + i0) * NDOF;  // first dof of the brick: always owned
  {
    const int col = tid & 127, half = tid >> 7;
    const int i = i0 - 1 + col / NDOF;
    const bool colin = col < ROWLEN && i >= 0 && i < g.NX;
    double xv[(TZ * TY + 1) / 2];
    unsigned char mk[(TZ * TY + 1) / 2];
    bool inb[(TZ * TY + 1) / 2];
#pragma unroll
    for (int q = 0; q < (TZ * TY + 1) / 2; ++q) {
      const int row = 2 * q + half;
      const int ty = row % TY, tz = row / TY;
      const int j = j0 - 1 + ty, k = DIM3 ? (g.kz0 + kl0 - 1 + tz) : 0;
      // planes beyond the slab's upper halo (k > kz0 + nzl) are never needed and may not be mapped
      inb[q] = colin && row < TZ * TY && j >= 0 && j < g.NY && k >= 0 && k < g.NZ && k <= g.kz0 + g.nzl;
      const long long idx = inb[q] ? (((long long)(k - g.kz0) * g.NY + j) * g.NX + (i0 - 1)) * NDOF + col : safe;
      xv[q] = __ldg(x + idx);
      mk[q] = mask ? __ldg(mask + idx) : (unsigned char)0;
    }
#pragma unroll
    for (int q = 0; q < (TZ * TY + 1) / 2; ++q) {
      const int row = 2 * q + half;
      if (col < ROWLEN && row < TZ * TY) su[row / TY][row % TY][col] = (inb[q] && !mk[q]) ? xv[q] : 0.0;
    }
  }
  // ---- stage the element densities: slot (tz,ty,tx) = element (i0-1+tx, j0-1+ty, k-1+tz)
  {
    constexpr int NS = SZ * (BY + 1) * (BX + 1);
    double sv[(NS + 255) / 256];
    bool sin[(NS + 255) / 256];
#pragma unroll
    for (int q = 0; q < (NS + 255) / 256; ++q) {
      const int p = tid + 256 * q;
      const int tx = p % (BX + 1), ty = (p / (BX + 1)) % (BY + 1), tz = p / ((BX + 1) * (BY + 1));
      const int ei = i0 - 1 + tx, ej = j0 - 1 + ty, ek = DIM3 ? (g.kz0 + kl0 - 1 + tz) : 0;
      // element layers above the last owned node plane belong to the next rank and are not needed
      sin[q] = p < NS && ei >= 0 && ei < g.nx && ej >= 0 && ej < g.ny && ek >= 0 && ek < g.nzE && (!DIM3 || ek < g.kz0 + g.nzl);
      // safe address: an element of the brick's own first node (exists unless the brick starts on the far faces)
      const long long sidx = sin[q] ? ((long long)(ek - (DIM3 ? g.kz0 : 0)) * g.ny + ej) * g.nx + ei : 0;
      sv[q] = __ldg(s + sidx);
    }
#pragma unroll
    for (int q = 0; q < (NS + 255) / 256; ++q) {
      const int p = tid + 256 * q;
      if (p < NS) ss[p / ((BX + 1) * (BY + 1))][(p / (BX + 1)) % (BY + 1)][p % (BX + 1)] = sin[q] ? sv[q] : 0.0;
    }
  }

  const int tx = tid % BX, ty = (tid / BX) % BY, tz = tid / (BX * BY);
  const int i = i0 + tx, j = j0 + ty, kl = kl0 + tz;
  const bool valid = i < g.NX && j < g.NY && kl < g.nzl;
  const long long ln = ((long long)kl * g.NY + j) * g.NX + i;

  // ---- epilogue operands of this thread's rows: issued now, consumed after the FMA phase
  double xr[NDOF], br[NDOF], dr[NDOF], dvr[NDOF];
  bool mr[NDOF];
#pragma unroll
  for (int d = 0; d < NDOF; ++d) {
    const long long r = valid ? ln * NDOF + d : safe;
    mr[d] = mask && __ldg(mask + r);
    xr[d] = __ldg(x + r);
    br[d] = (MODE != EMODE_SPMV) ? __ldg(b + r) : 0.0;
    dr[d] = (MODE == EMODE_JACOBI) ? __ldg(diag + r) : 1.0;
    dvr[d] = (partials && dotv) ? __ldg(dotv + r) : 0.0;
  }
  __syncthreads();

  // ---- t[e][d] = sum over the nodes m of element e of Ke[a(e), b(e,m)] u_m : every neighbour value is read from
  //      shared memory once and used by all elements that contain it (compile-time Ke indices -> constant-bank operands)
  constexpr int NE = DIM3 ? 8 : 4;
  double t[NE][NDOF];
#pragma unroll
  for (int e = 0; e < NE; ++e)
#pragma unroll
    for (int d = 0; d < NDOF; ++d) t[e][d] = 0.0;
  if (valid) {
#pragma unroll
    for (int dk = (DIM3 ? -1 : 0); dk <= (DIM3 ? 1 : 0); ++dk)
#pragma unroll
      for (int dj = -1; dj <= 1; ++dj)
#pragma unroll
        for (int di = -1; di <= 1; ++di) {
          const double* up = &su[DIM3 ? tz + 1 + dk : 0][ty + 1 + dj][(tx + 1 + di) * NDOF];
          double uv[NDOF];
#pragma unroll
          for (int c = 0; c < NDOF; ++c) uv[c] = up[c];
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            const int ox = e & 1, oy = (e >> 1) & 1, oz = (e >> 2) & 1;  // element (i-1+ox, j-1+oy, k-1+oz)
            const int ax = 1 - ox, ay = 1 - oy, az = DIM3 ? 1 - oz : 0;  // local position of this node in it
            const int bx = ax + di, by = ay + dj, bz = az + dk;           // local position of the neighbour in it
            if (bx < 0 || bx > 1 || by < 0 || by > 1 || bz < 0 || bz > (DIM3 ? 1 : 0)) continue;
            const int a = ax + 2 * ay + 4 * az, bn = bx + 2 * by + 4 * bz;
#pragma unroll
            for (int c = 0; c < NDOF; ++c)
#pragma unroll
              for (int d = 0; d < NDOF; ++d) t[e][d] = fma(ke.v[(a * NDOF + d) * LD + bn * NDOF + c], uv[c], t[e][d]);
          }
        }
  }
  double acc[NDOF];
#pragma unroll
  for (int d = 0; d < NDOF; ++d) acc[d] = 0.0;
  if (valid) {
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int ox = e & 1, oy = (e >> 1) & 1, oz = (e >> 2) & 1;
      const double se = ss[DIM3 ? tz + oz : 0][ty + oy][tx + ox];
#pragma unroll
      for (int d = 0; d < NDOF; ++d) acc[d] = fma(se, t[e][d], acc[d]);
    }
  }

  double d0 = 0.0, d1 = 0.0, d2 = 0.0;
  if (valid) {
#pragma unroll
    for (int d = 0; d < NDOF; ++d) {
      const long long r = ln * NDOF + d;
      const double ax = mr[d] ? bcdiag * xr[d] : acc[d];
      double out;
      if (MODE == EMODE_SPMV) out = ax;
      else if (MODE == EMODE_RESID) out = br[d] - ax;
      else out = xr[d] + w * ((br[d] - ax) / dr[d]);
      y[r] = out;
      if (partials) {
        d0 = fma(out, xr[d], d0);
        d1 = fma(xr[d], dvr[d], d1);
        d2 = fma(out, dvr[d], d2);
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
      for (int v = 0; v < 8; ++v) s0 += wred[0][v], s1 += wred[1][v], s2 += wred[2][v];
      const long long bid = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      partials[3 * bid] = s0;
      partials[3 * bid + 1] = s1;
      partials[3 * bid + 2] = s2;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// z-marching layout (3-D).  Source-level stall sampling of elem_kernel (profiles/ncu_elem_variants_r1.txt) puts ~60 % of
// the warp samples and 43 % of the executed instructions in the brick staging prologue (index arithmetic, predicates,
// global-load latency before the barrier), not in the FMA phase.  Here a CTA owns a 32 x BY column of nodes and marches
// over ZM_L consecutive node planes: the masked x planes live in a 4-slot shared-memory ring (3 in use, 1 being filled by
// cp.async), the element densities in a second ring; staging offsets and validity are computed once per CTA and the
// copies of the NEXT plane are in flight during the FMA phase.  Per element the same (dk, dj, di) FMA order as
// elem_kernel, so y is bit-identical.  Measured (B200, 256x128x128): ndof = 3 0.37 ms vs 0.35 ms for the brick kernel
// (the FMA phase dominates there and the brick kernel's higher occupancy wins); ndof = 1 0.52 ms vs 0.70 ms.
// ---------------------------------------------------------------------------------------------------------
constexpr int ZM_BX = 32, ZM_L = 8;

// 8-byte asynchronous global -> shared copy (LDGSTS); `valid == false` zero-fills the destination without reading src
__device__ __forceinline__ void cp_async8(double* dst, const double* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 8 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <int NDOF, int MODE, int BY, int MINB>
__global__ void __launch_bounds__(32 * BY, MINB)
    elem_kernel_zm(Geo g, const __grid_constant__ KeParam<NDOF, true> ke, const double* __restrict__ s,
                   const unsigned char* __restrict__ mask, double bcdiag, const double* __restrict__ x,
                   const double* __restrict__ b, const double* __restrict__ diag, double w, double* __restrict__ y,
                   const double* __restrict__ dotv, double* __restrict__ partials) {
  constexpr int BX = ZM_BX, NT = 32 * BY;
  constexpr int TX = BX + 2, TY = BY + 2;
  constexpr int ROWLEN = TX * NDOF, PLANE = TY * ROWLEN;   // doubles per staged x plane
  constexpr int SLAY = (BY + 1) * (BX + 1);                // densities per staged element layer
  constexpr int NQ = (PLANE + NT - 1) / NT, NQS = (SLAY + NT - 1) / NT;
  constexpr int LD = KeParam<NDOF, true>::LD;
  __shared__ double su[4][PLANE];
  __shared__ double ss[4][SLAY];
  __shared__ double wred[3][NT / 32];

  const int tid = threadIdx.x;
  const int i0 = blockIdx.x * BX, j0 = blockIdx.y * BY;
  const int kA = blockIdx.z * ZM_L;                                  // first owned local plane of this CTA
  const int kB = min(kA + ZM_L, g.nzl);                              // one past its last
  const long long xplane = (long long)g.NX * g.NY * NDOF;            // dofs per node plane
  const long long slayer = (long long)g.nx * g.ny;                   // elements per layer

  // ---- per-thread staging slots: offsets inside a plane / layer and in-grid flags, computed once
  int xoff[NQ], soff[NQS];
  bool xok[NQ], sok[NQS];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int p = tid + NT * q;
    const int row = p / ROWLEN, col = p - row * ROWLEN;
    const int i = i0 - 1 + col / NDOF, j = j0 - 1 + row;
    xok[q] = p < PLANE && i >= 0 && i < g.NX && j >= 0 && j < g.NY;
    xoff[q] = (j * g.NX + (i0 - 1)) * NDOF + col;
  }
#pragma unroll
  for (int q = 0; q < NQS; ++q) {
    const int p = tid + NT * q;
    const int ty = p / (BX + 1), tx = p - ty * (BX + 1);
    const int ei = i0 - 1 + tx, ej = j0 - 1 + ty;
    sok[q] = p < SLAY && ei >= 0 && ei < g.nx && ej >= 0 && ej < g.ny;
    soff[q] = ej * g.nx + ei;
  }
  // plane kl (local, may be -1 or nzl = halo) exists and may be read
  auto plane_ok = [&](int kl) {
    const int k = g.kz0 + kl;
    return k >= 0 && k < g.NZ && kl <= g.nzl;
  };
  auto layer_ok = [&](int el) {  // element layer el (local): layers above the last owned node plane are never needed
    const int ek = g.kz0 + el;
    return ek >= 0 && ek < g.nzE && el < g.nzl;
  };
  // Asynchronous staging: cp.async copies x / s straight into the ring slot (zero-fill outside the grid), so no register
  // holds a staged value across the FMA phase; only the Dirichlet mask bytes of the slots travel in registers and are
  // applied (slot zeroed) once the copies have landed.
  auto issue_plane = [&](int kl, int slot, unsigned char (&mk)[NQ]) {
    const bool pok = plane_ok(kl);
    const double* xp = x + (long long)kl * xplane;
    const unsigned char* mp = mask + (long long)kl * xplane;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int p = tid + NT * q;
      const bool ok = pok && xok[q];
      if (p < PLANE) cp_async8(&su[slot][p], ok ? xp + xoff[q] : x, ok);
      mk[q] = (mask && ok) ? __ldg(mp + xoff[q]) : (unsigned char)0;
    }
  };
  auto mask_plane = [&](int slot, const unsigned char (&mk)[NQ]) {
#pragma unroll
    for (int q = 0; q < NQ; ++q)
      if (mk[q]) su[slot][tid + NT * q] = 0.0;
  };
  auto issue_layer = [&](int el, int slot) {
    const bool lok = layer_ok(el);
    const double* sp = s + (long long)el * slayer;
#pragma unroll
    for (int q = 0; q < NQS; ++q) {
      const int p = tid + NT * q;
      const bool ok = lok && sok[q];
      if (p < SLAY) cp_async8(&ss[slot][p], ok ? sp + soff[q] : s, ok);
    }
  };

  // ---- prime the rings: planes kA-1, kA, kA+1 -> slots 0, 1, 2; layers kA-1, kA -> slots 0, 1 (all copies in flight at once)
  {
    unsigned char m0[NQ], m1[NQ], m2[NQ];
    issue_plane(kA - 1, 0, m0);
    issue_plane(kA, 1, m1);
    issue_plane(kA + 1, 2, m2);
    issue_layer(kA - 1, 0);
    issue_layer(kA, 1);
    cp_async_wait_all();
    mask_plane(0, m0);
    mask_plane(1, m1);
    mask_plane(2, m2);
  }
  __syncthreads();

  const int tx = tid % BX, ty = tid / BX;
  const int i = i0 + tx, j = j0 + ty;
  const bool valid = i < g.NX && j < g.NY;
  const int c0 = (ty + 1) * ROWLEN + (tx + 1) * NDOF;  // this thread's node inside a staged plane
  const int e0 = ty * (BX + 1) + tx;                    // its (ox, oy) = (0, 0) element inside a staged layer
  const long long rrow = valid ? ((long long)j * g.NX + i) * NDOF : 0;  // dof offset of the node inside a plane
  double d0 = 0.0, d1 = 0.0, d2 = 0.0;

  for (int kl = kA, t = 0; kl < kB; ++kl, ++t) {
    const bool more = kl + 1 < kB;
    // ---- issued now, landing during the FMA phase: next step's plane / layer (cp.async into the free ring slots) and
    //      L1 prefetches of this step's epilogue operands (no registers held across the FMA phase)
    unsigned char mk[NQ];
    if (more) {
      issue_plane(kl + 2, (t + 3) & 3, mk);  // slot held plane kl-2: last read in the previous step, which ended with a barrier
      issue_layer(kl + 1, (t + 2) & 3);
    }
    const long long r0 = (long long)kl * xplane + rrow;
    if (valid) {
      if (MODE != EMODE_SPMV) prefetch_l1(b + r0), prefetch_l1(b + r0 + NDOF - 1);
      if (MODE == EMODE_JACOBI) prefetch_l1(diag + r0), prefetch_l1(diag + r0 + NDOF - 1);
      if (mask) prefetch_l1(mask + r0);
    }
    const double* pl[3] = {su[t & 3], su[(t + 1) & 3], su[(t + 2) & 3]};
    const double* sl[2] = {ss[t & 3], ss[(t + 1) & 3]};
    if (valid) {
      // elements below the node plane (oz = 0: neighbour planes dk = -1, 0) then above it (oz = 1: dk = 0, 1): 12 live
      // accumulators instead of 24; per element the (dk, dj, di) order of elem_kernel is kept, so y is bit-identical
      double acc[NDOF];
#pragma unroll
      for (int d = 0; d < NDOF; ++d) acc[d] = 0.0;
#pragma unroll
      for (int oz = 0; oz < 2; ++oz) {
        double tacc[4][NDOF];
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int d = 0; d < NDOF; ++d) tacc[e][d] = 0.0;
        const int az = 1 - oz;
#pragma unroll
        for (int dk = -az; dk <= 1 - az; ++dk)
#pragma unroll
          for (int dj = -1; dj <= 1; ++dj)
#pragma unroll
            for (int di = -1; di <= 1; ++di) {
              const double* up = pl[dk + 1] + c0 + dj * ROWLEN + di * NDOF;
              double uv[NDOF];
#pragma unroll
              for (int c = 0; c < NDOF; ++c) uv[c] = up[c];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int ox = e & 1, oy = (e >> 1) & 1;
                const int ax = 1 - ox, ay = 1 - oy;
                const int bx = ax + di, by = ay + dj, bz = az + dk;
                if (bx < 0 || bx > 1 || by < 0 || by > 1) continue;
                const int a = ax + 2 * ay + 4 * az, bn = bx + 2 * by + 4 * bz;
#pragma unroll
                for (int c = 0; c < NDOF; ++c)
#pragma unroll
                  for (int d = 0; d < NDOF; ++d)
                    tacc[e][d] = fma(ke.v[(a * NDOF + d) * LD + bn * NDOF + c], uv[c], tacc[e][d]);
              }
            }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ox = e & 1, oy = (e >> 1) & 1;
          const double se = sl[oz][e0 + oy * (BX + 1) + ox];
#pragma unroll
          for (int d = 0; d < NDOF; ++d) acc[d] = fma(se, tacc[e][d], acc[d]);
        }
      }
#pragma unroll
      for (int d = 0; d < NDOF; ++d) {
        const long long r = r0 + d;
        // unmasked rows: x is the centre value of the staged plane; Dirichlet rows (rare) re-read it from memory
        const bool mr = mask && __ldg(mask + r);
        const double xr = mr ? __ldg(x + r) : pl[1][c0 + d];
        const double ax = mr ? bcdiag * xr : acc[d];
        double out;
        if (MODE == EMODE_SPMV) out = ax;
        else if (MODE == EMODE_RESID) out = __ldg(b + r) - ax;
        else out = xr + w * ((__ldg(b + r) - ax) / __ldg(diag + r));
        y[r] = out;
        if (partials) {
          const double dvv = dotv ? __ldg(dotv + r) : 0.0;
          d0 = fma(out, xr, d0);
          d1 = fma(xr, dvv, d1);
          d2 = fma(out, dvv, d2);
        }
      }
    }
    if (more) {
      cp_async_wait_all();
      mask_plane((t + 3) & 3, mk);
      __syncthreads();
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
      for (int v = 0; v < NT / 32; ++v) s0 += wred[0][v], s1 += wred[1][v], s2 += wred[2][v];
      const long long bid = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      partials[3 * bid] = s0;
      partials[3 * bid + 1] = s1;
      partials[3 * bid + 2] = s2;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Tensor-core layout (3-D, ndof = 3): the element products Ke u_e as FP64 DMMA (mma.sync m8n8k4).
//
// The layouts above issue one DFMA per multiply-add plus ~0.5 uniform constant loads and shared-memory loads around it:
// they saturate the instruction issue / operand delivery long before the FP64 pipe (45-50 % busy).  DMMA.8x8x4 runs at
// the same 37 TFLOP/s (scripts/probe_dmma.cu) with 8x fewer instructions and keeps Ke in registers: for 8 elements at a
// time a warp forms T = Ke (24 x 24) x U (24 x 8) as 3 x 6 DMMAs, the 18 A fragments (Ke) held in registers (read once
// from a shared-memory copy: lane-indexed reads of the parameter constant bank serialise 32-fold and the compiler
// re-materialises them inside the loop), the B fragments (element displacements) read straight from the staged node
// planes.  A CTA owns a column of 31 x 7 nodes = 32 x 8 elements per layer and marches over the element layers (same
// cp.async plane ring as elem_kernel_zm): per layer the 8 warps write s_e T_e to shared memory, then one thread per node
// gathers the 4 + 4 element contributions (the upper four are carried to the next step in registers) and applies the
// epilogue.  Summation order differs from
// elem_kernel (tensor-core accumulation, per-element grouping), so y agrees to rounding, not bit for bit.
// ---------------------------------------------------------------------------------------------------------
constexpr int MM_EX = 32, MM_EY = 8, MM_NXT = MM_EX - 1, MM_NYT = MM_EY - 1;   // elements / owned nodes of a CTA per layer
constexpr int MM_NT = 256, MM_NEL = MM_EX * MM_EY;                              // 8 warps x 4 batches of 8 elements
constexpr int MM_SX = MM_EX + 1, MM_SY = MM_EY + 1, MM_ROW = MM_SX * 3;         // staged node plane: 33 x 9 nodes
constexpr int MM_PLANE = (MM_SY * MM_ROW + 1) / 2 * 2, MM_TLD = MM_NEL + 8;     // T row stride: 64 B shift between rows
constexpr int MM_SMEM_DOUBLES = 3 * MM_PLANE + 2 * MM_NEL + 24 * MM_TLD;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void __launch_bounds__(MM_NT, 2)
    elem_kernel_mma(Geo g, const __grid_constant__ KeParam<3, true> ke, int zl, const double* __restrict__ s,
                    const unsigned char* __restrict__ mask, double bcdiag, const double* __restrict__ x,
                    const double* __restrict__ b, const double* __restrict__ diag, double w, double* __restrict__ y,
                    const double* __restrict__ dotv, double* __restrict__ partials) {
  constexpr int NDOF = 3, NT = MM_NT;
  constexpr int NQ = (MM_SY * MM_ROW + NT - 1) / NT;
  extern __shared__ __align__(16) double smm[];
  double* su = smm;                         // [3][MM_PLANE] ring of masked x planes
  double* ss = smm + 3 * MM_PLANE;          // [2][MM_NEL]   element densities of the current / next layer
  double* sT = ss + 2 * MM_NEL;             // [24][MM_TLD]  s_e (Ke u_e), row = local dof, column = element
  __shared__ double wred[3][NT / 32];
  __shared__ double sKe[24 * 24];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i0 = blockIdx.x * MM_NXT, j0 = blockIdx.y * MM_NYT;   // first owned node of the column
  for (int p = tid; p < 24 * 24; p += NT) sKe[p] = ke.v[p];
  __syncthreads();
  const int kA = blockIdx.z * zl, kB = min(kA + zl, g.nzl);       // owned local planes [kA, kB)
  const long long xplane = (long long)g.NX * g.NY * NDOF, slayer = (long long)g.nx * g.ny;

  // ---- A fragments: Ke[8 mt + lane/4][4 ks + lane%4], loaded once
  double af[3][6];
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int ks = 0; ks < 6; ++ks) af[mt][ks] = sKe[(8 * mt + (lane >> 2)) * 24 + 4 * ks + (lane & 3)];
  // ---- B fragment offsets inside a staged plane: local dof kk = 4 ks + lane%4 = 3 b + c, node b = (bx, by, bz = ks >= 3)
  int boff[6];
#pragma unroll
  for (int ks = 0; ks < 6; ++ks) {
    const int kk = 4 * ks + (lane & 3), bn = kk / 3, c = kk - 3 * bn;
    boff[ks] = ((bn >> 1) & 1) * MM_ROW + (bn & 1) * NDOF + c;
  }
  // ---- staging slots (computed once): staged node (sx, sy) = global node (i0 - 1 + sx, j0 - 1 + sy)
  int xoff[NQ];
  bool xok[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int p = tid + NT * q;
    const int row = p / MM_ROW, col = p - row * MM_ROW;
    const int i = i0 - 1 + col / NDOF, j = j0 - 1 + row;
    xok[q] = p < MM_SY * MM_ROW && i >= 0 && i < g.NX && j >= 0 && j < g.NY;
    xoff[q] = (j * g.NX + (i0 - 1)) * NDOF + col;
  }
  const int sex = tid % MM_EX, sey = tid / MM_EX;               // this thread stages the density of tile element tid
  const int sei = i0 - 1 + sex, sej = j0 - 1 + sey;
  const bool sokk = sei >= 0 && sei < g.nx && sej >= 0 && sej < g.ny;
  const int soff = sej * g.nx + sei;
  auto plane_ok = [&](int kl) {
    const int k = g.kz0 + kl;
    return k >= 0 && k < g.NZ && kl <= g.nzl;
  };
  auto layer_ok = [&](int el) {
    const int ek = g.kz0 + el;
    return ek >= 0 && ek < g.nzE && el < g.nzl;
  };
  // x plane kl -> ring slot by cp.async (zero-fill outside the grid); the Dirichlet mask bytes are only prefetched to L1
  // here and applied by mask_plane once the copies have landed (no register carries them across the DMMA phase)
  auto issue_plane = [&](int kl, int slot) {
    const bool pok = plane_ok(kl);
    const double* xp = x + (long long)kl * xplane;
    const unsigned char* mp = mask + (long long)kl * xplane;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int p = tid + NT * q;
      const bool ok = pok && xok[q];
      if (p < MM_SY * MM_ROW) cp_async8(su + slot * MM_PLANE + p, ok ? xp + xoff[q] : x, ok);
      if (mask && ok) prefetch_l1(mp + xoff[q]);
    }
  };
  auto mask_plane = [&](int kl, int slot) {
    if (!mask || !plane_ok(kl)) return;
    const unsigned char* mp = mask + (long long)kl * xplane;
#pragma unroll
    for (int q = 0; q < NQ; ++q)
      if (xok[q] && __ldg(mp + xoff[q])) su[slot * MM_PLANE + tid + NT * q] = 0.0;
  };
  auto issue_layer = [&](int el, int slot) {
    const bool ok = layer_ok(el) && sokk;
    cp_async8(ss + slot * MM_NEL + tid, ok ? s + (long long)el * slayer + soff : s, ok);
  };

  // ---- prime: planes kA-1, kA -> ring slots 0, 1; layer kA-1 -> density slot 0
  issue_plane(kA - 1, 0);
  issue_plane(kA, 1);
  issue_layer(kA - 1, 0);
  cp_async_wait_all();
  mask_plane(kA - 1, 0);
  mask_plane(kA, 1);
  __syncthreads();

  // gather role: one thread per owned node of the column
  const int nx = tid % MM_NXT, ny = tid / MM_NXT;
  const int gi = i0 + nx, gj = j0 + ny;
  const bool valid = ny < MM_NYT && gi < g.NX && gj < g.NY;
  const long long rrow = valid ? ((long long)gj * g.NX + gi) * NDOF : 0;
  const int ecol = ny * MM_EX + nx;  // element (ox, oy) = (0, 0) of this node in the tile; (ox, oy) adds ox + oy * MM_EX
  double hi[NDOF] = {0.0, 0.0, 0.0};
  double d0 = 0.0, d1 = 0.0, d2 = 0.0;

  for (int L = kA - 1, t = 0; L < kB; ++L, ++t) {
    const bool more = L + 1 < kB;
    if (more) {  // next layer needs node plane L+2 and densities L+1: in flight during the DMMA phase
      issue_plane(L + 2, (t + 2) % 3);
      issue_layer(L + 1, (t + 1) & 1);
    }
    const long long r0 = (long long)L * xplane + rrow;  // rows of this thread's node in plane L (owned iff L >= kA)
    if (valid && L >= kA) {
      if (MODE != EMODE_SPMV) prefetch_l1(b + r0), prefetch_l1(b + r0 + NDOF - 1);
      if (MODE == EMODE_JACOBI) prefetch_l1(diag + r0), prefetch_l1(diag + r0 + NDOF - 1);
      if (mask) prefetch_l1(mask + r0);
    }
    const double* pl0 = su + (t % 3) * MM_PLANE;        // node plane L   (bottom face of the layer's elements)
    const double* pl1 = su + ((t + 1) % 3) * MM_PLANE;  // node plane L+1 (top face)
    const double* sl = ss + (t & 1) * MM_NEL;

    // ---- DMMA phase: warp w takes the element batches w, w + 8, w + 16, w + 24 of the layer
#pragma unroll 1
    for (int bt = warp; bt < MM_NEL / 8; bt += NT / 32) {
      const int e = bt * 8 + (lane >> 2);            // B-fragment column of this lane
      const int ebase = (e / MM_EX) * MM_ROW + (e % MM_EX) * NDOF;
      double c[3][2];
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) c[mt][0] = c[mt][1] = 0.0;
#pragma unroll
      for (int ks = 0; ks < 6; ++ks) {
        const double bv = (ks < 3 ? pl0 : pl1)[ebase + boff[ks]];
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) dmma884(c[mt][0], c[mt][1], af[mt][ks], bv);
      }
      const int e0 = bt * 8 + 2 * (lane & 3);        // C-fragment columns e0, e0 + 1; rows 8 mt + lane/4
      const double2 sv = *reinterpret_cast<const double2*>(sl + e0);
#pragma unroll
      for (int mt = 0; mt < 3; ++mt)
        *reinterpret_cast<double2*>(sT + (8 * mt + (lane >> 2)) * MM_TLD + e0) = make_double2(sv.x * c[mt][0], sv.y * c[mt][1]);
    }
    __syncthreads();

    // ---- gather: node plane L takes the az = 0 rows of its four elements in layer L plus `hi` from layer L-1;
    //      the az = 1 rows are the contribution of layer L to node plane L+1
    if (valid) {
      double lo[NDOF], up[NDOF];
#pragma unroll
      for (int d = 0; d < NDOF; ++d) lo[d] = hi[d], up[d] = 0.0;
#pragma unroll
      for (int oy = 0; oy < 2; ++oy)
#pragma unroll
        for (int ox = 0; ox < 2; ++ox) {
          const double* tp = sT + ecol + oy * MM_EX + ox;
          const int a0 = (1 - ox) + 2 * (1 - oy);
#pragma unroll
          for (int d = 0; d < NDOF; ++d) {
            lo[d] += tp[(a0 * NDOF + d) * MM_TLD];
            up[d] += tp[((a0 + 4) * NDOF + d) * MM_TLD];
          }
        }
#pragma unroll
      for (int d = 0; d < NDOF; ++d) hi[d] = up[d];
      if (L >= kA) {
        const int c0 = (ny + 1) * MM_ROW + (nx + 1) * NDOF;  // this node inside staged plane L
#pragma unroll
        for (int d = 0; d < NDOF; ++d) {
          const long long r = r0 + d;
          const bool mr = mask && __ldg(mask + r);
          const double xr = mr ? __ldg(x + r) : pl0[c0 + d];
          const double ax = mr ? bcdiag * xr : lo[d];
          double out;
          if (MODE == EMODE_SPMV) out = ax;
          else if (MODE == EMODE_RESID) out = __ldg(b + r) - ax;
          else out = xr + w * ((__ldg(b + r) - ax) / __ldg(diag + r));
          y[r] = out;
          if (partials) {
            const double dvv = dotv ? __ldg(dotv + r) : 0.0;
            d0 = fma(out, xr, d0);
            d1 = fma(xr, dvv, d1);
            d2 = fma(out, dvv, d2);
          }
        }
      }
    }
    if (more) {
      cp_async_wait_all();
      mask_plane(L + 2, (t + 2) % 3);
    }
    __syncthreads();  // T may be overwritten, the next plane / layer are visible
  }
  if (partials) {
    d0 = warp_sum(d0);
    d1 = warp_sum(d1);
    d2 = warp_sum(d2);
    if (lane == 0) wred[0][warp] = d0, wred[1][warp] = d1, wred[2][warp] = d2;
    __syncthreads();
    if (tid == 0) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      for (int v = 0; v < NT / 32; ++v) s0 += wred[0][v], s1 += wred[1][v], s2 += wred[2][v];
      const long long bid = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      partials[3 * bid] = s0;
      partials[3 * bid + 1] = s1;
      partials[3 * bid + 2] = s2;
    }
  }
}

// planes per CTA of the tensor-core layout: few enough CTAs per wave lost to the tail, little redundant layer work
// (every CTA computes one extra element layer)
static int mma_zl(const Geo& g) {
  const long long tiles = (long long)((g.NX + MM_NXT - 1) / MM_NXT) * ((g.NY + MM_NYT - 1) / MM_NYT);
  const long long slots = 2LL * 148;
  int best = 8;
  double best_cost = 1e300;
  for (int zl = 4; zl <= 32; ++zl) {
    const long long ctas = tiles * ((g.nzl + zl - 1) / zl);
    const long long waves = (ctas + slots - 1) / slots;
    const double cost = (double)waves * slots / (double)ctas * (zl + 1.0) / zl;  // tail loss x extra-layer overhead
    if (cost < best_cost - 1e-12) best_cost = cost, best = zl;
  }
  return best < g.nzl ? best : (g.nzl > 0 ? g.nzl : 1);
}

template <bool DIM3>
static dim3 elem_grid(const Geo& g) {
  constexpr int BX = 32, BY = DIM3 ? 4 : 8, BZ = DIM3 ? 2 : 1;
  return dim3((g.NX + BX - 1) / BX, (g.NY + BY - 1) / BY, (g.nzl + BZ - 1) / BZ);
}

// ---- layout selection (3-D, ndof 1 or 3): 0 = one node per thread on a 32x4x2 brick (elem_kernel), 1 / 2 = z-marching
//      32x8 / 32x4 columns with ring-buffered planes (elem_kernel_zm), 3 = FP64 tensor-core layout (elem_kernel_mma,
//      ndof 3 only; other ndof fall back to 0).  Chosen per process by pmb_elem_set_variant() /
//      PMB_ELEM_VARIANT or measured by pmb_elem_autotune(); all layouts produce bit-identical y.
enum { PMB_ELEM_VARIANTS = 4 };
static int g_elem_variant[4] = {-1, -1, -1, -1};  // per dofs-per-node (the FMA phase is 9x heavier for ndof = 3 than for 1)
static int& elem_variant(int ndof) {
  if (g_elem_variant[0] < 0) {
    const char* e = getenv("PMB_ELEM_VARIANT");
    const int v = (e && e[0] >= '0' && e[0] < '0' + PMB_ELEM_VARIANTS) ? e[0] - '0' : 0;
    for (int i = 0; i < 4; ++i) g_elem_variant[i] = v;
  }
  return g_elem_variant[ndof >= 1 && ndof <= 3 ? ndof : 0];
}
extern "C" int pmb_elem_set_variant(int v) {
  PMB_REQUIRE(v >= 0 && v < PMB_ELEM_VARIANTS, "pmb_elem_set_variant: variant %d not in 0..%d", v, PMB_ELEM_VARIANTS - 1);
  for (int i = 0; i < 4; ++i) g_elem_variant[i] = v;
  return 0;
}
extern "C" int pmb_elem_get_variant(int ndof) { return elem_variant(ndof); }
extern "C" int pmb_elem_num_variants(void) { return PMB_ELEM_VARIANTS; }

static dim3 elem_grid_any(const Geo& g, int variant) {
  if (!g.dim3) return elem_grid<false>(g);
  switch (variant) {
    case 1: return dim3((g.NX + ZM_BX - 1) / ZM_BX, (g.NY + 7) / 8, (g.nzl + ZM_L - 1) / ZM_L);
    case 2: return dim3((g.NX + ZM_BX - 1) / ZM_BX, (g.NY + 3) / 4, (g.nzl + ZM_L - 1) / ZM_L);
    case 3:
      if (g.ndof == 3) {
        const int zl = mma_zl(g);
        return dim3((g.NX + MM_NXT - 1) / MM_NXT, (g.NY + MM_NYT - 1) / MM_NYT, (g.nzl + zl - 1) / zl);
      }
  }
  return elem_grid<true>(g);
}

extern "C" long long pmb_elem_ws_doubles(const pmb_grid* p) {
  if (validate_grid(p, "pmb_elem_ws_doubles")) return -1;
  Geo g = make_geo(p);
  long long m = 0;  // large enough for every variant, so the choice may change after workspaces were sized
  for (int v = 0; v < PMB_ELEM_VARIANTS; ++v) {
    dim3 gr = elem_grid_any(g, v);
    const long long c = 3LL * gr.x * gr.y * gr.z;
    m = c > m ? c : m;
  }
  return m;
}

template <int NDOF, bool DIM3, int MODE>
static int launch_elem(const Geo& g, const double* Ke_host, const double* s, const unsigned char* mask, double bcdiag,
                       const double* x, const double* b, const double* diag, double w, double* y, const double* dotv,
                       double* dot_out, double* ws, cudaStream_t st) {
  KeParam<NDOF, DIM3> ke;
  memcpy(ke.v, Ke_host, sizeof(ke.v));
  const int variant = (DIM3 && NDOF != 2) ? elem_variant(NDOF) : 0;
  dim3 grid = elem_grid_any(g, variant);
  double* part = dot_out ? ws : nullptr;
  if constexpr (DIM3 && NDOF != 2) {
    if (variant == 1)
      elem_kernel_zm<NDOF, MODE, 8, 2><<<grid, 256, 0, st>>>(g, ke, s, mask, bcdiag, x, b, diag, w, y, dotv, part);
    else if (variant == 2)
      elem_kernel_zm<NDOF, MODE, 4, 4><<<grid, 128, 0, st>>>(g, ke, s, mask, bcdiag, x, b, diag, w, y, dotv, part);
    else if (variant == 3 && NDOF == 3) {
      if constexpr (NDOF == 3) {
        constexpr size_t smem = sizeof(double) * MM_SMEM_DOUBLES;
        static bool configured = false;
        if (!configured) {
          cudaError_t e = cudaFuncSetAttribute(elem_kernel_mma<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return pmb_set_error("elem_kernel_mma attribute: %s", cudaGetErrorString(e));
          configured = true;
        }
        elem_kernel_mma<MODE><<<grid, MM_NT, smem, st>>>(g, ke, mma_zl(g), s, mask, bcdiag, x, b, diag, w, y, dotv, part);
      }
    }
    else
      elem_kernel<NDOF, DIM3, MODE><<<grid, 256, 0, st>>>(g, ke, s, mask, bcdiag, x, b, diag, w, y, dotv, part);
  } else {
    elem_kernel<NDOF, DIM3, MODE><<<grid, 256, 0, st>>>(g, ke, s, mask, bcdiag, x, b, diag, w, y, dotv, part);
  }
  PMB_CHECK_LAUNCH("pmb_elem_spmv");
  if (dot_out) {
    reduce_triples_kernel<<<1, 1024, 0, st>>>(ws, (long long)grid.x * grid.y * grid.z, dot_out);
    PMB_CHECK_LAUNCH("pmb_elem_spmv(reduce)");
  }
  return 0;
}

template <int NDOF, bool DIM3>
static int dispatch_elem(int mode, const Geo& g, const double* Ke_host, const double* s, const unsigned char* mask,
                         double bcdiag, const double* x, const double* b, const double* diag, double w, double* y,
                         const double* dotv, double* dot_out, double* ws, cudaStream_t st) {
  switch (mode) {
    case EMODE_SPMV: return launch_elem<NDOF, DIM3, EMODE_SPMV>(g, Ke_host, s, mask, bcdiag, x, b, diag, w, y, dotv, dot_out, ws, st);
    case EMODE_RESID: return launch_elem<NDOF, DIM3, EMODE_RESID>(g, Ke_host, s, mask, bcdiag, x, b, diag, w, y, dotv, dot_out, ws, st);
    case EMODE_JACOBI: return launch_elem<NDOF, DIM3, EMODE_JACOBI>(g, Ke_host, s, mask, bcdiag, x, b, diag, w, y, dotv, dot_out, ws, st);
  }
  return pmb_set_error("pmb_elem_spmv: unknown mode %d", mode);
}

extern "C" int pmb_elem_spmv(const pmb_grid* p, int mode, const double* Ke_host, const double* s, const unsigned char* bcmask,
                             double bcdiagval, const double* x, const double* b, const double* diag, double w, double* y,
                             const double* dotv, double* dot_out, double* ws, void* stream) {
  if (validate_grid(p, "pmb_elem_spmv")) return 1;
  PMB_REQUIRE(Ke_host && s && x && y, "pmb_elem_spmv: NULL pointer argument");
  PMB_REQUIRE(x != y, "pmb_elem_spmv: y must not alias x");
  PMB_REQUIRE(mode == EMODE_SPMV || b, "pmb_elem_spmv: b required for residual / Jacobi");
  PMB_REQUIRE(mode != EMODE_JACOBI || diag, "pmb_elem_spmv: diag required for Jacobi");
  PMB_REQUIRE(!dot_out || ws, "pmb_elem_spmv: workspace required for the fused dot products");
  Geo g = make_geo(p);
  cudaStream_t st = (cudaStream_t)stream;
  if (g.dim3) {
    switch (g.ndof) {
      case 1: return dispatch_elem<1, true>(mode, g, Ke_host, s, bcmask, bcdiagval, x, b, diag, w, y, dotv, dot_out, ws, st);
      case 2: return dispatch_elem<2, true>(mode, g, Ke_host, s, bcmask, bcdiagval, x, b, diag, w, y, dotv, dot_out, ws, st);
      case 3: return dispatch_elem<3, true>(mode, g, Ke_host, s, bcmask, bcdiagval, x, b, diag, w, y, dotv, dot_out, ws, st);
    }
  } else {
    switch (g.ndof) {
      case 1: return dispatch_elem<1, false>(mode, g, Ke_host, s, bcmask, bcdiagval, x, b, diag, w, y, dotv, dot_out, ws, st);
      case 2: return dispatch_elem<2, false>(mode, g, Ke_host, s, bcmask, bcdiagval, x, b, diag, w, y, dotv, dot_out, ws, st);
      case 3: return dispatch_elem<3, false>(mode, g, Ke_host, s, bcmask, bcdiagval, x, b, diag, w, y, dotv, dot_out, ws, st);
    }
  }
  return 1;
}

// Time every variant of the 3-D matrix-free kernel on the caller's buffers (Jacobi mode, y is scratch) and keep the
// fastest for this process and this number of dofs per node.  ms_out[pmb_elem_num_variants()] receives the average launch time of each variant.  Not capturable.
extern "C" int pmb_elem_autotune(const pmb_grid* p, const double* Ke_host, const double* s, const unsigned char* bcmask,
                                 double bcdiagval, const double* x, const double* b, const double* diag, double* y,
                                 double* ms_out, void* stream) {
  if (validate_grid(p, "pmb_elem_autotune")) return 1;
  PMB_REQUIRE(p->nz > 0, "pmb_elem_autotune: 3-D grids only");
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return pmb_set_error("pmb_elem_autotune: cudaEventCreate failed");
  int& slot = elem_variant(p->ndof);
  const int saved = slot;
  int best = saved, rc = 0;
  float best_ms = 1e30f;
  for (int v = 0; v < PMB_ELEM_VARIANTS && !rc; ++v) {
    slot = v;
    const int reps = 6;
    for (int r = 0; r < 2 + reps && !rc; ++r) {
      if (r == 2) cudaEventRecord(e0, st);
      rc = pmb_elem_spmv(p, EMODE_JACOBI, Ke_host, s, bcmask, bcdiagval, x, b, diag, 0.5, y, nullptr, nullptr, nullptr, stream);
    }
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) rc = pmb_set_error("pmb_elem_autotune: %s", cudaGetErrorString(cudaGetLastError()));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    if (ms_out) ms_out[v] = ms;
    // the tensor-core layout is timed and reported but never selected: its y equals the others' only to rounding
    if (!rc && ms < best_ms && v != 3) best_ms = ms, best = v;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  slot = rc ? saved : best;
  return rc;
}
