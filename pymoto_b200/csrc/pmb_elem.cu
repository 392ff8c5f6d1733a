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
#include <cuda.h>
#include <cstdlib>
#include <type_traits>
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
                                                    const unsigned char* __restrict__ mask_in,
                                                    const unsigned char* __restrict__ flags, double bcdiag,
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
  // (brick flags are NOT used here: reading the flag first puts one more dependent global load in front of the staging loads
  //  of every CTA and costs more than the mask bytes it saves -- measured 0.363 vs 0.348 ms at 256x128x128)
  const unsigned char* mask = mask_in;
  (void)flags;

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
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

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

// ---------------------------------------------------------------------------------------------------------
// Ring layout (variant 4, 3-D): persistent CTAs, bricks staged by TMA bulk copies into a 2-stage shared-memory ring.
//
// Source-level sampling of elem_kernel put 58 % of a warp's lifetime into the staging prologue (per-thread index
// arithmetic, predicates, __ldg latency before the barrier) while the FMA phase itself runs at the FP64 pipe rate.  Here
// the staging leaves the instruction stream: a CTA walks its bricks round-robin, and while brick t is in its FMA phase the
// 24 x-rows (34 nodes each) and 15 density rows of brick t+1 are already in flight as 1-D bulk copies
// (cp.async.bulk ... mbarrier::complete_tx::bytes; SASS UBLKCP), one row per lane of warp 0, landing in the other ring
// stage.  A tensor-map (cp.async.bulk.tensor) cannot describe the planes: its global strides must be multiples of 16 bytes
// and a node row is NX*NDOF doubles (771 for 257 nodes x 3 dofs), so every row is copied from its own 16-byte-aligned hull
// and read back with a per-row shift of 0 or 1 double.  Nothing is zero-filled or masked on the way in:
//   * out-of-grid node slots keep finite stale values and are only ever multiplied by the density of an out-of-grid
//     element, which IS forced to +0.0 (edge bricks patch their density tile after the copies land);
//   * Dirichlet columns are zeroed in shared memory only in bricks whose staged region holds a constrained dof
//     (`brickflags`, one byte per brick from pmb_elem_brickflags; bc sets are fixed over a design run).
// Per element the (dk, dj, di) FMA order of elem_kernel is kept, so y is bit-identical to variants 0-2.
// ---------------------------------------------------------------------------------------------------------
template <int NDOF>
struct RingCfg {
  static constexpr int BX = 32, BY = 4, BZ = 2, NT = BX * BY * BZ;
  static constexpr int TX = BX + 2, TY = BY + 2, TZ = BZ + 2, SZ = BZ + 1;
  static constexpr int XROWS = TZ * TY, XLEN = TX * NDOF, XPITCH = (XLEN + 2 + 1) / 2 * 2;
  static constexpr int SROWS = SZ * (BY + 1), SLEN = BX + 1, SPITCH = (SLEN + 2 + 1) / 2 * 2;
  static constexpr int STAGE = XROWS * XPITCH + SROWS * SPITCH;  // doubles per ring stage (even)
  static constexpr int STAGES = 2;
  static constexpr size_t SMEM = sizeof(double) * STAGE * STAGES;
  static_assert(STAGE % 2 == 0, "ring stages must stay 16-byte aligned");
  static_assert(XROWS + SROWS <= 64, "two staging rows per lane of the producer warp");
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int NDOF, int MODE, int CTAS>
__global__ void __launch_bounds__(RingCfg<NDOF>::NT, CTAS)
    elem_kernel_ring(Geo g, const __grid_constant__ KeParam<NDOF, true> ke, int nbx, int nby, int nbricks,
                     const double* __restrict__ s, const unsigned char* __restrict__ mask,
                     const unsigned char* __restrict__ flags, double bcdiag, const double* __restrict__ x,
                     const double* __restrict__ b, const double* __restrict__ diag, double w, double* __restrict__ y,
                     const double* __restrict__ dotv, double* __restrict__ partials) {
  using C = RingCfg<NDOF>;
  constexpr int BX = C::BX, BY = C::BY, BZ = C::BZ, NT = C::NT, TY = C::TY, XP = C::XPITCH, SP = C::SPITCH;
  constexpr int LD = KeParam<NDOF, true>::LD;
  extern __shared__ __align__(128) double ring[];
  __shared__ __align__(8) uint64_t full_bar[C::STAGES];
  __shared__ double wred[3][NT / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int my = (nbricks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // bricks blockIdx.x + t * gridDim.x
  const long long Dx = (long long)(reinterpret_cast<uintptr_t>(x) >> 3), Ds = (long long)(reinterpret_cast<uintptr_t>(s) >> 3);
  const long long xrow = (long long)g.NX * NDOF;  // doubles per node row

  // stale-but-finite contract: the ring starts as zeros, later it only ever holds copied vector / density values
  for (int p = tid; p < C::STAGE * C::STAGES; p += NT) ring[p] = 0.0;
  if (tid == 0) {
    for (int q = 0; q < C::STAGES; ++q) mbar_init(&full_bar[q], 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zeros before the async-proxy copies
  __syncthreads();

  // ---- producer (all lanes of warp 0): lane r copies staging rows r and r + 32 of brick t2 into ring stage t2 % 2
  auto issue = [&](int t2) {
    const int stg = t2 % C::STAGES;
    double* su = ring + (size_t)stg * C::STAGE;
    double* ss = su + C::XROWS * XP;
    const int brick = (int)blockIdx.x + t2 * (int)gridDim.x;
    const int bx = brick % nbx, rem = brick / nbx;
    const int i0 = bx * BX, j0 = (rem % nby) * BY, kl0 = (rem / nby) * BZ;
    uintptr_t src[2];
    double* dst[2];
    unsigned bytes[2] = {0u, 0u};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int r = lane + 32 * q;
      if (r < C::XROWS) {
        const int kl = kl0 - 1 + r / TY, j = j0 - 1 + r % TY, k = g.kz0 + kl;
        const int ia = max(i0 - 1, 0), ib = min(i0 + BX + 1, g.NX);
        // planes beyond the slab's upper halo (kl > nzl) are never needed and may not be mapped
        if (j >= 0 && j < g.NY && k >= 0 && k < g.NZ && kl <= g.nzl && ib > ia) {
          const long long rowb = ((long long)kl * g.NY + j) * xrow;
          const long long D0 = Dx + rowb + (long long)(i0 - 1) * NDOF;                       // wanted first double (absolute index)
          const long long lo = (Dx + rowb + (long long)ia * NDOF) & ~1LL;                    // 16-byte hull of the in-grid part
          const long long hi = (Dx + rowb + (long long)ib * NDOF + 1) & ~1LL;
          src[q] = (uintptr_t)lo << 3;
          dst[q] = su + r * XP + (int)(lo - D0 + (D0 & 1));  // absolute double D lands at row[D - D0 + (D0 & 1)]
          bytes[q] = (unsigned)(hi - lo) * 8u;
        }
      } else if (r < C::XROWS + C::SROWS) {
        const int rs = r - C::XROWS;
        const int el = kl0 - 1 + rs / (BY + 1), ej = j0 - 1 + rs % (BY + 1), ek = g.kz0 + el;
        const int ia = max(i0 - 1, 0), ib = min(i0 + BX, g.nx);
        // element layers above the last owned node plane belong to the next rank and are not needed
        if (ej >= 0 && ej < g.ny && ek >= 0 && ek < g.nzE && el < g.nzl && ib > ia) {
          const long long rowb = ((long long)el * g.ny + ej) * g.nx;
          const long long D0 = Ds + rowb + (i0 - 1);
          const long long lo = (Ds + rowb + ia) & ~1LL, hi = (Ds + rowb + ib + 1) & ~1LL;
          src[q] = (uintptr_t)lo << 3;
          dst[q] = ss + rs * SP + (int)(lo - D0 + (D0 & 1));
          bytes[q] = (unsigned)(hi - lo) * 8u;
        }
      }
    }
    const unsigned total = bytes[0] + bytes[1];
    if (total) mbar_expect_tx(&full_bar[stg], total);  // arrive + expect: my bytes are announced before they can complete
    else mbar_arrive(&full_bar[stg]);
#pragma unroll
    for (int q = 0; q < 2; ++q)
      if (bytes[q]) tma_load_1d(dst[q], reinterpret_cast<const void*>(src[q]), bytes[q], &full_bar[stg], false);
  };
  if (warp == 0)
    for (int t2 = 0; t2 < C::STAGES && t2 < my; ++t2) issue(t2);

  const int tx = tid % BX, ty = (tid / BX) % BY, tz = tid / (BX * BY);
  const int pkx = (int)(((long long)g.NY * xrow) & 1), pjx = (int)(xrow & 1);       // row-shift parity per +1 plane / +1 row
  const int pks = (int)(((long long)g.ny * g.nx) & 1), pjs = g.nx & 1;
  double d0 = 0.0, d1 = 0.0, d2 = 0.0;

  for (int t = 0; t < my; ++t) {
    const int stg = t % C::STAGES;
    double* su = ring + (size_t)stg * C::STAGE;
    double* ss = su + C::XROWS * XP;
    const int brick = (int)blockIdx.x + t * (int)gridDim.x;
    const int bx = brick % nbx, rem = brick / nbx;
    const int i0 = bx * BX, j0 = (rem % nby) * BY, kl0 = (rem / nby) * BZ;
    const int i = i0 + tx, j = j0 + ty, kl = kl0 + tz;
    const bool valid = i < g.NX && j < g.NY && kl < g.nzl;
    const long long ln = ((long long)kl * g.NY + j) * g.NX + i;
    const bool flagged = mask && (flags ? flags[brick] != 0 : true);  // CTA-uniform
    // density slots outside the grid / slab must read +0.0 (CTA-uniform test on the brick's element range)
    const bool edge = i0 == 0 || i0 + BX > g.nx || j0 == 0 || j0 + BY > g.ny || g.kz0 + kl0 == 0 ||
                      g.kz0 + kl0 + BZ > g.nzE || kl0 + BZ > g.nzl;

    // ---- epilogue operands of this thread's rows: issued now, consumed after the FMA phase
    double br[NDOF], dr[NDOF], dvr[NDOF], xm[NDOF];
    bool mr[NDOF];
#pragma unroll
    for (int d = 0; d < NDOF; ++d) {
      const long long r = valid ? ln * NDOF + d : 0;
      mr[d] = flagged && __ldg(mask + r);
      xm[d] = mr[d] ? __ldg(x + r) : 0.0;  // Dirichlet rows (rare) need the unmasked x
      br[d] = (MODE != EMODE_SPMV) ? __ldg(b + r) : 0.0;
      dr[d] = (MODE == EMODE_JACOBI) ? __ldg(diag + r) : 1.0;
      dvr[d] = (partials && dotv) ? __ldg(dotv + r) : 0.0;
    }

    mbar_wait(&full_bar[stg], (unsigned)((t / C::STAGES) & 1));

    if (edge || flagged) {
      if (edge) {
        for (int p = tid; p < C::SROWS * C::SLEN; p += NT) {
          const int rs = p / C::SLEN, c = p - rs * C::SLEN;
          const int el = kl0 - 1 + rs / (BY + 1), ej = j0 - 1 + rs % (BY + 1), ek = g.kz0 + el, ei = i0 - 1 + c;
          const bool ok = ei >= 0 && ei < g.nx && ej >= 0 && ej < g.ny && ek >= 0 && ek < g.nzE && el < g.nzl;
          if (!ok) ss[rs * SP + c + (int)((Ds + ((long long)el * g.ny + ej) * g.nx + (i0 - 1)) & 1)] = 0.0;
        }
      }
      if (flagged) {
        for (int p = tid; p < C::XROWS * C::XLEN; p += NT) {
          const int r = p / C::XLEN, c = p - r * C::XLEN;
          const int klr = kl0 - 1 + r / TY, jr = j0 - 1 + r % TY, kr = g.kz0 + klr, ir = i0 - 1 + c / NDOF;
          if (ir >= 0 && ir < g.NX && jr >= 0 && jr < g.NY && kr >= 0 && kr < g.NZ && klr <= g.nzl) {
            const long long rowb = ((long long)klr * g.NY + jr) * xrow + (long long)(i0 - 1) * NDOF;
            if (__ldg(mask + rowb + c)) su[r * XP + c + (int)((Dx + rowb) & 1)] = 0.0;
          }
        }
      }
      __syncthreads();
    }

    // ---- FMA phase (same (dk, dj, di) order per element as elem_kernel -> bit-identical y)
    // staged value of node (tx+1+di, ty+1+dj, tz+1+dk), component c: xb[|dk|&1][|dj|&1][((dk*TY + dj)*XP + di*NDOF + c]
    const long long q0x = Dx + ((long long)kl * g.NY + j) * xrow + (long long)(i0 - 1) * NDOF;
    const double* xc = su + ((tz + 1) * TY + (ty + 1)) * XP + (tx + 1) * NDOF;
    const double* xb[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) xb[a][c] = xc + (int)((q0x + a * pkx + c * pjx) & 1);
    // density of element (ox, oy, oz) of this node: sb[oz][oy][(oz*(BY+1) + oy)*SP + ox]
    const long long q0s = Ds + ((long long)(kl - 1) * g.ny + (j - 1)) * g.nx + (i0 - 1);
    const double* sc = ss + (tz * (BY + 1) + ty) * SP + tx;
    const double* sb[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) sb[a][c] = sc + (int)((q0s + a * pks + c * pjs) & 1);

    double tt[8][NDOF];
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int d = 0; d < NDOF; ++d) tt[e][d] = 0.0;
    double acc[NDOF];
#pragma unroll
    for (int d = 0; d < NDOF; ++d) acc[d] = 0.0;
    if (valid) {
#pragma unroll
      for (int dk = -1; dk <= 1; ++dk)
#pragma unroll
        for (int dj = -1; dj <= 1; ++dj)
#pragma unroll
          for (int di = -1; di <= 1; ++di) {
            const double* up = xb[dk & 1][dj & 1] + (dk * TY + dj) * XP + di * NDOF;
            double uv[NDOF];
#pragma unroll
            for (int c = 0; c < NDOF; ++c) uv[c] = up[c];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int ox = e & 1, oy = (e >> 1) & 1, oz = (e >> 2) & 1;
              const int ax = 1 - ox, ay = 1 - oy, az = 1 - oz;
              const int bxx = ax + di, byy = ay + dj, bzz = az + dk;
              if (bxx < 0 || bxx > 1 || byy < 0 || byy > 1 || bzz < 0 || bzz > 1) continue;
              const int a = ax + 2 * ay + 4 * az, bn = bxx + 2 * byy + 4 * bzz;
#pragma unroll
              for (int c = 0; c < NDOF; ++c)
#pragma unroll
                for (int d = 0; d < NDOF; ++d) tt[e][d] = fma(ke.v[(a * NDOF + d) * LD + bn * NDOF + c], uv[c], tt[e][d]);
            }
          }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int ox = e & 1, oy = (e >> 1) & 1, oz = (e >> 2) & 1;
        const double se = sb[oz][oy][(oz * (BY + 1) + oy) * SP + ox];
#pragma unroll
        for (int d = 0; d < NDOF; ++d) acc[d] = fma(se, tt[e][d], acc[d]);
      }
#pragma unroll
      for (int d = 0; d < NDOF; ++d) {
        const long long r = ln * NDOF + d;
        const double xr = mr[d] ? xm[d] : xb[0][0][d];
        const double ax = mr[d] ? bcdiag * xr : acc[d];
        double out;
        if (MODE == EMODE_SPMV) out = ax;
        else if (MODE == EMODE_RESID) out = br[d] - ax;
        else out = xr + w * ((br[d] - ax) / dr[d]);
        y[r] = out;
        if (partials) {
          d0 = fma(out, xr, d0);
          d1 = fma(xr, dvr[d], d1);
          d2 = fma(out, dvr[d], d2);
        }
      }
    }
    __syncthreads();  // every thread is done with ring stage stg
    if (warp == 0 && t + C::STAGES < my) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads / patches before the async-proxy refill
      issue(t + C::STAGES);
    }
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
      partials[3 * (long long)blockIdx.x] = s0;
      partials[3 * (long long)blockIdx.x + 1] = s1;
      partials[3 * (long long)blockIdx.x + 2] = s2;
    }
  }
}

// flags[brick] = 1 if the staged region of the ring layout's brick (32 x 4 x 2 nodes + 1-node apron) holds a masked dof
__global__ void __launch_bounds__(256) elem_brickflags_kernel(Geo g, int nbx, int nby, const unsigned char* __restrict__ mask,
                                                               unsigned char* __restrict__ flags) {
  constexpr int BX = 32, BY = 4, BZ = 2, TX = BX + 2, TY = BY + 2, TZ = BZ + 2;
  const int brick = blockIdx.x;
  const int i0 = (brick % nbx) * BX, j0 = ((brick / nbx) % nby) * BY, kl0 = (brick / nbx / nby) * BZ;
  const int len = TX * g.ndof;
  int any = 0;
  for (int p = threadIdx.x; p < TZ * TY * len; p += 256) {
    const int r = p / len, c = p - r * len;
    const int kl = kl0 - 1 + r / TY, j = j0 - 1 + r % TY, k = g.kz0 + kl, i = i0 - 1 + c / g.ndof;
    if (i >= 0 && i < g.NX && j >= 0 && j < g.NY && k >= 0 && k < g.NZ && kl <= g.nzl)
      any |= mask[(((long long)kl * g.NY + j) * g.NX + (i0 - 1)) * g.ndof + c] != 0;
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) flags[brick] = any ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------
// y-marching tensor-core layout (variant 6, 3-D, ndof = 3): FP64 DMMA with register-resident Ke and in-register
// accumulation -- no scatter buffer, no gather phase.
//
// The DFMA layouts spend 3 instructions per multiply-add (DFMA + the LDCU that feeds its Ke operand + shared loads /
// addressing; ncu: issue slots saturated at 48 % FP64-pipe activity).  DMMA.8x8x4 does 256 multiply-adds per instruction
// with the Ke operand held in registers (each lane keeps 1/32 of every fragment).  Arrangement:
//   * the 8 columns of a DMMA are 8 CONSECUTIVE ELEMENT ROWS (j .. j+7) of one element column (ei, ek); B fragments are
//     the element displacements, read from staged node rows in shared memory;
//   * the 8 rows are the 6 rows (ay, d) of Ke that belong to one (ax, az) corner group (+ 2 zero rows): the product of
//     group (ax, az) with element column (ei, ek) is that column's contribution to NODE column (ei + ax, ek + az), so the
//     <= 4 contributions of a node column accumulate in the same registers (scaled per element by s_e on the way in);
//   * what remains is the shift by ay along the DMMA columns (= along y): node row j takes the ay = 0 row of column j
//     and the ay = 1 row of column j - 1 -- two warp shuffles, and one carried value into the next batch of 8 rows.
//     A CTA owns a strip of 32 x 2 node columns and marches along y through the whole grid, so that carry never crosses
//     a CTA and no element is computed twice.
// Staging is the bulk-copy ring of the layouts above (per step 36 node rows + 24 density rows, one per lane of warp 0).
// Tensor-core accumulation order differs from the DFMA layouts: y agrees to rounding, not bit for bit.
// ---------------------------------------------------------------------------------------------------------
#ifndef PMB_YM_DBG
#define PMB_YM_DBG 0   // diagnostic builds only (scripts/ablate_ym.sh): 1 = no epilogue, 2 = no DMMA, 4 = no shared loads, 8 = no finalize
#endif
struct YmCfg {
  static constexpr int BX = 32, BZ = 2, JB = 8, NT = 256, NDOF = 3;
  static constexpr int XLEN = (BX + 2) * NDOF;               // doubles of a staged node row (34 nodes)
  static constexpr int XROWS = (BZ + 2) * (JB + 1);          // node rows a step needs: 4 planes x 9 rows
  // tensor-map TMA staging: per plane two boxes of 5 rows x 102 doubles (rows of even / odd parity, see below), each
  // padded to 512 doubles (4096 B); one box of 34 x 8 x 3 element densities
  // (a TMA box must START on a 16-byte boundary -- an odd double coordinate traps with "illegal instruction",
  //  scripts/probe_tma_f64.cu -- so boxes start one double / one element early where needed and are 104 / 36 wide)
  static constexpr int XBOX_ROWS = 5, XBOX = XLEN + 2, XREG = 528, XDOUBLES = (BZ + 2) * 2 * XREG;
  static constexpr int SBOX_X = BX + 4, SDOUBLES = SBOX_X * JB * (BZ + 1);
  static constexpr int STAGE = XDOUBLES + SDOUBLES, STAGES = 2;
  static constexpr int OUTW = 8 * NDOF;                      // outputs of a warp per node row
  static constexpr int SMEM_DOUBLES = STAGE * STAGES + (NT / 32) * (JB * OUTW + OUTW) + 24 * 24;
  static_assert(XBOX_ROWS * XBOX <= XREG && (XREG * 8) % 128 == 0 && (XDOUBLES * 8) % 128 == 0 && (STAGE * 8) % 128 == 0,
                "TMA boxes land 128-byte aligned");
  static_assert((XBOX * 8) % 16 == 0 && (SBOX_X * 8) % 16 == 0, "TMA box rows are whole 16-byte granules");
};

// TMA tensor-map loads (cp.async.bulk.tensor, SASS UTMALDG): coordinates are ELEMENT indices, innermost first; the part of
// a box outside the tensor is zero-filled
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}

template <int MODE, int CTAS>
__global__ void __launch_bounds__(YmCfg::NT, CTAS)
    elem_kernel_ym(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tms, Geo g,
                   const __grid_constant__ KeParam<3, true> ke, int nsteps, int nbz, int szoff,
                   const unsigned char* __restrict__ mask, const unsigned char* __restrict__ flags, double bcdiag,
                   const double* __restrict__ x, const double* __restrict__ b, const double* __restrict__ diag, double w,
                   double* __restrict__ y, const double* __restrict__ dotv, double* __restrict__ partials) {
  using C = YmCfg;
  constexpr int NDOF = 3, JB = C::JB, NT = C::NT, XL = C::XBOX, XREG = C::XREG;
  extern __shared__ __align__(128) double ring[];
  double* sOut = ring + C::STAGE * C::STAGES;                 // [warp][JB][24] finished rows, then [warp][24] carry
  double* sCarry = sOut + (NT / 32) * JB * C::OUTW;
  double* sKe = sCarry + (NT / 32) * C::OUTW;
  __shared__ __align__(8) uint64_t full_bar[C::STAGES];
  __shared__ int done_cnt[C::STAGES];
  __shared__ double wred[3][NT / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 1-D grid, strip-major: the CTAs of strip 0 (the ones that usually carry the Dirichlet face and patch their stages) start
  // in the first wave, the nearly empty last strip runs last
  const int bxs = blockIdx.x / nbz, bzs = blockIdx.x - bxs * nbz;
  const int i0 = bxs * C::BX, kl0 = bzs * C::BZ;
  if (tid < C::STAGES) done_cnt[tid] = 0;
  const long long xrow = (long long)g.NX * NDOF;

  for (int p = tid; p < C::STAGE * C::STAGES; p += NT) ring[p] = 0.0;   // planes that are never staged read as zeros
  for (int p = tid; p < (NT / 32) * C::OUTW; p += NT) sCarry[p] = 0.0;
  for (int p = tid; p < 24 * 24; p += NT) sKe[p] = ke.v[p];
  if (tid == 0) {
    for (int q = 0; q < C::STAGES; ++q) mbar_init(&full_bar[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  // (the tensor maps are the FIRST kernel parameters: a descriptor that sits beyond the classic 4 KB parameter window --
  //  behind the 4.6 KB element matrix -- makes cp.async.bulk.tensor trap with "illegal instruction")
  // ---- producer (one thread): 8 + 1 tensor-map TMA loads per step instead of 60 row copies (the copy engine retires only
  //      ~10 small bulk copies per microsecond and SM, which capped every row-staged layout at ~0.4 ms).
  //      Node vector: a row of nodes is NX*3 doubles -- an odd number, so neither the row nor the plane stride is a multiple
  //      of 16 bytes as a tensor map requires; TWO consecutive rows are.  The map describes the (plane-padded) vector as
  //      [super-row = 2 node rows][2 * NX * 3 doubles]; the 9 rows of a plane that a step needs are fetched as two boxes of 5
  //      super-rows x 102 doubles, one for the rows in the first half of their super-row and one for those in the second.
  //      Row r of plane p lands in region (p, r & 1), slot r >> 1.  Columns that fall off a row read its neighbour row (finite,
  //      only ever multiplied by the density of an out-of-grid element); the density box is zero-filled outside the grid by
  //      the TMA unit itself.
  auto issue = [&](int t2) {
    const int stg = t2 % C::STAGES;
    double* su = ring + (size_t)stg * C::STAGE;
    unsigned bytes = C::SDOUBLES * 8u;
#pragma unroll
    for (int p = 0; p < C::BZ + 2; ++p) {
      const int klp = kl0 - 1 + p, k = g.kz0 + klp;
      if (k >= 0 && k < g.NZ && klp <= g.nzl) bytes += 2u * C::XBOX_ROWS * XL * 8u;  // full boxes count, zero-filled parts included
    }
    mbar_expect_tx(&full_bar[stg], bytes);
#pragma unroll
    for (int p = 0; p < C::BZ + 2; ++p) {
      const int klp = kl0 - 1 + p, k = g.kz0 + klp;
      if (k >= 0 && k < g.NZ && klp <= g.nzl) {   // planes outside the grid / beyond the upper halo stay zero
        const int R0 = (klp + 1) * g.NY + JB * t2, a = R0 & 1;   // row index in the map (its first plane is local plane -1)
        // box starts floored to an even double: the wanted first double then sits at offset (start & 1) of every box row
        tma_load_2d(su + (2 * p) * XREG, &tmx, (a * (int)xrow + (i0 - 1) * NDOF) & ~1, R0 >> 1, &full_bar[stg]);
        tma_load_2d(su + (2 * p + 1) * XREG, &tmx, ((1 - a) * (int)xrow + (i0 - 1) * NDOF) & ~1, (R0 + 1) >> 1, &full_bar[stg]);
      }
    }
    tma_load_3d(su + C::XDOUBLES, &tms, i0 - 2, JB * t2, kl0 - 1 + szoff, &full_bar[stg]);
  };
  if (tid == 0)
    for (int t2 = 0; t2 < C::STAGES && t2 < nsteps; ++t2) issue(t2);

  // ---- consumer role of this warp: node columns i = iw .. iw+7 of plane kl
  const int wx = warp & 3, wz = warp >> 2;
  const int iw = i0 + 8 * wx, kl = kl0 + wz;
  const bool wactive = iw < g.NX && kl < g.nzl;
  const int gq = lane >> 2, q4 = lane & 3;   // DMMA fragment coordinates: row / column group, k index
  // A fragments: group (ax, az), row gq = (ay, d) (rows 6, 7 are zero), k = 4 ks + q4
  double af[2][2][6];
#pragma unroll
  for (int az = 0; az < 2; ++az)
#pragma unroll
    for (int ax = 0; ax < 2; ++ax)
#pragma unroll
      for (int ks = 0; ks < 6; ++ks) {
        const int ay = gq / 3, d = gq - 3 * ay;
        const int a = ax + 2 * ay + 4 * az;
        af[az][ax][ks] = gq < 6 ? sKe[(a * NDOF + d) * 24 + 4 * ks + q4] : 0.0;
      }
  // offset of the wanted first double inside the box rows of region (plane p, row parity q): parity of the box start
  auto xshift = [&](int p, int q) {
    const int a = ((kl0 + p) * g.NY) & 1;  // (JB * t is even: the parity does not depend on the step)
    return ((q ? 1 - a : a) * (int)xrow + (i0 - 1) * NDOF) & 1;
  };
  // B fragment offsets (doubles inside a ring stage, element column offset excluded): k index kk = 4 ks + q4 = 3 bn + c,
  // column gq = element row; node (bx, gq + by, bz) of the element, layer selector az (element layer = kl - az)
  int boff[2][6];
#pragma unroll
  for (int az = 0; az < 2; ++az)
#pragma unroll
    for (int ks = 0; ks < 6; ++ks) {
      const int kk = 4 * ks + q4, bn = kk / 3, c = kk - 3 * bn;
      const int bx = bn & 1, by = (bn >> 1) & 1, bz = bn >> 2;
      const int p = wz + 1 - az + bz, row = gq + by;
      boff[az][ks] = (2 * p + (row & 1)) * XREG + (row >> 1) * XL + bx * NDOF + c + xshift(p, row & 1);
    }
  int soff[2][2];
#pragma unroll
  for (int az = 0; az < 2; ++az)
#pragma unroll
    for (int h = 0; h < 2; ++h) soff[az][h] = C::XDOUBLES + ((wz + 1 - az) * JB + 2 * q4 + h) * C::SBOX_X + 1;  // box starts at element i0 - 2
  double* myOut = sOut + warp * JB * C::OUTW;
  double* myCarry = sCarry + warp * C::OUTW;
  const int nbxg = gridDim.x / nbz;
  const size_t flag0 = ((size_t)bzs * nbxg + bxs) * nsteps;
  double d0 = 0.0, d1 = 0.0, d2 = 0.0;
  bool fnext = flags ? __ldg(flags + flag0) != 0 : true;

  for (int t = 0; t <= nsteps; ++t) {
    const bool flush = t == nsteps;       // last pass: only the carried row (node row 8 nsteps, when it exists)
    if (flush && (JB * nsteps >= g.NY)) break;
    const int stg = t % C::STAGES;
    double* su = ring + (size_t)stg * C::STAGE;
    const int j0 = JB * t;
    bool flagged = false;
    if (!flush) {
      flagged = mask && fnext;
      // next step's flag and this step's epilogue operands start their trip through the memory system now
      if (flags && t + 1 < nsteps) fnext = __ldg(flags + flag0 + t + 1) != 0;
      if (wactive && lane < 16 && j0 + (lane >> 1) < g.NY) {
        const long long pr = (((long long)kl * g.NY + j0 + (lane >> 1)) * g.NX + iw) * NDOF + 16 * (lane & 1);
        if (MODE != EMODE_SPMV) prefetch_l2(b + pr);
        if (MODE == EMODE_JACOBI) prefetch_l2(diag + pr), prefetch_l2(x + pr);
      }
      mbar_wait(&full_bar[stg], (unsigned)((t / C::STAGES) & 1));
      if (flagged) {   // Dirichlet columns of the staged rows are zeroed in place (CTA-uniform, rare).  All mask bytes of a
                       // thread are loaded before the first is used: one memory round trip per step instead of 15
        constexpr int NQ = (C::XROWS * C::XLEN + NT - 1) / NT;
        unsigned char mk[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int p = tid + q * NT, rr = p / C::XLEN, c = p - rr * C::XLEN;
          const int pl = rr / (JB + 1), r = rr - pl * (JB + 1);
          const int klr = kl0 - 1 + pl, jr = j0 + r, kr = g.kz0 + klr, ir = i0 - 1 + c / NDOF;
          const bool ok = p < C::XROWS * C::XLEN && ir >= 0 && ir < g.NX && jr < g.NY && kr >= 0 && kr < g.NZ && klr <= g.nzl;
          mk[q] = ok ? __ldg(mask + ((long long)klr * g.NY + jr) * xrow + (long long)(i0 - 1) * NDOF + c) : (unsigned char)0;
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q)
          if (mk[q]) {
            const int p = tid + q * NT, rr = p / C::XLEN, c = p - rr * C::XLEN;
            const int pl = rr / (JB + 1), r = rr - pl * (JB + 1);
            su[(2 * pl + (r & 1)) * XREG + (r >> 1) * XL + c + xshift(pl, r & 1)] = 0.0;
          }
        __syncthreads();
      }
    }

    if (wactive) {
      if (!flush) {
        // ---- DMMA phase: element columns ei = iw-1 .. iw+7, layers kl-1 (az = 1) and kl (az = 0); node column ei is
        //      finished once element column ei has been processed
        double ynext[2] = {0.0, 0.0};
#pragma unroll
        for (int ex = -1; ex < 8; ++ex) {
          double ycur[2] = {ynext[0], ynext[1]};
          ynext[0] = ynext[1] = 0.0;
          const int coloff = (8 * wx + ex + 1) * NDOF;
          // both element layers of the column at once, every 24-long dot product split into two 12-long halves: up to 8
          // independent DMMA chains of depth 3 per warp (a dependent DMMA issues only every ~100 cycles, and the FP64 tensor
          // pipe needs ~8 independent accumulations per scheduler to stay busy)
          double bv[2][6], sv[2][2];
#pragma unroll
          for (int az = 0; az < 2; ++az) {
#pragma unroll
            for (int ks = 0; ks < 6; ++ks) bv[az][ks] = (PMB_YM_DBG & 4) ? 1.0 + ks + lane : su[boff[az][ks] + coloff];
            sv[az][0] = (PMB_YM_DBG & 4) ? 1.0 : su[soff[az][0] + 8 * wx + ex + 1];
            sv[az][1] = (PMB_YM_DBG & 4) ? 1.0 : su[soff[az][1] + 8 * wx + ex + 1];
          }
          double cc[2][2][2][2];  // [az][ax][half][c0 / c1]
#pragma unroll
          for (int az = 0; az < 2; ++az)
#pragma unroll
            for (int ax = 0; ax < 2; ++ax)
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) cc[az][ax][hf][0] = cc[az][ax][hf][1] = 0.0;
#pragma unroll
          for (int ks = 0; ks < 3; ++ks)
#pragma unroll
            for (int az = 0; az < 2; ++az)
#pragma unroll
              for (int ax = 0; ax < 2; ++ax) {
                if ((ax == 0 && ex < 0) || (ax == 1 && ex > 6)) continue;  // node column outside this warp's eight
                if (PMB_YM_DBG & 2) continue;
                dmma884(cc[az][ax][0][0], cc[az][ax][0][1], af[az][ax][ks], bv[az][ks]);
                dmma884(cc[az][ax][1][0], cc[az][ax][1][1], af[az][ax][ks + 3], bv[az][ks + 3]);
              }
#pragma unroll
          for (int az = 0; az < 2; ++az) {
            if (ex >= 0) {   // ax = 0: this element column's own node column
              ycur[0] = fma(sv[az][0], cc[az][0][0][0] + cc[az][0][1][0], ycur[0]);
              ycur[1] = fma(sv[az][1], cc[az][0][0][1] + cc[az][0][1][1], ycur[1]);
            }
            if (ex < 7) {    // ax = 1: the node column to the right
              ynext[0] = fma(sv[az][0], cc[az][1][0][0] + cc[az][1][1][0], ynext[0]);
              ynext[1] = fma(sv[az][1], cc[az][1][0][1] + cc[az][1][1][1], ynext[1]);
            }
          }
          if (ex >= 0 && !(PMB_YM_DBG & 8)) {
            // ---- node column di = ex finished: lane (row gq = (ay, d), cols 2 q4, 2 q4 + 1).  Node row c takes (ay = 0,
            //      col c) + (ay = 1, col c - 1); col -1 is the value carried from the previous step
            const int srcrow = (gq % 3 + 3) * 4;                              // lane base of row (ay = 1, d)
            const double t0 = __shfl_sync(0xffffffffu, ycur[1], srcrow + ((q4 + 3) & 3));   // (1, d), col 2 q4 - 1
            const double t1 = __shfl_sync(0xffffffffu, ycur[0], srcrow + q4);               // (1, d), col 2 q4
            if (gq < 3) {
              const double cin = myCarry[ex * NDOF + gq];
              myOut[(2 * q4) * C::OUTW + ex * NDOF + gq] = ycur[0] + (q4 == 0 ? cin : t0);
              myOut[(2 * q4 + 1) * C::OUTW + ex * NDOF + gq] = ycur[1] + t1;
            }
            __syncwarp();
            if (gq >= 3 && gq < 6 && q4 == 3) myCarry[ex * NDOF + gq - 3] = ycur[1];   // (1, d), col 7 -> next step's row 0
          }
        }
      } else {
        for (int p = lane; p < JB * C::OUTW; p += 32) myOut[p] = p < C::OUTW ? myCarry[p] : 0.0;
      }
      __syncwarp();
    }
    if (!flush) {
      // ---- release ring stage stg without a CTA barrier: the LAST warp to leave it refills it for step t + 2 (the epilogue
      //      below reads global memory only), so fast warps run on into their epilogue / next step instead of waiting
      __threadfence_block();  // this warp's reads of the stage are complete before it signs off
      int prev = 0;
      if (lane == 0) prev = atomicAdd(&done_cnt[stg], 1);
      prev = __shfl_sync(0xffffffffu, prev, 0);
      if (prev == NT / 32 - 1) {
        if (lane == 0) {
          done_cnt[stg] = 0;  // next touched after the refill below has landed and been consumed
          if (t + C::STAGES < nsteps) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(t + C::STAGES);
          }
        }
      }
    }
    if (wactive && !(PMB_YM_DBG & 1)) {
      // ---- epilogue: 8 node rows x 24 contiguous doubles (8 nodes x 3 dofs) of this warp; lane < 24 owns one column of
      //      them.  All loads of 4 rows are issued before the first use.
      const int nrows = flush ? 1 : JB;
      const bool act = lane < C::OUTW && iw + lane / NDOF < g.NX;
      const long long r0 = (((long long)kl * g.NY + j0) * g.NX + iw) * NDOF + lane;
      const bool masked_step = (flagged || flush) && mask;   // warp-uniform
#pragma unroll
      for (int h = 0; h < JB; h += 4) {
        double xr[4], br[4], dr[4], dv[4];
        bool ok[4], mr[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          ok[c] = act && h + c < nrows && j0 + h + c < g.NY;
          const long long r = ok[c] ? r0 + (long long)(h + c) * xrow : 0;
          xr[c] = (MODE == EMODE_JACOBI || partials || masked_step) ? __ldg(x + r) : 0.0;
          br[c] = (MODE != EMODE_SPMV) ? __ldg(b + r) : 0.0;
          dr[c] = (MODE == EMODE_JACOBI) ? __ldg(diag + r) : 1.0;
          dv[c] = (partials && dotv) ? __ldg(dotv + r) : 0.0;
          mr[c] = masked_step && __ldg(mask + r);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (ok[c]) {
            const double ax = mr[c] ? bcdiag * xr[c] : myOut[(h + c) * C::OUTW + lane];
            double out;
            if (MODE == EMODE_SPMV) out = ax;
            else if (MODE == EMODE_RESID) out = br[c] - ax;
            else out = xr[c] + w * ((br[c] - ax) / dr[c]);
            y[r0 + (long long)(h + c) * xrow] = out;
            if (partials) {
              d0 = fma(out, xr[c], d0);
              d1 = fma(xr[c], dv[c], d1);
              d2 = fma(out, dv[c], d2);
            }
          }
        }
      }
    }
    if (flush) break;
    __syncwarp();  // the epilogue's reads of myOut are done before the next step's lanes overwrite it (racecheck: WAR hazard)
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
      const long long bid = blockIdx.x;
      partials[3 * bid] = s0;
      partials[3 * bid + 1] = s1;
      partials[3 * bid + 2] = s2;
    }
  }
}

// flags[(bz * nbx + bx) * nsteps + t] = 1 iff the rows staged by CTA (bx, bz) of the y-marching layout in step t hold a masked dof
__global__ void __launch_bounds__(256) elem_ymflags_kernel(Geo g, int nsteps, const unsigned char* __restrict__ mask,
                                                            unsigned char* __restrict__ flags) {
  using C = YmCfg;
  const int t = blockIdx.x, i0 = blockIdx.y * C::BX, kl0 = blockIdx.z * C::BZ;   // grid (nsteps, nbx, nbz)
  int any = 0;
  for (int p = threadIdx.x; p < C::XROWS * C::XLEN; p += 256) {
    const int r = p / C::XLEN, c = p - r * C::XLEN;
    const int kl = kl0 - 1 + r / (C::JB + 1), j = C::JB * t + r % (C::JB + 1), k = g.kz0 + kl, i = i0 - 1 + c / 3;
    if (i >= 0 && i < g.NX && j < g.NY && k >= 0 && k < g.NZ && kl <= g.nzl)
      any |= mask[(((long long)kl * g.NY + j) * g.NX + (i0 - 1)) * 3 + c] != 0;
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) flags[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = any ? 1 : 0;
}

#include "pmb_elem_par.cuh"

static int sm_count_elem() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}
static void ring_bricks(const Geo& g, int& nbx, int& nby, int& nbz) {
  nbx = (g.NX + 31) / 32;
  nby = (g.NY + 3) / 4;
  nbz = (g.nzl + 1) / 2;
}

static void ym_grid(const Geo& g, int& nbx, int& nbz, int& nsteps) {
  nbx = (g.NX + YmCfg::BX - 1) / YmCfg::BX;
  nbz = (g.nzl + YmCfg::BZ - 1) / YmCfg::BZ;
  nsteps = (g.ny + YmCfg::JB - 1) / YmCfg::JB;
}

extern "C" long long pmb_elem_brickflags_bytes(const pmb_grid* p, int variant) {
  if (validate_grid(p, "pmb_elem_brickflags_bytes")) return -1;
  Geo g = make_geo(p);
  if (variant == 6 || variant == 7) {
    int nbx, nbz, nsteps;
    ym_grid(g, nbx, nbz, nsteps);
    return (long long)nbx * nbz * nsteps;
  }
  if (variant >= 8 && variant < 8 + PAR_NCFG) {
    const int ey = PAR_CFG[variant - 8].ey;
    return (long long)((g.NX + PAR_EX - 2) / (PAR_EX - 1)) * ((g.NY + ey - 2) / (ey - 1)) * (g.nzl + 2);
  }
  int nbx, nby, nbz;
  ring_bricks(g, nbx, nby, nbz);
  return (long long)nbx * nby * nbz;
}

extern "C" int pmb_elem_brickflags(const pmb_grid* p, int variant, const unsigned char* bcmask, unsigned char* flags, void* stream) {
  if (validate_grid(p, "pmb_elem_brickflags")) return 1;
  PMB_REQUIRE(bcmask && flags, "pmb_elem_brickflags: NULL pointer argument");
  PMB_REQUIRE(p->nz > 0, "pmb_elem_brickflags: 3-D grids only");
  PMB_REQUIRE(variant == 0 || (variant >= 4 && variant < 8 + PAR_NCFG), "pmb_elem_brickflags: layout %d takes no flags", variant);
  Geo g = make_geo(p);
  if (variant >= 8) {
    PMB_REQUIRE(g.ndof != 2, "pmb_elem_brickflags: the parity-block layouts are ndof = 1, 3 only");
    const int ey = PAR_CFG[variant - 8].ey;
    const int nbx = (g.NX + PAR_EX - 2) / (PAR_EX - 1), nby = (g.NY + ey - 2) / (ey - 1);
    PMB_REQUIRE(nby <= 65535 && g.nzl + 2 <= 65535, "pmb_elem_brickflags: grid too large");
    elem_parflags_kernel<<<dim3(nbx, nby, g.nzl + 2), 128, 0, (cudaStream_t)stream>>>(g, ey, bcmask, flags);
    PMB_CHECK_LAUNCH("pmb_elem_brickflags");
    return 0;
  }
  if (variant == 6 || variant == 7) {
    PMB_REQUIRE(g.ndof == 3, "pmb_elem_brickflags: layouts 6, 7 are ndof = 3 only");
    int nbx, nbz, nsteps;
    ym_grid(g, nbx, nbz, nsteps);
    PMB_REQUIRE(nbx <= 65535 && nbz <= 65535, "pmb_elem_brickflags: grid too large");
    elem_ymflags_kernel<<<dim3(nsteps, nbx, nbz), 256, 0, (cudaStream_t)stream>>>(g, nsteps, bcmask, flags);
    PMB_CHECK_LAUNCH("pmb_elem_brickflags");
    return 0;
  }
  int nbx, nby, nbz;
  ring_bricks(g, nbx, nby, nbz);
  const long long nb = (long long)nbx * nby * nbz;
  PMB_REQUIRE(nb < 2147483647LL, "pmb_elem_brickflags: too many bricks");
  elem_brickflags_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(g, nbx, nby, bcmask, flags);
  PMB_CHECK_LAUNCH("pmb_elem_brickflags");
  return 0;
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

// ---- layouts of the 3-D kernel (ndof 1 or 3; 2-D grids and ndof 2 always run layout 0): 0 = one node per thread on a
//      32x4x2 brick (elem_kernel), 1 / 2 = z-marching 32x8 / 32x4 columns with ring-buffered planes (elem_kernel_zm),
//      3 = FP64 tensor-core layout (elem_kernel_mma, ndof 3 only; other ndof run 0), 4 / 5 = persistent CTAs with bulk-copy
//      staged bricks (elem_kernel_ring) at 3 / 2 CTAs per SM, 6 = y-marching FP64 tensor-core layout with in-register
//      accumulation (elem_kernel_ym, ndof 3 only).  The layout is a field of the caller's pmb_elem_op: no process-wide state.
//      Layouts 0, 1, 2, 4, 5 produce bit-identical y; 3 and 6 (tensor-core accumulation order) agree to rounding.
//      8 / 9 = parity-block layout (elem_kernel_par, ndof 1 or 3; needs an element matrix with the reflection symmetry of
//      a cuboid voxel, otherwise layout 0 runs): 32 x 8 / 32 x 4 element columns per CTA; agrees with layout 0 to rounding.
enum { PMB_ELEM_VARIANTS = 8 + PAR_NCFG };
extern "C" int pmb_elem_num_variants(void) { return PMB_ELEM_VARIANTS; }

static int effective_variant(const Geo& g, int variant) {
  if (!g.dim3 || g.ndof == 2) return 0;
  if ((variant == 3 || variant == 6 || variant == 7) && g.ndof != 3) return 0;
  return variant;
}

static dim3 elem_grid_any(const Geo& g, int variant) {
  if (!g.dim3) return elem_grid<false>(g);
  variant = effective_variant(g, variant);
  switch (variant) {
    case 1: return dim3((g.NX + ZM_BX - 1) / ZM_BX, (g.NY + 7) / 8, (g.nzl + ZM_L - 1) / ZM_L);
    case 2: return dim3((g.NX + ZM_BX - 1) / ZM_BX, (g.NY + 3) / 4, (g.nzl + ZM_L - 1) / ZM_L);
    case 3: {
      const int zl = mma_zl(g);
      return dim3((g.NX + MM_NXT - 1) / MM_NXT, (g.NY + MM_NYT - 1) / MM_NYT, (g.nzl + zl - 1) / zl);
    }
    case 4:
    case 5: {
      int nbx, nby, nbz;
      ring_bricks(g, nbx, nby, nbz);
      const long long nb = (long long)nbx * nby * nbz, cap = (variant == 4 ? 3LL : 2LL) * sm_count_elem();
      return dim3((unsigned)(nb < cap ? nb : cap), 1, 1);
    }
    case 6:
    case 7: {
      int nbx, nbz, nsteps;
      ym_grid(g, nbx, nbz, nsteps);
      return dim3(nbx * nbz, 1, 1);
    }
  }
  if (variant >= 8 && variant < 8 + PAR_NCFG) {
    const ParLaunchCfg c = PAR_CFG[variant - 8];
    const int zl = par_zl(g, c.ey, c.minb, sm_count_elem());
    return dim3((g.NX + PAR_EX - 2) / (PAR_EX - 1), (g.NY + c.ey - 2) / (c.ey - 1), (g.nzl + zl - 1) / zl);
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

// ---- tensor maps of the y-marching layout (see elem_kernel_ym): the node vector as [2-row super-rows][2 * NX * 3] starting at
//      the plane below the slab (x must live in plane-padded storage whose first pad plane is 16-byte aligned:
//      DeviceCSR.new_vec()), the element scaling vector as a plain [layers][ny][nx] tensor of its valid layers
typedef CUresult (*pmb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static pmb_encode_tiled_fn encode_tiled_entry() {
  static pmb_encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (pmb_encode_tiled_fn)p;
  }
  return fn;
}

static bool ym_layout_applicable(const Geo& g, const double* x, const double* s) {
  const long long plane = (long long)g.NX * g.NY * 3;
  return g.dim3 && g.ndof == 3 && (g.nx % 2) == 0 && ((reinterpret_cast<uintptr_t>(x) - 8ull * plane) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(s) & 15) == 0 && encode_tiled_entry() != nullptr;
}

static int ym_tensor_maps(const Geo& g, const double* x, const double* s, CUtensorMap* tmx, CUtensorMap* tms, int* szoff) {
  pmb_encode_tiled_fn enc = encode_tiled_entry();
  PMB_REQUIRE(enc, "pmb_elem_spmv: cuTensorMapEncodeTiled not available in this driver");
  const cuuint64_t xrow = (cuuint64_t)g.NX * 3, plane = xrow * g.NY;
  const cuuint64_t rows = (cuuint64_t)(g.nzl + 2) * g.NY;
  {
    cuuint64_t dims[2] = {2 * xrow, (rows + 1) / 2};
    cuuint64_t strides[1] = {2 * xrow * sizeof(double)};
    cuuint32_t box[2] = {(cuuint32_t)YmCfg::XBOX, (cuuint32_t)YmCfg::XBOX_ROWS};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(x) - plane, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return pmb_set_error("pmb_elem_spmv: tensor map of the node vector rejected (CUresult %d)", (int)r);
  }
  {
    const int below = g.kz0 > 0 ? 1 : 0;   // the halo layer under the slab exists only inside the grid
    const int nlay = (g.nzl < g.nzE - g.kz0 ? g.nzl : g.nzE - g.kz0) + below;
    cuuint64_t dims[3] = {(cuuint64_t)g.nx, (cuuint64_t)g.ny, (cuuint64_t)(nlay > 0 ? nlay : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)g.nx * sizeof(double), (cuuint64_t)g.nx * g.ny * sizeof(double)};
    cuuint32_t box[3] = {(cuuint32_t)YmCfg::SBOX_X, (cuuint32_t)YmCfg::JB, (cuuint32_t)(YmCfg::BZ + 1)};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(tms, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(s) - (size_t)below * g.nx * g.ny, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return pmb_set_error("pmb_elem_spmv: tensor map of the element scaling vector rejected (CUresult %d)", (int)r);
    *szoff = below;
  }
  return 0;
}

template <int NDOF, bool DIM3, int MODE>
static int launch_elem(const Geo& g, const pmb_elem_op* op, const double* x, const double* b, const double* diag, double w,
                       double* y, const double* dotv, double* dot_out, double* ws, cudaStream_t st) {
  KeParam<NDOF, DIM3> ke;
  memcpy(ke.v, op->Ke_host, sizeof(ke.v));
  const double* s = op->s;
  const unsigned char* mask = op->bcmask;
  const double bcdiag = op->bcdiagval;
  int variant = effective_variant(g, op->variant);
  if ((variant == 6 || variant == 7) && !ym_layout_applicable(g, x, s)) variant = 0;  // odd nx / unpadded or misaligned storage
  [[maybe_unused]] ParBlocks<NDOF> kb;
  if (variant >= 8) {
    bool ok = false;
    if constexpr (DIM3 && NDOF != 2) ok = par_blocks<NDOF>(op->Ke_host, kb);
    if (!ok) variant = 0;  // element matrix without the reflection symmetry of a cuboid voxel
  }
  dim3 grid = elem_grid_any(g, variant);
  double* part = dot_out ? ws : nullptr;
  // the caller's flags follow the layout it ASKED for: a layout that falls back to the brick kernel must not read them
  const unsigned char* flags_brick = (op->variant == 0 || op->variant == 4 || op->variant == 5) ? op->brickflags : nullptr;
  if constexpr (DIM3 && NDOF != 2) {
    if (variant == 1)
      elem_kernel_zm<NDOF, MODE, 8, 2><<<grid, 256, 0, st>>>(g, ke, s, mask, bcdiag, x, b, diag, w, y, dotv, part);
    else if (variant == 2)
      elem_kernel_zm<NDOF, MODE, 4, 4><<<grid, 128, 0, st>>>(g, ke, s, mask, bcdiag, x, b, diag, w, y, dotv, part);
    else if (variant == 3) {
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
    } else if (variant == 6 || variant == 7) {
      if constexpr (NDOF == 3) {
        constexpr size_t smem = sizeof(double) * YmCfg::SMEM_DOUBLES;
        static bool configured = false;
        if (!configured) {
          cudaError_t e = cudaFuncSetAttribute(elem_kernel_ym<MODE, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e == cudaSuccess) e = cudaFuncSetAttribute(elem_kernel_ym<MODE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return pmb_set_error("elem_kernel_ym attribute: %s", cudaGetErrorString(e));
          configured = true;
        }
        int nbx, nbz, nsteps;
        ym_grid(g, nbx, nbz, nsteps);
        CUtensorMap tmx, tms;
        int szoff = 0;
        if (ym_tensor_maps(g, x, s, &tmx, &tms, &szoff)) return 1;
        if (variant == 6)
          elem_kernel_ym<MODE, 2><<<grid, YmCfg::NT, smem, st>>>(tmx, tms, g, ke, nsteps, nbz, szoff, mask, op->brickflags, bcdiag, x, b,
                                                                 diag, w, y, dotv, part);
        else
          elem_kernel_ym<MODE, 1><<<grid, YmCfg::NT, smem, st>>>(tmx, tms, g, ke, nsteps, nbz, szoff, mask, op->brickflags, bcdiag, x, b,
                                                                 diag, w, y, dotv, part);
      }
    } else if (variant >= 8) {
      const unsigned char* pflags = op->variant == variant ? op->brickflags : nullptr;   // flags follow the layout asked for
      int rc = 1;
      auto launch = [&](auto tag) {
        constexpr int I = decltype(tag)::value;
        constexpr ParLaunchCfg c = PAR_CFG[I];
        auto kern = elem_kernel_par<NDOF, MODE, c.ey, c.minb>;
        constexpr size_t smem = ParCfg<NDOF, c.ey>::SMEM;
        static bool configured = false;
        if (!configured) {
          cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) {
            rc = pmb_set_error("elem_kernel_par attribute: %s", cudaGetErrorString(e));
            return;
          }
          configured = true;
        }
        kern<<<grid, PAR_EX * c.ey, smem, st>>>(g, kb, par_zl(g, c.ey, c.minb, sm_count_elem()), s, mask, pflags, bcdiag, x, b, diag, w,
                                                y, dotv, part);
        rc = 0;
      };
      switch (variant - 8) {
        case 0: launch(std::integral_constant<int, 0>{}); break;
        case 1: launch(std::integral_constant<int, 1>{}); break;
      }
      if (rc) return rc;
    } else if (variant == 4 || variant == 5) {
      using C = RingCfg<NDOF>;
      static bool configured = false;
      if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(elem_kernel_ring<NDOF, MODE, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e == cudaSuccess)
          e = cudaFuncSetAttribute(elem_kernel_ring<NDOF, MODE, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return pmb_set_error("elem_kernel_ring attribute: %s", cudaGetErrorString(e));
        configured = true;
      }
      int nbx, nby, nbz;
      ring_bricks(g, nbx, nby, nbz);
      const long long nb = (long long)nbx * nby * nbz;
      PMB_REQUIRE(nb < 2147483647LL, "pmb_elem_spmv: too many bricks");
      if (variant == 4)
        elem_kernel_ring<NDOF, MODE, 3><<<grid, C::NT, C::SMEM, st>>>(g, ke, nbx, nby, (int)nb, s, mask, op->brickflags, bcdiag, x, b, diag,
                                                                      w, y, dotv, part);
      else
        elem_kernel_ring<NDOF, MODE, 2><<<grid, C::NT, C::SMEM, st>>>(g, ke, nbx, nby, (int)nb, s, mask, op->brickflags, bcdiag, x, b, diag,
                                                                      w, y, dotv, part);
    } else
      elem_kernel<NDOF, DIM3, MODE><<<grid, 256, 0, st>>>(g, ke, s, mask, flags_brick, bcdiag, x, b, diag, w, y, dotv, part);
  } else {
    elem_kernel<NDOF, DIM3, MODE><<<grid, 256, 0, st>>>(g, ke, s, mask, nullptr, bcdiag, x, b, diag, w, y, dotv, part);
  }
  PMB_CHECK_LAUNCH("pmb_elem_spmv");
  if (dot_out) {
    reduce_triples_kernel<<<1, 1024, 0, st>>>(ws, (long long)grid.x * grid.y * grid.z, dot_out);
    PMB_CHECK_LAUNCH("pmb_elem_spmv(reduce)");
  }
  return 0;
}

template <int NDOF, bool DIM3>
static int dispatch_elem(int mode, const Geo& g, const pmb_elem_op* op, const double* x, const double* b, const double* diag,
                         double w, double* y, const double* dotv, double* dot_out, double* ws, cudaStream_t st) {
  switch (mode) {
    case EMODE_SPMV: return launch_elem<NDOF, DIM3, EMODE_SPMV>(g, op, x, b, diag, w, y, dotv, dot_out, ws, st);
    case EMODE_RESID: return launch_elem<NDOF, DIM3, EMODE_RESID>(g, op, x, b, diag, w, y, dotv, dot_out, ws, st);
    case EMODE_JACOBI: return launch_elem<NDOF, DIM3, EMODE_JACOBI>(g, op, x, b, diag, w, y, dotv, dot_out, ws, st);
  }
  return pmb_set_error("pmb_elem_spmv: unknown mode %d", mode);
}

extern "C" int pmb_elem_spmv(const pmb_grid* p, int mode, const pmb_elem_op* op, const double* x, const double* b,
                             const double* diag, double w, double* y, const double* dotv, double* dot_out, double* ws,
                             void* stream) {
  if (validate_grid(p, "pmb_elem_spmv")) return 1;
  PMB_REQUIRE(op && op->Ke_host && op->s && x && y, "pmb_elem_spmv: NULL pointer argument");
  PMB_REQUIRE(op->variant >= 0 && op->variant < PMB_ELEM_VARIANTS, "pmb_elem_spmv: variant %d not in 0..%d", op->variant,
              PMB_ELEM_VARIANTS - 1);
  PMB_REQUIRE(x != y, "pmb_elem_spmv: y must not alias x");
  PMB_REQUIRE(mode == EMODE_SPMV || b, "pmb_elem_spmv: b required for residual / Jacobi");
  PMB_REQUIRE(mode != EMODE_JACOBI || diag, "pmb_elem_spmv: diag required for Jacobi");
  PMB_REQUIRE(!dot_out || ws, "pmb_elem_spmv: workspace required for the fused dot products");
  Geo g = make_geo(p);
  cudaStream_t st = (cudaStream_t)stream;
  if (g.dim3) {
    switch (g.ndof) {
      case 1: return dispatch_elem<1, true>(mode, g, op, x, b, diag, w, y, dotv, dot_out, ws, st);
      case 2: return dispatch_elem<2, true>(mode, g, op, x, b, diag, w, y, dotv, dot_out, ws, st);
      case 3: return dispatch_elem<3, true>(mode, g, op, x, b, diag, w, y, dotv, dot_out, ws, st);
    }
  } else {
    switch (g.ndof) {
      case 1: return dispatch_elem<1, false>(mode, g, op, x, b, diag, w, y, dotv, dot_out, ws, st);
      case 2: return dispatch_elem<2, false>(mode, g, op, x, b, diag, w, y, dotv, dot_out, ws, st);
      case 3: return dispatch_elem<3, false>(mode, g, op, x, b, diag, w, y, dotv, dot_out, ws, st);
    }
  }
  return 1;
}

// Time every layout of the 3-D matrix-free kernel on the caller's buffers (Jacobi mode, y is scratch).  ms_out[
// pmb_elem_num_variants()] receives the average launch time of each layout, *best the fastest one among those whose y is
// bit-identical to layout 0 (`allow_rounding` != 0 also admits the tensor-core layouts, whose y agrees to rounding).
// flags_scratch: pmb_elem_autotune_flag_bytes(g) bytes (the brick flags are layout-specific and recomputed per layout;
// may be NULL when op->bcmask is NULL).  The caller stores *best in its pmb_elem_op.  Not capturable into a CUDA graph.
extern "C" long long pmb_elem_autotune_flag_bytes(const pmb_grid* p) {
  long long m = pmb_elem_brickflags_bytes(p, 4);
  const long long b = p->ndof == 3 ? pmb_elem_brickflags_bytes(p, 6) : 0;
  m = b > m ? b : m;
  for (int v = 8; v < PMB_ELEM_VARIANTS; ++v) {
    const long long c = pmb_elem_brickflags_bytes(p, v);
    m = c > m ? c : m;
  }
  return m;
}

extern "C" int pmb_elem_autotune(const pmb_grid* p, const pmb_elem_op* op, const double* x, const double* b, const double* diag,
                                 double* y, unsigned char* flags_scratch, int allow_rounding, double* ms_out, int* best_out,
                                 void* stream) {
  if (validate_grid(p, "pmb_elem_autotune")) return 1;
  PMB_REQUIRE(p->nz > 0, "pmb_elem_autotune: 3-D grids only");
  PMB_REQUIRE(op && best_out, "pmb_elem_autotune: NULL pointer argument");
  PMB_REQUIRE(!op->bcmask || flags_scratch, "pmb_elem_autotune: flags_scratch required with a Dirichlet mask");
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return pmb_set_error("pmb_elem_autotune: cudaEventCreate failed");
  pmb_elem_op trial = *op;
  int best = op->variant, rc = 0;
  float best_ms = 1e30f;
  for (int v = 0; v < PMB_ELEM_VARIANTS && !rc; ++v) {
    trial.variant = v;
    trial.brickflags = nullptr;
    const bool wants_flags = (v == 4 || v == 5 || ((v == 6 || v == 7) && p->ndof == 3) || v >= 8) && p->ndof != 2;
    if (op->bcmask && wants_flags) {
      if (v == 4 || v == 6 || v >= 8) rc = pmb_elem_brickflags(p, v, op->bcmask, flags_scratch, stream);  // 5 / 7 reuse 4 / 6
      trial.brickflags = flags_scratch;
    }
    const int reps = 6;
    for (int r = 0; r < 2 + reps && !rc; ++r) {
      if (r == 2) cudaEventRecord(e0, st);
      rc = pmb_elem_spmv(p, EMODE_JACOBI, &trial, x, b, diag, 0.5, y, nullptr, nullptr, nullptr, stream);
    }
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) rc = pmb_set_error("pmb_elem_autotune: %s", cudaGetErrorString(cudaGetLastError()));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    if (ms_out) ms_out[v] = ms;
    const bool rounding = ((v == 3 || v == 6 || v == 7) && p->ndof == 3) || v >= 8;
    if (!rc && ms < best_ms && (allow_rounding || !rounding)) best_ms = ms, best = v;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *best_out = best;
  return rc;
}
