// libpmb: parity-block layout of the matrix-free finest-level operator (variants 8, 9 of pmb_elem_spmv; 3-D, ndof 1 or 3).
//
// The layouts of pmb_elem.cu evaluate y_e = Ke u_e as a dense 24 x 24 product: 576 multiply-adds per element, and they are
// bound by the FP64 issue rate (0.39 of the pipe peak).  This layout does less arithmetic instead of issuing it faster.
//
// The element matrix of a cuboid voxel with an isotropic (or orthotropic) material commutes with the three reflections of
// the element (x -> -x flips the node positions along x AND the sign of the x displacement).  In the basis of node parities
//     u^[p][c] = sum_a H[p][a] u[a][c],   H[p][a] = prod_axes (p_axis ? (a_axis ? +1 : -1) : 1)   (sums / differences along x, y, z)
// a dof (parity pattern p, component c) belongs to the irreducible representation q = p XOR e_c of Z2^3 (ndof 3) or q = p
// (scalar problems), and Ke only couples dofs of the same representation:
//     Ke = H'^T Kh H',   Kh = H' Ke H'^T / 64,   H' = H (x) I_ndof,   Kh = 8 diagonal blocks of ndof x ndof.
// Per element: forward butterflies (3 stages), 8 small blocks (72 multiply-adds for ndof 3, 8 multiplies for ndof 1), the
// transposed butterflies -- 264 FP64 instructions instead of 600; sharing the x / y stages of a node plane between the two
// element layers that touch it brings it to ~215.  The host verifies the block structure of the caller's Ke
// (par_blocks: off-block entries <= 1e-13 of the largest entry; the reference's hex8 stiffness and conductivity matrices
// pass at 1e-16, a generic element matrix falls back to layout 0), so the result equals the other layouts to rounding.
//
// Work distribution: one thread per ELEMENT column (ei, ej), marching over the element layers of a z-chunk.  A CTA holds
// 32 x EY element columns and owns the 31 x (EY - 1) node columns interior to them.  Per layer a thread
//   * reads the 4 corner nodes of the layer's top plane from a 2-slot shared-memory ring (cp.async, zero-fill outside the
//     grid, Dirichlet columns zeroed in shared memory), applies the x / y butterflies to them (the bottom plane's are carried
//     in registers from the previous layer), the z butterfly, the blocks, the density, the transposed z butterfly;
//   * adds the carried top-face result of the layer below -- still in the parity basis -- and applies the transposed x / y
//     butterflies to the sum: the contributions of layers el - 1 and el to its 4 corner nodes of plane el;
//   * publishes 3 of the 4 corners in shared memory; after ONE barrier per layer the thread that owns node (ei + 1, ej + 1)
//     adds the other three element columns' parts in a fixed order (deterministic) and applies the epilogue.
#pragma once

template <int NDOF>
struct ParBlocks {
  double v[8][NDOF][NDOF];  // v[q][c][c']: block of representation q, rows / columns ordered by component
};

// dof (p, c) -> representation
template <int NDOF>
__host__ __device__ constexpr int par_rep(int p, int c) {
  return NDOF == 3 ? (p ^ (1 << c)) : p;
}
// member c of representation q -> parity pattern
template <int NDOF>
__host__ __device__ constexpr int par_member(int q, int c) {
  return NDOF == 3 ? (q ^ (1 << c)) : q;
}

// Kh = H' Ke H'^T / 64 by in-place butterflies over the node index of the rows, then of the columns; returns false when Ke
// couples different representations (no reflection symmetry: anisotropic material, distorted element, arbitrary matrix).
template <int NDOF>
static bool par_blocks(const double* Ke, ParBlocks<NDOF>& kb) {
  constexpr int LD = 8 * NDOF;
  double A[LD * LD];
  for (int i = 0; i < LD * LD; ++i) A[i] = Ke[i];
  for (int pass = 0; pass < 2; ++pass) {  // pass 0: rows (index a of A[a*NDOF+d][.]), pass 1: columns
    for (int bit = 1; bit < 8; bit <<= 1)
      for (int a = 0; a < 8; ++a) {
        if (a & bit) continue;
        for (int d = 0; d < NDOF; ++d)
          for (int o = 0; o < LD; ++o) {
            const int lo = pass == 0 ? ((a * NDOF + d) * LD + o) : (o * LD + a * NDOF + d);
            const int hi = pass == 0 ? (((a | bit) * NDOF + d) * LD + o) : (o * LD + (a | bit) * NDOF + d);
            const double l = A[lo], h = A[hi];
            A[lo] = l + h;
            A[hi] = h - l;
          }
      }
  }
  double amax = 0.0, off = 0.0;
  for (int p = 0; p < 8; ++p)
    for (int c = 0; c < NDOF; ++c)
      for (int p2 = 0; p2 < 8; ++p2)
        for (int c2 = 0; c2 < NDOF; ++c2) {
          const double v = fabs(A[(p * NDOF + c) * LD + p2 * NDOF + c2]) / 64.0;
          if (!(v == v)) return false;
          amax = v > amax ? v : amax;
          if (par_rep<NDOF>(p, c) != par_rep<NDOF>(p2, c2)) off = v > off ? v : off;
        }
  if (off > 1e-13 * amax) return false;
  for (int q = 0; q < 8; ++q)
    for (int c = 0; c < NDOF; ++c)
      for (int c2 = 0; c2 < NDOF; ++c2)
        kb.v[q][c][c2] = A[(par_member<NDOF>(q, c) * NDOF + c) * LD + par_member<NDOF>(q, c2) * NDOF + c2] / 64.0;
  return true;
}

constexpr int PAR_EX = 32, PAR_RING = 4;   // element columns per CTA row; slots of the node-plane ring (3 planes in flight)
template <int NDOF, int EY, bool CSM>
constexpr size_t par_smem_bytes() {
  // plane ring + 2 corner-exchange buffers (3 corners) [+ with CSM the values carried from layer to layer: top-face result
  // and x / y-transformed top plane, 4 parity patterns each], per thread and dof
  return sizeof(double) * (PAR_RING * (EY + 1) * (PAR_EX + 1) * NDOF + (2 * 3 + (CSM ? 8 : 0)) * NDOF * PAR_EX * EY);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_pending() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// flags[(by * nbx + bx) * (nzl + 2) + kl + 1] = 1 iff the nodes of local plane kl (-1 .. nzl) that CTA tile (bx, by) of the
// parity-block layout stages hold a masked dof
__global__ void __launch_bounds__(128) elem_parflags_kernel(Geo g, int ey, const unsigned char* __restrict__ mask,
                                                            unsigned char* __restrict__ flags) {
  const int i0 = blockIdx.x * (PAR_EX - 1), j0 = blockIdx.y * (ey - 1), kl = (int)blockIdx.z - 1, k = g.kz0 + kl;
  const int len = (PAR_EX + 1) * g.ndof;
  int any = 0;
  if (k >= 0 && k < g.NZ && kl <= g.nzl)
    for (int p = threadIdx.x; p < (ey + 1) * len; p += 128) {
      const int r = p / len, c = p - r * len;
      const int j = j0 - 1 + r, i = i0 - 1 + c / g.ndof;
      if (i >= 0 && i < g.NX && j >= 0 && j < g.NY) any |= mask[(((long long)kl * g.NY + j) * g.NX + (i0 - 1)) * g.ndof + c] != 0;
    }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) flags[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * gridDim.z + blockIdx.z] = any ? 1 : 0;
}

template <int NDOF, int MODE, int EY, int MINB, bool CSM>
__global__ void __launch_bounds__(PAR_EX* EY, MINB)
    elem_kernel_par(Geo g, const __grid_constant__ ParBlocks<NDOF> kb, int zl, int zero, const double* __restrict__ s,
                    const unsigned char* __restrict__ mask, const unsigned char* __restrict__ flags, double bcdiag,
                    const double* __restrict__ x, const double* __restrict__ b, const double* __restrict__ diag, double w,
                    double* __restrict__ y, const double* __restrict__ dotv, double* __restrict__ partials) {
  constexpr int EX = PAR_EX, NT = EX * EY, R = PAR_RING, D = R - 1;
  constexpr int ROW = (EX + 1) * NDOF, PLANE = (EY + 1) * ROW;   // staged node plane: 33 x (EY + 1) nodes
  constexpr int NQ = (PLANE + NT - 1) / NT;
  constexpr int OC = NDOF * NT;                                  // one corner array of the exchange buffer: [c][ty][tx]
  static_assert(NQ <= 4, "mask bytes of a plane travel in one 32-bit register");
  extern __shared__ __align__(16) double par_smem[];
  double* su = par_smem;                    // [R][PLANE]   ring of masked x planes
  double* so = su + R * PLANE;              // [2][3 * OC]  corner exchange
  double* sc = so + 2 * 3 * OC;             // CSM: [4][OC] carried top-face values (parity basis) of the layer below,
  double* sb = sc + 4 * OC;                 //      [4][OC] x / y-transformed node plane under the current layer
  __shared__ double wred[3][NT / 32];

  const int tid = threadIdx.x, tx = tid % EX, ty = tid / EX;
  const int i0 = blockIdx.x * (EX - 1), j0 = blockIdx.y * (EY - 1);   // first owned node column
  const int kA = blockIdx.z * zl, kB = min(kA + zl, g.nzl);          // owned local planes [kA, kB)
  const long long xplane = (long long)g.NX * g.NY * NDOF, slayer = (long long)g.nx * g.ny;
  const unsigned char* fl = flags ? flags + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (g.nzl + 2) + 1 : nullptr;

  // ---- staging slots of this thread inside a plane (computed once): staged node (col / NDOF, row) = (i0 - 1 + ., j0 - 1 + .)
  int xoff[NQ];
  bool xok[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int p = tid + NT * q;
    const int row = p / ROW, col = p - row * ROW;
    const int i = i0 - 1 + col / NDOF, j = j0 - 1 + row;
    xok[q] = p < PLANE && i >= 0 && i < g.NX && j >= 0 && j < g.NY;
    xoff[q] = (j * g.NX + (i0 - 1)) * NDOF + col;
  }
  auto plane_ok = [&](int kl) {  // plane kl (local; -1 and nzl are the halo planes) exists and may be read
    const int k = g.kz0 + kl;
    return k >= 0 && k < g.NZ && kl <= g.nzl;
  };
  // the staged part of plane kl may hold Dirichlet dofs (CTA-uniform; without flags: whenever there is a mask)
  auto flagged = [&](int kl) -> bool { return mask && kl <= kB && plane_ok(kl) && (fl ? __ldg(fl + kl) != 0 : true); };
  auto issue_plane = [&](int kl) {  // plane kl -> ring slot (kl - kA + R) % R (cp.async, zero-fill outside the grid)
    const bool pok = plane_ok(kl) && kl <= kB;
    const double* xp = x + (long long)kl * xplane;
    double* dst = su + ((kl - kA + R) % R) * PLANE;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int p = tid + NT * q;
      const bool ok = pok && xok[q];
      if (p < PLANE) cp_async8(dst + p, ok ? xp + xoff[q] : x, ok);
    }
  };
  auto mask_bytes = [&](int kl) -> unsigned {  // Dirichlet flags of this thread's staging slots in plane kl, one per byte
    unsigned m = 0;
    const unsigned char* mp = mask + (long long)kl * xplane;
#pragma unroll
    for (int q = 0; q < NQ; ++q)
      if (xok[q]) m |= (unsigned)(__ldg(mp + xoff[q]) != 0) << (8 * q);
    return m;
  };
  auto mask_plane = [&](int kl, unsigned m) {
    double* dst = su + ((kl - kA + R) % R) * PLANE;
#pragma unroll
    for (int q = 0; q < NQ; ++q)
      if (m & (1u << (8 * q))) dst[tid + NT * q] = 0.0;
  };

  // ---- this thread's element column and the node column it owns
  const int ei = i0 - 1 + tx, ej = j0 - 1 + ty;
  const bool eok = ei >= 0 && ei < g.nx && ej >= 0 && ej < g.ny;
  const long long soff = eok ? (long long)ej * g.nx + ei : 0;
  auto density = [&](int el) -> double {  // layers above the last owned node plane belong to the next rank: not needed
    const int ek = g.kz0 + el;
    return (eok && ek >= 0 && ek < g.nzE && el < g.nzl) ? __ldg(s + (long long)el * slayer + soff) : 0.0;
  };
  const int ni = i0 + tx, nj = j0 + ty;
  const bool owner = tx < EX - 1 && ty < EY - 1 && ni < g.NX && nj < g.NY;
  const long long rrow = owner ? ((long long)nj * g.NX + ni) * NDOF : 0;
  const int c00 = ty * ROW + tx * NDOF;   // corner (0, 0) of the element inside a staged plane

  // x / y butterflies of the 4 corner nodes of a staged plane: t[px + 2 py][c]; xc = the (masked) value of the owned node
  auto plane_xy = [&](const double* pl, double (&t)[4][NDOF], double (&xc)[NDOF]) {
#pragma unroll
    for (int c = 0; c < NDOF; ++c) {
      const double a00 = pl[c00 + c], a10 = pl[c00 + NDOF + c], a01 = pl[c00 + ROW + c], a11 = pl[c00 + ROW + NDOF + c];
      const double sx0 = a00 + a10, dx0 = a10 - a00, sx1 = a01 + a11, dx1 = a11 - a01;
      t[0][c] = sx0 + sx1;
      t[1][c] = dx0 + dx1;
      t[2][c] = sx1 - sx0;
      t[3][c] = dx1 - dx0;
      xc[c] = a11;
    }
  };

  // ---- prime: planes kA - 1 .. kA + D - 1 in flight (kA - 1 and kA as the first group), carry = 0
  issue_plane(kA - 1);
  issue_plane(kA);
  cp_async_commit();
#pragma unroll
  for (int d = 1; d < D; ++d) {
    issue_plane(kA + d);
    cp_async_commit();
  }
  // Dirichlet flags of the planes: f0 = plane el (epilogue rows), f3 = plane el + 3 (its mask bytes are fetched one step
  // before it lands); mnext = mask bytes of the plane that lands at the end of the NEXT step
  bool f0 = flagged(kA - 1), f1 = flagged(kA), f2 = flagged(kA + 1), f3 = flagged(kA + 2);
  const unsigned mA = f0 ? mask_bytes(kA - 1) : 0u, mB = f1 ? mask_bytes(kA) : 0u;
  unsigned mnext = f2 ? mask_bytes(kA + 1) : 0u;
  double s_cur = density(kA - 1);
  double carry[4][NDOF];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < NDOF; ++c) {
      carry[p][c] = 0.0;
      if (CSM) sc[(p * NDOF + c) * NT + tid] = 0.0;
    }
  cp_async_wait_pending<D - 1>();
  mask_plane(kA - 1, mA);
  mask_plane(kA, mB);
  __syncthreads();
  double bt[4][NDOF], xprev[NDOF];
  plane_xy(su + (R - 1) * PLANE, bt, xprev);
  if (CSM) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int c = 0; c < NDOF; ++c) sb[(p * NDOF + c) * NT + tid] = bt[p][c];
  }
  __syncthreads();  // slot R - 1 is refilled in the first step

  double d0 = 0.0, d1 = 0.0, d2 = 0.0;
  for (int el = kA - 1, t = 0; el < kB; ++el, ++t) {
    // ---- in flight during the arithmetic: the plane D steps ahead, next density, the mask bytes of the plane after next,
    //      this step's epilogue operands
    issue_plane(el + 1 + D);
    cp_async_commit();
    const unsigned mland = mnext;                    // plane el + 2 lands at the end of this step
    mnext = f3 ? mask_bytes(el + 3) : 0u;
    const bool f4 = flagged(el + 4);
    const double s_next = density(el + 1);
    const bool emit = owner && t > 0;                // plane el >= kA is owned by this CTA
    const long long r0 = (long long)el * xplane + rrow;
    bool mr[NDOF];
#pragma unroll
    for (int c = 0; c < NDOF; ++c) mr[c] = emit && f0 && __ldg(mask + r0 + c);
    if (emit) {
      if (MODE != EMODE_SPMV) prefetch_l1(b + r0), prefetch_l1(b + r0 + NDOF - 1);
      if (MODE == EMODE_JACOBI) prefetch_l1(diag + r0), prefetch_l1(diag + r0 + NDOF - 1);
      if (partials && dotv) prefetch_l1(dotv + r0), prefetch_l1(dotv + r0 + NDOF - 1);
    }

    // ---- forward: top plane x / y butterflies, z butterfly with the carried bottom plane
    double tt[4][NDOF], xtop[NDOF];
    plane_xy(su + (t % R) * PLANE, tt, xtop);
    double uh[8][NDOF];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int c = 0; c < NDOF; ++c) {
        double* bp = sb + (p * NDOF + c) * NT + tid;
        const double bot = CSM ? *bp : bt[p][c];
        uh[p][c] = bot + tt[p][c];
        uh[p + 4][c] = tt[p][c] - bot;
        if (CSM) *bp = tt[p][c];
        else bt[p][c] = tt[p][c];
      }
    // ---- blocks: the dofs of representation q are (par_member(q, c), c).  The block entries are read from the parameter
    //      bank INSIDE the loop (the offset `zero * t`, zero = 0 from the host, defeats loop-invariant hoisting: hoisted, the 72 values do not fit
    //      the uniform register file and end up in local memory -- measured 1.7 GB of spill traffic per launch)
    const double* kbp = &kb.v[0][0][0] + zero * t;
    double vh[8][NDOF];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
#pragma unroll
      for (int c = 0; c < NDOF; ++c) {
        double acc = kbp[(q * NDOF + c) * NDOF] * uh[par_member<NDOF>(q, 0)][0];
#pragma unroll
        for (int c2 = 1; c2 < NDOF; ++c2) acc = fma(kbp[(q * NDOF + c) * NDOF + c2], uh[par_member<NDOF>(q, c2)][c2], acc);
        vh[par_member<NDOF>(q, c)][c] = s_cur * acc;
      }
    }
    // ---- transposed z butterfly; bottom face + carried top face of the layer below; transposed y / x butterflies
    double own[NDOF];   // contribution to the owned node = corner (1, 1)
    double* ob = so + (t & 1) * 3 * OC;
#pragma unroll
    for (int c = 0; c < NDOF; ++c) {
      double wv[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        double* cp = sc + (p * NDOF + c) * NT + tid;
        wv[p] = (vh[p][c] - vh[p + 4][c]) + (CSM ? *cp : carry[p][c]);
        if (CSM) *cp = vh[p][c] + vh[p + 4][c];
        else carry[p][c] = vh[p][c] + vh[p + 4][c];
      }
      const double r0s = wv[0] - wv[2], r0d = wv[1] - wv[3];   // row dy = 0: px = 0, 1
      const double r1s = wv[0] + wv[2], r1d = wv[1] + wv[3];   // row dy = 1
      ob[0 * OC + c * NT + tid] = r0s - r0d;                   // corner (dx, dy) = (0, 0)
      ob[1 * OC + c * NT + tid] = r0s + r0d;                   // (1, 0)
      ob[2 * OC + c * NT + tid] = r1s - r1d;                   // (0, 1)
      own[c] = r1s + r1d;
    }
    cp_async_wait_pending<D - 1>();   // plane el + 2 has landed (this thread's part)
    mask_plane(el + 2, mland);
    __syncthreads();

    // ---- node (ni, nj) of plane el: own corner (1, 1) + corner (0, 1) of column (tx + 1, ty) + corner (1, 0) of
    //      (tx, ty + 1) + corner (0, 0) of (tx + 1, ty + 1)
    if (emit) {
#pragma unroll
      for (int c = 0; c < NDOF; ++c) {
        const double acc = ((own[c] + ob[2 * OC + c * NT + tid + 1]) + ob[1 * OC + c * NT + tid + EX]) + ob[0 * OC + c * NT + tid + EX + 1];
        const long long r = r0 + c;
        const double xr = mr[c] ? __ldg(x + r) : xprev[c];
        const double ax = mr[c] ? bcdiag * xr : acc;
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
#pragma unroll
    for (int c = 0; c < NDOF; ++c) xprev[c] = xtop[c];   // (read before the barrier: the slot is refilled in the next step)
    s_cur = s_next;
    f0 = f1, f1 = f2, f2 = f3, f3 = f4;
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

// launch configurations of the parity-block layout: variant 8 + i
struct ParLaunchCfg {
  int ey, minb;
  bool csm;
};
constexpr int PAR_NCFG = 4;
constexpr ParLaunchCfg PAR_CFG[PAR_NCFG] = {{8, 2, true}, {12, 1, false}, {6, 2, false}, {4, 3, false}};

// planes per CTA: few CTAs lost to the last wave, little redundant layer work (every CTA computes one extra layer)
static int par_zl(const Geo& g, int ey, int ctas_per_sm, int sms) {
  const long long tiles = (long long)((g.NX + PAR_EX - 2) / (PAR_EX - 1)) * ((g.NY + ey - 2) / (ey - 1));
  const long long slots = (long long)ctas_per_sm * sms;
  int best = g.nzl;
  double best_cost = 1e300;
  for (int chunks = 1; chunks <= g.nzl; ++chunks) {
    const int zl = (g.nzl + chunks - 1) / chunks;
    const long long ctas = tiles * ((g.nzl + zl - 1) / zl);
    const long long waves = (ctas + slots - 1) / slots;
    const double cost = (double)waves * (zl + 1.5);   // time ~ waves x layers per CTA (+ prologue)
    if (cost < best_cost - 1e-12) best_cost = cost, best = zl;
  }
  return best > 0 ? best : 1;
}
