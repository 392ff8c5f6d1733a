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
struct alignas(16) ParBlocks {
  static constexpr int NB = NDOF * (NDOF + 1) / 2;   // upper triangle of a symmetric block: (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
  double v[8 * NB];
};
template <int NDOF>
__host__ __device__ constexpr int par_tri(int c, int c2) {   // index of entry (c, c2) = (c2, c) inside a packed block
  const int lo = c < c2 ? c : c2, hi = c < c2 ? c2 : c;
  return lo * NDOF - lo * (lo - 1) / 2 + (hi - lo);
}

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
// couples different representations (no reflection symmetry: anisotropic material, distorted element, arbitrary matrix)
// or is not symmetric (the kernel keeps the upper triangles of the blocks).
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
  for (int i = 0; i < LD; ++i)
    for (int j = 0; j < LD; ++j) {
      const double v = fabs(A[i * LD + j]) / 64.0;
      if (!(v == v)) return false;
      amax = v > amax ? v : amax;
      const double asym = fabs(A[i * LD + j] - A[j * LD + i]) / 64.0;
      const bool same = par_rep<NDOF>(i / NDOF, i % NDOF) == par_rep<NDOF>(j / NDOF, j % NDOF);
      const double bad = same ? asym : v;
      off = bad > off ? bad : off;
    }
  if (off > 1e-13 * amax) return false;
  for (int q = 0; q < 8; ++q)
    for (int c = 0; c < NDOF; ++c)
      for (int c2 = c; c2 < NDOF; ++c2) {
        const int i = par_member<NDOF>(q, c) * NDOF + c, j = par_member<NDOF>(q, c2) * NDOF + c2;
        kb.v[q * ParBlocks<NDOF>::NB + par_tri<NDOF>(c, c2)] = 0.5 * (A[i * LD + j] + A[j * LD + i]) / 64.0;
      }
  return true;
}

#ifndef PMB_PAR_DBG
#define PMB_PAR_DBG 0   // timing experiments only (wrong results): 1 = no CTA barrier per layer, 2 = no wait for the staged plane
#endif
constexpr int PAR_EX = 32, PAR_RING = 3;   // element columns per CTA row; slots of the node-plane ring (2 planes in flight)
constexpr int PAR_ZL_MAX = 60;             // planes per CTA (the Dirichlet flags of its planes travel as one 64-bit mask)
template <int NDOF, int EY>
struct ParCfg {
  static constexpr int NT = PAR_EX * EY;
  static constexpr int ROW = (PAR_EX + 1) * NDOF;            // doubles of a staged node row (33 nodes)
  static constexpr int PITCH = (ROW + 2 + 1) / 2 * 2;        // + 16-byte hull, even: every row starts 16-byte aligned
  static constexpr int PLANE = (EY + 1) * PITCH;
  static constexpr int OC = NDOF * NT;                       // one per-thread array: [c][ty][tx]
  // plane ring + 2 corner-exchange buffers (3 corners) + the values carried from layer to layer (top-face result and
  // x / y-transformed top plane, 4 parity patterns each)
  static constexpr size_t SMEM = sizeof(double) * (PAR_RING * PLANE + (2 * 3 + 8) * OC);
};

// n / d for a normal, finite d by Newton iterations on the hardware reciprocal seed: the compiler's IEEE division spends
// ~35 instructions per quotient on range checks and a slow path that a matrix diagonal never takes (last-bit differences
// against '/' are possible; this layout agrees with the others to rounding anyway)
__device__ __forceinline__ double par_div(double n, double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  const double q = n * r;
  return fma(fma(-d, q, n), r, q);
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

// Staging: the EY + 1 node rows of a plane are 1-D TMA bulk copies (one row per lane of warp 0, each from its 16-byte
// aligned hull: a node row of 3 NX doubles starts on an odd double every other row, so the wanted first double sits at
// offset 0 or 1 of the staged row -- `shift`).  Rows / planes / columns outside the grid are not copied: their slots keep
// finite stale values (the ring starts as zeros) that only ever meet the density of an out-of-grid element, which is 0.
template <int NDOF, int MODE, int EY, int MINB>
__global__ void __launch_bounds__(PAR_EX* EY, MINB)
    elem_kernel_par(Geo g, const __grid_constant__ ParBlocks<NDOF> kb, int zl, const double* __restrict__ s,
                    const unsigned char* __restrict__ mask, const unsigned char* __restrict__ flags, double bcdiag,
                    const double* __restrict__ x, const double* __restrict__ b, const double* __restrict__ diag, double w,
                    double* __restrict__ y, const double* __restrict__ dotv, double* __restrict__ partials) {
  using C = ParCfg<NDOF, EY>;
  constexpr int EX = PAR_EX, NT = C::NT, R = PAR_RING, ROW = C::ROW, PITCH = C::PITCH, PLANE = C::PLANE, OC = C::OC;
  constexpr int NB = ParBlocks<NDOF>::NB;
  extern __shared__ __align__(128) double par_smem[];
  double* su = par_smem;                    // [R][EY + 1][PITCH] ring of x planes
  double* so = su + R * PLANE;              // [2][3 * OC]  corner exchange
  // values carried from layer to layer, as double2 (parity patterns px = 0 / 1 of one (py, c)) -> 128-bit shared accesses:
  double2* sc = reinterpret_cast<double2*>(so + 2 * 3 * OC);   // [2 NDOF][NT] top-face result of the layer below (parity basis)
  double2* sb = sc + 2 * OC;                                    // [2 NDOF][NT] x / y-transformed node plane under the layer
  // The packed blocks stay in the kernel-parameter constant bank and reach the FP64 pipe through uniform registers (SASS: LDCU.64
  // + DFMA R, R, UR): no shared-memory instruction.  Read with plain indices the compiler hoists the 48 values out of the layer
  // loop, they do not fit the uniform registers and end up in local memory (1.7 GB of traffic per launch, measured); the index
  // used below carries a term of the loop counter that is always 0 but not provably so, which keeps the loads inside the loop.
  // (First kept in shared memory and read as broadcast LDS.128: 24 more shared-memory instructions per thread and layer on a
  // kernel whose top stall reasons are the shared-memory queue and the barrier -- 0.253 -> 0.225 ms at 256x128x128, ndof 3;
  // carrying the per-thread values in registers instead of shared memory on top of that gains nothing: at 128 registers it
  // spills, and with 12 or 8 warps per SM the kernel is 20 - 30 % slower.)
  const double2* ckb = reinterpret_cast<const double2*>(kb.v);
  __shared__ __align__(8) uint64_t full_bar[R];
  __shared__ double wred[3][NT / 32];

  const int tid = threadIdx.x, tx = tid % EX, ty = tid / EX;
  const int i0 = blockIdx.x * (EX - 1), j0 = blockIdx.y * (EY - 1);   // first owned node column
  const int kA = blockIdx.z * zl, kB = min(kA + zl, g.nzl);          // owned local planes [kA, kB)
  const long long xrow = (long long)g.NX * NDOF, xplane = xrow * g.NY, slayer = (long long)g.nx * g.ny;
  const long long Dx = (long long)(reinterpret_cast<uintptr_t>(x) >> 3);

  for (int p = tid; p < R * PLANE; p += NT) su[p] = 0.0;   // stale-but-finite contract
  if (tid == 0) {
    for (int q = 0; q < R; ++q) mbar_init(&full_bar[q], 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zeros before the async-proxy copies
  __syncthreads();

  // plane m (counted from local plane kA - 1) lives in ring slot m % R; absolute double index of its wanted first double in
  // staged row r: D0 = Dx + ((kl NY + j0 - 1 + r) xrow + (i0 - 1) NDOF); the row is copied so that double D lands at
  // row[D - D0 + (D0 & 1)]
  auto plane_ok = [&](int kl) {  // plane kl (local; -1 and nzl are the halo planes) exists, may be read and is needed
    const int k = g.kz0 + kl;
    return k >= 0 && k < g.NZ && kl <= g.nzl && kl <= kB;
  };
  // ---- producer (all lanes of warp 0): lane r copies row r of plane m into slot m % R
  auto issue = [&](int m) {
    const int kl = kA - 1 + m, slot = m % R, lane = tid;
    unsigned bytes = 0;
    uintptr_t src = 0;
    double* dst = nullptr;
    const int j = j0 - 1 + lane;
    if (lane <= EY && plane_ok(kl) && j >= 0 && j < g.NY) {
      const int ia = max(i0 - 1, 0), ib = min(i0 + EX, g.NX);
      const long long rowb = ((long long)kl * g.NY + j) * xrow;
      const long long D0 = Dx + rowb + (long long)(i0 - 1) * NDOF;
      const long long lo = (Dx + rowb + (long long)ia * NDOF) & ~1LL, hi = (Dx + rowb + (long long)ib * NDOF + 1) & ~1LL;
      src = (uintptr_t)lo << 3;
      dst = su + slot * PLANE + lane * PITCH + (int)(lo - D0 + (D0 & 1));
      bytes = (unsigned)(hi - lo) * 8u;
    }
    if (bytes) {
      mbar_expect_tx(&full_bar[slot], bytes);  // arrive + expect: my bytes are announced before they can complete
      tma_load_1d(dst, reinterpret_cast<const void*>(src), bytes, &full_bar[slot], false);
    } else {
      mbar_arrive(&full_bar[slot]);
    }
  };
  // zero the Dirichlet entries of staged plane m (rare: only CTAs / planes that carry constrained dofs)
  auto mask_plane = [&](int m) {
    const int kl = kA - 1 + m;
    double* pl = su + (m % R) * PLANE;
    for (int p = tid; p < (EY + 1) * ROW; p += NT) {
      const int r = p / ROW, c = p - r * ROW;
      const int i = i0 - 1 + c / NDOF, j = j0 - 1 + r;
      if (i >= 0 && i < g.NX && j >= 0 && j < g.NY) {
        const long long rowb = ((long long)kl * g.NY + j) * xrow + (long long)(i0 - 1) * NDOF;
        if (__ldg(mask + rowb + c)) pl[r * PITCH + c + (int)((Dx + rowb) & 1)] = 0.0;
      }
    }
  };

  if (tid < 32)
    for (int m = 0; m < R; ++m) issue(m);

  // ---- Dirichlet flags of the planes m = 0 .. kB - kA + 1 of this CTA as one bit mask (bit m: the staged part of plane m
  //      may hold constrained dofs; without a flag array: whenever there is a mask)
  unsigned long long fmask = 0;
  if (mask) {
    const unsigned char* fl = flags ? flags + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (g.nzl + 2) + 1 : nullptr;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = (tid & 31) + 32 * h, kl = kA - 1 + m;
      const bool f = plane_ok(kl) && (fl ? __ldg(fl + kl) != 0 : true);
      fmask |= (unsigned long long)__ballot_sync(0xffffffffu, f) << (32 * h);
    }
  }

  // ---- this thread's element column and the node column it owns
  const int ei = i0 - 1 + tx, ej = j0 - 1 + ty;
  const bool eok = ei >= 0 && ei < g.nx && ej >= 0 && ej < g.ny;
  // element layers [elo, ehi) exist and are needed (layers above the last owned node plane belong to the next rank)
  const int elo = max(kA - 1, -g.kz0), ehi = min(min(kB, g.nzl), g.nzE - g.kz0);
  const double* sp = s + ((long long)(kA - 1) * slayer + (eok ? (long long)ej * g.nx + ei : 0));   // density of layer el
  const int ni = i0 + tx, nj = j0 + ty;
  const bool owner = tx < EX - 1 && ty < EY - 1 && ni < g.NX && nj < g.NY;
  long long r0 = (long long)(kA - 1) * xplane + (owner ? ((long long)nj * g.NX + ni) * NDOF : 0);   // first dof of the node in plane el
  // shift of staged rows ty / ty + 1 in plane m = 0, and how it changes from plane to plane / row to row
  const int pj = (int)(xrow & 1), pk = (int)(xplane & 1);
  int sh0 = (int)((Dx + ((long long)(kA - 1) * g.NY + (j0 - 1 + ty)) * xrow + (long long)(i0 - 1) * NDOF) & 1);
  const int c00 = ty * PITCH + tx * NDOF;   // corner (0, 0) of the element inside a staged plane (shift excluded)

  // x / y butterflies of the 4 corner nodes of a staged plane: t[px + 2 py][c]; xc = the value of the owned node
  auto plane_xy = [&](const double* pl, int shift0, double (&t)[4][NDOF], double (&xc)[NDOF]) {
    const double* p0 = pl + c00 + shift0;
    const double* p1 = pl + c00 + PITCH + (shift0 ^ pj);
#pragma unroll
    for (int c = 0; c < NDOF; ++c) {
      const double a00 = p0[c], a10 = p0[NDOF + c], a01 = p1[c], a11 = p1[NDOF + c];
      const double sx0 = a00 + a10, dx0 = a10 - a00, sx1 = a01 + a11, dx1 = a11 - a01;
      t[0][c] = sx0 + sx1;
      t[1][c] = dx0 + dx1;
      t[2][c] = sx1 - sx0;
      t[3][c] = dx1 - dx0;
      xc[c] = a11;
    }
  };

  // ---- prime: planes m = 0, 1 landed and masked; carry = 0; the transformed plane m = 0 is the first "bottom" plane
  double s_cur = (eok && kA - 1 >= elo && kA - 1 < ehi) ? __ldg(sp) : 0.0;
  mbar_wait(&full_bar[0], 0u);
  mbar_wait(&full_bar[1], 0u);
  if (fmask & 1ull) mask_plane(0);
  if (fmask & 2ull) mask_plane(1);
  __syncthreads();
  double xprev[NDOF];
  {
    double bt[4][NDOF];
    plane_xy(su, sh0, bt, xprev);
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int c = 0; c < NDOF; ++c) {
        sb[(py * NDOF + c) * NT + tid] = make_double2(bt[2 * py][c], bt[2 * py + 1][c]);
        sc[(py * NDOF + c) * NT + tid] = make_double2(0.0, 0.0);
      }
  }
  __syncthreads();  // slot 0 is refilled in the first step

  double d0 = 0.0, d1 = 0.0, d2 = 0.0;
  int slot = 1;                                       // ring slot of the top plane m = t + 1
  const int nsteps = kB - kA + 1;
  for (int t = 0; t < nsteps; ++t) {
    const int el = kA - 1 + t;
    // ---- in flight during the arithmetic: plane m = t + 3 (into the slot of the bottom plane m = t, last read before the
    //      previous barrier), next density, this step's epilogue operands (L1 prefetch)
    if (tid < 32 && t + 3 <= nsteps) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads / patches before the async-proxy refill
      issue(t + 3);
    }
    sp += slayer;
    const double s_next = (eok && el + 1 >= elo && el + 1 < ehi) ? __ldg(sp) : 0.0;
    const bool emit = owner && t > 0;                // plane el >= kA is owned by this CTA
    sh0 ^= pk;                                       // shift of row ty in the top plane
    if (emit) {
      if (MODE != EMODE_SPMV) prefetch_l1(b + r0), prefetch_l1(b + r0 + NDOF - 1);
      if (MODE == EMODE_JACOBI) prefetch_l1(diag + r0), prefetch_l1(diag + r0 + NDOF - 1);
      if (partials && dotv) prefetch_l1(dotv + r0), prefetch_l1(dotv + r0 + NDOF - 1);
    }

    // ---- forward: top plane x / y butterflies, z butterfly with the carried bottom plane
    double tt[4][NDOF], xtop[NDOF];
    plane_xy(su + slot * PLANE, sh0, tt, xtop);
    double uh[8][NDOF];
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int c = 0; c < NDOF; ++c) {
        double2* bp = sb + (py * NDOF + c) * NT + tid;
        const double2 bot = *bp;
        uh[2 * py][c] = bot.x + tt[2 * py][c];
        uh[2 * py + 4][c] = tt[2 * py][c] - bot.x;
        uh[2 * py + 1][c] = bot.y + tt[2 * py + 1][c];
        uh[2 * py + 5][c] = tt[2 * py + 1][c] - bot.y;
        *bp = make_double2(tt[2 * py][c], tt[2 * py + 1][c]);
      }
    // ---- blocks: the dofs of representation q are (par_member(q, c), c)
    double vh[8][NDOF];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      double kv[NB + 1];
#pragma unroll
      for (int i = 0; i < (NB + 1) / 2; ++i) {
        const double2 k2 = ckb[(q * NB) / 2 + i + (t >> 28)];   // NB even (ndof 3) or q * NB read pairwise (ndof 1, see below)
        kv[2 * i] = k2.x;
        kv[2 * i + 1] = k2.y;
      }
      const int k0 = (NB & 1) ? (q & 1) : 0;           // ndof 1: block q is the (q & 1)-th double of pair q / 2
#pragma unroll
      for (int c = 0; c < NDOF; ++c) {
        double acc = kv[k0 + par_tri<NDOF>(c, 0)] * uh[par_member<NDOF>(q, 0)][0];
#pragma unroll
        for (int c2 = 1; c2 < NDOF; ++c2) acc = fma(kv[k0 + par_tri<NDOF>(c, c2)], uh[par_member<NDOF>(q, c2)][c2], acc);
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
      for (int py = 0; py < 2; ++py) {
        double2* cp = sc + (py * NDOF + c) * NT + tid;
        const double2 cv = *cp;
        wv[2 * py] = (vh[2 * py][c] - vh[2 * py + 4][c]) + cv.x;
        wv[2 * py + 1] = (vh[2 * py + 1][c] - vh[2 * py + 5][c]) + cv.y;
        *cp = make_double2(vh[2 * py][c] + vh[2 * py + 4][c], vh[2 * py + 1][c] + vh[2 * py + 5][c]);
      }
      const double r0s = wv[0] - wv[2], r0d = wv[1] - wv[3];   // row dy = 0: px = 0, 1
      const double r1s = wv[0] + wv[2], r1d = wv[1] + wv[3];   // row dy = 1
      ob[0 * OC + c * NT + tid] = r0s - r0d;                   // corner (dx, dy) = (0, 0)
      ob[1 * OC + c * NT + tid] = r0s + r0d;                   // (1, 0)
      ob[2 * OC + c * NT + tid] = r1s - r1d;                   // (0, 1)
      own[c] = r1s + r1d;
    }
    // ---- epilogue operands travel (from L1) while the CTA meets at the barrier
    double br[NDOF], dr[NDOF], dvr[NDOF];
#pragma unroll
    for (int c = 0; c < NDOF; ++c) {
      br[c] = (MODE != EMODE_SPMV && emit) ? __ldg(b + r0 + c) : 0.0;
      dr[c] = (MODE == EMODE_JACOBI && emit) ? __ldg(diag + r0 + c) : 1.0;
      dvr[c] = (partials && dotv && emit) ? __ldg(dotv + r0 + c) : 0.0;
    }
    if (t + 1 < nsteps) {   // plane m = t + 2 is the next top plane
      if (!(PMB_PAR_DBG & 2)) mbar_wait(&full_bar[(t + 2) % R], (unsigned)(((t + 2) / R) & 1));
      if (fmask & 4ull) mask_plane(t + 2);
    }
    if (!(PMB_PAR_DBG & 1)) __syncthreads();

    // ---- node (ni, nj) of plane el: own corner (1, 1) + corner (0, 1) of column (tx + 1, ty) + corner (1, 0) of
    //      (tx, ty + 1) + corner (0, 0) of (tx + 1, ty + 1)
    if (emit) {
      double ax[NDOF], xr[NDOF];
#pragma unroll
      for (int c = 0; c < NDOF; ++c) {
        ax[c] = ((own[c] + ob[2 * OC + c * NT + tid + 1]) + ob[1 * OC + c * NT + tid + EX]) + ob[0 * OC + c * NT + tid + EX + 1];
        xr[c] = xprev[c];
      }
      if (fmask & 1ull) {   // the plane may hold Dirichlet rows (rare): their staged value was zeroed, the row is bcdiag * x
#pragma unroll
        for (int c = 0; c < NDOF; ++c)
          if (__ldg(mask + r0 + c)) {
            xr[c] = __ldg(x + r0 + c);
            ax[c] = bcdiag * xr[c];
          }
      }
#pragma unroll
      for (int c = 0; c < NDOF; ++c) {
        double out;
        if (MODE == EMODE_SPMV) out = ax[c];
        else if (MODE == EMODE_RESID) out = br[c] - ax[c];
        else out = xr[c] + w * par_div(br[c] - ax[c], dr[c]);
        y[r0 + c] = out;
        if (partials) {
          d0 = fma(out, xr[c], d0);
          d1 = fma(xr[c], dvr[c], d1);
          d2 = fma(out, dvr[c], d2);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NDOF; ++c) xprev[c] = xtop[c];   // (read before the barrier: the slot is refilled in the next step)
    s_cur = s_next;
    fmask >>= 1;
    r0 += xplane;
    slot = slot == R - 1 ? 0 : slot + 1;
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
};
// (measured at 256x128x128, ndof 3: 32 x 8 columns, 2 CTAs / SM 0.254 ms; 32 x 16, 1 CTA / SM 0.254 ms; 12 / 10 warps per
//  SM -- 32 x 12 or 32 x 10 columns in one CTA -- 0.30 ms: the kernel is latency bound and wants every warp it can get)
constexpr int PAR_NCFG = 2;
constexpr ParLaunchCfg PAR_CFG[PAR_NCFG] = {{8, 2}, {16, 1}};

// planes per CTA: few CTAs lost to the last wave, little redundant layer work (every CTA computes one extra layer)
static int par_zl(const Geo& g, int ey, int ctas_per_sm, int sms) {
  const long long tiles = (long long)((g.NX + PAR_EX - 2) / (PAR_EX - 1)) * ((g.NY + ey - 2) / (ey - 1));
  const long long slots = (long long)ctas_per_sm * sms;
  int best = g.nzl;
  double best_cost = 1e300;
  for (int chunks = (g.nzl + PAR_ZL_MAX - 1) / PAR_ZL_MAX; chunks <= g.nzl; ++chunks) {
    const int zl = (g.nzl + chunks - 1) / chunks;
    const long long ctas = tiles * ((g.nzl + zl - 1) / zl);
    const long long waves = (ctas + slots - 1) / slots;
    const double cost = (double)waves * (zl + 1.5);   // time ~ waves x layers per CTA (+ prologue)
    if (cost < best_cost - 1e-12) best_cost = cost, best = zl;
  }
  return best > 0 ? best : 1;
}
