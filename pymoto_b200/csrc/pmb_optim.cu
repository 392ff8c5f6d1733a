// libpmb: MMA design update on the device and the FP32 packer of the VTI writer (8f row 4).
//
// MMA (pymoto/common/mma.py): every n-sized expression of mmasub / subsolv runs here -- one fused pass per phase of the
// primal-dual Newton iteration with deterministic two-stage reductions -- while the m-sized unknowns (y, z, lam, mu, zet,
// s) and the (m+1)x(m+1) Newton system stay on the host like in the reference.  Per Newton iteration: one "sums" pass,
// one "direction" pass and one line-search pass per trial step; the host reads back <= 2 + 2m + m^2 doubles per pass.
// All passes are HBM-bound streams over 9-17 vectors of n doubles.
#include "pmb_common.cuh"
#include "pmb_mma_math.h"

static constexpr int OPT_BLOCKS = 592;  // 4 CTAs per SM on 148 SMs
static constexpr int OPT_THREADS = 256;
static constexpr int OPT_MAXRED = 2 * PMB_MMA_MAXM + PMB_MMA_MAXM * PMB_MMA_MAXM;   // sums + maxima per pass (Newton sums: 2m + m^2)

extern "C" long long pmb_mma_ws_doubles(void) { return 8 + (long long)OPT_MAXRED * OPT_BLOCKS; }

// NS sums followed by NX maxima: block stage, then the last CTA combines the per-CTA partials in index order
// (run-to-run deterministic).  ws: [0] ticket counter (zero-initialised by the caller once, re-armed here), [8..] partials.
template <int NS, int NX>
__device__ __forceinline__ void opt_reduce(double (&s)[NS], double (&mx)[NX > 0 ? NX : 1], double* ws, double* out) {
  constexpr int NR = NS + NX;
  static_assert(NR <= OPT_MAXRED, "too many reductions in one pass");
  __shared__ double sm[NR][OPT_THREADS / 32];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NS; ++q)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
#pragma unroll
  for (int q = 0; q < NX; ++q)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx[q] = fmax(mx[q], __shfl_xor_sync(0xffffffffu, mx[q], o));
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NS; ++q) sm[q][wid] = s[q];
#pragma unroll
    for (int q = 0; q < NX; ++q) sm[NS + q][wid] = mx[q];
  }
  __syncthreads();
  unsigned* counter = reinterpret_cast<unsigned*>(ws);
  double* partials = ws + 8;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NR; ++q) {
      double a = sm[q][0];
      for (int w = 1; w < OPT_THREADS / 32; ++w) a = q < NS ? a + sm[q][w] : fmax(a, sm[q][w]);
      partials[q * OPT_BLOCKS + blockIdx.x] = a;
    }
    __threadfence();
    last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    if (threadIdx.x < NR) {  // one thread per reduced quantity, partials combined in CTA order
      const int q = threadIdx.x;
      double a = __ldcg(partials + q * OPT_BLOCKS);
      for (int b = 1; b < (int)gridDim.x; ++b) {
        const double p = __ldcg(partials + q * OPT_BLOCKS + b);
        a = q < NS ? a + p : fmax(a, p);
      }
      out[q] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) *counter = 0u;
  }
}

static int opt_blocks(long long n) {
  long long b = (n + OPT_THREADS - 1) / OPT_THREADS;
  return (int)(b < 1 ? 1 : (b > OPT_BLOCKS ? OPT_BLOCKS : b));
}

__device__ __forceinline__ double bound_at(const pmb_bound& b, long long j) { return b.v ? b.v[j] : b.s; }

template <int M>
__device__ __forceinline__ void load_var(const pmb_mma_vecs& a, long long n, long long j, MmaVar& v) {
  v.x = a.x[j]; v.xsi = a.xsi[j]; v.eta = a.eta[j];
  v.low = a.low[j]; v.upp = a.upp[j]; v.alfa = a.alfa[j]; v.beta = a.beta[j];
#pragma unroll
  for (int i = 0; i <= M; ++i) {
    v.P[i] = a.P[i * n + j];
    v.Q[i] = a.Q[i * n + j];
  }
}

// ------------------------------------------------------------------------------------------------- asymptote offsets
__global__ void __launch_bounds__(OPT_THREADS) mma_asymptotes_kernel(long long n, const double* __restrict__ x, const double* __restrict__ xold1,
                                                                     const double* __restrict__ xold2, double asyincr, double asydecr,
                                                                     double asybound, double* __restrict__ offset) {
  const long long stride = (long long)gridDim.x * OPT_THREADS;
  for (long long j = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; j < n; j += stride)
    offset[j] = mma_offset_update(offset[j], x[j], xold1[j], xold2[j], asyincr, asydecr, asybound);
}

extern "C" int pmb_mma_asymptotes(long long n, const double* x, const double* xold1, const double* xold2, double asyincr,
                                  double asydecr, double asybound, double* offset, void* stream) {
  PMB_REQUIRE(x && xold1 && xold2 && offset, "pmb_mma_asymptotes: NULL pointer argument");
  if (n <= 0) return 0;
  mma_asymptotes_kernel<<<opt_blocks(n), OPT_THREADS, 0, (cudaStream_t)stream>>>(n, x, xold1, xold2, asyincr, asydecr, asybound, offset);
  PMB_CHECK_LAUNCH("pmb_mma_asymptotes");
  return 0;
}

// ------------------------------------------------------------------------------------------------- subproblem set-up
struct MmaRows {
  const double* r[PMB_MMA_MAXM + 1];
};

template <int M>
__global__ void __launch_bounds__(OPT_THREADS) mma_setup_kernel(long long n, const double* __restrict__ xval, MmaRows dg,
                                                                const double* __restrict__ offset, pmb_bound xmin, pmb_bound xmax,
                                                                pmb_bound move, double albefa, MmaSmall rho, int version, pmb_mma_vecs a,
                                                                double* out, double* ws) {
  double s[M + 1], mx[1] = {0.0};
#pragma unroll
  for (int i = 0; i <= M; ++i) s[i] = 0.0;
  const long long stride = (long long)gridDim.x * OPT_THREADS;
  for (long long j = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; j < n; j += stride) {
    double dgj[M + 1];
#pragma unroll
    for (int i = 0; i <= M; ++i) dgj[i] = dg.r[i][j];
    MmaVar v;
    const double sinv = mma_setup_pt<M>(xval[j], dgj, offset[j], bound_at(xmin, j), bound_at(xmax, j), bound_at(move, j), albefa,
                                        rho, version, v);
    a.x[j] = v.x; a.xsi[j] = v.xsi; a.eta[j] = v.eta;
    a.low[j] = v.low; a.upp[j] = v.upp; a.alfa[j] = v.alfa; a.beta[j] = v.beta;
#pragma unroll
    for (int i = 0; i <= M; ++i) {
      a.P[i * n + j] = v.P[i];
      a.Q[i * n + j] = v.Q[i];
      s[i] += v.P[i] * sinv + v.Q[i] * sinv;
    }
  }
  opt_reduce<M + 1, 0>(s, mx, ws, out);
}

static int check_vecs(const pmb_mma_vecs* v, const char* who) {
  if (!v) return pmb_set_error("%s: vecs is NULL", who);
  if (!(v->x && v->xsi && v->eta && v->xo && v->xsio && v->etao && v->dx && v->dxsi && v->deta && v->low && v->upp && v->alfa &&
        v->beta && v->P && v->Q))
    return pmb_set_error("%s: NULL vector in pmb_mma_vecs", who);
  return 0;
}

#define MMA_DISPATCH(m, CALL)                                                       \
  switch (m) {                                                                      \
    case 1: { constexpr int M = 1; CALL; } break;                                   \
    case 2: { constexpr int M = 2; CALL; } break;                                   \
    case 3: { constexpr int M = 3; CALL; } break;                                   \
    case 4: { constexpr int M = 4; CALL; } break;                                   \
    case 5: { constexpr int M = 5; CALL; } break;                                   \
    case 6: { constexpr int M = 6; CALL; } break;                                   \
    default: return pmb_set_error("pmb_mma: m=%d not in 1..%d", m, PMB_MMA_MAXM);   \
  }

static MmaSmall small_from(const double* h, int count) {
  MmaSmall s;
  for (int i = 0; i <= PMB_MMA_MAXM; ++i) s.v[i] = (h && i < count) ? h[i] : 0.0;
  return s;
}

extern "C" int pmb_mma_setup(long long n, int m, const double* xval, const double* const* dg, const double* offset, pmb_bound xmin,
                             pmb_bound xmax, pmb_bound move, double albefa, const double* rho, int version, const pmb_mma_vecs* v,
                             double* out, double* ws, void* stream) {
  if (check_vecs(v, "pmb_mma_setup")) return 1;
  PMB_REQUIRE(n > 0 && xval && dg && offset && rho && out && ws, "pmb_mma_setup: invalid argument");
  PMB_REQUIRE(version == 1987 || version == 2007, "pmb_mma_setup: version must be 1987 or 2007 (GCMMA: 2007 with per-response rho >= 1e-6)");
  PMB_REQUIRE(m >= 1 && m <= PMB_MMA_MAXM, "pmb_mma_setup: m=%d not in 1..%d", m, PMB_MMA_MAXM);
  MmaRows rows;
  for (int i = 0; i <= PMB_MMA_MAXM; ++i) rows.r[i] = i <= m ? dg[i] : nullptr;
  for (int i = 0; i <= m; ++i) PMB_REQUIRE(rows.r[i], "pmb_mma_setup: NULL sensitivity row %d", i);
  const MmaSmall rh = small_from(rho, m + 1);
  MMA_DISPATCH(m, (mma_setup_kernel<M><<<opt_blocks(n), OPT_THREADS, 0, (cudaStream_t)stream>>>(n, xval, rows, offset, xmin, xmax, move,
                                                                                                albefa, rh, version, *v, out, ws)));
  PMB_CHECK_LAUNCH("pmb_mma_setup");
  return 0;
}

// ------------------------------------------------------------------------------------------------- residual
// out = [sum of squared residuals (x, xsi, eta parts), gvec[0..m), max squared residual]
template <int M>
__global__ void __launch_bounds__(OPT_THREADS) mma_residual_kernel(long long n, pmb_mma_vecs a, MmaSmall lam, double epsi, double* out, double* ws) {
  double s[M + 1], mx[1] = {0.0};
#pragma unroll
  for (int i = 0; i <= M; ++i) s[i] = 0.0;
  const long long stride = (long long)gridDim.x * OPT_THREADS;
  for (long long j = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; j < n; j += stride) {
    MmaVar v;
    load_var<M>(a, n, j, v);
    mma_resid_pt<M>(v, lam, epsi, s[0], mx[0], s + 1);
  }
  opt_reduce<M + 1, 1>(s, mx, ws, out);
}

extern "C" int pmb_mma_residual(long long n, int m, const pmb_mma_vecs* v, const double* lam, double epsi, double* out, double* ws,
                                void* stream) {
  if (check_vecs(v, "pmb_mma_residual")) return 1;
  PMB_REQUIRE(n > 0 && lam && out && ws, "pmb_mma_residual: invalid argument");
  const MmaSmall l = small_from(lam, m);
  MMA_DISPATCH(m, (mma_residual_kernel<M><<<opt_blocks(n), OPT_THREADS, 0, (cudaStream_t)stream>>>(n, *v, l, epsi, out, ws)));
  PMB_CHECK_LAUNCH("pmb_mma_residual");
  return 0;
}

// ------------------------------------------------------------------------------------------------- Newton sums
// out = [gvec[m], GG (delx/diagx) [m], (GG/diagx) GG^T [m*m row-major]]
template <int M>
__global__ void __launch_bounds__(OPT_THREADS) mma_newton_sums_kernel(long long n, pmb_mma_vecs a, MmaSmall lam, double epsi, double* out,
                                                                      double* ws) {
  constexpr int NS = 2 * M + M * M;
  double s[NS], mx[1] = {0.0};
#pragma unroll
  for (int i = 0; i < NS; ++i) s[i] = 0.0;
  const long long stride = (long long)gridDim.x * OPT_THREADS;
  for (long long j = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; j < n; j += stride) {
    MmaVar v;
    load_var<M>(a, n, j, v);
    double delx, diagx, GG[M], gterm[M];
    mma_newton_pt<M>(v, lam, epsi, delx, diagx, GG, gterm);
    const double r = delx / diagx;
#pragma unroll
    for (int i = 0; i < M; ++i) {
      s[i] += gterm[i];
      s[M + i] += GG[i] * r;
      const double gd = GG[i] / diagx;
#pragma unroll
      for (int k = 0; k < M; ++k) s[2 * M + i * M + k] += gd * GG[k];
    }
  }
  opt_reduce<NS, 0>(s, mx, ws, out);
}

extern "C" int pmb_mma_newton_sums(long long n, int m, const pmb_mma_vecs* v, const double* lam, double epsi, double* out, double* ws,
                                   void* stream) {
  if (check_vecs(v, "pmb_mma_newton_sums")) return 1;
  PMB_REQUIRE(n > 0 && lam && out && ws, "pmb_mma_newton_sums: invalid argument");
  const MmaSmall l = small_from(lam, m);
  MMA_DISPATCH(m, (mma_newton_sums_kernel<M><<<opt_blocks(n), OPT_THREADS, 0, (cudaStream_t)stream>>>(n, *v, l, epsi, out, ws)));
  PMB_CHECK_LAUNCH("pmb_mma_newton_sums");
  return 0;
}

// ------------------------------------------------------------------------------------------------- Newton direction
// stores dx, dxsi, deta and the line-search base point (xo, xsio, etao); out = 4 maxima (see mma_dir_pt)
template <int M>
__global__ void __launch_bounds__(OPT_THREADS) mma_newton_dir_kernel(long long n, pmb_mma_vecs a, MmaSmall lam, MmaSmall dlam, double epsi,
                                                                     double* out, double* ws) {
  double s[1] = {0.0}, mx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) mx[i] = -INFINITY;
  const long long stride = (long long)gridDim.x * OPT_THREADS;
  for (long long j = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; j < n; j += stride) {
    MmaVar v;
    load_var<M>(a, n, j, v);
    double dx, dxsi, deta, cand[4];
    mma_dir_pt<M>(v, lam, dlam, epsi, dx, dxsi, deta, cand);
    a.dx[j] = dx; a.dxsi[j] = dxsi; a.deta[j] = deta;
    a.xo[j] = v.x; a.xsio[j] = v.xsi; a.etao[j] = v.eta;
#pragma unroll
    for (int i = 0; i < 4; ++i) mx[i] = fmax(mx[i], cand[i]);
  }
  opt_reduce<1, 4>(s, mx, ws, out);
}

extern "C" int pmb_mma_newton_dir(long long n, int m, const pmb_mma_vecs* v, const double* lam, const double* dlam, double epsi,
                                  double* out, double* ws, void* stream) {
  if (check_vecs(v, "pmb_mma_newton_dir")) return 1;
  PMB_REQUIRE(n > 0 && lam && dlam && out && ws, "pmb_mma_newton_dir: invalid argument");
  const MmaSmall l = small_from(lam, m), dl = small_from(dlam, m);
  MMA_DISPATCH(m, (mma_newton_dir_kernel<M><<<opt_blocks(n), OPT_THREADS, 0, (cudaStream_t)stream>>>(n, *v, l, dl, epsi, out, ws)));
  PMB_CHECK_LAUNCH("pmb_mma_newton_dir");
  return 0;
}

// ------------------------------------------------------------------------------------------------- line-search trial
// x = xo + steg dx (and xsi, eta), then the residual of the trial point with the trial lam: out as pmb_mma_residual
template <int M>
__global__ void __launch_bounds__(OPT_THREADS) mma_linesearch_kernel(long long n, pmb_mma_vecs a, MmaSmall lam, double steg, double epsi,
                                                                     double* out, double* ws) {
  double s[M + 1], mx[1] = {0.0};
#pragma unroll
  for (int i = 0; i <= M; ++i) s[i] = 0.0;
  const long long stride = (long long)gridDim.x * OPT_THREADS;
  for (long long j = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; j < n; j += stride) {
    MmaVar v;
    load_var<M>(a, n, j, v);
    v.x = a.xo[j] + steg * a.dx[j];
    v.xsi = a.xsio[j] + steg * a.dxsi[j];
    v.eta = a.etao[j] + steg * a.deta[j];
    a.x[j] = v.x; a.xsi[j] = v.xsi; a.eta[j] = v.eta;
    mma_resid_pt<M>(v, lam, epsi, s[0], mx[0], s + 1);
  }
  opt_reduce<M + 1, 1>(s, mx, ws, out);
}

extern "C" int pmb_mma_linesearch(long long n, int m, const pmb_mma_vecs* v, const double* lam, double steg, double epsi, double* out,
                                  double* ws, void* stream) {
  if (check_vecs(v, "pmb_mma_linesearch")) return 1;
  PMB_REQUIRE(n > 0 && lam && out && ws, "pmb_mma_linesearch: invalid argument");
  const MmaSmall l = small_from(lam, m);
  MMA_DISPATCH(m, (mma_linesearch_kernel<M><<<opt_blocks(n), OPT_THREADS, 0, (cudaStream_t)stream>>>(n, *v, l, steg, epsi, out, ws)));
  PMB_CHECK_LAUNCH("pmb_mma_linesearch");
  return 0;
}

// ------------------------------------------------------------------------------------------------- GCMMA passes
// out[i] = sum_j (xmax_j - xmin_j) |dg_i[j]|, i = 0..m   (mma.py:151; the caller scales by 0.1 / n)
template <int M>
__global__ void __launch_bounds__(OPT_THREADS) mma_gcmma_rho_kernel(long long n, MmaRows dg, pmb_bound xmin, pmb_bound xmax, double* out,
                                                                    double* ws) {
  double s[M + 1], mx[1] = {0.0};
#pragma unroll
  for (int i = 0; i <= M; ++i) s[i] = 0.0;
  const long long stride = (long long)gridDim.x * OPT_THREADS;
  for (long long j = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; j < n; j += stride) {
    const double lo = bound_at(xmin, j), hi = bound_at(xmax, j);
#pragma unroll
    for (int i = 0; i <= M; ++i) s[i] += mma_rho_term(dg.r[i][j], lo, hi);
  }
  opt_reduce<M + 1, 0>(s, mx, ws, out);
}

extern "C" int pmb_mma_gcmma_rho(long long n, int m, const double* const* dg, pmb_bound xmin, pmb_bound xmax, double* out, double* ws,
                                 void* stream) {
  PMB_REQUIRE(n > 0 && dg && out && ws, "pmb_mma_gcmma_rho: invalid argument");
  PMB_REQUIRE(m >= 1 && m <= PMB_MMA_MAXM, "pmb_mma_gcmma_rho: m=%d not in 1..%d", m, PMB_MMA_MAXM);
  MmaRows rows;
  for (int i = 0; i <= PMB_MMA_MAXM; ++i) rows.r[i] = i <= m ? dg[i] : nullptr;
  for (int i = 0; i <= m; ++i) PMB_REQUIRE(rows.r[i], "pmb_mma_gcmma_rho: NULL sensitivity row %d", i);
  MMA_DISPATCH(m, (mma_gcmma_rho_kernel<M><<<opt_blocks(n), OPT_THREADS, 0, (cudaStream_t)stream>>>(n, rows, xmin, xmax, out, ws)));
  PMB_CHECK_LAUNCH("pmb_mma_gcmma_rho");
  return 0;
}

// out[0..m] = sum_j P_ij/(upp_j - x_j) + Q_ij/(x_j - low_j) at the subproblem solution x (mma.py:236 without "- rhs"),
// out[m+1] = dk = sum_j (upp-low)(x-xval)^2 / ((upp-x)(x-low)(xmax-xmin))   (:239)
template <int M>
__global__ void __launch_bounds__(OPT_THREADS) mma_gcmma_estimate_kernel(long long n, pmb_mma_vecs a, const double* __restrict__ xval,
                                                                         pmb_bound xmin, pmb_bound xmax, double* out, double* ws) {
  double s[M + 2], mx[1] = {0.0};
#pragma unroll
  for (int i = 0; i < M + 2; ++i) s[i] = 0.0;
  const long long stride = (long long)gridDim.x * OPT_THREADS;
  for (long long j = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; j < n; j += stride) {
    MmaVar v;
    load_var<M>(a, n, j, v);
    s[M + 1] += mma_estimate_pt<M>(v, xval[j], bound_at(xmin, j), bound_at(xmax, j), s);
  }
  opt_reduce<M + 2, 0>(s, mx, ws, out);
}

extern "C" int pmb_mma_gcmma_estimate(long long n, int m, const pmb_mma_vecs* v, const double* xval, pmb_bound xmin, pmb_bound xmax,
                                      double* out, double* ws, void* stream) {
  if (check_vecs(v, "pmb_mma_gcmma_estimate")) return 1;
  PMB_REQUIRE(n > 0 && xval && out && ws, "pmb_mma_gcmma_estimate: invalid argument");
  MMA_DISPATCH(m, (mma_gcmma_estimate_kernel<M><<<opt_blocks(n), OPT_THREADS, 0, (cudaStream_t)stream>>>(n, *v, xval, xmin, xmax, out, ws)));
  PMB_CHECK_LAUNCH("pmb_mma_gcmma_estimate");
  return 0;
}

// ------------------------------------------------------------------------------------------------- VTI payload packer
// out[i*co + c] = (float) in[i*ci + c] for c < ci, 0 for ci <= c < co   (pymoto/common/domain.py:541-548: astype(float32) and the
// 2 -> 3 component padding of 2-D nodal vectors); round-to-nearest-even like numpy's astype
__global__ void __launch_bounds__(256) pack_f32_kernel(long long nitems, int ci, int co, const double* __restrict__ in, float* __restrict__ out) {
  const long long total = nitems * co;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long long i = t / co;
    const int c = (int)(t - i * co);
    out[t] = c < ci ? __double2float_rn(in[i * ci + c]) : 0.0f;
  }
}

extern "C" int pmb_pack_f32(long long nitems, int ncomp_in, int ncomp_out, const double* in, float* out, void* stream) {
  PMB_REQUIRE(in && out, "pmb_pack_f32: NULL pointer argument");
  PMB_REQUIRE(ncomp_in >= 1 && ncomp_out >= ncomp_in, "pmb_pack_f32: components %d -> %d", ncomp_in, ncomp_out);
  if (nitems <= 0) return 0;
  long long b = (nitems * ncomp_out + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  pack_f32_kernel<<<(unsigned)b, 256, 0, (cudaStream_t)stream>>>(nitems, ncomp_in, ncomp_out, in, out);
  PMB_CHECK_LAUNCH("pmb_pack_f32");
  return 0;
}
