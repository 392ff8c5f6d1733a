// libpmb: CG reductions and vector updates (K8, K9), Dirichlet split (K12), SIMP glue.
//
// Replaces the numpy dot / norm / axpy calls of pymoto/solvers/iterative.py:359-395 (CG) and the Dirichlet
// handling of pymoto/solvers/solvers.py:175-176, 213-218 (LDAS).  Scalars stay in device memory: kernels that
// need alpha / beta read the dot products directly, so the host only polls the residual norm.
#include "pmb_common.cuh"

static constexpr int RED_BLOCKS = 592;  // 4 CTAs per SM on 148 SMs
static constexpr int RED_THREADS = 256;

__device__ __forceinline__ double warp_sum_v(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-level sum of up to 4 values; result valid in thread 0
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* sm /* K*8 */) {
#pragma unroll
  for (int q = 0; q < K; ++q) v[q] = warp_sum_v(v[q]);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0)
#pragma unroll
    for (int q = 0; q < K; ++q) sm[q * 8 + wid] = v[q];
  __syncthreads();
  if (threadIdx.x == 0)
#pragma unroll
    for (int q = 0; q < K; ++q) {
      double s = 0.0;
      for (int w = 0; w < RED_THREADS / 32; ++w) s += sm[q * 8 + w];
      v[q] = s;
    }
}

// second stage: the last CTA to finish sums the per-CTA partials in index order (deterministic)
template <int K>
__device__ __forceinline__ void finish_reduction(double (&v)[K], double* ws, double* out) {
  // ws layout: [0] = ticket counter (as unsigned), [8 ...] = partials[K][RED_BLOCKS]
  __shared__ bool last;
  unsigned* counter = reinterpret_cast<unsigned*>(ws);
  double* partials = ws + 8;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < K; ++q) partials[q * RED_BLOCKS + blockIdx.x] = v[q];
    __threadfence();
    unsigned ticket = atomicAdd(counter, 1u);
    last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    __shared__ double sm2[4 * 8];
    double a[K];
#pragma unroll
    for (int q = 0; q < K; ++q) {
      a[q] = 0.0;
      for (int bI = threadIdx.x; bI < (int)gridDim.x; bI += RED_THREADS) a[q] += __ldcg(partials + q * RED_BLOCKS + bI);
    }
    block_sum<K>(a, sm2);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int q = 0; q < K; ++q) out[q] = a[q];
      *counter = 0u;  // re-arm for the next call on this stream
    }
  }
}

extern "C" long long pmb_ws_doubles(void) { return 8 + 4 * RED_BLOCKS; }

// ------------------------------------------------------------------------------------------------- K8
template <int K>
__global__ void __launch_bounds__(RED_THREADS) dots_kernel(long long n, const double* __restrict__ a0, const double* __restrict__ b0,
                                                           const double* __restrict__ a1, const double* __restrict__ b1,
                                                           const double* __restrict__ a2, const double* __restrict__ b2,
                                                           const double* __restrict__ a3, const double* __restrict__ b3,
                                                           double* __restrict__ out, double* ws) {
  __shared__ double sm[4 * 8];
  double v[K];
#pragma unroll
  for (int q = 0; q < K; ++q) v[q] = 0.0;
  const long long stride = (long long)gridDim.x * RED_THREADS;
  for (long long t = (long long)blockIdx.x * RED_THREADS + threadIdx.x; t < n; t += stride) {
    v[0] = fma(a0[t], b0[t], v[0]);
    if (K > 1) v[K > 1 ? 1 : 0] = fma(a1[t], b1[t], v[K > 1 ? 1 : 0]);
    if (K > 2) v[K > 2 ? 2 : 0] = fma(a2[t], b2[t], v[K > 2 ? 2 : 0]);
    if (K > 3) v[K > 3 ? 3 : 0] = fma(a3[t], b3[t], v[K > 3 ? 3 : 0]);
  }
  block_sum<K>(v, sm);
  finish_reduction<K>(v, ws, out);
}

static int red_blocks_for(long long n) {
  long long b = (n + RED_THREADS - 1) / RED_THREADS;
  return (int)(b < 1 ? 1 : (b > RED_BLOCKS ? RED_BLOCKS : b));
}

extern "C" int pmb_dots(long long n, int k, const double* a0, const double* b0, const double* a1, const double* b1,
                        const double* a2, const double* b2, const double* a3, const double* b3, double* out, double* ws,
                        void* stream) {
  PMB_REQUIRE(k >= 1 && k <= 4, "pmb_dots: k=%d not in 1..4", k);
  PMB_REQUIRE(out && ws && a0 && b0, "pmb_dots: NULL pointer argument");
  PMB_REQUIRE(n >= 0, "pmb_dots: negative length");
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = red_blocks_for(n);
  switch (k) {
    case 1: dots_kernel<1><<<blocks, RED_THREADS, 0, st>>>(n, a0, b0, a1, b1, a2, b2, a3, b3, out, ws); break;
    case 2: PMB_REQUIRE(a1 && b1, "pmb_dots: NULL pair 1");
      dots_kernel<2><<<blocks, RED_THREADS, 0, st>>>(n, a0, b0, a1, b1, a2, b2, a3, b3, out, ws); break;
    case 3: PMB_REQUIRE(a1 && b1 && a2 && b2, "pmb_dots: NULL pair");
      dots_kernel<3><<<blocks, RED_THREADS, 0, st>>>(n, a0, b0, a1, b1, a2, b2, a3, b3, out, ws); break;
    case 4: PMB_REQUIRE(a1 && b1 && a2 && b2 && a3 && b3, "pmb_dots: NULL pair");
      dots_kernel<4><<<blocks, RED_THREADS, 0, st>>>(n, a0, b0, a1, b1, a2, b2, a3, b3, out, ws); break;
  }
  PMB_CHECK_LAUNCH("pmb_dots");
  return 0;
}

// ------------------------------------------------------------------------------------------------- K9
__device__ __forceinline__ double coef_value(const pmb_coef& c) {
  double v = c.c;
  if (c.num) v *= *c.num;
  if (c.den) v /= (c.sqrt_den ? sqrt(*c.den) : *c.den);
  return v;
}

__global__ void __launch_bounds__(256) lincomb_kernel(long long n, double* out, pmb_coef ca, const double* a, pmb_coef cb,
                                                       const double* b) {
  const double va = coef_value(ca);
  const double vb = b ? coef_value(cb) : 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    double r = va * a[t];
    if (b) r = fma(vb, b[t], r);
    out[t] = r;
  }
}

static unsigned ew_blocks(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = 148LL * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" int pmb_lincomb(long long n, double* out, pmb_coef ca, const double* a, pmb_coef cb, const double* b, void* stream) {
  PMB_REQUIRE(out && a, "pmb_lincomb: NULL pointer argument");
  if (n <= 0) return 0;
  lincomb_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(n, out, ca, a, cb, b);
  PMB_CHECK_LAUNCH("pmb_lincomb");
  return 0;
}

// alpha = pr/pq ; x += alpha p ; (optionally) r -= alpha q and rr = r.r      (iterative.py:378-386)
__global__ void __launch_bounds__(RED_THREADS) cg_xr_kernel(long long n, double* __restrict__ x, double* __restrict__ r,
                                                            const double* __restrict__ p, const double* __restrict__ q,
                                                            const double* __restrict__ pr, const double* __restrict__ pq,
                                                            double* __restrict__ rr_out, double* ws) {
  __shared__ double sm[4 * 8];
  const double alpha = *pr / *pq;
  double v[1] = {0.0};
  const long long stride = (long long)gridDim.x * RED_THREADS;
  for (long long t = (long long)blockIdx.x * RED_THREADS + threadIdx.x; t < n; t += stride) {
    x[t] = fma(alpha, p[t], x[t]);
    if (q) {
      double rn = fma(-alpha, q[t], r[t]);
      r[t] = rn;
      v[0] = fma(rn, rn, v[0]);
    }
  }
  if (q && rr_out) {
    block_sum<1>(v, sm);
    finish_reduction<1>(v, ws, rr_out);
  }
}

extern "C" int pmb_cg_xr_update(long long n, double* x, double* r, const double* p, const double* q, const double* pr,
                                const double* pq, double* rr_out, double* ws, void* stream) {
  PMB_REQUIRE(x && p && pr && pq, "pmb_cg_xr_update: NULL pointer argument");
  PMB_REQUIRE(!q || (r && (!rr_out || ws)), "pmb_cg_xr_update: r / workspace required with q");
  cg_xr_kernel<<<red_blocks_for(n), RED_THREADS, 0, (cudaStream_t)stream>>>(n, x, r, p, q, pr, pq, rr_out, ws);
  PMB_CHECK_LAUNCH("pmb_cg_xr_update");
  return 0;
}

// ------------------------------------------------------------------------------------------------- K12
__global__ void __launch_bounds__(256) bc_split_kernel(long long n, const unsigned char* __restrict__ mask, const double* __restrict__ rhs,
                                                        const double* __restrict__ diag, double* __restrict__ sol,
                                                        double* __restrict__ rhs_loc) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const bool m = mask[t];
    const double f = rhs[t];
    sol[t] = m ? f / diag[t] : 0.0;
    rhs_loc[t] = m ? 0.0 : f;
  }
}

extern "C" int pmb_bc_split(long long n, const unsigned char* mask, const double* rhs, const double* diag, double* sol,
                            double* rhs_loc, void* stream) {
  PMB_REQUIRE(mask && rhs && diag && sol && rhs_loc, "pmb_bc_split: NULL pointer argument");
  if (n <= 0) return 0;
  bc_split_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(n, mask, rhs, diag, sol, rhs_loc);
  PMB_CHECK_LAUNCH("pmb_bc_split");
  return 0;
}

__global__ void __launch_bounds__(256) mask_zero_kernel(long long n, const unsigned char* __restrict__ mask, const double* in, double* out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) out[t] = mask[t] ? 0.0 : in[t];
}

extern "C" int pmb_mask_zero(long long n, const unsigned char* mask, const double* in, double* out, void* stream) {
  PMB_REQUIRE(mask && in && out, "pmb_mask_zero: NULL pointer argument");
  if (n <= 0) return 0;
  mask_zero_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(n, mask, in, out);
  PMB_CHECK_LAUNCH("pmb_mask_zero");
  return 0;
}

// mask[r] = (diag[r] != 0 && nnz_offdiag[r] == 0)      (get_diagonal_indices, solvers.py:88-96)
__global__ void __launch_bounds__(256) diag_mask_kernel(long long n, const double* __restrict__ diag, const int* __restrict__ nnz_off,
                                                         unsigned char* __restrict__ mask) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
    mask[t] = (diag[t] != 0.0 && nnz_off[t] == 0) ? 1 : 0;
}

extern "C" int pmb_diag_mask(long long n, const double* diag, const int* nnz_offdiag, unsigned char* mask, void* stream) {
  PMB_REQUIRE(diag && nnz_offdiag && mask, "pmb_diag_mask: NULL pointer argument");
  if (n <= 0) return 0;
  diag_mask_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(n, diag, nnz_offdiag, mask);
  PMB_CHECK_LAUNCH("pmb_diag_mask");
  return 0;
}

// ------------------------------------------------------------------------------------------------- elementwise glue
__global__ void __launch_bounds__(256) vec_div_kernel(long long n, const double* a, const double* b, double* out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) out[t] = a[t] / b[t];
}

extern "C" int pmb_vec_div(long long n, const double* a, const double* b, double* out, void* stream) {
  PMB_REQUIRE(a && b && out, "pmb_vec_div: NULL pointer argument");
  if (n <= 0) return 0;
  vec_div_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(n, a, b, out);
  PMB_CHECK_LAUNCH("pmb_vec_div");
  return 0;
}

__device__ __forceinline__ double ipow(double y, int p) {
  double r = 1.0;
  for (int q = 0; q < p; ++q) r *= y;
  return r;
}

__global__ void __launch_bounds__(256) simp_kernel(long long n, double xmin, int p, const double* __restrict__ y, double* __restrict__ s) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
    s[t] = xmin + (1.0 - xmin) * ipow(y[t], p);
}

__global__ void __launch_bounds__(256) simp_bwd_kernel(long long n, double xmin, int p, const double* __restrict__ y,
                                                        const double* __restrict__ ds, double* __restrict__ dy) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
    dy[t] = ds[t] * ((double)p * (1.0 - xmin) * ipow(y[t], p - 1));
}

extern "C" int pmb_simp(long long n, double xmin, int p, const double* y, double* s, void* stream) {
  PMB_REQUIRE(y && s && p >= 1, "pmb_simp: invalid argument");
  if (n <= 0) return 0;
  simp_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(n, xmin, p, y, s);
  PMB_CHECK_LAUNCH("pmb_simp");
  return 0;
}

extern "C" int pmb_simp_bwd(long long n, double xmin, int p, const double* y, const double* ds, double* dy, void* stream) {
  PMB_REQUIRE(y && ds && dy && p >= 1, "pmb_simp_bwd: invalid argument");
  if (n <= 0) return 0;
  simp_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(n, xmin, p, y, ds, dy);
  PMB_CHECK_LAUNCH("pmb_simp_bwd");
  return 0;
}

// ------------------------------------------------------------------------------------------------- OC update
// One bisection candidate of the optimality-criteria update (pymoto/common/optimizers.py:425-435):
//   xnew_i = clip(x_i sqrt(-min(dg_i, 0) / lmid), max(xmin, x_i - move), min(xmax, x_i + move)),  sum_out = sum_i xnew_i
// xnew may be NULL (only the volume is needed while bisecting).
__global__ void __launch_bounds__(RED_THREADS) oc_candidate_kernel(long long n, const double* __restrict__ x, const double* __restrict__ dg,
                                                                   double move, double xmin, double xmax, double lmid,
                                                                   double* __restrict__ xnew, double* __restrict__ sum_out, double* ws) {
  __shared__ double sm[4 * 8];
  double v[1] = {0.0};
  const long long stride = (long long)gridDim.x * RED_THREADS;
  for (long long t = (long long)blockIdx.x * RED_THREADS + threadIdx.x; t < n; t += stride) {
    const double xi = x[t];
    const double g = fmin(dg[t], 0.0);
    const double cand = xi * sqrt(-g / lmid);
    const double lb = fmax(xmin, xi - move), ub = fmin(xmax, xi + move);
    const double c = fmin(fmax(cand, lb), ub);
    if (xnew) xnew[t] = c;
    v[0] += c;
  }
  block_sum<1>(v, sm);
  finish_reduction<1>(v, ws, sum_out);
}

extern "C" int pmb_oc_candidate(long long n, const double* x, const double* dg, double move, double xmin, double xmax, double lmid,
                                double* xnew, double* sum_out, double* ws, void* stream) {
  PMB_REQUIRE(x && dg && sum_out && ws, "pmb_oc_candidate: NULL pointer argument");
  PMB_REQUIRE(lmid > 0.0, "pmb_oc_candidate: lmid must be positive");
  oc_candidate_kernel<<<red_blocks_for(n), RED_THREADS, 0, (cudaStream_t)stream>>>(n, x, dg, move, xmin, xmax, lmid, xnew, sum_out, ws);
  PMB_CHECK_LAUNCH("pmb_oc_candidate");
  return 0;
}

// ------------------------------------------------------------------------------------------------- halo mailboxes (multi-GPU)
// Neighbour halo exchange over NVLink peer memory: every rank owns a mailbox in symmetric memory; pmb_halo_pack stores
// this rank's boundary planes straight into the mailboxes of its z-neighbours (peer pointers, plain st.global over
// NVLink), a device-side barrier follows (torch symmetric-memory signal pads), then pmb_halo_unpack moves the received
// planes into the halo planes of the destination vector.  Either destination / source may be NULL (domain ends or a
// one-sided exchange).
__global__ void __launch_bounds__(256) halo_copy2_kernel(long long n, const double* __restrict__ src0, double* __restrict__ dst0,
                                                          const double* __restrict__ src1, double* __restrict__ dst1) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < 2 * n; t += stride) {
    if (t < n) {
      if (src0 && dst0) dst0[t] = src0[t];
    } else {
      if (src1 && dst1) dst1[t - n] = src1[t - n];
    }
  }
}

extern "C" int pmb_halo_copy2(long long n, const double* src0, double* dst0, const double* src1, double* dst1, void* stream) {
  PMB_REQUIRE(n > 0, "pmb_halo_copy2: invalid length");
  if (!(src0 && dst0) && !(src1 && dst1)) return 0;
  long long blocks = (2 * n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  halo_copy2_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, src0, dst0, src1, dst1);
  PMB_CHECK_LAUNCH("pmb_halo_copy2");
  return 0;
}
