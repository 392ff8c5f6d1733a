// TEST INFRASTRUCTURE ONLY -- never linked into libpmb.so, never imported by pymoto_b200.
//
// Host loops over the SAME per-variable arithmetic the CUDA kernels of pymoto_b200/csrc/pmb_optim.cu use
// (pymoto_b200/csrc/pmb_mma_math.h), with the same entry-point semantics and output layouts, so the CPU test suite can run
// the product's Newton driver (pymoto_b200.optimizers.mma_subsolv) against the reference's MMA without a GPU and pin the
// formulas before they are exercised on the device.  Built by tests/test_mma_cpu.py with g++.
#include <algorithm>
#include <cmath>
#include "../pymoto_b200/csrc/pmb_mma_math.h"

struct HVecs {
  double *x, *xsi, *eta, *xo, *xsio, *etao, *dx, *dxsi, *deta, *low, *upp, *alfa, *beta, *P, *Q;
};
struct HBound {
  double s;
  const double* v;
};
static inline double bat(const HBound& b, long long j) { return b.v ? b.v[j] : b.s; }

template <int M>
static void load(const HVecs& a, long long n, long long j, MmaVar& v) {
  v.x = a.x[j]; v.xsi = a.xsi[j]; v.eta = a.eta[j];
  v.low = a.low[j]; v.upp = a.upp[j]; v.alfa = a.alfa[j]; v.beta = a.beta[j];
  for (int i = 0; i <= M; ++i) v.P[i] = a.P[i * n + j], v.Q[i] = a.Q[i * n + j];
}
static MmaSmall small_from(const double* h, int count) {
  MmaSmall s;
  for (int i = 0; i <= PMB_MMA_MAXM; ++i) s.v[i] = (h && i < count) ? h[i] : 0.0;
  return s;
}

template <int M>
static void setup_t(long long n, const double* xval, const double* const* dg, const double* offset, HBound xmin, HBound xmax, HBound move,
                    double albefa, const double* rho, int version, HVecs a, double* out) {
  const MmaSmall rh = small_from(rho, M + 1);
  for (int i = 0; i <= M; ++i) out[i] = 0.0;
  for (long long j = 0; j < n; ++j) {
    double dgj[M + 1];
    for (int i = 0; i <= M; ++i) dgj[i] = dg[i][j];
    MmaVar v;
    const double sinv = mma_setup_pt<M>(xval[j], dgj, offset[j], bat(xmin, j), bat(xmax, j), bat(move, j), albefa, rh, version, v);
    a.x[j] = v.x; a.xsi[j] = v.xsi; a.eta[j] = v.eta; a.low[j] = v.low; a.upp[j] = v.upp; a.alfa[j] = v.alfa; a.beta[j] = v.beta;
    for (int i = 0; i <= M; ++i) {
      a.P[i * n + j] = v.P[i];
      a.Q[i * n + j] = v.Q[i];
      out[i] += v.P[i] * sinv + v.Q[i] * sinv;
    }
  }
}
template <int M>
static void residual_t(long long n, HVecs a, const double* lam, double epsi, double* out) {
  const MmaSmall l = small_from(lam, M);
  double s[M + 1] = {0.0}, mx = 0.0;
  for (long long j = 0; j < n; ++j) {
    MmaVar v;
    load<M>(a, n, j, v);
    mma_resid_pt<M>(v, l, epsi, s[0], mx, s + 1);
  }
  for (int i = 0; i <= M; ++i) out[i] = s[i];
  out[M + 1] = mx;
}
template <int M>
static void sums_t(long long n, HVecs a, const double* lam, double epsi, double* out) {
  const MmaSmall l = small_from(lam, M);
  constexpr int NS = 2 * M + M * M;
  for (int i = 0; i < NS; ++i) out[i] = 0.0;
  for (long long j = 0; j < n; ++j) {
    MmaVar v;
    load<M>(a, n, j, v);
    double delx, diagx, GG[M], gterm[M];
    mma_newton_pt<M>(v, l, epsi, delx, diagx, GG, gterm);
    const double r = delx / diagx;
    for (int i = 0; i < M; ++i) {
      out[i] += gterm[i];
      out[M + i] += GG[i] * r;
      const double gd = GG[i] / diagx;
      for (int k = 0; k < M; ++k) out[2 * M + i * M + k] += gd * GG[k];
    }
  }
}
template <int M>
static void dir_t(long long n, HVecs a, const double* lam, const double* dlam, double epsi, double* out) {
  const MmaSmall l = small_from(lam, M), dl = small_from(dlam, M);
  double mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  for (long long j = 0; j < n; ++j) {
    MmaVar v;
    load<M>(a, n, j, v);
    double dx, dxsi, deta, cand[4];
    mma_dir_pt<M>(v, l, dl, epsi, dx, dxsi, deta, cand);
    a.dx[j] = dx; a.dxsi[j] = dxsi; a.deta[j] = deta; a.xo[j] = v.x; a.xsio[j] = v.xsi; a.etao[j] = v.eta;
    for (int i = 0; i < 4; ++i) mx[i] = std::max(mx[i], cand[i]);
  }
  out[0] = 0.0;
  for (int i = 0; i < 4; ++i) out[1 + i] = mx[i];
}
template <int M>
static void ls_t(long long n, HVecs a, const double* lam, double steg, double epsi, double* out) {
  const MmaSmall l = small_from(lam, M);
  double s[M + 1] = {0.0}, mx = 0.0;
  for (long long j = 0; j < n; ++j) {
    MmaVar v;
    load<M>(a, n, j, v);
    v.x = a.xo[j] + steg * a.dx[j];
    v.xsi = a.xsio[j] + steg * a.dxsi[j];
    v.eta = a.etao[j] + steg * a.deta[j];
    a.x[j] = v.x; a.xsi[j] = v.xsi; a.eta[j] = v.eta;
    mma_resid_pt<M>(v, l, epsi, s[0], mx, s + 1);
  }
  for (int i = 0; i <= M; ++i) out[i] = s[i];
  out[M + 1] = mx;
}

template <int M>
static void rho_t(long long n, const double* const* dg, HBound xmin, HBound xmax, double* out) {
  for (int i = 0; i <= M; ++i) out[i] = 0.0;
  for (long long j = 0; j < n; ++j)
    for (int i = 0; i <= M; ++i) out[i] += mma_rho_term(dg[i][j], bat(xmin, j), bat(xmax, j));
}
template <int M>
static void est_t(long long n, HVecs a, const double* xval, HBound xmin, HBound xmax, double* out) {
  double s[M + 2] = {0.0};
  for (long long j = 0; j < n; ++j) {
    MmaVar v;
    load<M>(a, n, j, v);
    s[M + 1] += mma_estimate_pt<M>(v, xval[j], bat(xmin, j), bat(xmax, j), s);
  }
  for (int i = 0; i < M + 2; ++i) out[i] = s[i];
}

#define DISPATCH(m, CALL) \
  switch (m) { case 1: { constexpr int M = 1; CALL; } break; case 2: { constexpr int M = 2; CALL; } break; \
               case 3: { constexpr int M = 3; CALL; } break; case 4: { constexpr int M = 4; CALL; } break; \
               case 5: { constexpr int M = 5; CALL; } break; case 6: { constexpr int M = 6; CALL; } break; default: return 1; }

extern "C" {
int hmma_asymptotes(long long n, const double* x, const double* xold1, const double* xold2, double asyincr, double asydecr, double asybound,
                    double* offset) {
  for (long long j = 0; j < n; ++j) offset[j] = mma_offset_update(offset[j], x[j], xold1[j], xold2[j], asyincr, asydecr, asybound);
  return 0;
}
int hmma_setup(long long n, int m, const double* xval, const double* const* dg, const double* offset, HBound xmin, HBound xmax, HBound move,
               double albefa, const double* rho, int version, const HVecs* v, double* out) {
  DISPATCH(m, (setup_t<M>(n, xval, dg, offset, xmin, xmax, move, albefa, rho, version, *v, out)));
  return 0;
}
int hmma_residual(long long n, int m, const HVecs* v, const double* lam, double epsi, double* out) {
  DISPATCH(m, (residual_t<M>(n, *v, lam, epsi, out)));
  return 0;
}
int hmma_newton_sums(long long n, int m, const HVecs* v, const double* lam, double epsi, double* out) {
  DISPATCH(m, (sums_t<M>(n, *v, lam, epsi, out)));
  return 0;
}
int hmma_newton_dir(long long n, int m, const HVecs* v, const double* lam, const double* dlam, double epsi, double* out) {
  DISPATCH(m, (dir_t<M>(n, *v, lam, dlam, epsi, out)));
  return 0;
}
int hmma_gcmma_rho(long long n, int m, const double* const* dg, HBound xmin, HBound xmax, double* out) {
  DISPATCH(m, (rho_t<M>(n, dg, xmin, xmax, out)));
  return 0;
}
int hmma_gcmma_estimate(long long n, int m, const HVecs* v, const double* xval, HBound xmin, HBound xmax, double* out) {
  DISPATCH(m, (est_t<M>(n, *v, xval, xmin, xmax, out)));
  return 0;
}
int hmma_linesearch(long long n, int m, const HVecs* v, const double* lam, double steg, double epsi, double* out) {
  DISPATCH(m, (ls_t<M>(n, *v, lam, steg, epsi, out)));
  return 0;
}
}
