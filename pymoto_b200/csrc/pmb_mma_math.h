// Per-variable arithmetic of the MMA subproblem, shared by the CUDA kernels (pmb_optim.cu) and by
// the host harness the CPU tests build from it (tests/mma_host_harness.cpp) -- plain C++, no CUDA types.
//
// Restates, variable by variable, the n-sized expressions of pymoto/common/mma.py: asymptote / bound / P, Q set-up
// (mmasub, :170-224) and the primal-dual Newton iteration of subsolv (:246-474).  The m-sized quantities (y, z, lam, mu,
// zet, s and the (m+1)x(m+1) Newton system) stay on the host exactly as in the reference.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MMA_HD __host__ __device__ __forceinline__
#else
#define MMA_HD inline
#endif

#ifndef PMB_MMA_MAXM
#define PMB_MMA_MAXM 6  // general constraints m handled by the device path; P and Q have m+1 <= 7 rows
#endif

struct MmaVar {  // everything of ONE design variable the Newton kernels need
  double x, xsi, eta, low, upp, alfa, beta;
  double P[PMB_MMA_MAXM + 1], Q[PMB_MMA_MAXM + 1];
};

struct MmaSmall {  // m-sized host vectors passed to kernels by value
  double v[PMB_MMA_MAXM + 1];
};

// mma.py:129-140: offset *= asyincr where (x-xold1)(xold1-xold2) > 0, *= asydecr where < 0, then clip
MMA_HD double mma_offset_update(double offset, double x, double xold1, double xold2, double asyincr, double asydecr, double asybound) {
  const double zzz = (x - xold1) * (xold1 - xold2);
  if (zzz > 0.0) offset *= asyincr;
  if (zzz < 0.0) offset *= asydecr;
  return fmin(fmax(offset, 1.0 / (asybound * asybound)), asybound);
}

// mma.py:178-217 for one variable; rows 0..m of dg in dgj.  version: 1987 or 2007.  Returns 1/shift.
template <int M>
MMA_HD double mma_setup_pt(double xval, const double* dgj, double offset, double xmin, double xmax, double move, double albefa,
                           const MmaSmall& rho, int version, MmaVar& v) {
  const double dxr = xmax - xmin;
  const double shift = offset * dxr;
  v.low = xval - shift;
  v.upp = xval + shift;
  v.alfa = fmax(fmax(v.low + albefa * shift, xval - move * dxr), xmin);
  v.beta = fmin(fmin(v.upp - albefa * shift, xval + move * dxr), xmax);
  const double dx2 = shift * shift;
  for (int i = 0; i <= M; ++i) {
    const double gp = fmax(dgj[i], 0.0), gm = fmax(-dgj[i], 0.0);
    if (version == 1987) {
      v.P[i] = dx2 * gp;
      v.Q[i] = dx2 * gm;
    } else {
      v.P[i] = dx2 * (1.001 * gp + 0.001 * gm + rho.v[i] / dxr);
      v.Q[i] = dx2 * (0.001 * gp + 1.001 * gm + rho.v[i] / dxr);
    }
  }
  // subsolv start point (:288-293)
  v.x = fmin(fmax(xval, v.alfa + 1e-10), v.beta - 1e-10);
  v.xsi = fmax(1.0 / (v.x - v.alfa), 1.0);
  v.eta = fmax(1.0 / (v.beta - v.x), 1.0);
  return 1.0 / shift;
}

template <int M>
MMA_HD void mma_plam_qlam(const MmaVar& v, const MmaSmall& lam, double& plam, double& qlam) {
  double ps = 0.0, qs = 0.0;
  for (int i = 0; i < M; ++i) {
    ps += lam.v[i] * v.P[i + 1];
    qs += lam.v[i] * v.Q[i + 1];
  }
  plam = v.P[0] + ps;
  qlam = v.Q[0] + qs;
}

// residual contributions of one variable (:313-336 / :439-459): squares of rex, rexsi, reeta; gvec_i += P_i/ux1 + Q_i/xl1
template <int M>
MMA_HD void mma_resid_pt(const MmaVar& v, const MmaSmall& lam, double epsi, double& sumsq, double& maxsq, double* gvec) {
  const double ux1 = v.upp - v.x, xl1 = v.x - v.low;
  double plam, qlam;
  mma_plam_qlam<M>(v, lam, plam, qlam);
  const double dpsidx = plam / (ux1 * ux1) - qlam / (xl1 * xl1);
  const double rex = dpsidx - v.xsi + v.eta;
  const double rexsi = v.xsi * (v.x - v.alfa) - epsi;
  const double reeta = v.eta * (v.beta - v.x) - epsi;
  const double a = rex * rex, b = rexsi * rexsi, c = reeta * reeta;
  sumsq += a + b + c;
  maxsq = fmax(maxsq, fmax(a, fmax(b, c)));
  for (int i = 0; i < M; ++i) gvec[i] += v.P[i + 1] / ux1 + v.Q[i + 1] / xl1;
}

// Newton right-hand side and diagonal of one variable (:349-383): delx, diagx, GG_i, and gvec_i terms
template <int M>
MMA_HD void mma_newton_pt(const MmaVar& v, const MmaSmall& lam, double epsi, double& delx, double& diagx, double* GG, double* gterm) {
  const double ux1 = v.upp - v.x, xl1 = v.x - v.low;
  const double ux2 = ux1 * ux1, xl2 = xl1 * xl1;
  const double uxinv1 = 1.0 / ux1, xlinv1 = 1.0 / xl1, uxinv2 = 1.0 / ux2, xlinv2 = 1.0 / xl2;
  double plam, qlam;
  mma_plam_qlam<M>(v, lam, plam, qlam);
  for (int i = 0; i < M; ++i) {
    GG[i] = v.P[i + 1] * uxinv2 - v.Q[i + 1] * xlinv2;
    gterm[i] = v.P[i + 1] * uxinv1 + v.Q[i + 1] * xlinv1;
  }
  const double dpsidx = plam / ux2 - qlam / xl2;
  delx = dpsidx - epsi / (v.x - v.alfa) + epsi / (v.beta - v.x);
  diagx = 2.0 * (plam / (ux1 * ux2) + qlam / (xl1 * xl2)) + v.xsi / (v.x - v.alfa) + v.eta / (v.beta - v.x);
}

// Newton direction of one variable (:401-406) and its step-length candidates (:409-423), all as "larger is worse":
// cand[0] = -dxsi/xsi, cand[1] = -deta/eta, cand[2] = -dx/(x-alfa), cand[3] = dx/(beta-x)
template <int M>
MMA_HD void mma_dir_pt(const MmaVar& v, const MmaSmall& lam, const MmaSmall& dlam, double epsi, double& dx, double& dxsi, double& deta,
                       double* cand) {
  double delx, diagx, GG[M], gterm[M];
  mma_newton_pt<M>(v, lam, epsi, delx, diagx, GG, gterm);
  double dg = 0.0;
  for (int i = 0; i < M; ++i) dg += dlam.v[i] * GG[i];
  dx = -delx / diagx - dg / diagx;
  dxsi = -v.xsi + epsi / (v.x - v.alfa) - (v.xsi * dx) / (v.x - v.alfa);
  deta = -v.eta + epsi / (v.beta - v.x) + (v.eta * dx) / (v.beta - v.x);
  cand[0] = -dxsi / v.xsi;
  cand[1] = -deta / v.eta;
  cand[2] = -dx / (v.x - v.alfa);
  cand[3] = dx / (v.beta - v.x);
}

// ---- GCMMA (mma.py:104-160, 232-242)
// initial conservativeness rho_i = 0.1/n sum_j dx_j |dg_ij| (:151): the term of one variable and one response
MMA_HD double mma_rho_term(double dgij, double xmin, double xmax) { return (xmax - xmin) * fabs(dgij); }

// value of the convex approximations at the subproblem solution (:236, before "- rhs"): est_i += P_i/(upp-x) + Q_i/(x-low) for
// i = 0..M, and the variable's term of the normalised step measure dk (:239)
template <int M>
MMA_HD double mma_estimate_pt(const MmaVar& v, double xval, double xmin, double xmax, double* est) {
  const double ux1 = v.upp - v.x, xl1 = v.x - v.low;
  for (int i = 0; i <= M; ++i) est[i] += v.P[i] / ux1 + v.Q[i] / xl1;
  const double dxm = v.x - xval;
  return (v.upp - v.low) * (dxm * dxm) / (ux1 * xl1 * (xmax - xmin));
}
