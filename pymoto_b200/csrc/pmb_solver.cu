// libpmb: C-side driver of the whole linear solve (pmb_pcg_solve) -- preconditioned CG with one geometric
// multigrid V-cycle per iteration, every kernel launched from C and the host polling ONE scalar (the residual norm) per
// iteration.  It issues exactly the launches of the Python driver (pymoto_b200/solvers.py: CG._solve1,
// GeometricMultigrid._solve_eager) in the same order, so iterates and iteration counts are bit-identical; it removes the
// interpreter / ctypes overhead (~10 us per launch), which dominates small grids.
//
// Reference: pymoto/solvers/iterative.py:340-403 (CG.solve) and :222-256 (GeometricMultigrid.solve).  Single GPU.
#include <cmath>
#include "pmb_common.cuh"

static long long level_rows(const pmb_grid& g) {
  return (long long)(g.nx + 1) * (g.ny + 1) * g.nzl * g.ndof;
}

// y = A x | b - A x | x + w (b - A x) / diag on level l (level 0 from the element densities when a generator is given)
static int level_apply(const pmb_mg_desc* mg, int l, int mode, const double* x, const double* b, double w, double* y,
                       const double* dotv, double* dot_out, double* ws, void* st) {
  const pmb_mg_level& L = mg->level[l];
  if (l == 0 && mg->gen.Ke_host)
    return pmb_elem_spmv(&L.grid, mode, &mg->gen, x, b, L.diag, w, y, dotv, dot_out, ws, st);
  return pmb_spmv(&L.grid, mode, L.A, x, b, L.diag, w, y, dotv, dot_out, ws, st);
}

// one V-cycle on levels l .. nlevels-1 + the dense coarsest solve; the result is left in one of the level's scratch vectors
static int vcycle(const pmb_mg_desc* mg, int l, const double* rhs, double** out, void* st) {
  const pmb_mg_level& L = mg->level[l];
  const long long n = level_rows(L.grid);
  const pmb_grid* gc = (l + 1 < mg->nlevels) ? &mg->level[l + 1].grid : &mg->coarse_grid;
  double *u = L.u, *u2 = L.u2;
  if (pmb_smooth0(n, L.w, rhs, L.diag, u, st)) return 1;  // first pre-sweep from zero: u = w r / D
  for (int k = 1; k < L.smooth_steps; ++k) {
    if (level_apply(mg, l, PMB_JACOBI, u, rhs, L.w, u2, nullptr, nullptr, nullptr, st)) return 1;
    double* tmp = u; u = u2; u2 = tmp;
  }
  if (level_apply(mg, l, PMB_RESIDUAL, u, rhs, 0.0, L.t, nullptr, nullptr, nullptr, st)) return 1;
  if (pmb_restrict(&L.grid, gc, L.t, L.rc, st)) return 1;
  double* uc = nullptr;
  if (l + 1 < mg->nlevels) {
    if (vcycle(mg, l + 1, L.rc, &uc, st)) return 1;
  } else {
    const int nc = (int)level_rows(mg->coarse_grid);
    if (pmb_dense_gemv(nc, mg->coarse_inv, L.rc, mg->coarse_out, st)) return 1;
    uc = mg->coarse_out;
  }
  if (pmb_prolong_add(&L.grid, gc, uc, u, st)) return 1;
  for (int k = 0; k < L.smooth_steps; ++k) {
    if (level_apply(mg, l, PMB_JACOBI, u, rhs, L.w, u2, nullptr, nullptr, nullptr, st)) return 1;
    double* tmp = u; u = u2; u2 = tmp;
  }
  *out = u;
  return 0;
}

static int check_desc(const pmb_mg_desc* mg, const char* who) {
  if (!mg) return pmb_set_error("%s: descriptor is NULL", who);
  PMB_REQUIRE(mg->nlevels >= 1 && mg->nlevels <= PMB_MAX_LEVELS, "%s: nlevels=%d not in 1..%d", who, mg->nlevels, PMB_MAX_LEVELS);
  for (int l = 0; l < mg->nlevels; ++l) {
    const pmb_mg_level& L = mg->level[l];
    if (validate_grid(&L.grid, who)) return 1;
    PMB_REQUIRE(L.grid.kz0 == 0 && L.grid.nzl == L.grid.nz + 1, "%s: level %d is a slab (the C driver is single-GPU)", who, l);
    PMB_REQUIRE((L.A || (l == 0 && mg->gen.Ke_host)) && L.diag && L.u && L.u2 && L.t && L.rc, "%s: NULL pointer in level %d", who, l);
    PMB_REQUIRE(L.smooth_steps >= 1 && L.w > 0.0 && L.w <= 1.0, "%s: level %d smoother (steps %d, w %g)", who, l, L.smooth_steps, L.w);
    const pmb_grid& c = (l + 1 < mg->nlevels) ? mg->level[l + 1].grid : mg->coarse_grid;
    PMB_REQUIRE(L.grid.nx == 2 * c.nx && L.grid.ny == 2 * c.ny && L.grid.nz == 2 * c.nz && L.grid.ndof == c.ndof,
                "%s: level %d is not followed by its 2:1 coarsening", who, l);
  }
  if (validate_grid(&mg->coarse_grid, who)) return 1;
  PMB_REQUIRE(mg->coarse_inv && mg->coarse_out, "%s: coarsest-level inverse / output missing", who);
  PMB_REQUIRE(!mg->gen.Ke_host || mg->gen.s, "%s: generator without element scaling vector", who);
  return 0;
}

extern "C" int pmb_vcycle(const pmb_mg_desc* mg, const double* r, double* z, void* stream) {
  if (check_desc(mg, "pmb_vcycle")) return 1;
  PMB_REQUIRE(r && z, "pmb_vcycle: NULL pointer argument");
  double* out = nullptr;
  if (vcycle(mg, 0, r, &out, stream)) return 1;
  const long long n = level_rows(mg->level[0].grid);
  cudaError_t e = cudaMemcpyAsync(z, out, sizeof(double) * n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) return pmb_set_error("pmb_vcycle: %s", cudaGetErrorString(e));
  return 0;
}

static int read_scalar(const double* dev, double* host, cudaStream_t st) {
  cudaError_t e = cudaMemcpyAsync(host, dev, sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return pmb_set_error("pmb_pcg_solve: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int pmb_pcg_solve(const pmb_mg_desc* mg, const double* b, double* x, double* r, double* q, double* p, double tol,
                             int maxit, int restart, double* scal, double* ws_red, double* ws_spmv, int* iters, double* relres,
                             void* stream) {
  if (check_desc(mg, "pmb_pcg_solve")) return 1;
  PMB_REQUIRE(b && x && r && q && p && scal && ws_red && ws_spmv && iters && relres, "pmb_pcg_solve: NULL pointer argument");
  PMB_REQUIRE(maxit >= 0 && restart >= 1, "pmb_pcg_solve: maxit %d / restart %d", maxit, restart);
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = level_rows(mg->level[0].grid);
  // device scalars: [0..2] = p.q, p.r, q.r   [4] = r.r   [5] = b.b   [6] = z.z   [7] = q.z
  double *d3 = scal, *rr = scal + 4, *zz = scal + 6, *qz = scal + 7;
  const pmb_coef one = {1.0, nullptr, nullptr, 0}, zero = {0.0, nullptr, nullptr, 0};

  if (level_apply(mg, 0, PMB_RESIDUAL, x, b, 0.0, r, nullptr, nullptr, nullptr, stream)) return 1;
  if (pmb_dots(n, 2, r, r, b, b, nullptr, nullptr, nullptr, nullptr, rr, ws_red, stream)) return 1;  // rr[0] = r.r, rr[1] = b.b
  double h[2];
  if (read_scalar(rr, &h[0], st) || read_scalar(rr + 1, &h[1], st)) return 1;
  const double bnorm = sqrt(h[1]);
  double tval = sqrt(h[0]) / bnorm;
  *iters = 0;
  *relres = tval;
  if (tval <= tol) return 0;

  double* z = nullptr;
  if (vcycle(mg, 0, r, &z, stream)) return 1;
  if (pmb_dots(n, 1, z, z, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, zz, ws_red, stream)) return 1;
  const pmb_coef inv_norm = {1.0, nullptr, zz, 1};
  if (pmb_lincomb(n, p, inv_norm, z, zero, nullptr, stream)) return 1;  // p = z / |z|
  for (int i = 0; i < maxit; ++i) {
    if (level_apply(mg, 0, PMB_SPMV, p, nullptr, 0.0, q, r, d3, ws_spmv, stream)) return 1;  // q = A p; d3 = [q.p, p.r, q.r]
    const double *pq = d3, *pr = d3 + 1;
    if (i % restart == 0) {  // explicit residual (iterative.py:369-376), including the first iteration
      if (pmb_cg_xr_update(n, x, nullptr, p, nullptr, pr, pq, nullptr, nullptr, stream)) return 1;
      if (level_apply(mg, 0, PMB_RESIDUAL, x, b, 0.0, r, nullptr, nullptr, nullptr, stream)) return 1;
      if (pmb_dots(n, 1, r, r, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, rr, ws_red, stream)) return 1;
    } else {
      if (pmb_cg_xr_update(n, x, r, p, q, pr, pq, rr, ws_red, stream)) return 1;
    }
    if (read_scalar(rr, &h[0], st)) return 1;  // the only host synchronisation of the iteration
    tval = sqrt(h[0]) / bnorm;
    *iters = i + 1;
    *relres = tval;
    if (tval <= tol) break;
    if (!std::isfinite(tval)) return pmb_set_error("pmb_pcg_solve: residual became non-finite in iteration %d (singular operator or preconditioner)", i);
    if (vcycle(mg, 0, r, &z, stream)) return 1;
    if (pmb_dots(n, 1, q, z, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, qz, ws_red, stream)) return 1;
    const pmb_coef beta = {-1.0, qz, pq, 0};
    if (pmb_lincomb(n, p, one, z, beta, p, stream)) return 1;  // p = z - (q.z / p.q) p
  }
  return 0;
}
