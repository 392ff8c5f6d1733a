// libpmb: C-side driver of the whole linear solve (pmb_pcg_solve) -- preconditioned CG with one geometric
// multigrid V-cycle per iteration, every kernel launched from C and the host polling ONE scalar (the residual norm) per
// iteration.  It issues exactly the launches of the Python driver (pymoto_b200/solvers.py: CG._solve1,
// GeometricMultigrid._solve_eager) in the same order, so iterates and iteration counts are bit-identical; it removes the
// interpreter / ctypes overhead (~10 us per launch), which dominates small grids.
//
// Reference: pymoto/solvers/iterative.py:340-403 (CG.solve) and :222-256 (GeometricMultigrid.solve).  Single GPU.
#include <cmath>
#include <cstdlib>
#include <new>
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

// ---------------------------------------------------------------------------------------------------------
// Plan = descriptor + bound vectors + (lazily captured) CUDA graph of one CG iteration.
// Between two host polls a non-restart iteration is always the same launch sequence on the same addresses:
//     z = V-cycle(r);  q.z;  p = z - (q.z / p.q) p;  q = A p (+ p.q, p.r, q.r);  x += a p, r -= a q (+ r.r)
// (~75 kernels at 64x32x32, where launch latency, not bandwidth, sets the pace).  The plan executes that sequence once
// eagerly (first-use kernel attributes), captures it into a graph on its second occurrence and replays it afterwards -- in
// this solve and in every later solve through the same plan (the graph reads the operator VALUES from memory, so it stays
// valid across design iterations as long as the addresses, the element matrix and the layout in the descriptor do).
// ---------------------------------------------------------------------------------------------------------
struct pmb_pcg_plan {
  pmb_mg_desc mg;
  const double* b;
  double *x, *r, *q, *p, *scal, *ws_red, *ws_spmv;
  cudaGraphExec_t exec;
  cudaStream_t own;      // the solve runs on the plan's own stream (the caller's may be the legacy default stream, which cannot
  cudaEvent_t ev_in, ev_out;   // be captured), ordered after / before the caller's stream by two events
  int eager_bodies;   // bodies executed eagerly so far
  int capture_failed;
  long long graph_replays;
};

struct PcgBufs {
  const double* b;
  double *x, *r, *q, *p, *scal, *ws_red, *ws_spmv;
};

// the body between two host polls for a non-restart iteration; *zout = where the V-cycle leaves z
static int pcg_body(const pmb_mg_desc* mg, const PcgBufs& B, long long n, void* stream) {
  double *d3 = B.scal, *rr = B.scal + 4, *qz = B.scal + 7;
  const double *pq = d3, *pr = d3 + 1;
  const pmb_coef one = {1.0, nullptr, nullptr, 0};
  double* z = nullptr;
  if (vcycle(mg, 0, B.r, &z, stream)) return 1;
  if (pmb_dots(n, 1, B.q, z, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, qz, B.ws_red, stream)) return 1;
  const pmb_coef beta = {-1.0, qz, pq, 0};
  if (pmb_lincomb(n, B.p, one, z, beta, B.p, stream)) return 1;  // p = z - (q.z / p.q) p
  if (level_apply(mg, 0, PMB_SPMV, B.p, nullptr, 0.0, B.q, B.r, d3, B.ws_spmv, stream)) return 1;  // q = A p; d3 = [q.p, p.r, q.r]
  if (pmb_cg_xr_update(n, B.x, B.r, B.p, B.q, pr, pq, rr, B.ws_red, stream)) return 1;
  return 0;
}

static int pcg_run(const pmb_mg_desc* mg, const PcgBufs& B, double tol, int maxit, int restart, int* iters, double* relres,
                   pmb_pcg_plan* plan, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = level_rows(mg->level[0].grid);
  // device scalars: [0..2] = p.q, p.r, q.r   [4] = r.r   [5] = b.b   [6] = z.z   [7] = q.z
  double *d3 = B.scal, *rr = B.scal + 4, *zz = B.scal + 6;
  const pmb_coef zero = {0.0, nullptr, nullptr, 0};

  if (level_apply(mg, 0, PMB_RESIDUAL, B.x, B.b, 0.0, B.r, nullptr, nullptr, nullptr, stream)) return 1;
  if (pmb_dots(n, 2, B.r, B.r, B.b, B.b, nullptr, nullptr, nullptr, nullptr, rr, B.ws_red, stream)) return 1;  // rr[0] = r.r, rr[1] = b.b
  double h[2];
  if (read_scalar(rr, &h[0], st) || read_scalar(rr + 1, &h[1], st)) return 1;
  const double bnorm = sqrt(h[1]);
  double tval = sqrt(h[0]) / bnorm;
  *iters = 0;
  *relres = tval;
  if (tval <= tol) return 0;

  double* z = nullptr;
  if (vcycle(mg, 0, B.r, &z, stream)) return 1;
  if (pmb_dots(n, 1, z, z, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, zz, B.ws_red, stream)) return 1;
  const pmb_coef inv_norm = {1.0, nullptr, zz, 1};
  if (pmb_lincomb(n, B.p, inv_norm, z, zero, nullptr, stream)) return 1;  // p = z / |z|
  // iteration 0 is a restart iteration (explicit residual, iterative.py:369-376)
  for (int i = 0; i < maxit; ++i) {
    if (i % restart == 0) {
      const double *pq = d3, *pr = d3 + 1;
      if (i > 0) {  // direction update of the previous iteration (for i = 0 p was set above)
        double* qz = B.scal + 7;
        const pmb_coef one = {1.0, nullptr, nullptr, 0};
        if (vcycle(mg, 0, B.r, &z, stream)) return 1;
        if (pmb_dots(n, 1, B.q, z, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, qz, B.ws_red, stream)) return 1;
        const pmb_coef beta = {-1.0, qz, pq, 0};
        if (pmb_lincomb(n, B.p, one, z, beta, B.p, stream)) return 1;
      }
      if (level_apply(mg, 0, PMB_SPMV, B.p, nullptr, 0.0, B.q, B.r, d3, B.ws_spmv, stream)) return 1;
      if (pmb_cg_xr_update(n, B.x, nullptr, B.p, nullptr, pr, pq, nullptr, nullptr, stream)) return 1;
      if (level_apply(mg, 0, PMB_RESIDUAL, B.x, B.b, 0.0, B.r, nullptr, nullptr, nullptr, stream)) return 1;
      if (pmb_dots(n, 1, B.r, B.r, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, rr, B.ws_red, stream)) return 1;
    } else if (plan && plan->exec) {
      cudaError_t e = cudaGraphLaunch(plan->exec, st);
      if (e != cudaSuccess) return pmb_set_error("pmb_pcg_plan_solve: cudaGraphLaunch: %s", cudaGetErrorString(e));
      plan->graph_replays++;
    } else if (plan && !plan->capture_failed && plan->eager_bodies >= 1) {
      // second occurrence of the body: capture it while it is being issued, then launch the captured graph
      cudaGraph_t graph = nullptr;
      cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
      int rc = 1;
      if (e == cudaSuccess) {
        rc = pcg_body(mg, B, n, stream);
        e = cudaStreamEndCapture(st, &graph);
      }
      if (e == cudaSuccess && rc == 0 && graph) e = cudaGraphInstantiate(&plan->exec, graph, 0);
      if (graph) cudaGraphDestroy(graph);
      if (e != cudaSuccess || rc != 0 || !plan->exec) {   // not capturable on this driver / stream: stay eager
        if (getenv("PMB_DEBUG"))
          fprintf(stderr, "pmb_pcg_plan_solve: iteration graph not captured (cuda: %s; body rc %d: %s)\n", cudaGetErrorString(e), rc,
                  rc ? pmb_last_error() : "-");
        plan->exec = nullptr;
        plan->capture_failed = 1;
        cudaGetLastError();
        if (pcg_body(mg, B, n, stream)) return 1;
      } else {
        e = cudaGraphLaunch(plan->exec, st);
        if (e != cudaSuccess) return pmb_set_error("pmb_pcg_plan_solve: cudaGraphLaunch: %s", cudaGetErrorString(e));
        plan->graph_replays++;
      }
    } else {
      if (pcg_body(mg, B, n, stream)) return 1;
      if (plan) plan->eager_bodies++;
    }
    if (read_scalar(rr, &h[0], st)) return 1;  // the only host synchronisation of the iteration
    tval = sqrt(h[0]) / bnorm;
    *iters = i + 1;
    *relres = tval;
    if (tval <= tol) break;
    if (!std::isfinite(tval)) return pmb_set_error("pmb_pcg_solve: residual became non-finite in iteration %d (singular operator or preconditioner)", i);
  }
  return 0;
}

extern "C" int pmb_pcg_solve(const pmb_mg_desc* mg, const double* b, double* x, double* r, double* q, double* p, double tol,
                             int maxit, int restart, double* scal, double* ws_red, double* ws_spmv, int* iters, double* relres,
                             void* stream) {
  if (check_desc(mg, "pmb_pcg_solve")) return 1;
  PMB_REQUIRE(b && x && r && q && p && scal && ws_red && ws_spmv && iters && relres, "pmb_pcg_solve: NULL pointer argument");
  PMB_REQUIRE(maxit >= 0 && restart >= 1, "pmb_pcg_solve: maxit %d / restart %d", maxit, restart);
  const PcgBufs B = {b, x, r, q, p, scal, ws_red, ws_spmv};
  return pcg_run(mg, B, tol, maxit, restart, iters, relres, nullptr, stream);
}

extern "C" int pmb_pcg_plan_create(const pmb_mg_desc* mg, const double* b, double* x, double* r, double* q, double* p, double* scal,
                                   double* ws_red, double* ws_spmv, pmb_pcg_plan** plan) {
  if (check_desc(mg, "pmb_pcg_plan_create")) return 1;
  PMB_REQUIRE(b && x && r && q && p && scal && ws_red && ws_spmv && plan, "pmb_pcg_plan_create: NULL pointer argument");
  pmb_pcg_plan* P = new (std::nothrow) pmb_pcg_plan();
  PMB_REQUIRE(P, "pmb_pcg_plan_create: out of host memory");
  P->mg = *mg;
  P->b = b, P->x = x, P->r = r, P->q = q, P->p = p, P->scal = scal, P->ws_red = ws_red, P->ws_spmv = ws_spmv;
  P->exec = nullptr;
  P->eager_bodies = 0, P->capture_failed = 0, P->graph_replays = 0;
  cudaError_t e = cudaStreamCreateWithFlags(&P->own, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&P->ev_in, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&P->ev_out, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    delete P;
    return pmb_set_error("pmb_pcg_plan_create: %s", cudaGetErrorString(e));
  }
  *plan = P;
  return 0;
}

extern "C" int pmb_pcg_plan_solve(pmb_pcg_plan* plan, double tol, int maxit, int restart, int* iters, double* relres, void* stream) {
  PMB_REQUIRE(plan && iters && relres, "pmb_pcg_plan_solve: NULL pointer argument");
  PMB_REQUIRE(maxit >= 0 && restart >= 1, "pmb_pcg_plan_solve: maxit %d / restart %d", maxit, restart);
  const PcgBufs B = {plan->b, plan->x, plan->r, plan->q, plan->p, plan->scal, plan->ws_red, plan->ws_spmv};
  cudaStream_t user = (cudaStream_t)stream;
  cudaError_t e = cudaEventRecord(plan->ev_in, user);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(plan->own, plan->ev_in, 0);
  if (e != cudaSuccess) return pmb_set_error("pmb_pcg_plan_solve: %s", cudaGetErrorString(e));
  const int rc = pcg_run(&plan->mg, B, tol, maxit, restart, iters, relres, plan, (void*)plan->own);
  e = cudaEventRecord(plan->ev_out, plan->own);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(user, plan->ev_out, 0);
  if (rc) return rc;
  if (e != cudaSuccess) return pmb_set_error("pmb_pcg_plan_solve: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" long long pmb_pcg_plan_graph_replays(const pmb_pcg_plan* plan) { return plan ? plan->graph_replays : -1; }

extern "C" int pmb_pcg_plan_destroy(pmb_pcg_plan* plan) {
  if (!plan) return 0;
  if (plan->exec) cudaGraphExecDestroy(plan->exec);
  cudaStreamSynchronize(plan->own);
  cudaEventDestroy(plan->ev_in);
  cudaEventDestroy(plan->ev_out);
  cudaStreamDestroy(plan->own);
  delete plan;
  return 0;
}
