// Shared device/host helpers for libpmb (sm_100a): grid geometry, closed-form CSR offsets, error handling.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include "../../include/pmb.h"

#define PMB_HD __host__ __device__ __forceinline__

// ---------------------------------------------------------------------------------------------------------
// Geometry of a structured voxel grid slab.  Numbering follows pymoto/common/domain.py:200-251 (x fastest,
// then y, then z); the CSR pattern is the 27-/9-point block stencil of pymoto/modules/assembly.py:130-206.
// ---------------------------------------------------------------------------------------------------------
struct Geo {
  int NX, NY, NZ;   // nodes per direction
  int nx, ny, nzE;  // elements per direction (nzE = max(nz, 1))
  int dim3;         // 1 if 3-D
  int ndof, kz0, nzl;
  long long Sx, Sy;  // neighbour-slot totals along x and y: 3*NX-2, 3*NY-2
  long long plane;   // nodes per z-plane
  long long bo0;     // global block offset of the first owned node
  long long nOwned;  // owned nodes
};

// number of neighbours of index i in [i-1, i+1] clipped to [0, M-1]
PMB_HD int cnt1(int i, int M) { return 3 - (i == 0) - (i == M - 1); }
// exclusive prefix sum of cnt1 (valid for 0 <= i <= M)
PMB_HD long long pre1(int i, int M) { return 3LL * i - (i > 0) - (i == M); }

PMB_HD long long block_offset(const Geo& g, int i, int j, int k) {
  return pre1(k, g.NZ) * g.Sy * g.Sx + (long long)cnt1(k, g.NZ) * (pre1(j, g.NY) * g.Sx + (long long)cnt1(j, g.NY) * pre1(i, g.NX));
}

static inline Geo make_geo(const pmb_grid* p) {
  Geo g;
  g.nx = p->nx; g.ny = p->ny; g.nzE = p->nz > 0 ? p->nz : 1;
  g.dim3 = p->nz > 0;
  g.NX = p->nx + 1; g.NY = p->ny + 1; g.NZ = p->nz + 1;
  g.ndof = p->ndof; g.kz0 = p->kz0; g.nzl = p->nzl;
  g.Sx = 3LL * g.NX - 2; g.Sy = 3LL * g.NY - 2;
  g.plane = (long long)g.NX * g.NY;
  g.bo0 = pre1(g.kz0, g.NZ) * g.Sy * g.Sx;
  g.nOwned = g.plane * g.nzl;
  return g;
}

// local (slab) node index -> global (i, j, k)
__device__ __forceinline__ void node_ijk(const Geo& g, long long ln, int& i, int& j, int& k) {
  int kk = (int)(ln / g.plane);
  int rem = (int)(ln - (long long)kk * g.plane);
  j = rem / g.NX;
  i = rem - j * g.NX;
  k = kk + g.kz0;
}

// entry offset (in doubles, relative to the slab's first entry) of the first entry of local node ln;
// also valid for ln == nOwned (one past the end)
__device__ __forceinline__ long long node_entry_offset(const Geo& g, long long ln) {
  int i, j, k;
  node_ijk(g, ln, i, j, k);
  return (long long)(g.ndof * g.ndof) * (block_offset(g, i, j, k) - g.bo0);
}

// ---------------------------------------------------------------------------------------------------------
// error handling: never throw / exit across the C ABI
// ---------------------------------------------------------------------------------------------------------
extern thread_local char pmb_err_buf[512];
int pmb_set_error(const char* fmt, ...);

#define PMB_CHECK_LAUNCH(name)                                                        \
  do {                                                                                \
    cudaError_t e_ = cudaGetLastError();                                              \
    if (e_ != cudaSuccess) return pmb_set_error("%s: %s", name, cudaGetErrorString(e_)); \
  } while (0)

#define PMB_REQUIRE(cond, ...)                         \
  do {                                                 \
    if (!(cond)) return pmb_set_error(__VA_ARGS__);    \
  } while (0)

static inline int validate_grid(const pmb_grid* g, const char* who) {
  if (!g) return pmb_set_error("%s: grid is NULL", who);
  if (g->nx < 1 || g->ny < 1 || g->nz < 0) return pmb_set_error("%s: invalid grid %d x %d x %d", who, g->nx, g->ny, g->nz);
  if (g->ndof < 1 || g->ndof > 3) return pmb_set_error("%s: ndof=%d not in 1..3", who, g->ndof);
  if (g->kz0 < 0 || g->nzl < 1 || g->kz0 + g->nzl > g->nz + 1)
    return pmb_set_error("%s: slab [%d, %d) outside node planes [0, %d)", who, g->kz0, g->kz0 + g->nzl, g->nz + 1);
  return 0;
}
