// libpmb: error plumbing and size queries.
#include <cstdarg>
#include "pmb_common.cuh"

thread_local char pmb_err_buf[512] = "";

int pmb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(pmb_err_buf, sizeof(pmb_err_buf), fmt, ap);
  va_end(ap);
  return 1;
}

extern "C" const char* pmb_last_error(void) { return pmb_err_buf; }
extern "C" int pmb_version(void) { return 100; }

extern "C" long long pmb_nnz(const pmb_grid* p) {
  if (validate_grid(p, "pmb_nnz")) return -1;
  Geo g = make_geo(p);
  long long b1 = pre1(g.kz0 + g.nzl, g.NZ) * g.Sy * g.Sx;
  return (long long)(g.ndof * g.ndof) * (b1 - g.bo0);
}

extern "C" long long pmb_nrows(const pmb_grid* p) {
  if (validate_grid(p, "pmb_nrows")) return -1;
  Geo g = make_geo(p);
  return g.nOwned * g.ndof;
}
