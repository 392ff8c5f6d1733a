// libpmb: slab communication entry points of SURVEY.md 8b (pmb_comm_init / pmb_halo_exchange / pmb_allreduce) over NCCL, so
// that a host without torch.distributed can drive the z-slab path: one communicator handle per process / GPU, created
// from an ncclUniqueId the caller distributes (file, socket, MPI ...).
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- inside a PyTorch process that resolves to the copy torch already
// loaded), so libpmb.so itself carries no NCCL link dependency and loads on machines without it; the entry points then
// fail with a clear message.  The reference has no communication layer (single process); the exchange steps are the ones
// listed in pymoto_b200/slab.py: neighbour halo planes (ncclSend / ncclRecv pairs inside one group) and the sum
// all-reduce of the CG / LDAS dot products.
#include <dlfcn.h>
#include "pmb_common.cuh"

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_FLOAT64 = 8, NCCL_SUM = 0 };  // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since NCCL 2.0)

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names)
      if ((api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (api.lib) {
      api.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(api.lib, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(api.lib, "ncclCommInitRank");
      api.CommDestroy = (int (*)(ncclComm_t))dlsym(api.lib, "ncclCommDestroy");
      api.GroupStart = (int (*)())dlsym(api.lib, "ncclGroupStart");
      api.GroupEnd = (int (*)())dlsym(api.lib, "ncclGroupEnd");
      api.Send = (int (*)(const void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(api.lib, "ncclSend");
      api.Recv = (int (*)(void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(api.lib, "ncclRecv");
      api.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
      api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.GroupStart || !api.GroupEnd || !api.Send || !api.Recv ||
          !api.AllReduce) {
        dlclose(api.lib);
        api.lib = nullptr;
      }
    }
  }
  return api.lib ? &api : nullptr;
}

int nccl_fail(const char* who, NcclApi* a, int rc) {
  return pmb_set_error("%s: NCCL error %d (%s)", who, rc, a->GetErrorString ? a->GetErrorString(rc) : "?");
}
}  // namespace

struct pmb_comm {
  ncclComm_t comm;
  int rank, nranks;
};

extern "C" int pmb_comm_unique_id(void* id128) {
  PMB_REQUIRE(id128, "pmb_comm_unique_id: NULL output");
  NcclApi* a = nccl_api();
  PMB_REQUIRE(a, "pmb_comm_unique_id: libnccl.so.2 not found");
  ncclUniqueId id;
  const int rc = a->GetUniqueId(&id);
  if (rc) return nccl_fail("pmb_comm_unique_id", a, rc);
  memcpy(id128, &id, sizeof(id));
  return 0;
}

extern "C" int pmb_comm_init(const void* id128, int rank, int nranks, pmb_comm** out) {
  PMB_REQUIRE(id128 && out, "pmb_comm_init: NULL pointer argument");
  PMB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "pmb_comm_init: rank %d of %d", rank, nranks);
  NcclApi* a = nccl_api();
  PMB_REQUIRE(a, "pmb_comm_init: libnccl.so.2 not found");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  pmb_comm* c = new pmb_comm{nullptr, rank, nranks};
  const int rc = a->CommInitRank(&c->comm, nranks, id, rank);
  if (rc) {
    delete c;
    return nccl_fail("pmb_comm_init", a, rc);
  }
  *out = c;
  return 0;
}

extern "C" int pmb_comm_destroy(pmb_comm* c) {
  if (!c) return 0;
  NcclApi* a = nccl_api();
  if (a && c->comm) a->CommDestroy(c->comm);
  delete c;
  return 0;
}

extern "C" int pmb_comm_rank(const pmb_comm* c) { return c ? c->rank : -1; }
extern "C" int pmb_comm_size(const pmb_comm* c) { return c ? c->nranks : -1; }

// Fill the halo planes of a padded device buffer: base[own_offset, own_offset + own_len) are the owned doubles, the n doubles
// below / above them are the halos.  lower != 0: receive my lower halo (the rank below sends its top n owned doubles);
// upper != 0: receive my upper halo.  Every rank calls it with the same flags (z-neighbours = rank +- 1).
extern "C" int pmb_halo_exchange(pmb_comm* c, double* base, long long own_offset, long long own_len, long long n, int lower, int upper,
                                 void* stream) {
  PMB_REQUIRE(c && base, "pmb_halo_exchange: NULL pointer argument");
  PMB_REQUIRE(n >= 0 && own_len >= n && own_offset >= n, "pmb_halo_exchange: halo of %lld doubles does not fit (offset %lld, owned %lld)", n,
              own_offset, own_len);
  if (c->nranks == 1 || n == 0) return 0;
  NcclApi* a = nccl_api();
  cudaStream_t st = (cudaStream_t)stream;
  const int lo = c->rank - 1, hi = c->rank + 1;
  double* own = base + own_offset;
  int rc = a->GroupStart();
  if (upper) {  // data flows downwards: my bottom planes go to the rank below, my upper halo comes from the rank above
    if (!rc && lo >= 0) rc = a->Send(own, (size_t)n, NCCL_FLOAT64, lo, c->comm, st);
    if (!rc && hi < c->nranks) rc = a->Recv(own + own_len, (size_t)n, NCCL_FLOAT64, hi, c->comm, st);
  }
  if (lower) {  // data flows upwards
    if (!rc && hi < c->nranks) rc = a->Send(own + own_len - n, (size_t)n, NCCL_FLOAT64, hi, c->comm, st);
    if (!rc && lo >= 0) rc = a->Recv(own - n, (size_t)n, NCCL_FLOAT64, lo, c->comm, st);
  }
  const int rc2 = a->GroupEnd();
  if (rc || rc2) return nccl_fail("pmb_halo_exchange", a, rc ? rc : rc2);
  return 0;
}

// in-place sum all-reduce of `count` device doubles (CG / LDAS dot products, compliance, volume)
extern "C" int pmb_allreduce(pmb_comm* c, double* buf, long long count, void* stream) {
  PMB_REQUIRE(c && buf && count > 0, "pmb_allreduce: invalid argument");
  if (c->nranks == 1) return 0;
  NcclApi* a = nccl_api();
  const int rc = a->AllReduce(buf, buf, (size_t)count, NCCL_FLOAT64, NCCL_SUM, c->comm, (cudaStream_t)stream);
  if (rc) return nccl_fail("pmb_allreduce", a, rc);
  return 0;
}
