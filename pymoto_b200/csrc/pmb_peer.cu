// libpmb: the exchange steps of the z-slab path as ONE launch each, over NVLink peer memory.
//
// The reference is single-process; these are the data-path exchanges of section 8e.  The buffers are
// symmetric-memory allocations mapped into every rank (peer pointers).  A launch stores straight into the neighbours' memory
// and polls its own memory for what the neighbours store -- no host involvement, no collective library, no separate barrier,
// ordinary stream work (capturable in a CUDA graph).
//
// Flag-in-data transport: every double travels as one 16-byte unit {low word, number, high word, number} written with a single
// vector store; each 8-byte half (payload word + exchange number) is written atomically, so a reader that sees the expected
// number in both halves has the payload -- there is no fence and no flag round trip on the path, the latency of an exchange is
// one launch + one NVLink store flight.  (Twice the bytes on the wire: 1.6 MB for the largest plane, 2 us on NVLink 5.)
//
// Exchange numbers instead of barrier generations: every rank issues the same sequence of exchanges, a counter in the rank's
// ordinary device memory (ctl) numbers them, and exchange s uses mailbox slot s & 1.  A slot written in exchange s is rewritten
// in exchange s + 2; the writer gets there only after finishing s + 1, in which it polled the reader's HEADER unit s + 1, which
// the reader stored after its launch s (unpack included) had ended.  That is why every launch sends a header unit to BOTH
// neighbours and polls both headers even when the data flows one way only.
#include "pmb_common.cuh"

namespace {

__device__ __forceinline__ void st_unit(double* unit, double v, unsigned num) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(unit), "r"((unsigned)b), "r"(num), "r"((unsigned)(b >> 32)), "r"(num)
               : "memory");
}

// Poll a unit of this rank's own mailbox until both halves carry `num`.  A peer that never arrives (a rank died, the ranks
// issued different sequences) must not hang the GPU for good: after ~60 s the wait gives up, the launch records the exchange
// number in ctl (results are invalid from then on; the host checks the word at its own synchronisation points) and later
// launches do not wait at all.
__device__ __forceinline__ bool ld_unit(const double* unit, unsigned num, double& v) {
  unsigned lo, f0, hi, f1, polls = 0;
  unsigned long long t_start = 0;
  for (;;) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(unit) : "memory");
    if (f0 == num && f1 == num) break;
    if ((++polls & 0x3FFu) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t_start == 0) t_start = now;
      else if (now - t_start > 60000000000ULL) return false;
    }
  }
  v = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
  return true;
}

__device__ __forceinline__ unsigned long long ld_volatile(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}

// Mailbox of a rank, in 16-byte units: [slot 0 / 1][0: from the rank below, 1: from the rank above][header, cap payload units].
// ctl words: [0] exchanges completed, [1] CTAs done with the current one, [2] number of the first exchange that timed out.
__global__ void __launch_bounds__(256) peer_halo_kernel(pmb_peer_halo h, long long n, const double* __restrict__ send_lo,
                                                         const double* __restrict__ send_hi, double* __restrict__ recv_lo,
                                                         double* __restrict__ recv_hi) {
  const unsigned long long seq = ld_volatile(h.ctl) + 1;  // ctl[0] moves only after every CTA of this launch has finished
  const unsigned num = (unsigned)seq;
  const long long half = 2 * (h.cap + 1);  // doubles per (slot, direction)
  const long long slot = (long long)(seq & 1) * 2 * half;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long units = n + 1;  // header + payload
  // ---- send: my bottom planes arrive "from above" at the lower neighbour, my top planes "from below" at the upper one
  double* out_lo = h.box_lo ? h.box_lo + slot + half : nullptr;
  double* out_hi = h.box_hi ? h.box_hi + slot : nullptr;
  for (long long t = t0; t < 2 * units; t += stride) {
    const bool up = t >= units;
    const long long u = up ? t - units : t;
    double* out = up ? out_hi : out_lo;
    const double* src = up ? send_hi : send_lo;
    if (out && (u == 0 || src)) st_unit(out + 2 * u, u == 0 ? 0.0 : src[u - 1], num);
  }
  // ---- receive
  const double* in_lo = h.box_lo ? h.box + slot : nullptr;
  const double* in_hi = h.box_hi ? h.box + slot + half : nullptr;
  bool ok = ld_volatile(h.ctl + 2) == 0;
  for (long long t = t0; t < 2 * units; t += stride) {
    const bool up = t >= units;
    const long long u = up ? t - units : t;
    const double* in = up ? in_hi : in_lo;
    double* dst = up ? recv_hi : recv_lo;
    if (in && (u == 0 || dst)) {
      double v = 0.0;
      if (ok) ok = ld_unit(in + 2 * u, num, v);
      if (u > 0) dst[u - 1] = v;
    }
  }
  if (!ok) atomicCAS(h.ctl + 2, 0ULL, seq);
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long done = atomicAdd(h.ctl + 1, 1ULL);
    if (done == gridDim.x - 1) {
      h.ctl[1] = 0;
      h.ctl[0] = seq;
    }
  }
}

// One-shot all-reduce of <= PMB_PEER_COUNT_MAX doubles, same transport: every rank stores its values (as units) into its own
// column of every rank's table, polls the columns of its own table and combines them in rank order (the same bits on every
// rank).  One CTA; thread (p, i) sends value i to rank p and receives value i of rank p.
// Table of a rank, in units: [slot 0 / 1][source rank][PMB_PEER_COUNT_MAX].  ctl words: [0] all-reduces completed, [1] first timed out.
__global__ void __launch_bounds__(PMB_PEER_MAX* PMB_PEER_COUNT_MAX) peer_allreduce_kernel(pmb_peer_reduce r, double* __restrict__ val,
                                                                                         int count, int op) {
  __shared__ double got[PMB_PEER_MAX][PMB_PEER_COUNT_MAX];
  const int t = threadIdx.x, p = t / PMB_PEER_COUNT_MAX, i = t % PMB_PEER_COUNT_MAX;
  const unsigned long long seq = ld_volatile(r.ctl) + 1;
  const unsigned num = (unsigned)seq;
  const long long slot = (long long)(seq & 1) * r.world * PMB_PEER_COUNT_MAX;
  const bool active = p < r.world && i < count;
  if (active) {
    st_unit(r.slots[p] + 2 * (slot + r.rank * PMB_PEER_COUNT_MAX + i), val[i], num);
    double v = 0.0;
    if (ld_volatile(r.ctl + 1) != 0 || !ld_unit(r.slots[r.rank] + 2 * (slot + p * PMB_PEER_COUNT_MAX + i), num, v))
      atomicCAS(r.ctl + 1, 0ULL, seq);
    got[p][i] = v;
  }
  __syncthreads();
  if (t < count) {
    double s = got[0][t];
    for (int q = 1; q < r.world; ++q) s = op == 0 ? s + got[q][t] : fmax(s, got[q][t]);
    val[t] = s;
  }
  if (t == 0) r.ctl[0] = seq;
}

}  // namespace

extern "C" long long pmb_peer_halo_box_doubles(long long cap) { return cap < 1 ? -1 : 8 * (cap + 1); }
extern "C" long long pmb_peer_reduce_table_doubles(int world) {
  return (world < 1 || world > PMB_PEER_MAX) ? -1 : 4LL * world * PMB_PEER_COUNT_MAX;
}

extern "C" int pmb_peer_halo_exchange(const pmb_peer_halo* h, long long n, const double* send_lo, const double* send_hi,
                                      double* recv_lo, double* recv_hi, void* stream) {
  PMB_REQUIRE(h && h->box && h->ctl, "pmb_peer_halo_exchange: NULL mailbox / control pointer");
  PMB_REQUIRE(n > 0 && n <= h->cap, "pmb_peer_halo_exchange: %lld doubles do not fit the mailbox (capacity %lld)", n, h->cap);
  PMB_REQUIRE(((unsigned long long)h->box & 15) == 0 && ((unsigned long long)h->box_lo & 15) == 0 && ((unsigned long long)h->box_hi & 15) == 0,
              "pmb_peer_halo_exchange: mailboxes must be 16-byte aligned");
  // no CTA ever waits for another CTA of the same launch (a unit is polled by the thread that needs it and written by a
  // neighbour's send phase, which never blocks), so the grid may be as wide as the copy wants
  long long blocks = (2 * (n + 1) + 511) / 512;
  if (blocks > 2 * 148) blocks = 2 * 148;
  peer_halo_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*h, n, send_lo, send_hi, recv_lo, recv_hi);
  PMB_CHECK_LAUNCH("pmb_peer_halo_exchange");
  return 0;
}

extern "C" int pmb_peer_allreduce(const pmb_peer_reduce* r, double* val, int count, int op, void* stream) {
  PMB_REQUIRE(r && val && r->ctl, "pmb_peer_allreduce: NULL pointer argument");
  PMB_REQUIRE(r->world >= 1 && r->world <= PMB_PEER_MAX && r->rank >= 0 && r->rank < r->world,
              "pmb_peer_allreduce: rank %d / world %d outside 1..%d", r->rank, r->world, PMB_PEER_MAX);
  PMB_REQUIRE(count >= 1 && count <= PMB_PEER_COUNT_MAX, "pmb_peer_allreduce: count %d outside 1..%d", count, PMB_PEER_COUNT_MAX);
  PMB_REQUIRE(op == 0 || op == 1, "pmb_peer_allreduce: op must be 0 (sum) or 1 (max)");
  for (int p = 0; p < r->world; ++p)
    PMB_REQUIRE(r->slots[p] && ((unsigned long long)r->slots[p] & 15) == 0, "pmb_peer_allreduce: table of rank %d NULL or misaligned", p);
  peer_allreduce_kernel<<<1, PMB_PEER_MAX * PMB_PEER_COUNT_MAX, 0, (cudaStream_t)stream>>>(*r, val, count, op);
  PMB_CHECK_LAUNCH("pmb_peer_allreduce");
  return 0;
}
