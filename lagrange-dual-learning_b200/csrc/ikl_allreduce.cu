// ikl_allreduce.cu -- in-place sum over the ranks of one NVSwitch box of a flat fp32 vector, as ONE kernel over NVLink peer
// memory (include/ikl.h: ikl_allreduce_sum_f32).
//
// The reference has no distributed code; data-parallel training of its network adds exactly one exchange per step: the sum
// of the replicas' parameter gradients (BASELINE.json north_star: "allreduce of the model gradients ... so that lambda
// stays identical on every rank").  The network of scripts/trainer.py:60-62 has 2.5 k to 83 k parameters, so that exchange
// is latency, not bandwidth: a one-shot "everybody pushes everything to everybody" all-reduce costs one NVLink store
// latency plus the skew between the ranks, needs no NCCL launch (5 kernels and ~20 us in the train step's CUDA graph) and
// is capturable by construction -- the protocol state (a launch sequence number) lives on the device.
//
// Protocol (the same tagged-word idea as the K-sum exchange in ikl_lagrangian.cuh): every rank owns a mailbox in symmetric
// memory, slot[parity][rank r][element i] = one 8-byte word holding the fp32 bits of rank r's element i and, in the upper
// half, the 32-bit sequence number of the launch that wrote it.  A value and its "valid" flag therefore arrive in ONE
// atomic store: no fence, no separate flag.  Launch s (parity s & 1):
//   push  thread t stores its elements, tagged s, into slot [parity][my rank] of EVERY rank's mailbox (its own included);
//   pull  thread t polls its own mailbox for the same elements of every rank until the tags read s and adds the values in
//         rank order -- the same floats in the same order everywhere, so the result is bit-identical on all ranks.
// Two parities suffice: finishing launch s needs every peer's words of launch s, so a rank is never more than one launch
// ahead of a peer.  No block waits before it has pushed everything it owns, so blocks never wait for each other's pushes
// in a cycle as long as the whole grid is resident (grid <= SM count).  A peer that never shows up trips a poll limit and
// the kernel traps instead of hanging the GPU.
#include "ikl_common.h"

namespace ikl {
namespace {

constexpr int kArThreads = 256;
constexpr int kArHeaderWords = 16;                 // 128 bytes: [0] sequence number, [1] block ticket

struct AllReduceParams {
  int world, rank;
  long long stride;                                // words per (parity, rank) slot row: max_elems rounded up to even
  unsigned long long* peer[IKL_MAX_RANKS];         // mailbox of rank r as mapped into this process
};

struct U64x2 { unsigned long long x, y; };

__device__ __forceinline__ void st_sys_v2(unsigned long long* p, unsigned long long a, unsigned long long b) {
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ U64x2 ld_sys_v2(const unsigned long long* p) {
  U64x2 v;
  asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}

// Two elements per thread and step (one 16-byte store / load per rank: each 8-byte word is atomic on its own and carries its
// own tag).  The pull issues the loads of ALL ranks before it looks at any tag, so a thread pays one round trip to its own
// L2 per polling round, not one per rank.  An odd n is padded with a zero word (the slot rows are an even number of words).
__global__ void __launch_bounds__(kArThreads) peer_allreduce_kernel(float* __restrict__ data, long long n, const AllReduceParams p) {
  unsigned long long* own = p.peer[p.rank];
  const unsigned long long seq = own[0] + 1;       // written only by the last block of a launch, after every block's ticket
  const unsigned long long tag = (seq & 0xffffffffull) << 32;
  const long long par_off = kArHeaderWords + (long long)(seq & 1ull) * p.world * p.stride;
  const long long step = 2ll * gridDim.x * kArThreads;
  const long long first = 2ll * ((long long)blockIdx.x * kArThreads + threadIdx.x);
  // push
  for (long long i = first; i < n; i += step) {
    const bool two = i + 1 < n;
    const unsigned long long w0 = tag | (unsigned long long)__float_as_uint(data[i]);
    const unsigned long long w1 = tag | (unsigned long long)(two ? __float_as_uint(data[i + 1]) : 0u);
    const long long slot = par_off + (long long)p.rank * p.stride + i;
#pragma unroll
    for (int r = 0; r < IKL_MAX_RANKS; ++r)
      if (r < p.world) st_sys_v2(p.peer[r] + slot, w0, w1);
  }
  // pull: the peers' words are already on their way while this rank was still pushing
  for (long long i = first; i < n; i += step) {
    U64x2 w[IKL_MAX_RANKS];
    unsigned pending = (1u << p.world) - 1u;
    unsigned long long polls = 0;
    while (pending) {
#pragma unroll
      for (int r = 0; r < IKL_MAX_RANKS; ++r)
        if (pending >> r & 1u) w[r] = ld_sys_v2(own + par_off + (long long)r * p.stride + i);
#pragma unroll
      for (int r = 0; r < IKL_MAX_RANKS; ++r)
        if ((pending >> r & 1u) && (w[r].x & 0xffffffff00000000ull) == tag && (w[r].y & 0xffffffff00000000ull) == tag) pending &= ~(1u << r);
      if (++polls > (1ull << 27)) asm volatile("trap;");
    }
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int r = 0; r < IKL_MAX_RANKS; ++r) {
      if (r < p.world) {                            // rank order: the same floats in the same order on every rank
        s0 += __uint_as_float((unsigned)(w[r].x & 0xffffffffull));
        s1 += __uint_as_float((unsigned)(w[r].y & 0xffffffffull));
      }
    }
    data[i] = s0;
    if (i + 1 < n) data[i + 1] = s1;
  }
  // the last block to finish advances the sequence number (device-side, so CUDA-graph replays pair up across ranks)
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(own + 1);
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      *ticket = 0u;
      own[0] = seq;
    }
  }
}

}  // namespace
}  // namespace ikl

using namespace ikl;

extern "C" {

size_t ikl_allreduce_bytes(int64_t max_elems, int32_t world) {
  if (max_elems < 1 || world < 1 || world > IKL_MAX_RANKS) return 0;
  const long long stride = (max_elems + 1) & ~1ll;   // slot rows are an even number of words (16-byte vector accesses)
  return (size_t)(kArHeaderWords + 2ll * world * stride) * sizeof(unsigned long long);
}

IklStatus ikl_allreduce_sum_f32(float* data, int64_t n, const IklExchange* ex, int64_t max_elems, void* stream) {
  if (!ex || n < 0 || max_elems < 1 || n > max_elems) return IKL_ERR_INVALID_ARGUMENT;
  if (ex->world < 1 || ex->world > IKL_MAX_RANKS || ex->rank < 0 || ex->rank >= ex->world) return IKL_ERR_INVALID_ARGUMENT;
  if (n > 0 && !data) return IKL_ERR_INVALID_ARGUMENT;
  if (n > 0 && (reinterpret_cast<uintptr_t>(data) & 3u)) return IKL_ERR_INVALID_ARGUMENT;
  AllReduceParams p;
  p.world = ex->world;
  p.rank = ex->rank;
  p.stride = (max_elems + 1) & ~1ll;
  for (int r = 0; r < IKL_MAX_RANKS; ++r) p.peer[r] = nullptr;
  for (int r = 0; r < ex->world; ++r) {
    if (!ex->peers[r] || (reinterpret_cast<uintptr_t>(ex->peers[r]) & 15u)) return IKL_ERR_INVALID_ARGUMENT;
    p.peer[r] = reinterpret_cast<unsigned long long*>(ex->peers[r]);
  }
  if (n == 0 || ex->world == 1) return IKL_OK;      // nothing to exchange; every rank takes this branch alike
  long long blocks = (n + 2 * kArThreads - 1) / (2 * kArThreads);
  int cap = sm_count() - reserved_sms();
  if (cap < 1) cap = 1;
  if (blocks > cap) blocks = cap;                   // the whole grid must be resident (see the protocol note above)
  count_launch();
  peer_allreduce_kernel<<<(unsigned)blocks, kArThreads, 0, (cudaStream_t)stream>>>(data, (long long)n, p);
  return cuda_status(cudaGetLastError());
}

}  // extern "C"
