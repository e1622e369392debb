// ikl_lagrangian.cuh -- the fused FK + penalty + Lagrangian (+ analytic gradient) kernels.
//
// One persistent grid streams the batch once: per sample it reads theta, the target and the obstacle
// rows, evaluates every constraint term (ikl_packed.cuh / ikl_device.cuh), optionally writes
// d(sum_i w_i g_i)/d(theta), and keeps K running sums per thread.  Sums go thread (fp32, <= 64 samples) ->
// warp (fp32 butterfly) or thread (fp64 slot) -> fp64 per warp -> block -> one fp64 partial per block per
// constraint -> the last block to finish adds the partials in a fixed order, so results are bit-reproducible
// for a given launch geometry and no float atomics are used.
// (Reference: scripts/trainer.py:182-247 + :264 of miloknowles/lagrange-dual-learning.)
//
// Memory layouts (see IklBatch in include/ikl.h) and the kernels that serve them:
//   kPlanes   structure-of-arrays planes.  lagrangian_planes_tma_kernel: tiled TMA loads (tensor maps) into a
//             barrier-free shared-memory pipeline, thread v owns samples (2v, 2v+1) of a 512-sample tile; the staged
//             register pair IS the packed f32x2 operand.  Fallback (odd alignment): direct 8-byte loads.
//   kRows     the reference's own row-major tensors: theta (B,3), target (B,2|3), obstacles (B,4,4).
//             lagrangian_rows_tma_kernel: three bulk copies per 512-row tile into the same pipeline, thread t owns
//             samples (t, t + 256); ragged tail through the scalar code.  Fallback: one LDG.128 per obstacle row.
//   kStrided  any element strides, 2 links, runtime link lengths -- scalar code, always correct.
// The pipelines hand each pair of samples to a "body": LagrangianBody (sums + gradient) or StatsBody (the evaluation
// statistics of scripts/evaluation/visualize_ik_solutions.py:97-119).
#pragma once

#include <stdlib.h>

#include <type_traits>

#include "ikl_common.h"
#include "ikl_device.cuh"
#include "ikl_packed.cuh"

namespace ikl {

#ifndef IKL_THREADS
#define IKL_THREADS 256
#endif
constexpr int kThreads = IKL_THREADS;
constexpr int kWarps = kThreads / 32;
#ifndef IKL_FLUSH_SAMPLES
#define IKL_FLUSH_SAMPLES 64
#endif
constexpr int kFlushSamples = IKL_FLUSH_SAMPLES;   // samples a thread adds in fp32 before flushing into its fp64 slot in shared memory
constexpr int kFinalParts = kThreads / kMaxK;   // 16 partial chains in the last-block reduce

#ifndef IKL_ROWS_UNROLL
#define IKL_ROWS_UNROLL 2         // samples in flight per thread in the strided fallback
#endif
#ifndef IKL_MIN_BLOCKS
#define IKL_MIN_BLOCKS 2
#endif

enum Layout : int { kStrided = 0, kRows = 1, kPlanes = 2 };
// kGradOnly*: d/dtheta alone -- no running sums, no reduction, no viol_out / loss_out (staged kernels only; other kernels
// serve such a request with the corresponding kGrad* mode and discard the sums)
enum Mode : int { kFwd = 0, kGradHostW = 1, kGradDevW = 2, kGradOnlyHostW = 3, kGradOnlyDevW = 4 };

struct alignas(64) TensorMap { unsigned long long opaque[16]; };      // same size and alignment as the driver's CUtensorMap

// One rank's mailbox for the in-kernel exchange of the K sums (include/ikl.h: IklExchange).  Low-latency protocol: every
// 8-byte word carries 32 bits of payload and the 32-bit sequence number of the launch that wrote it, so a value and its
// "valid" flag arrive in ONE store -- no fence, no separate flag, one NVLink latency.  slot[parity][r][i] = the two halves of
// rank r's double-precision total of constraint i for the launch with sequence number seq (parity = seq & 1).  Two parities
// suffice: a rank can only be one launch ahead of a peer, because finishing launch s needs every peer's words of launch s.
// `seq` is the owner's private launch counter (device-side, so graph replays advance it).
constexpr int kMaxRanks = IKL_MAX_RANKS;
struct ExchangeBuf {
  unsigned long long slot[2][kMaxRanks][kMaxK][2];
  unsigned long long seq;
};

struct LagrangianParams {
  Uniforms u;
  PackedUniforms pu;
  StaticPairUniforms sp;
  long long batch;
  const float* theta;   long long th_sb, th_sj;
  const float* target;  long long tg_sb, tg_sc;
  const float* obst;    long long ob_sb, ob_so, ob_sc;
  float* dtheta;        long long dt_sb, dt_sj;
  float* q_ee;          // (B,3) or null
  float* err_sq;        // (B,2) or null
  const float* w_dev;   // K weights in device memory (kGradDevW, or the loss of a kFwd launch) or null
  double* partials;     // [gridDim.x][kMaxK]
  unsigned int* counter;
  float* viol_out;
  float* loss_out;
  double* viol_accum;   // or null
  int use_tma;          // planes layout: TMA-staged kernel (needs batch % 4 == 0, 16-byte aligned planes, plane strides % 4 == 0)
  // TMA descriptors of the planes as 2-D / 3-D tensors (sample index innermost), encoded by the host per call:
  // theta {B, 3}, target {B, 2}, obstacles {B, 3, n}; box = 256 samples x every plane of the group.
  TensorMap tm_theta, tm_target, tm_obst;
  // evaluation statistics (stats_*_tma_kernel): counts[0..K) = #(term > threshold), [16] any joint limit, [17] any obstacle
  float stats_threshold;
  unsigned long long* stats_counts;
  double* stats_pos_err;
  // in-kernel all-reduce of the K sums over peer memory (ex_world <= 1: off)
  int ex_world, ex_rank;
  ExchangeBuf* ex_peer[kMaxRanks];
  // epoch-closing launch of a dual step (ikl_lagrangian_*_dual_step): what is exchanged is viol_accum's LOCAL running sum plus
  // this launch's sums; viol_accum receives the GLOBAL epoch sums and, with dual_lam_out, the multipliers are updated
  int close_epoch;
  float dual_lr;
  const float* dual_lam_in;
  float* dual_lam_out;
};
static_assert(sizeof(LagrangianParams) <= 4096, "kernel parameter block too large");

template <int LAYOUT_, int J_, int NOBS_, bool DYN_, bool JL_, int MODE_, bool UNIT_>
struct Cfg {
  static constexpr int LAYOUT = LAYOUT_, J = J_, NOBS = NOBS_, MODE = MODE_;
  static constexpr bool DYN = DYN_, JL = JL_, UNIT = UNIT_, GRAD = MODE_ != kFwd;
  static constexpr bool SUMS = MODE_ != kGradOnlyHostW && MODE_ != kGradOnlyDevW;      // the K running sums are formed
  static constexpr bool DEVW = MODE_ == kGradDevW || MODE_ == kGradOnlyDevW;           // multipliers read from device memory
  using WithSums = Cfg<LAYOUT_, J_, NOBS_, DYN_, JL_, MODE_ == kGradOnlyHostW ? kGradHostW : (MODE_ == kGradOnlyDevW ? kGradDevW : MODE_), UNIT_>;
  static constexpr int K = num_constraints(J_, NOBS_, JL_);
  static constexpr int NOB = NOBS_ > 0 ? NOBS_ : 1;
};

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
  float r;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// ---- reduction tail ----------------------------------------------------------------------------------
// Second-level accumulators: one fp64 slot per (constraint, thread) in shared memory, so the hot loop only
// carries the K fp32 running sums in registers.
template <int K>
struct SharedAcc {
  double v[K][kThreads];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < K; ++i) v[i][threadIdx.x] = 0.0;
  }
  __device__ __forceinline__ void flush(float (&acc)[K]) {
#pragma unroll
    for (int i = 0; i < K; ++i) {
      v[i][threadIdx.x] += (double)acc[i];
      acc[i] = 0.f;
    }
  }
  __device__ __forceinline__ void flush(f2 (&acc)[K]) {
#pragma unroll
    for (int i = 0; i < K; ++i) {
      v[i][threadIdx.x] += (double)(f2_lo(acc[i]) + f2_hi(acc[i]));
      acc[i] = f2_splat(0.f);
    }
  }
};

// ---- in-kernel all-reduce over NVLink peer memory (IklExchange) ---------------------------------------------------------
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Called by the whole LAST block of a rank's launch; thread i < K passes this rank's total of constraint i and gets the
// global one back.  Push: thread i splits its double into two 32-bit halves, tags each with the launch's sequence number and
// stores the two 8-byte words into slot [parity][rank][i] of EVERY rank's mailbox (its own included) -- plain NVLink P2P
// stores, each atomic, so no fence and no flag are needed.  Pull: thread i polls its own mailbox's slot [parity][r][i] for
// r = 0 .. world-1 until both words carry the sequence number, reassembles the doubles and adds them in rank order -- the
// same order of the same doubles on every rank, hence bit-identical results.  A peer that never shows up trips the poll
// limit (seconds) and the kernel traps rather than hanging.
template <int K>
__device__ __forceinline__ double exchange_sums(const LagrangianParams& p, double mine) {
  const int tid = threadIdx.x;
  ExchangeBuf* own = p.ex_peer[p.ex_rank];
  const unsigned long long seq = own->seq + 1;      // only this block (the last one) of this rank touches own->seq
  const int par = (int)(seq & 1ull);
  const unsigned long long tag = (seq & 0xffffffffull) << 32;
  double tot = 0.0;
  if (tid < K) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(mine);
    const unsigned long long w0 = tag | (bits & 0xffffffffull), w1 = tag | (bits >> 32);
    for (int r = 0; r < p.ex_world; ++r) {
      unsigned long long* dst = p.ex_peer[r]->slot[par][p.ex_rank][tid];
      st_relaxed_sys_u64(dst, w0);
      st_relaxed_sys_u64(dst + 1, w1);
    }
    for (int r = 0; r < p.ex_world; ++r) {
      const unsigned long long* src = own->slot[par][r][tid];
      unsigned long long a, b, polls = 0;
      for (;;) {
        a = ld_relaxed_sys_u64(src);
        b = ld_relaxed_sys_u64(src + 1);
        if ((a & 0xffffffff00000000ull) == tag && (b & 0xffffffff00000000ull) == tag) break;
        if (++polls > (1ull << 27)) asm volatile("trap;");
      }
      tot += __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
    }
  }
  __syncthreads();
  if (tid == 0) own->seq = seq;
  return tot;
}

// Block-level tail shared by every kernel: s_red[warp][i] holds each warp's fp64 sum of constraint i.
template <int K>
__device__ __forceinline__ void block_finalize(const LagrangianParams& p, double (*s_red)[kMaxK]) {
  __shared__ double s_fin[kFinalParts][kMaxK];
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  __syncthreads();
  if (tid < K) {
    double v = 0.0;
#pragma unroll
    for (int wq = 0; wq < kWarps; ++wq) v += s_red[wq][tid];
    p.partials[(size_t)blockIdx.x * kMaxK + tid] = v;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(p.counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {
    const int i = tid % kMaxK, part = tid / kMaxK;
    // same order of additions as a plain loop, but up to twenty loads in flight at a time (a grid of 296 blocks is ONE batch per
    // thread): as a chain of dependent L2 round trips this was ~3 us of every launch, in batches of eight still three round
    // trips -- nothing at 2^24 samples, several per cent of a launch on a 2^21-sample shard (strong scaling).  Slots past the
    // grid contribute +0.0, which leaves the sum's bits alone.
    double v = 0.0;
    if (i < K) {
      constexpr int kBatch = 20;
      for (unsigned blk = part; blk < gridDim.x; blk += kBatch * kFinalParts) {
        double t[kBatch];
#pragma unroll
        for (int uu = 0; uu < kBatch; ++uu) {
          const unsigned b = blk + uu * kFinalParts;
          t[uu] = b < gridDim.x ? __ldcg(p.partials + (size_t)b * kMaxK + i) : 0.0;
        }
#pragma unroll
        for (int uu = 0; uu < kBatch; ++uu) v += t[uu];
      }
    }
    s_fin[part][i] = v;
  }
  __syncthreads();
  double tot = 0.0;
  if (tid < K) {
#pragma unroll
    for (int part = 0; part < kFinalParts; ++part) tot += s_fin[part][tid];
  }
  if (p.close_epoch) {
    // Last launch of a dual step (scripts/trainer.py:267-286): this launch's sums join the rank's running sum, the running
    // sums -- not the per-launch ones -- cross NVLink once per epoch, and the dual ascent is applied here, by the block that
    // holds the global sums: compute, collective and multiplier update in one kernel.  viol_out / loss_out stay this
    // launch's LOCAL values (with the multipliers of BEFORE the update, as process_batch computes them).
    double run = tid < K ? tot + p.viol_accum[tid] : 0.0;
    if (p.ex_world > 1) run = exchange_sums<K>(p, run);     // block-uniform branch; contains block barriers
    if (tid < K) {
      const float wi = (p.w_dev != nullptr) ? p.w_dev[tid] : p.u.w[tid];     // read before lam_out (which may alias w_dev) is written
      p.viol_accum[tid] = run;
      p.viol_out[tid] = (float)tot;
      s_fin[0][tid] = tot * (double)wi;
      if (p.dual_lam_out) {
        // the arithmetic of dual_update_kernel (trainer.py:280-284): fp32 lam + lr * sum, clamp(min=0); NaN propagates
        const float next = __fadd_rn(p.dual_lam_in[tid], __fmul_rn(p.dual_lr, (float)run));
        p.dual_lam_out[tid] = next < 0.f ? 0.f : next;
      }
    }
  } else {
    if (p.ex_world > 1) tot = exchange_sums<K>(p, tot);       // block-uniform branch; contains block barriers
    if (tid < K) {
      p.viol_out[tid] = (float)tot;
      if (p.viol_accum) p.viol_accum[tid] += tot;
      const float wi = (p.w_dev != nullptr) ? p.w_dev[tid] : p.u.w[tid];
      s_fin[0][tid] = tot * (double)wi;
    }
  }
  __syncthreads();
  if (tid == 0) {
    double loss = 0.0;
    for (int i = 0; i < K; ++i) loss += s_fin[0][i];
    *p.loss_out = (float)loss;
    *p.counter = 0u;            // leave the workspace ready for the next launch
  }
}

template <int K>
__device__ __forceinline__ void finalize_sums(const LagrangianParams& p, SharedAcc<K>& sacc) {
  __shared__ double s_red[kWarps][kMaxK];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    double v = sacc.v[i][tid];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) s_red[warp][i] = v;
  }
  block_finalize<K>(p, s_red);
}

// Warp-level second level for the TMA-staged kernels: every kFlushSamples samples per thread the
// K packed fp32 running sums of a warp are reduced across its lanes with a transposing butterfly (16 + 8 + 4 + 2 + 1
// shuffles instead of 5 per constraint) and lane i adds constraint i's warp total to one fp64 slot in shared memory.
// 1 KiB of shared memory per block instead of SharedAcc's 32 KiB, and no extra registers in the hot loop.
template <int K>
struct WarpSharedAcc {
  double (*s_warp)[kMaxK];
  __device__ __forceinline__ explicit WarpSharedAcc(double (*sw)[kMaxK]) : s_warp(sw) {
    if (threadIdx.x < kWarps * kMaxK) (&sw[0][0])[threadIdx.x] = 0.0;
  }
  static constexpr int N = K <= 1 ? 1 : K <= 2 ? 2 : K <= 4 ? 4 : K <= 8 ? 8 : 16;      // vector length: next power of two >= K
  __device__ __forceinline__ void flush(f2 (&acc)[K]) {
    float v[N];
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = i < K ? f2_lo(acc[i]) + f2_hi(acc[i]) : 0.f;
#pragma unroll
    for (int i = 0; i < K; ++i) acc[i] = f2_splat(0.f);
    reduce(v);
  }
  __device__ __forceinline__ void flush(float (&acc)[K]) {
    float v[N];
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = i < K ? acc[i] : 0.f;
#pragma unroll
    for (int i = 0; i < K; ++i) acc[i] = 0.f;
    reduce(v);
  }
  // Rows pipeline: the thread keeps the sums of obstacle (k ^ r) in obstacle slot k (see run_rows_pipe); put them back
  // in constraint order before the lanes (which have different r) are added together.
  template <int OBASE, int J>
  __device__ __forceinline__ void flush_xor4(f2 (&acc)[K], int r) {
    static_assert(K == OBASE + 4 * J, "four obstacles x J tips");
    float v[N];
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = i < K ? f2_lo(acc[i]) + f2_hi(acc[i]) : 0.f;
#pragma unroll
    for (int i = 0; i < K; ++i) acc[i] = f2_splat(0.f);
#pragma unroll
    for (int bit = 1; bit <= 2; bit <<= 1) {
      const bool sw = (r & bit) != 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k & bit) continue;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const float a = v[OBASE + J * k + j], b = v[OBASE + J * (k | bit) + j];
          v[OBASE + J * k + j] = sw ? b : a;
          v[OBASE + J * (k | bit) + j] = sw ? a : b;
        }
      }
    }
    reduce(v);
  }
  // all 32 lanes of the warp must call this together
  __device__ __forceinline__ void reduce(float (&v)[N]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // transposing steps: a lane keeps the half of the vector selected by its `off` bit and adds its partner's copy of it
    int idx = 0, off = 16;
#pragma unroll
    for (int n = N / 2; n >= 1; n >>= 1, off >>= 1) {
      const bool up = (lane & off) != 0;
      idx += up ? n : 0;
#pragma unroll
      for (int i = 0; i < n; ++i) {
        const float keep = up ? v[i + n] : v[i];
        const float send = up ? v[i] : v[i + n];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    // plain butterfly over the remaining lane bits
    const int rest = 2 * off - 1;            // mask of the lane bits not used for transposing
#pragma unroll
    for (; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    if ((lane & rest) == 0 && idx < K) s_warp[warp][idx] += (double)v[0];
  }
};

// ---- per-sample IO for the one-sample-per-thread layouts ----------------------------------------------------
template <class C>
__device__ __forceinline__ void load_sample(const LagrangianParams& p, long long b, float (&th)[C::J], float& tx, float& ty,
                                            float (&ob)[C::NOB][3]) {
  if (C::LAYOUT == kRows) {
    const float* t = p.theta + b * C::J;
#pragma unroll
    for (int j = 0; j < C::J; ++j) th[j] = ldg_stream1(t + j);
    const float* g = p.target + b * p.tg_sb;
    tx = ldg_stream1(g);
    ty = ldg_stream1(g + 1);
    if (C::DYN) {
      const float* o = p.obst + b * (4 * kMaxObst);
#pragma unroll
      for (int k = 0; k < C::NOBS; ++k) {
        const float4 v = ldg_stream4(o + 4 * k);
        ob[k][0] = v.x; ob[k][1] = v.y; ob[k][2] = v.z;
      }
    }
  } else {
    const float* t = p.theta + b * p.th_sb;
#pragma unroll
    for (int j = 0; j < C::J; ++j) th[j] = __ldg(t + j * p.th_sj);
    const float* g = p.target + b * p.tg_sb;
    tx = __ldg(g);
    ty = __ldg(g + p.tg_sc);
    if (C::DYN) {
      const float* o = p.obst + b * p.ob_sb;
#pragma unroll
      for (int k = 0; k < C::NOBS; ++k) {
#pragma unroll
        for (int c = 0; c < 3; ++c) ob[k][c] = __ldg(o + k * p.ob_so + c * p.ob_sc);
      }
    }
  }
}

template <class C>
__device__ __forceinline__ void static_obstacles(const LagrangianParams& p, float (&ob)[C::NOB][3]) {
#pragma unroll
  for (int k = 0; k < C::NOB; ++k) {
#pragma unroll
    for (int c = 0; c < 3; ++c) ob[k][c] = p.u.stat_ob[k][c];
  }
}

template <class C>
__device__ __forceinline__ void store_sample(const LagrangianParams& p, long long b, const float (&dth)[C::J], const float (&ee)[3],
                                             float tx, float ty) {
  if (C::GRAD) {
    if (C::LAYOUT == kRows) {
      float* d = p.dtheta + b * C::J;
#pragma unroll
      for (int j = 0; j < C::J; ++j) d[j] = dth[j];
    } else {
      float* d = p.dtheta + b * p.dt_sb;
#pragma unroll
      for (int j = 0; j < C::J; ++j) d[j * p.dt_sj] = dth[j];
    }
  } else {
    if (p.q_ee) {
      float* q = p.q_ee + b * 3;
      q[0] = ee[0]; q[1] = ee[1]; q[2] = ee[2];
    }
    if (p.err_sq) {
      const float ex = tx - ee[0], ey = ty - ee[1];
      p.err_sq[b * 2] = 0.5f * (ex * ex);
      p.err_sq[b * 2 + 1] = 0.5f * (ey * ey);
    }
  }
}

template <class C, class W>
__device__ __forceinline__ void run_scalar_layout(const LagrangianParams& p, const W& w, SharedAcc<C::K>& sacc) {
  constexpr int U = IKL_ROWS_UNROLL;
  const long long B = p.batch;
  const long long step = (long long)gridDim.x * (kThreads * U);
  float acc[C::K];
#pragma unroll
  for (int i = 0; i < C::K; ++i) acc[i] = 0.f;
  int since_flush = 0;
  for (long long base = (long long)blockIdx.x * (kThreads * U) + threadIdx.x; base < B; base += step) {
    float th[U][C::J], tx[U], ty[U], ob[U][C::NOB][3];
#pragma unroll
    for (int uu = 0; uu < U; ++uu) {
      const long long b = base + uu * kThreads;
      if (!C::DYN) static_obstacles<C>(p, ob[uu]);
      if (b < B) load_sample<C>(p, b, th[uu], tx[uu], ty[uu], ob[uu]);
    }
#pragma unroll
    for (int uu = 0; uu < U; ++uu) {
      const long long b = base + uu * kThreads;
      if (b < B) {
        float dth[C::J], ee[3];
        eval_sample<C::J, C::NOBS, C::JL, C::GRAD, C::UNIT>(p.u, w, th[uu], tx[uu], ty[uu], ob[uu], acc, dth, ee);
        store_sample<C>(p, b, dth, ee, tx[uu], ty[uu]);
      }
    }
    if (++since_flush == kFlushSamples / U) {
      since_flush = 0;
      sacc.flush(acc);
    }
  }
  sacc.flush(acc);
}

// ---- packed layouts: two samples per thread on the f32x2 pipe ----------------------------------------------------
__device__ __forceinline__ f2 ldg_stream_f2(const float* p) {
  f2 r;
  asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(r) : "l"(p));
  return r;
}
// IKL_STORE_CS=1: d/dtheta is stored with the evict-first (.cs) policy instead of L1::no_allocate.  Measured and rejected
// (profiles/r02/kernel_table_cache_policy.jsonl): 0.2116 -> 0.2136 ms on the benchmark kernel, nothing elsewhere.
#ifndef IKL_STORE_CS
#define IKL_STORE_CS 0
#endif
__device__ __forceinline__ void stg_stream_f2(float* p, f2 v) {
#if IKL_STORE_CS
  asm volatile("st.global.cs.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
  asm volatile("st.global.L1::no_allocate.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}

template <class C>
__device__ __forceinline__ void static_obstacles_packed(const LagrangianParams& p, f2 (&ob)[C::NOB][3]) {
#pragma unroll
  for (int k = 0; k < C::NOB; ++k) {
#pragma unroll
    for (int c = 0; c < 3; ++c) ob[k][c] = as_f2(p.pu.stat_ob[k][c]);
  }
}

// side outputs of a forward-only launch for the pair (bA, bB): row-major (B,3) / (B,2), rarely requested
__device__ __forceinline__ void store_side_outputs(const LagrangianParams& p, long long bA, long long bB, const f2 (&ee)[3], f2 tx, f2 ty) {
  if (p.q_ee) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      p.q_ee[bA * 3 + c] = f2_lo(ee[c]);
      p.q_ee[bB * 3 + c] = f2_hi(ee[c]);
    }
  }
  if (p.err_sq) {
    const f2 ex = f2_sub(tx, ee[0]), ey = f2_sub(ty, ee[1]);
    const f2 qx = f2_mul(f2_splat(0.5f), f2_mul(ex, ex)), qy = f2_mul(f2_splat(0.5f), f2_mul(ey, ey));
    p.err_sq[bA * 2] = f2_lo(qx);  p.err_sq[bA * 2 + 1] = f2_lo(qy);
    p.err_sq[bB * 2] = f2_hi(qx);  p.err_sq[bB * 2 + 1] = f2_hi(qy);
  }
}

// planes: thread v owns samples (2v, 2v+1); every plane access is one coalesced 8-byte load/store.
template <class C, class W2>
__device__ __forceinline__ void run_planes_packed(const LagrangianParams& p, const W2& w2, SharedAcc<C::K>& sacc) {
  const long long npair = p.batch >> 1;             // the API guarantees an even batch for this layout
  const long long step = (long long)gridDim.x * kThreads;
  f2 acc[C::K];
#pragma unroll
  for (int i = 0; i < C::K; ++i) acc[i] = f2_splat(0.f);
  int since_flush = 0;
  for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < npair; v += step) {
    const long long b = v << 1;
    f2 th[C::J], ob[C::NOB][3], dth[C::J], ee[3];
#pragma unroll
    for (int j = 0; j < C::J; ++j) th[j] = ldg_stream_f2(p.theta + j * p.th_sj + b);
    const f2 tx = ldg_stream_f2(p.target + b);
    const f2 ty = ldg_stream_f2(p.target + p.tg_sc + b);
    if (C::DYN) {
#pragma unroll
      for (int k = 0; k < C::NOBS; ++k) {
#pragma unroll
        for (int c = 0; c < 3; ++c) ob[k][c] = ldg_stream_f2(p.obst + k * p.ob_so + c * p.ob_sc + b);
      }
    } else {
      static_obstacles_packed<C>(p, ob);
    }
    eval_pair<C::J, C::UNIT, C::NOBS, C::JL, C::GRAD, true, C::DYN, true>(p.pu, w2, th, tx, ty, ob, acc, dth, ee, as_f2(p.pu.ob_ndiff));
    if (C::GRAD) {
#pragma unroll
      for (int j = 0; j < C::J; ++j) stg_stream_f2(p.dtheta + j * p.dt_sj + b, dth[j]);
    } else {
      store_side_outputs(p, b, b + 1, ee, tx, ty);
    }
    if (++since_flush == kFlushSamples / 2) {
      since_flush = 0;
      sacc.flush(acc);
    }
  }
  sacc.flush(acc);
}

// rows: thread t of a 512-sample tile owns samples (base + t, base + t + 256); the ragged tail (< 512 samples)
// goes through the scalar per-sample code so any batch size is accepted.
template <class C, class W2, class W>
__device__ __forceinline__ void run_rows_packed(const LagrangianParams& p, const W2& w2, const W& w, SharedAcc<C::K>& sacc) {
  constexpr int kTile = 2 * kThreads;
  const long long B = p.batch;
  const long long full = (B / kTile) * kTile;
  f2 acc[C::K];
#pragma unroll
  for (int i = 0; i < C::K; ++i) acc[i] = f2_splat(0.f);
  int since_flush = 0;
  for (long long base = (long long)blockIdx.x * kTile; base < full; base += (long long)gridDim.x * kTile) {
    const long long bA = base + threadIdx.x, bB = bA + kThreads;
    f2 th[C::J], ob[C::NOB][3], dth[C::J], ee[3];
    const float* tA = p.theta + bA * C::J;
    const float* tB = p.theta + bB * C::J;
#pragma unroll
    for (int j = 0; j < C::J; ++j) th[j] = f2_make(ldg_stream1(tA + j), ldg_stream1(tB + j));
    const float* gA = p.target + bA * p.tg_sb;
    const float* gB = p.target + bB * p.tg_sb;
    const f2 tx = f2_make(ldg_stream1(gA), ldg_stream1(gB));
    const f2 ty = f2_make(ldg_stream1(gA + 1), ldg_stream1(gB + 1));
    if (C::DYN) {
      const float* oA = p.obst + bA * (4 * kMaxObst);
      const float* oB = p.obst + bB * (4 * kMaxObst);
#pragma unroll
      for (int k = 0; k < C::NOBS; ++k) {
        const float4 a = ldg_stream4(oA + 4 * k), b4 = ldg_stream4(oB + 4 * k);
        ob[k][0] = f2_make(a.x, b4.x);
        ob[k][1] = f2_make(a.y, b4.y);
        ob[k][2] = f2_make(a.z, b4.z);
      }
    } else {
      static_obstacles_packed<C>(p, ob);
    }
    eval_pair<C::J, C::UNIT, C::NOBS, C::JL, C::GRAD, true, C::DYN, true>(p.pu, w2, th, tx, ty, ob, acc, dth, ee, as_f2(p.pu.ob_ndiff));
    if (C::GRAD) {
      float* dA = p.dtheta + bA * C::J;
      float* dB = p.dtheta + bB * C::J;
#pragma unroll
      for (int j = 0; j < C::J; ++j) {
        dA[j] = f2_lo(dth[j]);
        dB[j] = f2_hi(dth[j]);
      }
    } else {
      store_side_outputs(p, bA, bB, ee, tx, ty);
    }
    if (++since_flush == kFlushSamples / 2) {
      since_flush = 0;
      sacc.flush(acc);
    }
  }
  sacc.flush(acc);
  // ragged tail, one sample per thread
  float sacc1[C::K];
#pragma unroll
  for (int i = 0; i < C::K; ++i) sacc1[i] = 0.f;
  for (long long b = full + (long long)blockIdx.x * kThreads + threadIdx.x; b < B; b += (long long)gridDim.x * kThreads) {
    float th[C::J], tx, ty, ob[C::NOB][3], dth[C::J], ee[3];
    if (!C::DYN) static_obstacles<C>(p, ob);
    load_sample<C>(p, b, th, tx, ty, ob);
    eval_sample<C::J, C::NOBS, C::JL, C::GRAD, C::UNIT>(p.u, w, th, tx, ty, ob, sacc1, dth, ee);
    store_sample<C>(p, b, dth, ee, tx, ty);
  }
  sacc.flush(sacc1);
}

}  // namespace ikl

#include "ikl_pipeline.cuh"      // staged kernels: run_planes_pipe / run_rows_pipe, LagrangianBody, StatsBody

namespace ikl {

template <class C>
__global__ void __launch_bounds__(kThreads, IKL_MIN_BLOCKS) lagrangian_kernel(const __grid_constant__ LagrangianParams p) {
  __shared__ SharedAcc<C::K> sacc;
  __shared__ float2 s_lam2[kMaxK];
  sacc.clear();
  if constexpr (C::LAYOUT == kStrided) {
    if constexpr (C::MODE == kGradDevW) {
      const DeviceWeights<C::K> w(p.u, p.w_dev);
      run_scalar_layout<C>(p, w, sacc);
    } else {
      const HostWeights w(p.u, nullptr);
      run_scalar_layout<C>(p, w, sacc);
    }
  } else if constexpr (C::MODE == kGradDevW) {
    // device-resident multipliers: stage the duplicated pairs once per block
    if (threadIdx.x < C::K) {
      const float l = __ldg(p.w_dev + threadIdx.x);
      s_lam2[threadIdx.x] = make_float2(l, l);
    }
    __syncthreads();
    const PackedSmemWeights w2{s_lam2};
    if constexpr (C::LAYOUT == kPlanes) {
      run_planes_packed<C>(p, w2, sacc);
    } else {
      const DeviceWeights<C::K> w(p.u, p.w_dev);
      run_rows_packed<C>(p, w2, w, sacc);
    }
  } else {
    const PackedHostWeights w2{p.pu};
    if constexpr (C::LAYOUT == kPlanes) {
      run_planes_packed<C>(p, w2, sacc);
    } else {
      const HostWeights w(p.u, nullptr);
      run_rows_packed<C>(p, w2, w, sacc);
    }
  }
  finalize_sums<C::K>(p, sacc);
}

// ---- host-side launch ------------------------------------------------------------------------------------------
inline int grid_sms() {
  const int n = sm_count() - reserved_sms();
  return n > 0 ? n : 1;
}
// Blocks per SM a staged kernel's persistent grid is sized for.  The kernels without obstacles need so few registers that
// four or five blocks fit an SM, but the ones that also WRITE d/dtheta run best with three (same-run sweep at 2^24 samples,
// profiles/r02/kernel_table_block_cap.jsonl: unconstrained planes fwd+bwd 0.099 -> 0.090 ms, gradient-only 0.095 -> 0.084 ms =
// 98 % of the HBM peak; rows gradient-only 0.109 -> 0.093 ms; joint limits alike); forward-only launches keep every block
// they can get (0.066 vs 0.071 ms).  IKL_GRID_BLOCKS_PER_SM=n overrides the rule (measurements).
template <class C>
inline int staged_blocks_per_sm(int resident) {
  static const int forced = [] { const char* e = getenv("IKL_GRID_BLOCKS_PER_SM"); return e && *e ? atoi(e) : -1; }();
  if (forced > 0) return resident < forced ? resident : forced;
  if (forced == 0) return resident;
  constexpr int cap = (C::GRAD && C::NOBS == 0) ? 3 : 1 << 20;
  return resident < cap ? resident : cap;
}

// Resident blocks per SM of a kernel with `smem` bytes of dynamic shared memory, cached per (kernel instantiation, device):
// the opt-in to more than 48 KiB of dynamic shared memory is a per-device function attribute, so a process that drives
// several GPUs must set it on each of them.
template <class Kernel>
cudaError_t resident_blocks(Kernel kernel, size_t smem, int (&cache)[64], int* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const int slot = (dev >= 0 && dev < 64) ? dev : 0;
  if (cache[slot] == 0 || slot != dev) {
    if (smem > 0) {
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, smem);
    if (e != cudaSuccess) return e;
    cache[slot] = n > 0 ? n : 1;
  }
  *out = cache[slot];
  return cudaSuccess;
}

// Programmatic dependent launch (IKL_PDL=1).  Back-to-back launches of the staged kernels on one stream -- the steps of a CUDA
// graph on a small shard (strong scaling: 2^21 samples per rank last ~36 us, of which ~10 us do not shrink with the shard) --
// each pay a launch gap and a prologue (block scheduling, mbarrier set-up) in series with the previous kernel's tail.  With
// the attribute the NEXT kernel's blocks may become resident as soon as the previous grid has released its SM resources
// (every block calls griddepcontrol.launch_dependents at its start), run their prologue, and stop at griddepcontrol.wait --
// which returns when the previous grid has completed and its writes are visible -- before their first global read or write.
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("IKL_PDL"); return e && e[0] == '1'; }();
  return on;
}
template <class Kernel>
cudaError_t launch_staged(Kernel kernel, unsigned grid, size_t smem, cudaStream_t stream, const LagrangianParams& p) {
  if (!pdl_enabled()) {
    kernel<<<grid, kThreads, smem, stream>>>(p);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}

template <class C>
cudaError_t launch_cfg(const LagrangianParams& p, cudaStream_t stream, int max_blocks, const char* name) {
  if constexpr (!C::SUMS) {
    // gradient-only specialisations exist for the staged kernels; everything else runs the full kernel (sums discarded)
    if (!p.use_tma || C::LAYOUT == kStrided) return launch_cfg<typename C::WithSums>(p, stream, max_blocks, name);
  }
  if constexpr (C::LAYOUT == kPlanes) {
    if (p.use_tma) {
      static int cache[64] = {0};
      constexpr size_t smem = TmaLayout<C>::kSmemBytes;
      int tma_blocks_per_sm = 1;
      const cudaError_t e = resident_blocks(lagrangian_planes_tma_kernel<C>, smem, cache, &tma_blocks_per_sm);
      if (e != cudaSuccess) return e;
      long long tiles = (p.batch + kTmaTile - 1) / kTmaTile;
      if (tiles < 1) tiles = 1;
      long long grid = (long long)grid_sms() * staged_blocks_per_sm<C>(tma_blocks_per_sm);
      if (grid > tiles) grid = tiles;
      if (grid > max_blocks) grid = max_blocks;
      set_last_kernel_name(name);
      count_launch();
      return launch_staged(lagrangian_planes_tma_kernel<C>, (unsigned)grid, smem, stream, p);
    }
  }
  if constexpr (C::LAYOUT == kRows && RowsTmaLayout<C>::kSupported) {
    if (p.use_tma) {
      static int cache[64] = {0};
      constexpr size_t smem = RowsTmaLayout<C>::kSmemBytes;
      int rows_blocks_per_sm = 1;
      const cudaError_t e = resident_blocks(lagrangian_rows_tma_kernel<C>, smem, cache, &rows_blocks_per_sm);
      if (e != cudaSuccess) return e;
      long long tiles = p.batch / kTmaTile;
      if (tiles < 1) tiles = 1;
      long long grid = (long long)grid_sms() * staged_blocks_per_sm<C>(rows_blocks_per_sm);
      if (grid > tiles) grid = tiles;
      if (grid > max_blocks) grid = max_blocks;
      set_last_kernel_name(name);
      count_launch();
      return launch_staged(lagrangian_rows_tma_kernel<C>, (unsigned)grid, smem, stream, p);
    }
  }
  if constexpr (C::SUMS) {
    static int cache[64] = {0};
    int blocks_per_sm = 1;
    {
      const cudaError_t e = resident_blocks(lagrangian_kernel<C>, 0, cache, &blocks_per_sm);
      if (e != cudaSuccess) return e;
    }
    const long long per_block = (long long)kThreads * (C::LAYOUT == kStrided ? IKL_ROWS_UNROLL : 2);
    long long tiles = (p.batch + per_block - 1) / per_block;
    if (tiles < 1) tiles = 1;
    long long grid = (long long)grid_sms() * blocks_per_sm;
    if (grid > tiles) grid = tiles;
    if (grid > max_blocks) grid = max_blocks;
    set_last_kernel_name(name);
    count_launch();
    lagrangian_kernel<C><<<(unsigned)grid, kThreads, 0, stream>>>(p);
    return cudaGetLastError();
  } else {
    return cudaErrorInvalidValue;      // unreachable: handled by the redirection at the top
  }
}

// Key of a specialisation inside one (LAYOUT, J, UNIT) family.
struct Variant {
  int nobs;
  bool dyn, jl;
  int mode;
  int links;      // J
  bool unit;      // every link length is exactly 1 (the compile-time-folded families)
};

// Instantiated once per family in ikl_lagrangian_<family>.cu.
template <int LAYOUT, int J, bool UNIT>
cudaError_t launch_family(const Variant& v, const LagrangianParams& p, cudaStream_t stream, int max_blocks, bool* found);

// Evaluation statistics on the staged pipelines (p.use_tma must be set); instantiated in ikl_stats_<layout>.cu.
template <class C>
cudaError_t launch_stats_cfg(const LagrangianParams& p, cudaStream_t stream, int max_blocks) {
  constexpr bool planes = C::LAYOUT == kPlanes;
  auto* kernel = planes ? stats_planes_tma_kernel<C> : stats_rows_tma_kernel<C>;
  constexpr size_t smem = planes ? TmaLayout<C>::kSmemBytes : RowsTmaLayout<C>::kSmemBytes;
  static int cache[64] = {0};
  int blocks_per_sm = 1;
  {
    const cudaError_t e = resident_blocks(kernel, smem, cache, &blocks_per_sm);
    if (e != cudaSuccess) return e;
  }
  long long tiles = planes ? (p.batch + kTmaTile - 1) / kTmaTile : p.batch / kTmaTile;
  if (tiles < 1) tiles = 1;
  long long grid = (long long)grid_sms() * blocks_per_sm;
  if (grid > tiles) grid = tiles;
  if (grid > max_blocks) grid = max_blocks;
  count_launch();
  kernel<<<(unsigned)grid, kThreads, smem, stream>>>(p);
  return cudaGetLastError();
}
template <int LAYOUT>
cudaError_t launch_stats_family(const Variant& v, const LagrangianParams& p, cudaStream_t stream, int max_blocks, bool* found);

// Argument checks, constants, layout detection and TMA descriptors shared by the entry points (ikl_api.cu).
// Returns the layout (kStrided / kRows / kPlanes) and sets p.use_tma; a negative value is -IklStatus.
int prepare_lagrangian_params(const IklSpec* spec, const IklBatch* in, const float* w_host, bool grad, float* dtheta, int64_t dt_sb,
                              int64_t dt_sj, LagrangianParams& p, Variant& v);

}  // namespace ikl
