// ikl_pipeline.cuh -- the staged (TMA / bulk-copy) kernels: shared-memory pipelines for the planes and rows layouts, the
// per-pair "bodies" they drive (Lagrangian sums + gradient, evaluation statistics) and the kernels built from them.
// Included by ikl_lagrangian.cuh (which defines LagrangianParams, Cfg, the accumulators and the per-sample IO helpers).
#pragma once

namespace ikl {

// ---- planes layout, TMA-staged ---------------------------------------------------------------------------------------
// The planes of a batch are three tensors with the sample index innermost: theta {B, 3}, target {B, 2}, obstacles
// {B, 3, n}.  A 512-sample tile is two 256-sample boxes (the hardware limit per box dimension) of each tensor: six tiled
// TMA loads (SASS UTMALDG) land a whole tile of all 17 planes in one shared-memory stage and complete on the stage's
// mbarrier.  Threads read their pair of samples with LDS.64 at immediate offsets and write d/dtheta straight to HBM.
constexpr int kTmaTile = 2 * kThreads;                 // samples per tile

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Wait for a phase of an mbarrier.  The common case is one successful poll; the spinning path is a separate (non-inlined)
// function so its poll counter costs the hot loop no register.  A copy that never completes (a transaction-count mismatch
// would do it) must not hang the GPU: after 2^26 unsuccessful polls (each try_wait suspends for a hardware time slice;
// seconds in total, against microseconds for a copy) the kernel traps and the launch reports an error instead.
static __device__ __noinline__ void mbar_wait_spin(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .u32 n;\n"
      "mov.u32 n, 0;\n"
      "IKL_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni IKL_DONE;\n"
      "add.u32 n, n, 1;\n"
      "setp.lt.u32 p, n, 0x4000000;\n"
      "@p bra.uni IKL_WAIT;\n"
      "trap;\n"
      "IKL_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  if (!done) mbar_wait_spin(bar, parity);
}
// Non-blocking probe of a phase (test_wait never suspends the thread).  IKL_EARLY_PROBE=1 probes the NEXT tile's "full"
// barrier right after handing the current stage back, a whole tile of arithmetic before the answer is needed, to take the
// ~90-cycle phase check off the path between two tiles (the branch on its result held 6.6 % of the warp-stall samples of the
// static-obstacle kernel, profiles/r02/ncu_obs4_stat_planes_final).  Measured (profiles/r02/kernel_table_probe.jsonl, interleaved
// A/B): 1-5 % SLOWER on the static / unconstrained configs, +-0.5 % on the dynamic one -- the probe mostly comes back "not
// yet" (the refill is issued by the block's slowest warp) and the blocking wait follows anyway.  Rejected; off by default.
#ifndef IKL_EARLY_PROBE
#define IKL_EARLY_PROBE 0
#endif
__device__ __forceinline__ bool mbar_test(unsigned bar, unsigned parity) {
  unsigned done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
// Tiled TMA loads: one instruction moves a {256 samples x all planes of a group} box; rows past the end of the batch
// are zero-filled by the hardware, so ragged last tiles need no special case.
// IKL_TMA_EVICT_FIRST=1: the staged loads carry an L2 evict-first policy (the batch is streamed once and never re-read).
// Measured and rejected (profiles/r02/kernel_table_cache_policy.jsonl, same-run A/B at 2^24): the benchmark kernel gets 8 %
// SLOWER (0.2116 -> 0.2280 ms fwd+bwd, 0.204 -> 0.223 gradient-only), forward-only and the light kernels do not move.
#ifndef IKL_TMA_EVICT_FIRST
#define IKL_TMA_EVICT_FIRST 0
#endif
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const TensorMap* tm, int x, int y, unsigned bar) {
#if IKL_TMA_EVICT_FIRST
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
               "l"(tm), "r"(x), "r"(y), "r"(bar), "l"(l2_evict_first_policy()) : "memory");
#else
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(tm), "r"(x), "r"(y), "r"(bar) : "memory");
#endif
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const TensorMap* tm, int x, int y, int z, unsigned bar) {
#if IKL_TMA_EVICT_FIRST
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(dst),
               "l"(tm), "r"(x), "r"(y), "r"(z), "r"(bar), "l"(l2_evict_first_policy()) : "memory");
#else
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(tm), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
#endif
}
// Programmatic dependent launch (see launch_staged): no-ops when the kernel was launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

constexpr int kTmaBox = 256;                           // samples per TMA box (the hardware limit per box dimension)
constexpr int kTmaBoxBytes = kTmaBox * 4;

// Shared-memory budget: IKL_TMA_STAGES stages of [half (2)][plane][256 samples] + the mbarriers.  Measured on B200
// (profiles/r01/kernel_table_pipeline.jsonl): 2 stages 0.2166 ms, 3 stages 0.2200 ms, 1 stage 0.2199 ms (fwd+bwd, 2^24).
#ifndef IKL_TMA_STAGES
#define IKL_TMA_STAGES 2
#endif
#ifndef IKL_PIPE_PROXY_FENCE
#define IKL_PIPE_PROXY_FENCE 1
#endif
template <class C>
struct TmaLayout {
  static constexpr int kPlanes = C::J + 2 + (C::DYN ? 3 * C::NOBS : 0);
  static constexpr int kHalfBytes = kPlanes * kTmaBoxBytes;
  static constexpr int kStageBytes = (kTmaTile / kTmaBox) * kHalfBytes;
  static constexpr int kStages = IKL_TMA_STAGES;
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 128;
};

// The gradient kernels' slow path: a pair that eval_pair<!EXACT> reported as within ~8 ulp of a tie (a tip on an obstacle's
// circle) is re-read from global memory through the generic strides, evaluated with exact distances and its d/dtheta
// stored again.  Not inlined: it costs the hot loop no registers, and it runs for about one pair in a million.
template <class C, class W2>
static __device__ __noinline__ void redo_pair_exact(const LagrangianParams& p, const W2& w2, long long bA, long long bB) {
  f2 th[C::J], ob[C::NOB][3], acc[C::K], dth[C::J], ee[3];
#pragma unroll
  for (int j = 0; j < C::J; ++j) th[j] = f2_make(p.theta[bA * p.th_sb + j * p.th_sj], p.theta[bB * p.th_sb + j * p.th_sj]);
  const f2 tx = f2_make(p.target[bA * p.tg_sb], p.target[bB * p.tg_sb]);
  const f2 ty = f2_make(p.target[bA * p.tg_sb + p.tg_sc], p.target[bB * p.tg_sb + p.tg_sc]);
  if (C::DYN) {
#pragma unroll
    for (int k = 0; k < C::NOBS; ++k) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        ob[k][c] = f2_make(p.obst[bA * p.ob_sb + k * p.ob_so + c * p.ob_sc], p.obst[bB * p.ob_sb + k * p.ob_so + c * p.ob_sc]);
    }
  } else {
    static_obstacles_packed<C>(p, ob);
  }
#pragma unroll
  for (int i = 0; i < C::K; ++i) acc[i] = f2_splat(0.f);
  eval_pair<C::J, C::UNIT, C::NOBS, C::JL, true, true, C::DYN, false>(p.pu, w2, th, tx, ty, ob, acc, dth, ee, as_f2(p.pu.ob_ndiff));
#pragma unroll
  for (int j = 0; j < C::J; ++j) {
    p.dtheta[bA * p.dt_sb + j * p.dt_sj] = f2_lo(dth[j]);
    p.dtheta[bB * p.dt_sb + j * p.dt_sj] = f2_hi(dth[j]);
  }
}

// ---- what a pipeline does with a pair of samples ----------------------------------------------------------------------
// The two staged pipelines (run_planes_pipe, run_rows_pipe) hand every pair of samples to a "body":
//   pair(th, tx, ty, ob, bA, bB)   operands of samples bA and bB (packed: lo half = bA, hi half = bB)
//   tile_done()                    once per tile, by every thread of the block (also by threads without a valid pair)
//   finish()                       after the last tile
// LagrangianBody: the K running sums (+ d/dtheta).  `r`: the rows pipeline keeps obstacle (k ^ r) in slot k; `w2` are the
// multipliers in that slot order, `w2_plain` in constraint order (for the exact redo of a near-tie pair).
template <class C, class W2>
struct LagrangianBody {
  const LagrangianParams& p;
  const W2& w2;
  const W2& w2_plain;
  WarpSharedAcc<C::K>& sacc;
  const int r;
  const f2 ndiff;                 // lo - hi slope of the circle penalty, in registers (see eval_pair)
  f2 acc[C::K];
  int since_flush;
  static constexpr bool kPermuted = C::LAYOUT == kRows && C::DYN && C::NOBS == kMaxObst;
  static constexpr int kObstBase = 1 + (C::JL ? C::J : 0);
  // cheap distances (see ikl_packed.cuh) in the forward-only and the gradient kernels alike, so both report the same
  // sums bit for bit; the gradient kernels redo near-tie pairs exactly
  static constexpr bool kExact = C::NOBS == 0 || IKL_EXACT_GRAD;
  __device__ __forceinline__ LagrangianBody(const LagrangianParams& pp, const W2& w, const W2& w_plain, WarpSharedAcc<C::K>& sa, int rr,
                                            const volatile float2* s_ndiff)
      : p(pp), w2(w), w2_plain(w_plain), sacc(sa), r(rr), ndiff(C::DYN ? as_f2(pp.pu.ob_ndiff) : f2_make(s_ndiff->x, s_ndiff->y)), since_flush(0) {
#pragma unroll
    for (int i = 0; i < C::K; ++i) acc[i] = f2_splat(0.f);
  }
  __device__ __forceinline__ void pair(const f2 (&th)[C::J], f2 tx, f2 ty, const f2 (&ob)[C::NOB][3], long long bA, long long bB) {
    f2 dth[C::J], ee[3];
    const bool near_tie = eval_pair<C::J, C::UNIT, C::NOBS, C::JL, C::GRAD, kExact, C::DYN, C::SUMS>(p.pu, w2, th, tx, ty, ob, acc, dth, ee, ndiff);
    if (C::GRAD) {
      if (C::LAYOUT == kPlanes) {                       // bB == bA + 1: one 8-byte store per plane
#pragma unroll
        for (int j = 0; j < C::J; ++j) stg_stream_f2(p.dtheta + j * p.dt_sj + bA, dth[j]);
      } else {
        float* dA = p.dtheta + bA * C::J;
        float* dB = p.dtheta + bB * C::J;
#pragma unroll
        for (int j = 0; j < C::J; ++j) {
          dA[j] = f2_lo(dth[j]);
          dB[j] = f2_hi(dth[j]);
        }
      }
    } else {
      store_side_outputs(p, bA, bB, ee, tx, ty);
    }
    if constexpr (!kExact && C::GRAD) {
      if (near_tie) redo_pair_exact<C>(p, w2_plain, bA, bB);
    }
  }
  __device__ __forceinline__ void flush() {
    if constexpr (!C::SUMS) return;             // gradient-only: there are no sums
    else if constexpr (kPermuted) sacc.template flush_xor4<kObstBase, C::J>(acc, r);
    else sacc.flush(acc);
  }
  __device__ __forceinline__ void tile_done() {
    if constexpr (C::SUMS) {
      if (++since_flush == kFlushSamples / 2) {
        since_flush = 0;
        flush();
      }
    }
  }
  __device__ __forceinline__ void finish() { flush(); }
};

// SingleStaticBody: the Lagrangian body for configs with a STATIC obstacle table.  The pipelines still hand over a pair of
// samples; the body evaluates them one after the other (a two-trip loop that is deliberately not unrolled) with
// eval_single_static, whose packed lanes are obstacle pairs.  Half the live state of LagrangianBody -- 16 running sums
// instead of 32 -- so these FP32-latency-bound kernels fit a third / fourth resident block per SM (staged_min_blocks).
#ifndef IKL_STATIC_SINGLE
#define IKL_STATIC_SINGLE 0
#endif
struct SingleHostWeights {
  const LagrangianParams& p;
  __device__ __forceinline__ float lam_s(int i) const { return p.u.w[i]; }
  __device__ __forceinline__ f2 lam_ob(int m, int q) const { return as_f2(p.sp.lam[m][q]); }
};
struct SingleSmemWeights {
  const float* lam;                 // K multipliers
  const float2* lam_pairs;          // [3][kMaxObstPairs]
  __device__ __forceinline__ float lam_s(int i) const { return lam[i]; }
  __device__ __forceinline__ f2 lam_ob(int m, int q) const { return *reinterpret_cast<const f2*>(lam_pairs + m * kMaxObstPairs + q); }
};

// device-resident multipliers as obstacle pairs [tip m][pair q] (0 in a padding lane); threads 0 .. 3*kMaxObstPairs-1
template <class C>
__device__ __forceinline__ void stage_pair_weights(const float* w_dev, float2* lam_pairs) {
  constexpr int base = 1 + (C::JL ? 3 : 0);
  const int t = threadIdx.x;
  if (t < 3 * kMaxObstPairs) {
    const int m = t / kMaxObstPairs, q = t % kMaxObstPairs, j = 2 - m;
    const float w0 = 2 * q < C::NOBS ? __ldg(w_dev + base + 3 * (2 * q) + j) : 0.f;
    const float w1 = 2 * q + 1 < C::NOBS ? __ldg(w_dev + base + 3 * (2 * q + 1) + j) : 0.f;
    lam_pairs[t] = make_float2(w0, w1);
  }
}

template <class C>
constexpr bool use_single_static() { return IKL_STATIC_SINGLE && !IKL_EXACT_GRAD && C::SUMS && !C::DYN && C::NOBS > 0 && C::J == 3 && C::UNIT; }

template <class C, class W1, class W2>
struct SingleStaticBody {
  const LagrangianParams& p;
  const W1& w1;
  const W2& w2_plain;               // packed multipliers in constraint order, for the exact redo of a near-tie pair
  WarpSharedAcc<C::K>& sacc;
  const f2 ndiff;
  float acc_s[4];
  f2 acc_ob[3][kMaxObstPairs];
  int since_flush;
  static constexpr int kObstBase = 1 + (C::JL ? 3 : 0);
  static constexpr int NP = (C::NOBS + 1) / 2;
  __device__ __forceinline__ SingleStaticBody(const LagrangianParams& pp, const W1& w, const W2& w_plain, WarpSharedAcc<C::K>& sa,
                                              const volatile float2* s_ndiff)
      : p(pp), w1(w), w2_plain(w_plain), sacc(sa), ndiff(f2_make(s_ndiff->x, s_ndiff->y)), since_flush(0) {
    clear();
  }
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc_s[i] = 0.f;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
#pragma unroll
      for (int q = 0; q < kMaxObstPairs; ++q) acc_ob[m][q] = f2_splat(0.f);
    }
  }
  __device__ __forceinline__ void pair(const f2 (&th)[3], f2 tx, f2 ty, const f2 (&)[C::NOB][3], long long bA, long long bB) {
    bool near_tie = false;
    float keep[3];                    // sample A's d/dtheta (planes: stored together with sample B's as one 8-byte store per plane)
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      float t3[3], dth[3], ee[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) t3[j] = h ? f2_hi(th[j]) : f2_lo(th[j]);
      const float x = h ? f2_hi(tx) : f2_lo(tx), y = h ? f2_hi(ty) : f2_lo(ty);
      near_tie |= eval_single_static<C::NOBS, C::JL, C::GRAD>(p.pu, p.sp, w1, t3, x, y, acc_s, acc_ob, dth, ee, ndiff);
      const long long b = h ? bB : bA;
      if (C::GRAD) {
        if (C::LAYOUT == kPlanes) {
          if (h == 0) {
#pragma unroll
            for (int j = 0; j < 3; ++j) keep[j] = dth[j];
          } else {
#pragma unroll
            for (int j = 0; j < 3; ++j) stg_stream_f2(p.dtheta + j * p.dt_sj + bA, f2_make(keep[j], dth[j]));
          }
        } else {
          float* d = p.dtheta + b * 3;
#pragma unroll
          for (int j = 0; j < 3; ++j) d[j] = dth[j];
        }
      } else {
        if (p.q_ee) {
#pragma unroll
          for (int c = 0; c < 3; ++c) p.q_ee[b * 3 + c] = ee[c];
        }
        if (p.err_sq) {
          const float ex = x - ee[0], ey = y - ee[1];
          p.err_sq[b * 2] = 0.5f * (ex * ex);
          p.err_sq[b * 2 + 1] = 0.5f * (ey * ey);
        }
      }
    }
    if constexpr (C::GRAD) {
      if (near_tie) redo_pair_exact<C>(p, w2_plain, bA, bB);
    }
  }
  __device__ __forceinline__ void flush() {
    float v[C::K];
    v[0] = acc_s[0];
    if (C::JL) {
#pragma unroll
      for (int j = 0; j < 3; ++j) v[1 + j] = acc_s[1 + j];
    }
#pragma unroll
    for (int o = 0; o < C::NOBS; ++o) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const f2 a = acc_ob[2 - j][o >> 1];
        v[kObstBase + 3 * o + j] = (o & 1) ? f2_hi(a) : f2_lo(a);
      }
    }
    clear();
    sacc.flush(v);
  }
  __device__ __forceinline__ void tile_done() {
    if (++since_flush == kFlushSamples / 2) {
      since_flush = 0;
      flush();
    }
  }
  __device__ __forceinline__ void finish() { flush(); }
};

// StatsBody: the evaluation statistics of scripts/evaluation/visualize_ik_solutions.py:97-119 (reference) -- per-constraint
// counts of per-sample terms above a threshold, samples violating any joint limit / any obstacle (term > 0) and the
// summed end-effector error sqrt(sum(position_err_sq)).  Same per-sample maths as the Lagrangian (eval_pair with fresh sums).
constexpr int kStatSlots = kMaxK + 2;
template <class C>
struct StatsBody {
  const LagrangianParams& p;
  const int r;
  unsigned cnt[C::K];
  unsigned any_jl, any_ob;
  double err;
  static constexpr bool kPermuted = C::LAYOUT == kRows && C::DYN && C::NOBS == kMaxObst;
  static constexpr int kObstBase = 1 + (C::JL ? C::J : 0);
  __device__ __forceinline__ StatsBody(const LagrangianParams& pp, int rr) : p(pp), r(rr), any_jl(0), any_ob(0), err(0.0) {
#pragma unroll
    for (int i = 0; i < C::K; ++i) cnt[i] = 0;
  }
  __device__ __forceinline__ void add(const float (&v)[C::K]) {
    float mjl = -1.f, mob = -1.f;
#pragma unroll
    for (int i = 0; i < C::K; ++i) {
      cnt[i] += v[i] > p.stats_threshold;
      if (C::JL && i >= 1 && i < 1 + C::J) mjl = fmaxf(mjl, v[i]);
      if (i >= kObstBase) mob = fmaxf(mob, v[i]);
    }
    any_jl += mjl > 0.f;
    any_ob += mob > 0.f;
    err += (double)__fsqrt_rn(v[0]);      // sqrt(position_err_sq.sum()) with position_err_sq = 0.5*e^2: the reference's |e|/sqrt(2)
  }
  __device__ __forceinline__ void pair(const f2 (&th)[C::J], f2 tx, f2 ty, const f2 (&ob)[C::NOB][3], long long, long long) {
    f2 acc[C::K], dth[C::J], ee[3];
#pragma unroll
    for (int i = 0; i < C::K; ++i) acc[i] = f2_splat(0.f);
    const PackedHostWeights w2{p.pu};
    eval_pair<C::J, C::UNIT, C::NOBS, C::JL, false, true, C::DYN, true>(p.pu, w2, th, tx, ty, ob, acc, dth, ee, as_f2(p.pu.ob_ndiff));
    float v[C::K];
#pragma unroll
    for (int i = 0; i < C::K; ++i) v[i] = f2_lo(acc[i]);
    add(v);
#pragma unroll
    for (int i = 0; i < C::K; ++i) v[i] = f2_hi(acc[i]);
    add(v);
  }
  __device__ __forceinline__ void tile_done() {}
  // after the last tile: put the obstacle counters back in constraint order (rows pipeline), so that the caller can add
  // tail samples in plain order before commit()
  __device__ __forceinline__ void finish() {
    if constexpr (kPermuted) {
#pragma unroll
      for (int bit = 1; bit <= 2; bit <<= 1) {
        const bool sw = (r & bit) != 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k & bit) continue;
#pragma unroll
          for (int j = 0; j < C::J; ++j) {
            const unsigned a = cnt[kObstBase + C::J * k + j], b = cnt[kObstBase + C::J * (k | bit) + j];
            cnt[kObstBase + C::J * k + j] = sw ? b : a;
            cnt[kObstBase + C::J * (k | bit) + j] = sw ? a : b;
          }
        }
      }
    }
  }
  // warp sums -> the block's shared counters (s_counts zeroed by the caller before the pipeline's first barrier)
  __device__ __forceinline__ void commit(unsigned* s_counts, double* s_err) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < C::K; ++i) {
      const unsigned v = __reduce_add_sync(0xffffffffu, cnt[i]);
      if (lane == 0 && v) atomicAdd(&s_counts[i], v);
    }
    const unsigned vj = __reduce_add_sync(0xffffffffu, any_jl), vo = __reduce_add_sync(0xffffffffu, any_ob);
    double e = err;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
    if (lane == 0) {
      if (vj) atomicAdd(&s_counts[kMaxK], vj);
      if (vo) atomicAdd(&s_counts[kMaxK + 1], vo);
      s_err[threadIdx.x >> 5] = e;
    }
  }
};

// block totals -> global (64-bit counters, fp64 error sum)
__device__ __forceinline__ void stats_block_finalize(const LagrangianParams& p, const unsigned* s_counts, const double* s_err) {
  __syncthreads();
  if (threadIdx.x < kStatSlots && s_counts[threadIdx.x]) atomicAdd(p.stats_counts + threadIdx.x, (unsigned long long)s_counts[threadIdx.x]);
  if (threadIdx.x == 0) {
    double e = 0.0;
#pragma unroll
    for (int wq = 0; wq < kWarps; ++wq) e += s_err[wq];
    atomicAdd(p.stats_pos_err, e);
  }
}

// Arrive on an mbarrier (release: the caller's shared-memory reads are performed first) and report whether this was the
// arrival that completed the phase: the returned state is the one BEFORE the arrival, so a pending count of 1 means "last".
__device__ __forceinline__ bool mbar_arrive_is_last(unsigned bar) {
  unsigned long long state;
  unsigned pending;
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 %0, [%1];" : "=l"(state) : "r"(bar) : "memory");
  asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(pending) : "l"(state));
  return pending == 1u;
}

// Barrier-free pipeline.  A thread reads its operands from a stage at the START of a tile and then computes for ~450
// instructions without touching shared memory, so the stage is handed back as soon as every warp has done its loads --
// not a whole tile later, and without a block barrier.  Each warp arrives on the stage's "drained" mbarrier (count =
// warps per block) after its loads; the warp whose arrival completes the phase re-arms the stage's "full" mbarrier and
// issues the TMA loads of the tile that goes there next.  Nobody waits for anybody, only for data.  (Round-1 history,
// profiles/r01/README.md: with 17 one-dimensional bulk copies per tile the issuing warp ran ~45 % more instructions than
// its siblings and throttled the block whichever warp it was; six UTMALDG make the refill cheap enough to hand around.)
template <class C, class Body>
__device__ __forceinline__ void run_planes_pipe(const LagrangianParams& p, Body& body, unsigned char* stage_mem, unsigned long long* bars) {
  using L = TmaLayout<C>;
  constexpr int S = L::kStages;
  static_assert(kTmaTile % kTmaBox == 0, "a tile is a whole number of 256-sample boxes; thread v owns samples (2v, 2v+1): box = v / 128");
  // 32-bit indices: the host selects this kernel only for batches below 2^31 - 2 tiles (prepare_lagrangian_params)
  const int B = (int)p.batch;
  const int ntiles = (B + kTmaTile - 1) / kTmaTile;
  const int tid = threadIdx.x, lane = tid & 31;
  const unsigned bar0 = smem_u32(bars);              // S "full" barriers, then S "drained" barriers
  const unsigned drained0 = bar0 + 8 * S;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      mbar_init(bar0 + 8 * s, 1);
      mbar_init(drained0 + 8 * s, kWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();                                         // the previous kernel on the stream is complete: global memory may be touched

  // Stage layout: [half (2)][plane][256 samples]; each half is what one box of every group delivers.
  auto issue = [&](int tile, int stage) {            // one thread
    const int b0 = tile * kTmaTile;
    const unsigned bar = bar0 + 8 * stage;
    const unsigned dst = smem_u32(stage_mem + (size_t)stage * L::kStageBytes);
    mbar_expect_tx(bar, (unsigned)L::kStageBytes);    // boxes always deliver their full size (zero-filled past the batch end)
#pragma unroll
    for (int h = 0; h < kTmaTile / kTmaBox; ++h) {
      const unsigned d = dst + h * L::kHalfBytes;
      const int x = b0 + h * kTmaBox;
      tma_load_2d(d, &p.tm_theta, x, 0, bar);
      tma_load_2d(d + C::J * kTmaBoxBytes, &p.tm_target, x, 0, bar);
      if (C::DYN) tma_load_3d(d + (C::J + 2) * kTmaBoxBytes, &p.tm_obst, x, 0, 0, bar);
    }
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int t = (int)blockIdx.x + s * (int)gridDim.x;
      if (t < ntiles) issue(t, s);
    }
  }

  int stage = 0;
  unsigned parity = 0;
  bool ready = false;                                  // the next tile's data was seen complete by an early probe
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (!ready) mbar_wait(bar0 + 8 * stage, parity);
    const int b = tile * kTmaTile + 2 * tid;
    const unsigned char* base = stage_mem + (size_t)stage * L::kStageBytes + (tid >> 7) * L::kHalfBytes + 8 * (tid & 127);
    f2 th[C::J], ob[C::NOB][3];
    int pl = 0;
#pragma unroll
    for (int j = 0; j < C::J; ++j, ++pl) th[j] = *reinterpret_cast<const f2*>(base + pl * kTmaBoxBytes);
    const f2 tx = *reinterpret_cast<const f2*>(base + C::J * kTmaBoxBytes);
    const f2 ty = *reinterpret_cast<const f2*>(base + (C::J + 1) * kTmaBoxBytes);
    pl = C::J + 2;
    if (C::DYN) {
#pragma unroll
      for (int k = 0; k < C::NOBS; ++k) {
#pragma unroll
        for (int c = 0; c < 3; ++c, ++pl) ob[k][c] = *reinterpret_cast<const f2*>(base + pl * kTmaBoxBytes);
      }
    } else {
      static_obstacles_packed<C>(p, ob);
    }
    // hand the stage back: the last of the block's warps to get here refills it
    __syncwarp();
    if (lane == 0) {
      const bool last = mbar_arrive_is_last(drained0 + 8 * stage);
      const int next = tile + S * (int)gridDim.x;
      if (last && next < ntiles) {
#if IKL_PIPE_PROXY_FENCE
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
        issue(next, stage);
      }
    }
    __syncwarp();
#if IKL_EARLY_PROBE
    {
      const int nstage = stage + 1 == S ? 0 : stage + 1;
      ready = (tile + (int)gridDim.x < ntiles) && mbar_test(bar0 + 8 * nstage, nstage == 0 ? parity ^ 1u : parity);
    }
#endif
    if (b < B) body.pair(th, tx, ty, ob, b, b + 1);
    body.tile_done();
    if (++stage == S) {
      stage = 0;
      parity ^= 1u;
    }
  }
  body.finish();
}

// Resident blocks per SM the staged kernels are compiled for: the dynamic-obstacle kernels are HBM-bound and run best with
// 128 registers (2 blocks); the kernels without per-sample obstacles keep fewer values live (no obstacle operands) and are
// FP32-latency bound, so they trade registers for a third resident block (IKL_STATIC_MIN_BLOCKS).
#ifndef IKL_STATIC_MIN_BLOCKS
#define IKL_STATIC_MIN_BLOCKS 2
#endif
// The gradient-only kernels carry no running sums (up to 32 registers less).  Measured (profiles/r02/kernel_table_gradonly.jsonl,
// same-run A/B at 2^24 samples): with ONE static obstacle they fit a third resident block without spills (78-80 registers) and
// gain 3-9 % from it (planes 0.108 -> 0.098-0.101 ms, rows 0.107 -> 0.103 ms); with four static obstacles the third block
// costs 64-72 bytes of spills and 5 % of the time, and the kernels WITH sums lose up to 35 % (600 bytes of spills).
#ifndef IKL_GRADONLY_STATIC_MIN_BLOCKS
#define IKL_GRADONLY_STATIC_MIN_BLOCKS 0          // 0: three blocks for one static obstacle, IKL_STATIC_MIN_BLOCKS otherwise
#endif
template <class C>
constexpr int staged_min_blocks() {
  if (C::DYN) return IKL_MIN_BLOCKS;
  if (C::SUMS) return IKL_STATIC_MIN_BLOCKS;
  if (IKL_GRADONLY_STATIC_MIN_BLOCKS > 0) return IKL_GRADONLY_STATIC_MIN_BLOCKS;
  return C::NOBS == 1 ? 3 : IKL_STATIC_MIN_BLOCKS;
}

template <class C>
__global__ void __launch_bounds__(kThreads, staged_min_blocks<C>()) lagrangian_planes_tma_kernel(const __grid_constant__ LagrangianParams p) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  using L = TmaLayout<C>;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(dyn_smem + (size_t)L::kStages * L::kStageBytes);
  __shared__ float2 s_lam2[kMaxK];
  __shared__ double s_warp[kWarps][kMaxK];
  __shared__ float s_lam1[kMaxK];
  __shared__ float2 s_lam_pairs[3][kMaxObstPairs];
  pdl_launch_dependents();
  if constexpr (C::DEVW) pdl_wait();           // the multipliers are read from global memory right away
  if constexpr (C::DEVW) {
    if (threadIdx.x < C::K) {
      const float l = __ldg(p.w_dev + threadIdx.x);
      s_lam2[threadIdx.x] = make_float2(l, l);
      s_lam1[threadIdx.x] = l;
    }
    if constexpr (use_single_static<C>()) stage_pair_weights<C>(p.w_dev, &s_lam_pairs[0][0]);
  }
  __shared__ float2 s_ndiff;
  if (threadIdx.x == 0) s_ndiff = p.pu.ob_ndiff;
  WarpSharedAcc<C::K> sacc(s_warp);          // zeroes s_warp; run_planes_pipe starts with a block barrier
  __syncthreads();
  auto run = [&](const auto& w2) {
    if constexpr (use_single_static<C>()) {
      auto run1 = [&](const auto& w1) {
        SingleStaticBody<C, std::decay_t<decltype(w1)>, std::decay_t<decltype(w2)>> body(p, w1, w2, sacc, &s_ndiff);
        run_planes_pipe<C>(p, body, dyn_smem, bars);
      };
      if constexpr (C::DEVW) run1(SingleSmemWeights{s_lam1, &s_lam_pairs[0][0]});
      else run1(SingleHostWeights{p});
    } else {
      LagrangianBody<C, std::decay_t<decltype(w2)>> body(p, w2, w2, sacc, 0, &s_ndiff);
      run_planes_pipe<C>(p, body, dyn_smem, bars);
    }
  };
  if constexpr (C::DEVW) run(PackedSmemWeights{s_lam2});
  else run(PackedHostWeights{p.pu});
  if constexpr (C::SUMS) block_finalize<C::K>(p, s_warp);
}

// ---- rows layout, bulk-copy staged -------------------------------------------------------------------------------------
// The reference's own tensors are contiguous per tile: 512 rows of theta (6 KiB), of q_ee_desired (4 or 6 KiB) and of
// dynamic_obstacles (32 KiB) -- three one-dimensional bulk copies (SASS UBLKCP) per tile into the same barrier-free
// pipeline as the planes kernel.  Thread t owns samples (t, t + 256) of the tile, like the direct-load rows kernel.
// Shared-memory reads: theta / target rows are 3 (2) words apart -- conflict-free; an obstacle row is 16 words and a
// sample 64 words, so the four LDS.128 of a sample are issued in the order k ^ r, r = (t >> 1) & 3, which spreads the
// eight lanes of a quarter-warp over all 32 banks.  The thread therefore holds obstacle (k ^ r) in slot k; the maths is
// symmetric in the obstacles except for the multiplier and the sum each term belongs to: multipliers come from a
// per-r permuted table in shared memory, sums are put back in order when they are flushed (WarpSharedAcc::flush_xor4).
// With fewer than four dynamic obstacles (the reference still hands over (B,4,4) rows, kinematics/dataset.py:80-85) the
// whole tile is staged the same way and slot k simply holds obstacle k (r = 0): the one to three LDS.128 per sample then
// take 2- to 4-way bank conflicts, i.e. no more shared-memory cycles than the four conflict-free loads of the full case.
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar) : "memory");
}

template <class C>
struct RowsTmaLayout {
  static constexpr bool kSupported = true;
  static constexpr bool kPermuted = C::DYN && C::NOBS == kMaxObst;         // obstacle (k ^ r) in slot k
  static constexpr int kThetaBytes = kTmaTile * C::J * 4;
  static constexpr int kTargetBytes = kTmaTile * 3 * 4;                 // room for target rows of 2 or 3 floats
  static constexpr int kObstBytes = C::DYN ? kTmaTile * 4 * kMaxObst * 4 : 0;
  static constexpr int kStageBytes = kThetaBytes + kTargetBytes + kObstBytes;
  static constexpr int kStages = IKL_TMA_STAGES;
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 128;
  static constexpr int kObstBase = 1 + (C::JL ? C::J : 0);              // index of the first obstacle constraint
};

template <class C, class Body>
__device__ __forceinline__ void run_rows_pipe(const LagrangianParams& p, Body& body, unsigned char* stage_mem, unsigned long long* bars, int r) {
  using L = RowsTmaLayout<C>;
  constexpr int S = L::kStages;
  // 32-bit tile indices: the host selects this kernel only for batches below 2^31 (prepare_lagrangian_params)
  const int ntiles = (int)(p.batch / kTmaTile);      // full tiles only; the caller evaluates the ragged tail
  const int tid = threadIdx.x, lane = tid & 31;
  const int tg_sb = (int)p.tg_sb;                    // 2 or 3 (checked on the host)
  const unsigned bar0 = smem_u32(bars);              // S "full" barriers, then S "drained" barriers
  const unsigned drained0 = bar0 + 8 * S;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      mbar_init(bar0 + 8 * s, 1);
      mbar_init(drained0 + 8 * s, kWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();                                         // the previous kernel on the stream is complete: global memory may be touched

  auto issue = [&](int tile, int stage) {            // one thread
    const long long b0 = (long long)tile * kTmaTile;
    const unsigned bar = bar0 + 8 * stage;
    const unsigned dst = smem_u32(stage_mem + (size_t)stage * L::kStageBytes);
    const unsigned tg_bytes = (unsigned)(kTmaTile * 4) * (unsigned)tg_sb;
    mbar_expect_tx(bar, L::kThetaBytes + tg_bytes + L::kObstBytes);
    bulk_g2s(dst, p.theta + b0 * C::J, L::kThetaBytes, bar);
    bulk_g2s(dst + L::kThetaBytes, p.target + b0 * tg_sb, tg_bytes, bar);
    if (C::DYN) bulk_g2s(dst + L::kThetaBytes + L::kTargetBytes, p.obst + b0 * (4 * kMaxObst), L::kObstBytes, bar);
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int t = (int)blockIdx.x + s * (int)gridDim.x;
      if (t < ntiles) issue(t, s);
    }
  }

  int stage = 0;
  unsigned parity = 0;
  bool ready = false;                                  // the next tile's data was seen complete by an early probe
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (!ready) mbar_wait(bar0 + 8 * stage, parity);
    const unsigned char* st = stage_mem + (size_t)stage * L::kStageBytes;
    f2 th[C::J], ob[C::NOB][3];
    const float* tA = reinterpret_cast<const float*>(st) + C::J * tid;
    const float* tB = tA + C::J * kThreads;
#pragma unroll
    for (int j = 0; j < C::J; ++j) th[j] = f2_make(tA[j], tB[j]);
    const float* gA = reinterpret_cast<const float*>(st + L::kThetaBytes) + tg_sb * tid;
    const float* gB = gA + tg_sb * kThreads;
    const f2 tx = f2_make(gA[0], gB[0]);
    const f2 ty = f2_make(gA[1], gB[1]);
    if constexpr (C::DYN) {
      const float4* oA = reinterpret_cast<const float4*>(st + L::kThetaBytes + L::kTargetBytes) + kMaxObst * tid;
      const float4* oB = oA + kMaxObst * kThreads;
#pragma unroll
      for (int k = 0; k < C::NOBS; ++k) {
        const float4 a = oA[k ^ r], b4 = oB[k ^ r];
        ob[k][0] = f2_make(a.x, b4.x);
        ob[k][1] = f2_make(a.y, b4.y);
        ob[k][2] = f2_make(a.z, b4.z);
      }
    } else {
      static_obstacles_packed<C>(p, ob);
    }
    // hand the stage back: the last of the block's warps to get here refills it
    __syncwarp();
    if (lane == 0) {
      const bool last = mbar_arrive_is_last(drained0 + 8 * stage);
      const int next = tile + S * (int)gridDim.x;
      if (last && next < ntiles) {
#if IKL_PIPE_PROXY_FENCE
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
        issue(next, stage);
      }
    }
    __syncwarp();
#if IKL_EARLY_PROBE
    {
      const int nstage = stage + 1 == S ? 0 : stage + 1;
      ready = (tile + (int)gridDim.x < ntiles) && mbar_test(bar0 + 8 * nstage, nstage == 0 ? parity ^ 1u : parity);
    }
#endif
    const long long bA = (long long)tile * kTmaTile + tid;
    body.pair(th, tx, ty, ob, bA, bA + kThreads);
    body.tile_done();
    if (++stage == S) {
      stage = 0;
      parity ^= 1u;
    }
  }
  body.finish();
}

template <class C>
__global__ void __launch_bounds__(kThreads, staged_min_blocks<C>()) lagrangian_rows_tma_kernel(const __grid_constant__ LagrangianParams p) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  using L = RowsTmaLayout<C>;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(dyn_smem + (size_t)L::kStages * L::kStageBytes);
  __shared__ float2 s_lam_perm[4][kMaxK];      // multipliers as (w, w) pairs, obstacle entries permuted by k ^ r
  __shared__ double s_warp[kWarps][kMaxK];
  const int r = L::kPermuted ? ((threadIdx.x >> 1) & 3) : 0;
  constexpr bool kSmemWeights = C::GRAD && (C::DYN || C::DEVW);
  pdl_launch_dependents();
  if constexpr (C::DEVW) pdl_wait();           // the multipliers are read from global memory right away
  if constexpr (kSmemWeights) {
    if (threadIdx.x < 4 * kMaxK) {
      const int rr = threadIdx.x / kMaxK, i = threadIdx.x % kMaxK;
      int src = i;
      if (L::kPermuted && i >= L::kObstBase && i < C::K) {
        const int q = i - L::kObstBase, k = q / C::J, j = q % C::J;
        src = L::kObstBase + C::J * (k ^ rr) + j;
      }
      float l = 0.f;
      if (i < C::K) l = (C::DEVW) ? __ldg(p.w_dev + src) : p.u.w[src];
      s_lam_perm[rr][i] = make_float2(l, l);
    }
  }
  __shared__ float s_lam1[kMaxK];
  __shared__ float2 s_lam_pairs[3][kMaxObstPairs];
  if constexpr (C::DEVW && use_single_static<C>()) {
    if (threadIdx.x < C::K) s_lam1[threadIdx.x] = __ldg(p.w_dev + threadIdx.x);
    stage_pair_weights<C>(p.w_dev, &s_lam_pairs[0][0]);
  }
  WarpSharedAcc<C::K> sacc(s_warp);            // zeroes s_warp; run_rows_pipe starts with a block barrier
  __shared__ float2 s_ndiff;
  if (threadIdx.x == 0) s_ndiff = p.pu.ob_ndiff;
  __syncthreads();
  auto run = [&](const auto& w2, const auto& w2_plain) {
    if constexpr (use_single_static<C>()) {
      auto run1 = [&](const auto& w1) {
        SingleStaticBody<C, std::decay_t<decltype(w1)>, std::decay_t<decltype(w2_plain)>> body(p, w1, w2_plain, sacc, &s_ndiff);
        run_rows_pipe<C>(p, body, dyn_smem, bars, r);
      };
      if constexpr (C::DEVW) run1(SingleSmemWeights{s_lam1, &s_lam_pairs[0][0]});
      else run1(SingleHostWeights{p});
    } else {
      LagrangianBody<C, std::decay_t<decltype(w2)>> body(p, w2, w2_plain, sacc, r, &s_ndiff);
      run_rows_pipe<C>(p, body, dyn_smem, bars, r);
    }
  };
  if constexpr (kSmemWeights) run(PackedSmemWeights{s_lam_perm[r]}, PackedSmemWeights{s_lam_perm[0]});     // table 0: k ^ 0 = k
  else run(PackedHostWeights{p.pu}, PackedHostWeights{p.pu});
  // ragged tail (< 512 samples), one sample per thread through the scalar code
  {
    const long long full = (p.batch / kTmaTile) * kTmaTile;
    float sacc1[C::K];
#pragma unroll
    for (int i = 0; i < C::K; ++i) sacc1[i] = 0.f;
    auto tail = [&](const auto& w) {
      for (long long b = full + (long long)blockIdx.x * kThreads + threadIdx.x; b < p.batch; b += (long long)gridDim.x * kThreads) {
        float th[C::J], tx, ty, ob[C::NOB][3], dth[C::J], ee[3];
        if (!C::DYN) static_obstacles<C>(p, ob);
        load_sample<C>(p, b, th, tx, ty, ob);
        eval_sample<C::J, C::NOBS, C::JL, C::GRAD, C::UNIT>(p.u, w, th, tx, ty, ob, sacc1, dth, ee);
        store_sample<C>(p, b, dth, ee, tx, ty);
      }
    };
    if (full < p.batch) {
      if constexpr (C::DEVW) tail(DeviceWeights<C::K>(p.u, p.w_dev));
      else tail(HostWeights(p.u, nullptr));
    }
    if constexpr (C::SUMS) sacc.flush(sacc1);
  }
  if constexpr (C::SUMS) block_finalize<C::K>(p, s_warp);
}

// ---- evaluation statistics through the same pipelines -------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(kThreads, IKL_MIN_BLOCKS) stats_planes_tma_kernel(const __grid_constant__ LagrangianParams p) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  using L = TmaLayout<C>;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(dyn_smem + (size_t)L::kStages * L::kStageBytes);
  __shared__ unsigned s_counts[kStatSlots];
  __shared__ double s_err[kWarps];
  if (threadIdx.x < kStatSlots) s_counts[threadIdx.x] = 0;      // ordered by the pipeline's first block barrier
  StatsBody<C> body(p, 0);
  run_planes_pipe<C>(p, body, dyn_smem, bars);
  body.commit(s_counts, s_err);
  stats_block_finalize(p, s_counts, s_err);
}

template <class C>
__global__ void __launch_bounds__(kThreads, IKL_MIN_BLOCKS) stats_rows_tma_kernel(const __grid_constant__ LagrangianParams p) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  using L = RowsTmaLayout<C>;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(dyn_smem + (size_t)L::kStages * L::kStageBytes);
  __shared__ unsigned s_counts[kStatSlots];
  __shared__ double s_err[kWarps];
  if (threadIdx.x < kStatSlots) s_counts[threadIdx.x] = 0;
  const int r = L::kPermuted ? ((threadIdx.x >> 1) & 3) : 0;
  StatsBody<C> body(p, r);
  run_rows_pipe<C>(p, body, dyn_smem, bars, r);
  // ragged tail (< 512 rows), one sample per thread through the scalar code
  const long long full = (p.batch / kTmaTile) * kTmaTile;
  const HostWeights w(p.u, nullptr);
  for (long long b = full + (long long)blockIdx.x * kThreads + threadIdx.x; b < p.batch; b += (long long)gridDim.x * kThreads) {
    float th[C::J], tx, ty, ob[C::NOB][3], dth[C::J], ee[3], acc[C::K];
    if (!C::DYN) static_obstacles<C>(p, ob);
    load_sample<C>(p, b, th, tx, ty, ob);
#pragma unroll
    for (int i = 0; i < C::K; ++i) acc[i] = 0.f;
    eval_sample<C::J, C::NOBS, C::JL, false, C::UNIT>(p.u, w, th, tx, ty, ob, acc, dth, ee);
    body.add(acc);
  }
  body.commit(s_counts, s_err);
  stats_block_finalize(p, s_counts, s_err);
}

}  // namespace ikl
