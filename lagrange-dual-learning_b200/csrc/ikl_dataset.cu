// ikl_dataset.cu -- the callers either side of the hot path, on the device:
//   * dataset generation (reference: kinematics/dataset.py:52-121): uniform joint angles inside the limits, forward
//     kinematics for the targets, rejection sampling of per-sample obstacle centres (dynamic configs) or of the joint
//     angles themselves (static configs) with the reference's collision rule (kinematics/penalties.py:75-92: a joint
//     tip or the arm base within `radius`, `<=`).  One Philox stream per example, so the result depends only on
//     (seed, index) -- the reference's host loop is O(N) Python (3.5 h extrapolated for 2^24 examples).
//   * evaluation statistics (reference: scripts/evaluation/visualize_ik_solutions.py:97-119): per-constraint counts of
//     per-sample violations > 1e-4, how many samples violate any joint limit / any obstacle (> 0), and the summed
//     end-effector position error sqrt(sum(position_err_sq)).
#include <curand_kernel.h>
#include <math.h>

#include "ikl_common.h"
#include "ikl_device.cuh"
#include "ikl_lagrangian.cuh"

namespace ikl {
namespace {

constexpr int kDsThreads = 128;
constexpr int kMaxRedraws = 100000;     // an obstacle covering the arm base would loop forever in the reference

struct DatasetParams {
  long long n;
  unsigned long long seed;
  float lo, hi;
  int n_dynamic;                 // > 0: dynamic circles of radius dyn_radius
  float dyn_radius;
  int n_static;
  float stat[kMaxObst][3];
  float* theta;                  // (N,3)
  float* ee;                     // (N,3)  (x, y, phi) of the end effector
  float* obstacles;              // (N,4,4) rows [x, y, r, 0] or null
  unsigned int* exhausted;       // count of examples that hit kMaxRedraws
};

__device__ __forceinline__ float uniform_in(curandStatePhilox4_32_10_t& st, float lo, float hi) {
  const float u = 1.0f - curand_uniform(&st);       // curand_uniform is (0, 1]; the reference's uniform_ is [lo, hi)
  return fmaf(hi - lo, u, lo);
}

__device__ __forceinline__ void fk3(const float (&th)[3], float (&px)[3], float (&py)[3], float& phi3) {
  float phi = 0.f, x = 0.f, y = 0.f;
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    phi = (m == 0) ? th[0] : __fadd_rn(phi, th[m]);
    float s, c;
    sincosf(phi, &s, &c);
    x = (m == 0) ? c : __fadd_rn(x, c);
    y = (m == 0) ? s : __fadd_rn(y, s);
    px[m] = x;
    py[m] = y;
  }
  phi3 = phi;
}

__device__ __forceinline__ bool tips_hit(const float (&px)[3], const float (&py)[3], float cx, float cy, float r) {
  bool hit = false;
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    const float dx = __fsub_rn(px[m], cx), dy = __fsub_rn(py[m], cy);
    hit |= __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))) <= r;
  }
  return hit;
}

__global__ void __launch_bounds__(kDsThreads) dataset_kernel(DatasetParams p) {
  for (long long i = (long long)blockIdx.x * kDsThreads + threadIdx.x; i < p.n; i += (long long)gridDim.x * kDsThreads) {
    curandStatePhilox4_32_10_t st;
    curand_init(p.seed, (unsigned long long)i, 0ull, &st);
    float th[3], px[3], py[3], phi3;
#pragma unroll
    for (int j = 0; j < 3; ++j) th[j] = uniform_in(st, p.lo, p.hi);
    fk3(th, px, py, phi3);
    bool gave_up = false;
    // static obstacles: re-draw the joint angles until the CURRENT obstacle is cleared (the reference does not
    // re-check earlier obstacles after a re-draw either: kinematics/dataset.py:111-119)
    for (int o = 0; o < p.n_static; ++o) {
      const float cx = p.stat[o][0], cy = p.stat[o][1], r = p.stat[o][2];
      int tries = 0;
      while (tips_hit(px, py, cx, cy, r)) {
        if (++tries > kMaxRedraws) { gave_up = true; break; }
#pragma unroll
        for (int j = 0; j < 3; ++j) th[j] = uniform_in(st, p.lo, p.hi);
        fk3(th, px, py, phi3);
      }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) p.theta[i * 3 + j] = th[j];
    p.ee[i * 3 + 0] = px[2];
    p.ee[i * 3 + 1] = py[2];
    p.ee[i * 3 + 2] = phi3;
    if (p.obstacles) {
      float* row = p.obstacles + i * 16;
      const float r = p.dyn_radius;
      for (int o = 0; o < kMaxObst; ++o) {
        float cx = 0.f, cy = 0.f;
        if (o < p.n_dynamic) {
          int tries = 0;
          while (true) {
            cx = uniform_in(st, -1.2f, 1.2f);                 // kinematics/dataset.py:101
            cy = uniform_in(st, -1.2f, 1.2f);
            const bool base_hit = sqrt((double)__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy))) <= (double)r;
            if (!tips_hit(px, py, cx, cy, r) && !base_hit) break;
            if (++tries > kMaxRedraws) { gave_up = true; break; }
          }
        }
        row[4 * o + 0] = cx;
        row[4 * o + 1] = cy;
        row[4 * o + 2] = r;                                   // dataset.py:85 fills the radius of all four slots
        row[4 * o + 3] = 0.f;
      }
    }
    if (gave_up) atomicAdd(p.exhausted, 1u);
  }
}

// ---- evaluation statistics --------------------------------------------------------------------------------------
struct StatsParams {
  Uniforms u;
  long long batch;
  const float* theta;   long long th_sb, th_sj;
  const float* target;  long long tg_sb, tg_sc;
  const float* obst;    long long ob_sb, ob_so, ob_sc;
  float threshold;
  unsigned long long* counts;      // [K] per-constraint (term > threshold), then [K] any-JL, [K+1] any-obstacle
  double* pos_err_sum;             // sum_b sqrt(sum(position_err_sq_b))
};

template <int NOBS, bool DYN, bool JL>
__global__ void __launch_bounds__(256) stats_kernel(const __grid_constant__ StatsParams p) {
  constexpr int K = num_constraints(3, NOBS, JL);
  constexpr int NOB = NOBS > 0 ? NOBS : 1;
  __shared__ unsigned int s_counts[kMaxK + 2];
  __shared__ double s_err[8];
  if (threadIdx.x < kMaxK + 2) s_counts[threadIdx.x] = 0;
  __syncthreads();
  unsigned int cnt[K];
#pragma unroll
  for (int i = 0; i < K; ++i) cnt[i] = 0;
  unsigned int any_jl = 0, any_ob = 0;
  double err = 0.0;
  const HostWeights w(p.u, nullptr);
  for (long long b = (long long)blockIdx.x * 256 + threadIdx.x; b < p.batch; b += (long long)gridDim.x * 256) {
    float th[3], ob[NOB][3], acc[K], dth[3], ee[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) th[j] = p.theta[b * p.th_sb + j * p.th_sj];
    const float tx = p.target[b * p.tg_sb], ty = p.target[b * p.tg_sb + p.tg_sc];
#pragma unroll
    for (int k = 0; k < NOB; ++k) {
#pragma unroll
      for (int c = 0; c < 3; ++c) ob[k][c] = DYN ? p.obst[b * p.ob_sb + k * p.ob_so + c * p.ob_sc] : p.u.stat_ob[k][c];
    }
#pragma unroll
    for (int i = 0; i < K; ++i) acc[i] = 0.f;
    eval_sample<3, NOBS, JL, false, false>(p.u, w, th, tx, ty, ob, acc, dth, ee);
    bool jl = false, obv = false;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      cnt[i] += acc[i] > p.threshold;
      if (JL && i >= 1 && i < 4) jl |= acc[i] > 0.f;
      if (i >= (JL ? 4 : 1)) obv |= acc[i] > 0.f;
    }
    any_jl += jl;
    any_ob += obv;
    err += sqrt((double)acc[0]);     // sqrt(position_err_sq.sum()) with position_err_sq = 0.5*e^2: |e|/sqrt(2), the reference's quirk
  }
#pragma unroll
  for (int i = 0; i < K; ++i) {
    unsigned v = cnt[i];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_counts[i], v);
  }
  for (int off = 16; off > 0; off >>= 1) {
    any_jl += __shfl_xor_sync(0xffffffffu, any_jl, off);
    any_ob += __shfl_xor_sync(0xffffffffu, any_ob, off);
    err += __shfl_xor_sync(0xffffffffu, err, off);
  }
  if ((threadIdx.x & 31) == 0) {
    if (any_jl) atomicAdd(&s_counts[kMaxK], any_jl);
    if (any_ob) atomicAdd(&s_counts[kMaxK + 1], any_ob);
    s_err[threadIdx.x >> 5] = err;
  }
  __syncthreads();
  if (threadIdx.x < kMaxK + 2 && s_counts[threadIdx.x]) atomicAdd(p.counts + threadIdx.x, (unsigned long long)s_counts[threadIdx.x]);
  if (threadIdx.x == 0) {
    double e = 0.0;
    for (int wq = 0; wq < 8; ++wq) e += s_err[wq];
    atomicAdd(p.pos_err_sum, e);
  }
}

}  // namespace
}  // namespace ikl

using namespace ikl;

extern "C" {

IklStatus ikl_dataset_generate(const IklSpec* spec, int64_t n, uint64_t seed, float dynamic_radius, float* theta_out, float* ee_out,
                               float* obstacles_out, uint32_t* exhausted_out, void* stream) {
  if (!spec || n < 0) return IKL_ERR_INVALID_ARGUMENT;
  if (spec->num_links != 3) return IKL_ERR_UNSUPPORTED;
  if (n == 0) return IKL_OK;
  if (!theta_out || !ee_out || !exhausted_out) return IKL_ERR_INVALID_ARGUMENT;
  const bool dyn = spec->dynamic_obstacles != 0;
  if (dyn && !obstacles_out) return IKL_ERR_INVALID_ARGUMENT;
  DatasetParams p;
  p.n = n;
  p.seed = seed;
  p.lo = spec->limit_lo;
  p.hi = spec->limit_hi;
  p.n_dynamic = dyn ? spec->num_obstacles : 0;
  p.dyn_radius = dynamic_radius;
  p.n_static = dyn ? 0 : spec->num_obstacles;
  for (int o = 0; o < kMaxObst; ++o)
    for (int c = 0; c < 3; ++c) p.stat[o][c] = spec->static_obstacles[o][c];
  p.theta = theta_out;
  p.ee = ee_out;
  p.obstacles = dyn ? obstacles_out : nullptr;
  p.exhausted = exhausted_out;
  long long grid = (n + kDsThreads - 1) / kDsThreads;
  const long long cap = (long long)sm_count() * 16;
  if (grid > cap) grid = cap;
  count_launch();
  dataset_kernel<<<(unsigned)grid, kDsThreads, 0, (cudaStream_t)stream>>>(p);
  return cuda_status(cudaGetLastError());
}

IklStatus ikl_eval_stats(const IklSpec* spec, const IklBatch* in, float threshold, uint64_t* counts_out, double* pos_err_sum_out,
                         void* stream) {
  if (!spec || !in || !counts_out || !pos_err_sum_out || in->batch < 0) return IKL_ERR_INVALID_ARGUMENT;
  if (spec->num_links != 3 || spec->num_obstacles > IKL_MAX_OBSTACLES) return IKL_ERR_UNSUPPORTED;
  if (in->batch == 0) return IKL_OK;
  const bool dyn = spec->num_obstacles > 0 && spec->dynamic_obstacles;
  if (!in->theta || !in->target || (dyn && !in->obstacles)) return IKL_ERR_INVALID_ARGUMENT;
  {
    // the reference's own tensors or planes: same staged pipelines as the Lagrangian kernels
    LagrangianParams lp;
    Variant v;
    const int layout = prepare_lagrangian_params(spec, in, nullptr, false, nullptr, 0, 0, lp, v);
    if (layout < 0) return (IklStatus)(-layout);
    if ((layout == kRows || layout == kPlanes) && lp.use_tma && v.links == 3 && v.unit) {
      lp.stats_threshold = threshold;
      lp.stats_counts = reinterpret_cast<unsigned long long*>(counts_out);
      lp.stats_pos_err = pos_err_sum_out;
      bool found = false;
      const cudaError_t e = layout == kRows ? launch_stats_family<kRows>(v, lp, (cudaStream_t)stream, 2048, &found)
                                            : launch_stats_family<kPlanes>(v, lp, (cudaStream_t)stream, 2048, &found);
      if (found) {
        set_last_kernel_name(layout == kRows ? "ikl_eval_stats<rows_tma>" : "ikl_eval_stats<planes_tma>");
        return cuda_status(e);
      }
    }
  }
  set_last_kernel_name("ikl_eval_stats<strided>");
  StatsParams p;
  memset(&p, 0, sizeof(p));
  for (int m = 0; m < kMaxLinks; ++m) p.u.link[m] = spec->link_length[m];
  const float lo = spec->limit_lo, hi = spec->limit_hi;
  p.u.jl_mid = (lo + hi) / 2.0f;
  const float rng = fabsf(hi - lo);
  p.u.jl_in = spec->good_slope;
  p.u.jl_out = spec->bad_slope;
  p.u.jl_cin = (p.u.jl_in * 0.5f) * rng;
  p.u.jl_cout = (p.u.jl_out * 0.5f) * rng;
  p.u.ob_hi = fmaxf(spec->good_slope, spec->bad_slope);
  p.u.ob_lo = fminf(spec->good_slope, spec->bad_slope);
  for (int o = 0; o < kMaxObst; ++o)
    for (int c = 0; c < 3; ++c) p.u.stat_ob[o][c] = spec->static_obstacles[o][c];
  p.batch = in->batch;
  p.theta = in->theta;   p.th_sb = in->theta_sb;  p.th_sj = in->theta_sj;
  p.target = in->target; p.tg_sb = in->target_sb; p.tg_sc = in->target_sc;
  p.obst = in->obstacles; p.ob_sb = in->obst_sb;  p.ob_so = in->obst_so;  p.ob_sc = in->obst_sc;
  p.threshold = threshold;
  p.counts = reinterpret_cast<unsigned long long*>(counts_out);
  p.pos_err_sum = pos_err_sum_out;
  long long grid = (in->batch + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (grid > cap) grid = cap;
  cudaStream_t st = (cudaStream_t)stream;
  const bool jl = spec->enforce_joint_limits != 0;
  const int nobs = spec->num_obstacles;
  count_launch();
#define IKL_STATS_CASE(NOBS, DYN, JL) \
  if (nobs == NOBS && dyn == DYN && jl == JL) { stats_kernel<NOBS, DYN, JL><<<(unsigned)grid, 256, 0, st>>>(p); return cuda_status(cudaGetLastError()); }
  IKL_STATS_CASE(0, false, false)
  IKL_STATS_CASE(0, false, true)
  IKL_STATS_CASE(1, false, true)
  IKL_STATS_CASE(2, false, true)
  IKL_STATS_CASE(3, false, true)
  IKL_STATS_CASE(4, false, true)
  IKL_STATS_CASE(1, true, true)
  IKL_STATS_CASE(2, true, true)
  IKL_STATS_CASE(3, true, true)
  IKL_STATS_CASE(4, true, true)
  IKL_STATS_CASE(1, false, false)
  IKL_STATS_CASE(2, false, false)
  IKL_STATS_CASE(3, false, false)
  IKL_STATS_CASE(4, false, false)
  IKL_STATS_CASE(1, true, false)
  IKL_STATS_CASE(2, true, false)
  IKL_STATS_CASE(3, true, false)
  IKL_STATS_CASE(4, true, false)
#undef IKL_STATS_CASE
  return IKL_ERR_UNSUPPORTED;
}

}  // extern "C"
