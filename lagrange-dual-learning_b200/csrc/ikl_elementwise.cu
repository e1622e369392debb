// ikl_elementwise.cu -- function-level drop-in kernels: forward kinematics, the two penalties
// (forward + backward each) and the dual-ascent update.  These let an unchanged scripts/trainer.py
// reach CUDA op by op; the fused kernels in ikl_lagrangian.cuh are the fast path.
//
// The penalty kernels reproduce the reference's fp32 roundings exactly (explicit _rn intrinsics: no FMA
// contraction, IEEE sqrt/div), so their forward values are bit-identical to torch's CPU results.
#include <math.h>

#include "ikl_common.h"

namespace ikl {
namespace {

constexpr int kEwThreads = 256;

inline unsigned ew_grid(long long n) {
  long long g = (n + kEwThreads - 1) / kEwThreads;
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

constexpr int kFkMaxLinks = 8;          // the function-level FK also serves the reference's (broken) eight-link arm
struct Lengths { float v[kFkMaxLinks]; };

// Elements per thread per grid-stride step.  The penalty kernels read a few 4-byte elements per sample from strided column
// views; with one sample per thread an SM has too few bytes in flight to cover the HBM latency (jl_forward: 41 % of the
// measured peak on physical bytes), so every thread issues the loads of kEwUnroll samples before touching any of them.
constexpr int kEwUnroll = 4;

inline unsigned ew_grid_unrolled(long long n) { return ew_grid((n + kEwUnroll - 1) / kEwUnroll); }

// kinematics/forward_kinematics.py:64-81 -- tips_out block k = tip J-k (end effector first), rows (x, y, angle)
template <int J>
__device__ __forceinline__ void fk_sample(const float (&th)[J], const Lengths& L, float (&out)[J][3]) {
  float phi = 0.f, x = 0.f, y = 0.f;
#pragma unroll
  for (int m = 0; m < J; ++m) {
    phi = (m == 0) ? th[0] : __fadd_rn(phi, th[m]);
    float s, c;
    sincosf(phi, &s, &c);
    x = (m == 0) ? __fmul_rn(L.v[0], c) : __fadd_rn(x, __fmul_rn(L.v[m], c));
    y = (m == 0) ? __fmul_rn(L.v[0], s) : __fadd_rn(y, __fmul_rn(L.v[m], s));
    out[J - 1 - m][0] = x;
    out[J - 1 - m][1] = y;
    out[J - 1 - m][2] = phi;
  }
}

template <int J>
__global__ void __launch_bounds__(kEwThreads) fk_forward_kernel(const float* __restrict__ theta, long long sb, long long sj, long long B,
                                                               Lengths L, float* __restrict__ tips) {
  for (long long b = (long long)blockIdx.x * kEwThreads + threadIdx.x; b < B; b += (long long)gridDim.x * kEwThreads) {
    float th[J], out[J][3];
#pragma unroll
    for (int m = 0; m < J; ++m) th[m] = theta[b * sb + m * sj];
    fk_sample<J>(th, L, out);
#pragma unroll
    for (int k = 0; k < J; ++k) {
      float* o = tips + ((long long)k * B + b) * 3;
      o[0] = out[k][0]; o[1] = out[k][1]; o[2] = out[k][2];
    }
  }
}

// d(theta) from the gradients of all J tips (x, y and the angle column).
template <int J>
__global__ void __launch_bounds__(kEwThreads) fk_backward_kernel(const float* __restrict__ theta, long long sb, long long sj, long long B,
                                                                Lengths L, const float* __restrict__ gtips, float* __restrict__ dtheta) {
  for (long long b = (long long)blockIdx.x * kEwThreads + threadIdx.x; b < B; b += (long long)gridDim.x * kEwThreads) {
    float s[J], c[J];
    float phi = 0.f;
#pragma unroll
    for (int m = 0; m < J; ++m) {
      const float t = theta[b * sb + m * sj];
      phi = (m == 0) ? t : __fadd_rn(phi, t);
      sincosf(phi, &s[m], &c[m]);
    }
    float sx = 0.f, sy = 0.f, run = 0.f;
#pragma unroll
    for (int m = J - 1; m >= 0; --m) {
      const float* g = gtips + ((long long)(J - 1 - m) * B + b) * 3;
      sx += g[0];
      sy += g[1];
      run += L.v[m] * (c[m] * sy - s[m] * sx) + g[2];
      dtheta[b * J + m] = run;
    }
  }
}

__device__ __forceinline__ float torch_max(float a, float b) { return (a != a || b != b) ? NAN : fmaxf(a, b); }

// kinematics/penalties.py:12-17
__global__ void __launch_bounds__(kEwThreads) circle_forward_kernel(const float* __restrict__ jx, long long s_jx, const float* __restrict__ jy,
                                                                   long long s_jy, const float* __restrict__ ox, long long s_ox,
                                                                   const float* __restrict__ oy, long long s_oy,
                                                                   const float* __restrict__ rad, long long s_r, long long n, float a_in,
                                                                   float a_out, float* __restrict__ out) {
  const long long step = (long long)gridDim.x * kEwThreads;
  for (long long i0 = (long long)blockIdx.x * kEwThreads + threadIdx.x; i0 < n; i0 += step * kEwUnroll) {
    float a[kEwUnroll], b[kEwUnroll], c[kEwUnroll], e[kEwUnroll], r[kEwUnroll];
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long i = i0 + u * step;
      if (i < n) { a[u] = jx[i * s_jx]; b[u] = jy[i * s_jy]; c[u] = ox[i * s_ox]; e[u] = oy[i * s_oy]; r[u] = rad[i * s_r]; }
    }
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long i = i0 + u * step;
      if (i < n) {
        const float dx = __fsub_rn(a[u], c[u]);
        const float dy = __fsub_rn(b[u], e[u]);
        const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        const float t = __fsub_rn(d, r[u]);
        out[i] = torch_max(__fmul_rn(-a_in, t), __fmul_rn(-a_out, t));
      }
    }
  }
}

__global__ void __launch_bounds__(kEwThreads) circle_backward_kernel(const float* __restrict__ jx, long long s_jx, const float* __restrict__ jy,
                                                                    long long s_jy, const float* __restrict__ ox, long long s_ox,
                                                                    const float* __restrict__ oy, long long s_oy,
                                                                    const float* __restrict__ rad, long long s_r, long long n, float a_in,
                                                                    float a_out, const float* __restrict__ gout, long long s_g, float* __restrict__ d_jx,
                                                                    float* __restrict__ d_jy, float* __restrict__ d_ox,
                                                                    float* __restrict__ d_oy, float* __restrict__ d_r) {
  const long long step = (long long)gridDim.x * kEwThreads;
  for (long long i0 = (long long)blockIdx.x * kEwThreads + threadIdx.x; i0 < n; i0 += step * kEwUnroll) {
    float a[kEwUnroll], b[kEwUnroll], c[kEwUnroll], e[kEwUnroll], r[kEwUnroll], go[kEwUnroll];
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long i = i0 + u * step;
      if (i < n) { a[u] = jx[i * s_jx]; b[u] = jy[i * s_jy]; c[u] = ox[i * s_ox]; e[u] = oy[i * s_oy]; r[u] = rad[i * s_r]; go[u] = gout[i * s_g]; }
    }
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long i = i0 + u * step;
      if (i >= n) continue;
      const float dx = __fsub_rn(a[u], c[u]);
      const float dy = __fsub_rn(b[u], e[u]);
      const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      const float t = __fsub_rn(d, r[u]);
      const float vi = __fmul_rn(-a_in, t), vo = __fmul_rn(-a_out, t);
      // torch MaximumBackward: the larger branch gets the gradient, a tie splits it evenly
      const float slope = (vi > vo) ? -a_in : ((vi < vo) ? -a_out : -0.5f * (a_in + a_out));
      const float gt = go[u] * slope;                       // d/dt
      // d(t)/d(jx) = dx / d ; at the centre (d == 0) torch gives NaN, this library gives 0
      const float ux = d > 0.f ? __fdiv_rn(dx, d) : 0.f;
      const float uy = d > 0.f ? __fdiv_rn(dy, d) : 0.f;
      if (d_jx) d_jx[i] = gt * ux;
      if (d_jy) d_jy[i] = gt * uy;
      if (d_ox) d_ox[i] = -gt * ux;
      if (d_oy) d_oy[i] = -gt * uy;
      if (d_r) d_r[i] = -gt;
    }
  }
}

// kinematics/penalties.py:51-56
__global__ void __launch_bounds__(kEwThreads) jl_forward_kernel(const float* __restrict__ th, long long s_t, const float* __restrict__ lo,
                                                               long long s_lo, const float* __restrict__ hi, long long s_hi, long long n,
                                                               float a_in, float a_out, float* __restrict__ out) {
  const float h_in = a_in * 0.5f, h_out = a_out * 0.5f;     // exact: the reference folds slope*0.5 before touching the tensor
  const long long step = (long long)gridDim.x * kEwThreads;
  for (long long i0 = (long long)blockIdx.x * kEwThreads + threadIdx.x; i0 < n; i0 += step * kEwUnroll) {
    float l[kEwUnroll], h[kEwUnroll], t[kEwUnroll];
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long i = i0 + u * step;
      if (i < n) { l[u] = lo[i * s_lo]; h[u] = hi[i * s_hi]; t[u] = th[i * s_t]; }
    }
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long i = i0 + u * step;
      if (i >= n) continue;
      const float mid = __fdiv_rn(__fadd_rn(l[u], h[u]), 2.0f);
      const float rng = fabsf(__fsub_rn(h[u], l[u]));
      const float dev = fabsf(__fsub_rn(t[u], mid));
      const float vin = __fsub_rn(__fmul_rn(a_in, dev), __fmul_rn(h_in, rng));
      const float vout = __fsub_rn(__fmul_rn(a_out, dev), __fmul_rn(h_out, rng));
      out[i] = torch_max(vin, vout);
    }
  }
}

__device__ __forceinline__ float sgn(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

__global__ void __launch_bounds__(kEwThreads) jl_backward_kernel(const float* __restrict__ th, long long s_t, const float* __restrict__ lo,
                                                                long long s_lo, const float* __restrict__ hi, long long s_hi, long long n,
                                                                float a_in, float a_out, const float* __restrict__ gout, long long s_g,
                                                                float* __restrict__ d_th, float* __restrict__ d_lo, float* __restrict__ d_hi) {
  const float h_in = a_in * 0.5f, h_out = a_out * 0.5f;
  const long long step = (long long)gridDim.x * kEwThreads;
  for (long long i0 = (long long)blockIdx.x * kEwThreads + threadIdx.x; i0 < n; i0 += step * kEwUnroll) {
    float l[kEwUnroll], h[kEwUnroll], t[kEwUnroll], go[kEwUnroll];
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long i = i0 + u * step;
      if (i < n) { l[u] = lo[i * s_lo]; h[u] = hi[i * s_hi]; t[u] = th[i * s_t]; go[u] = gout[i * s_g]; }
    }
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long i = i0 + u * step;
      if (i >= n) continue;
      const float mid = __fdiv_rn(__fadd_rn(l[u], h[u]), 2.0f);
      const float span = __fsub_rn(h[u], l[u]);
      const float uu = __fsub_rn(t[u], mid);
      const float dev = fabsf(uu), rng = fabsf(span);
      const float vin = __fsub_rn(__fmul_rn(a_in, dev), __fmul_rn(h_in, rng));
      const float vout = __fsub_rn(__fmul_rn(a_out, dev), __fmul_rn(h_out, rng));
      const float slope = (vin > vout) ? a_in : ((vin < vout) ? a_out : 0.5f * (a_in + a_out));
      const float g = go[u] * slope;
      const float gu = g * sgn(uu);                // d/dtheta ; d/dmid = -gu
      const float gr = -0.5f * g * sgn(span);      // through -slope*0.5*|hi-lo| : d/dhi = gr, d/dlo = -gr
      if (d_th) d_th[i] = gu;
      if (d_lo) d_lo[i] = -0.5f * gu - gr;
      if (d_hi) d_hi[i] = -0.5f * gu + gr;
    }
  }
}

// scripts/trainer.py:280-284
__global__ void dual_update_kernel(const float* __restrict__ lam_in, const double* __restrict__ viol64, const float* __restrict__ viol32,
                                   float lr, int K, float* __restrict__ lam_out) {
  const int i = threadIdx.x;
  if (i < K) {
    // the reference sums the per-batch K-vectors in fp32 and does lam + lr*sum in fp32
    const float v = viol64 ? (float)viol64[i] : viol32[i];
    const float next = __fadd_rn(lam_in[i], __fmul_rn(lr, v));
    lam_out[i] = next < 0.f ? 0.f : next;        // clamp(min=0); NaN propagates like torch.clamp
  }
}

}  // namespace
}  // namespace ikl

using namespace ikl;

extern "C" {

IklStatus ikl_fk_forward(const float* theta, int64_t theta_sb, int64_t theta_sj, int64_t batch, int32_t num_links,
                         const float* link_length_host, float* tips_out, void* stream) {
  if (batch < 0 || num_links < 1) return IKL_ERR_INVALID_ARGUMENT;
  if (num_links != 2 && num_links != 3 && num_links != 8) return IKL_ERR_UNSUPPORTED;
  if (batch == 0) return IKL_OK;
  if (!theta || !tips_out) return IKL_ERR_INVALID_ARGUMENT;
  Lengths L;
  for (int m = 0; m < kFkMaxLinks; ++m) L.v[m] = (link_length_host && m < num_links) ? link_length_host[m] : 1.0f;
  cudaStream_t st = (cudaStream_t)stream;
  count_launch();
  if (num_links == 3) fk_forward_kernel<3><<<ew_grid(batch), kEwThreads, 0, st>>>(theta, theta_sb, theta_sj, batch, L, tips_out);
  else if (num_links == 2) fk_forward_kernel<2><<<ew_grid(batch), kEwThreads, 0, st>>>(theta, theta_sb, theta_sj, batch, L, tips_out);
  else fk_forward_kernel<8><<<ew_grid(batch), kEwThreads, 0, st>>>(theta, theta_sb, theta_sj, batch, L, tips_out);
  return cuda_status(cudaGetLastError());
}

IklStatus ikl_fk_backward(const float* theta, int64_t theta_sb, int64_t theta_sj, int64_t batch, int32_t num_links,
                          const float* link_length_host, const float* grad_tips, float* dtheta, void* stream) {
  if (batch < 0 || num_links < 1) return IKL_ERR_INVALID_ARGUMENT;
  if (num_links != 2 && num_links != 3 && num_links != 8) return IKL_ERR_UNSUPPORTED;
  if (batch == 0) return IKL_OK;
  if (!theta || !grad_tips || !dtheta) return IKL_ERR_INVALID_ARGUMENT;
  Lengths L;
  for (int m = 0; m < kFkMaxLinks; ++m) L.v[m] = (link_length_host && m < num_links) ? link_length_host[m] : 1.0f;
  cudaStream_t st = (cudaStream_t)stream;
  count_launch();
  if (num_links == 3) fk_backward_kernel<3><<<ew_grid(batch), kEwThreads, 0, st>>>(theta, theta_sb, theta_sj, batch, L, grad_tips, dtheta);
  else if (num_links == 2) fk_backward_kernel<2><<<ew_grid(batch), kEwThreads, 0, st>>>(theta, theta_sb, theta_sj, batch, L, grad_tips, dtheta);
  else fk_backward_kernel<8><<<ew_grid(batch), kEwThreads, 0, st>>>(theta, theta_sb, theta_sj, batch, L, grad_tips, dtheta);
  return cuda_status(cudaGetLastError());
}

IklStatus ikl_circle_penalty_forward(const float* jx, int64_t s_jx, const float* jy, int64_t s_jy, const float* ox, int64_t s_ox,
                                     const float* oy, int64_t s_oy, const float* radius, int64_t s_r, int64_t n, float inside_slope,
                                     float outside_slope, float* out, void* stream) {
  if (n < 0 || !(inside_slope >= 0.f) || !(outside_slope >= 0.f)) return IKL_ERR_INVALID_ARGUMENT;   // penalties.py:10
  if (n == 0) return IKL_OK;
  if (!jx || !jy || !ox || !oy || !radius || !out) return IKL_ERR_INVALID_ARGUMENT;
  count_launch();
  circle_forward_kernel<<<ew_grid_unrolled(n), kEwThreads, 0, (cudaStream_t)stream>>>(jx, s_jx, jy, s_jy, ox, s_ox, oy, s_oy, radius, s_r, n,
                                                                              inside_slope, outside_slope, out);
  return cuda_status(cudaGetLastError());
}

IklStatus ikl_circle_penalty_backward(const float* jx, int64_t s_jx, const float* jy, int64_t s_jy, const float* ox, int64_t s_ox,
                                      const float* oy, int64_t s_oy, const float* radius, int64_t s_r, int64_t n, float inside_slope,
                                      float outside_slope, const float* grad_out, int64_t s_g, float* d_jx, float* d_jy, float* d_ox,
                                      float* d_oy, float* d_radius, void* stream) {
  if (n < 0 || !(inside_slope >= 0.f) || !(outside_slope >= 0.f)) return IKL_ERR_INVALID_ARGUMENT;
  if (n == 0) return IKL_OK;
  if (!jx || !jy || !ox || !oy || !radius || !grad_out) return IKL_ERR_INVALID_ARGUMENT;
  count_launch();
  circle_backward_kernel<<<ew_grid_unrolled(n), kEwThreads, 0, (cudaStream_t)stream>>>(jx, s_jx, jy, s_jy, ox, s_ox, oy, s_oy, radius, s_r, n,
                                                                               inside_slope, outside_slope, grad_out, s_g, d_jx, d_jy, d_ox,
                                                                               d_oy, d_radius);
  return cuda_status(cudaGetLastError());
}

IklStatus ikl_joint_limit_penalty_forward(const float* theta, int64_t s_t, const float* lo, int64_t s_lo, const float* hi, int64_t s_hi,
                                          int64_t n, float inside_slope, float outside_slope, float* out, void* stream) {
  if (n < 0 || !(inside_slope >= 0.f) || !(outside_slope >= 0.f)) return IKL_ERR_INVALID_ARGUMENT;   // penalties.py:49
  if (n == 0) return IKL_OK;
  if (!theta || !lo || !hi || !out) return IKL_ERR_INVALID_ARGUMENT;
  count_launch();
  jl_forward_kernel<<<ew_grid_unrolled(n), kEwThreads, 0, (cudaStream_t)stream>>>(theta, s_t, lo, s_lo, hi, s_hi, n, inside_slope, outside_slope, out);
  return cuda_status(cudaGetLastError());
}

IklStatus ikl_joint_limit_penalty_backward(const float* theta, int64_t s_t, const float* lo, int64_t s_lo, const float* hi, int64_t s_hi,
                                           int64_t n, float inside_slope, float outside_slope, const float* grad_out, int64_t s_g,
                                           float* d_theta, float* d_lo, float* d_hi, void* stream) {
  if (n < 0 || !(inside_slope >= 0.f) || !(outside_slope >= 0.f)) return IKL_ERR_INVALID_ARGUMENT;
  if (n == 0) return IKL_OK;
  if (!theta || !lo || !hi || !grad_out) return IKL_ERR_INVALID_ARGUMENT;
  count_launch();
  jl_backward_kernel<<<ew_grid_unrolled(n), kEwThreads, 0, (cudaStream_t)stream>>>(theta, s_t, lo, s_lo, hi, s_hi, n, inside_slope, outside_slope,
                                                                           grad_out, s_g, d_theta, d_lo, d_hi);
  return cuda_status(cudaGetLastError());
}

IklStatus ikl_dual_update(const float* lam_in, const double* viol_sum, const float* viol_sum_f32, float multiplier_lr,
                          int32_t num_constraints, float* lam_out, void* stream) {
  if (num_constraints < 1 || num_constraints > IKL_MAX_CONSTRAINTS) return IKL_ERR_INVALID_ARGUMENT;
  if (!lam_in || !lam_out || (!viol_sum && !viol_sum_f32)) return IKL_ERR_INVALID_ARGUMENT;
  count_launch();
  dual_update_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(lam_in, viol_sum, viol_sum_f32, multiplier_lr, num_constraints, lam_out);
  return cuda_status(cudaGetLastError());
}

}  // extern "C"
