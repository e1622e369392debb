// ikl_api.cu -- C ABI of the fused Lagrangian path (include/ikl.h): argument validation, layout
// detection, constant preparation and dispatch to the kernel families.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include <cuda.h>          // CUtensorMap types only; the encoder is looked up at run time (no link against libcuda)

#include "ikl_lagrangian.cuh"

namespace ikl {

namespace {
thread_local cudaError_t g_last_cuda_error = cudaSuccess;
thread_local const char* g_last_kernel = "";
std::atomic<long long> g_launches{0};
std::atomic<int> g_reserved_sms{0};

constexpr int kMaxGrid = 2048;                                       // upper bound on persistent grid size
constexpr size_t kPartialsBytes = (size_t)kMaxGrid * kMaxK * sizeof(double);
constexpr size_t kWorkspaceBytes = kPartialsBytes + 256;             // + the completion counter
}  // namespace

void set_last_cuda_error(cudaError_t e) { g_last_cuda_error = e; }
void set_last_kernel_name(const char* name) { g_last_kernel = name; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int reserved_sms() { return g_reserved_sms.load(std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

namespace {

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

int forced_layout() {
  const char* e = getenv("IKL_FORCE_LAYOUT");
  if (!e || !*e) return -1;
  if (!strcmp(e, "strided")) return kStrided;
  if (!strcmp(e, "rows")) return kRows;
  if (!strcmp(e, "planes")) return kPlanes;
  return -1;
}

IklStatus validate(const IklSpec* spec, const IklBatch* in, const IklWeights* w, void* workspace, const float* viol_out,
                   const float* loss_out) {
  if (!spec || !in || !w || !workspace || !viol_out || !loss_out) return IKL_ERR_INVALID_ARGUMENT;
  if (spec->num_links < 1 || spec->num_obstacles < 0 || in->batch < 0) return IKL_ERR_INVALID_ARGUMENT;
  if (spec->num_links != 2 && spec->num_links != 3) return IKL_ERR_UNSUPPORTED;
  if (spec->num_obstacles > IKL_MAX_OBSTACLES) return IKL_ERR_UNSUPPORTED;
  if (!(spec->good_slope >= 0.f) || !(spec->bad_slope >= 0.f)) return IKL_ERR_INVALID_ARGUMENT;   // penalties.py:10,49
  if ((w->host == nullptr) == (w->device == nullptr)) return IKL_ERR_INVALID_ARGUMENT;
  if (in->batch > 0) {
    if (!in->theta || !in->target) return IKL_ERR_INVALID_ARGUMENT;
    if (spec->num_obstacles > 0 && spec->dynamic_obstacles && !in->obstacles) return IKL_ERR_INVALID_ARGUMENT;
  }
  return IKL_OK;
}

void fill_uniforms(const IklSpec* spec, const float* w_host, Uniforms& u) {
  memset(&u, 0, sizeof(u));
  const int J = spec->num_links;
  for (int m = 0; m < kMaxLinks; ++m) u.link[m] = m < J ? spec->link_length[m] : 1.0f;
  const float lo = spec->limit_lo, hi = spec->limit_hi;
  // penalties.py:51-52 in fp32, the way torch evaluates them on float32 tensors
  u.jl_mid = (lo + hi) / 2.0f;
  const float rng = fabsf(hi - lo);
  u.jl_in = spec->good_slope;          // trainer.py:203-204: inside = good, outside = bad
  u.jl_out = spec->bad_slope;
  u.jl_cin = (u.jl_in * 0.5f) * rng;   // penalties.py:53: inside_slope*0.5 is folded first, then * limit_range
  u.jl_cout = (u.jl_out * 0.5f) * rng;
  u.ob_hi = fmaxf(spec->good_slope, spec->bad_slope);
  u.ob_lo = fminf(spec->good_slope, spec->bad_slope);
  for (int o = 0; o < kMaxObst; ++o)
    for (int c = 0; c < 3; ++c) u.stat_ob[o][c] = spec->static_obstacles[o][c];
  if (w_host) {
    const int K = ikl_num_constraints(spec);
    for (int i = 0; i < K; ++i) u.w[i] = w_host[i];
    int idx = 1;
    if (spec->enforce_joint_limits) {
      for (int j = 0; j < J; ++j, ++idx) {
        u.w_jl_in[j] = u.w[idx] * u.jl_in;
        u.w_jl_out[j] = u.w[idx] * u.jl_out;
        u.w_jl_tie[j] = u.w[idx] * (0.5f * (u.jl_in + u.jl_out));
      }
    }
    for (int q = 0; q < J * spec->num_obstacles; ++q, ++idx) {
      u.w_ob_hi[q] = -u.w[idx] * u.ob_hi;
      u.w_ob_lo[q] = -u.w[idx] * u.ob_lo;
      u.w_ob_tie[q] = -u.w[idx] * (0.5f * (u.ob_hi + u.ob_lo));
    }
  }
}

void fill_packed_uniforms(const IklSpec* spec, const Uniforms& u, PackedUniforms& pu) {
  memset(&pu, 0, sizeof(pu));
  auto dup = [](float v) { return make_float2(v, v); };
  const float half = 0.5f * fabsf(spec->limit_hi - spec->limit_lo);
  const float smin = fminf(u.jl_in, u.jl_out), smax = fmaxf(u.jl_in, u.jl_out);
  pu.jl_mid = dup(u.jl_mid);
  pu.jl_half = dup(half);
  pu.jl_smin = dup(smin);
  pu.jl_sdiff = dup(smax - smin);
  pu.ob_nlo = dup(-u.ob_lo);
  pu.ob_ndiff = dup(u.ob_lo - u.ob_hi);
  pu.ob_nhi_sel = dup(fmaf(1.0f, u.ob_lo - u.ob_hi, -u.ob_lo));
  pu.jl_shi_sel = dup(fmaf(1.0f, smax - smin, smin));
  for (int o = 0; o < kMaxObst; ++o)
    for (int c = 0; c < 3; ++c) pu.stat_ob[o][c] = dup(u.stat_ob[o][c]);
  for (int o = 0; o < kMaxObst; ++o) pu.stat_thr[o] = dup(u.stat_ob[o][2] * kNearTie);      // fl(r * 2^-20): exact scaling
  for (int i = 0; i < kMaxK; ++i) pu.lam[i] = dup(u.w[i]);
  for (int i = 0; i < kMaxK; ++i) {
    pu.ob_nlo_lam[i] = dup(pu.ob_nlo.x * u.w[i]);          // fp32 products, the roundings the kernels' own FMUL2 would make
    pu.ob_nhi_lam[i] = dup(pu.ob_nhi_sel.x * u.w[i]);
  }
  for (int m = 0; m < kMaxLinks; ++m) pu.link[m] = dup(u.link[m]);
}


// Static obstacles as pairs (2q, 2q+1) for the single-sample kernels (eval_single_static); an odd count repeats the last
// obstacle in the padding lane with weight 0.
void fill_static_pairs(const IklSpec* spec, const Uniforms& u, StaticPairUniforms& sp) {
  memset(&sp, 0, sizeof(sp));
  const int n = spec->num_obstacles, J = 3;
  const int base = 1 + (spec->enforce_joint_limits ? J : 0);
  for (int q = 0; q < kMaxObstPairs; ++q) {
    const int o0 = 2 * q < n ? 2 * q : (n > 0 ? n - 1 : 0);
    const int o1 = 2 * q + 1 < n ? 2 * q + 1 : o0;
    sp.ox[q] = make_float2(u.stat_ob[o0][0], u.stat_ob[o1][0]);
    sp.oy[q] = make_float2(u.stat_ob[o0][1], u.stat_ob[o1][1]);
    sp.nr[q] = make_float2(-u.stat_ob[o0][2], -u.stat_ob[o1][2]);
    sp.thr[q] = make_float2(u.stat_ob[o0][2] * kNearTie, u.stat_ob[o1][2] * kNearTie);
    for (int m = 0; m < J; ++m) {
      const int j = J - 1 - m;                       // tip m is FK tuple index j (end effector first)
      const float w0 = 2 * q < n ? u.w[base + 3 * (2 * q) + j] : 0.f;
      const float w1 = 2 * q + 1 < n ? u.w[base + 3 * (2 * q + 1) + j] : 0.f;
      sp.lam[m][q] = make_float2(w0, w1);
    }
  }
}

// ---- TMA descriptors ---------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled lives in the driver; libikl.so links cudart statically and asks it for the entry point.
#ifndef IKL_TMA_L2_PROMOTION
#define IKL_TMA_L2_PROMOTION CU_TENSOR_MAP_L2_PROMOTION_L2_256B
#endif
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// A group of `n1 x n2` fp32 planes of `batch` samples each as a rank-2 (n2 == 0) or rank-3 tensor, sample index innermost;
// box = 256 samples x the whole group.  Strides are in elements.  Returns false when the hardware cannot describe it.
bool encode_planes(TensorMap* out, const float* base, int64_t batch, int n1, int64_t stride1, int n2, int64_t stride2) {
  static_assert(sizeof(TensorMap) == sizeof(CUtensorMap) && alignof(TensorMap) >= alignof(CUtensorMap), "TensorMap must mirror CUtensorMap");
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  if (stride1 <= 0 || (stride1 * 4) % 16 != 0 || (n2 > 0 && (stride2 <= 0 || (stride2 * 4) % 16 != 0))) return false;
  const cuuint32_t rank = n2 > 0 ? 3 : 2;
  const cuuint64_t dims[3] = {(cuuint64_t)batch, (cuuint64_t)n1, (cuuint64_t)(n2 > 0 ? n2 : 1)};
  const cuuint64_t strides[2] = {(cuuint64_t)stride1 * 4, (cuuint64_t)stride2 * 4};
  const cuuint32_t box[3] = {(cuuint32_t)kTmaBox, (cuuint32_t)n1, (cuuint32_t)(n2 > 0 ? n2 : 1)};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<float*>(base), dims, strides,
                         box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, IKL_TMA_L2_PROMOTION,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool unit_links(const IklSpec* spec) {
  for (int m = 0; m < spec->num_links; ++m)
    if (spec->link_length[m] != 1.0f) return false;
  return true;
}

// Which vectorised layout (if any) the strides describe.  batch == 1 always takes the strided kernel.
int detect_layout(const IklSpec* spec, const IklBatch* in, bool grad, const float* dtheta, int64_t dt_sb, int64_t dt_sj) {
  const int J = spec->num_links;
  const bool dyn = spec->num_obstacles > 0 && spec->dynamic_obstacles;
  if ((J != 2 && J != 3) || in->batch < 2) return kStrided;
  const bool rows = in->theta_sj == 1 && in->theta_sb == J && in->target_sc == 1 && in->target_sb >= 2 &&
                    (!dyn || (in->obst_sc == 1 && in->obst_so == 4 && in->obst_sb == 4 * kMaxObst && aligned16(in->obstacles))) &&
                    (!grad || (dt_sj == 1 && dt_sb == J));
  if (rows) return kRows;
  // planes: every plane is read two samples at a time with 8-byte accesses
  const bool planes = in->theta_sb == 1 && in->target_sb == 1 && in->batch % 2 == 0 && aligned8(in->theta) && in->theta_sj % 2 == 0 &&
                      aligned8(in->target) && in->target_sc % 2 == 0 &&
                      (!dyn || (in->obst_sb == 1 && aligned8(in->obstacles) && in->obst_so % 2 == 0 && in->obst_sc % 2 == 0)) &&
                      (!grad || (dt_sb == 1 && dt_sj % 2 == 0 && aligned8(dtheta)));
  if (planes) return kPlanes;
  return kStrided;
}

}  // namespace

int prepare_lagrangian_params(const IklSpec* spec, const IklBatch* in, const float* w_host, bool grad, float* dtheta, int64_t dt_sb,
                              int64_t dt_sj, LagrangianParams& p, Variant& v) {
  memset(&p, 0, sizeof(p));
  fill_uniforms(spec, w_host, p.u);
  fill_packed_uniforms(spec, p.u, p.pu);
  if (spec->num_links == 3 && !(spec->num_obstacles > 0 && spec->dynamic_obstacles)) fill_static_pairs(spec, p.u, p.sp);
  p.batch = in->batch;
  p.theta = in->theta;    p.th_sb = in->theta_sb;  p.th_sj = in->theta_sj;
  p.target = in->target;  p.tg_sb = in->target_sb; p.tg_sc = in->target_sc;
  p.obst = in->obstacles; p.ob_sb = in->obst_sb;   p.ob_so = in->obst_so;  p.ob_sc = in->obst_sc;
  p.dtheta = dtheta;      p.dt_sb = dt_sb;         p.dt_sj = dt_sj;

  v.nobs = spec->num_obstacles;
  v.dyn = spec->num_obstacles > 0 && spec->dynamic_obstacles != 0;
  v.jl = spec->enforce_joint_limits != 0;
  v.mode = kFwd;
  v.links = spec->num_links;
  v.unit = unit_links(spec);

  int layout = detect_layout(spec, in, grad, dtheta, dt_sb, dt_sj);
  if (layout == kPlanes) {
    // the TMA-staged kernel describes the planes as tensors: batch % 4 == 0, 16-byte aligned planes, strides % 4 == 0
    const bool dyn = v.dyn;
    const char* e = getenv("IKL_PLANES_TMA");
    const bool want = !(e && e[0] == '0');
    p.use_tma = want && in->batch % 4 == 0 && in->batch >= 4 && in->batch < (int64_t(1) << 31) - 2 * kTmaTile && aligned16(in->theta) &&
                in->theta_sj % 4 == 0 && aligned16(in->target) && in->target_sc % 4 == 0 &&
                (!dyn || (aligned16(in->obstacles) && in->obst_so % 4 == 0 && in->obst_sc % 4 == 0));
    if (p.use_tma) {
      p.use_tma = encode_planes(&p.tm_theta, in->theta, in->batch, spec->num_links, in->theta_sj, 0, 0) &&
                  encode_planes(&p.tm_target, in->target, in->batch, 2, in->target_sc, 0, 0) &&
                  (!dyn || encode_planes(&p.tm_obst, in->obstacles, in->batch, 3, in->obst_sc, v.nobs, in->obst_so));
    }
  }
  if (layout == kRows) {
    // bulk-copy staged rows kernel: whole 512-row tiles of each tensor are single 16-byte aligned runs
    const char* e = getenv("IKL_ROWS_TMA");
    const bool want = !(e && e[0] == '0');
    p.use_tma = want && in->batch >= kTmaTile && in->batch < (int64_t(1) << 31) - 2 * kTmaTile && aligned16(in->theta) && aligned16(in->target) &&
                (in->target_sb == 2 || in->target_sb == 3);
  }
  const int forced = forced_layout();
  if (forced == kStrided) layout = kStrided;                 // the generic kernel accepts anything
  else if (forced >= 0 && forced != layout) return -(int)IKL_ERR_INVALID_ARGUMENT;   // tests: demand a fast layout or fail loudly
  return layout;
}

namespace {

IklStatus run(const IklSpec* spec, const IklBatch* in, const IklWeights* w, void* workspace, float* viol_out, float* loss_out,
              double* viol_accum, float* q_ee_out, float* err_sq_out, float* dtheta, int64_t dt_sb, int64_t dt_sj, bool grad,
              void* stream, const IklExchange* ex = nullptr, bool sums = true, bool close_epoch = false,
              const IklDualStep* dual = nullptr) {
  IklStatus st = validate(spec, in, w, workspace, viol_out, loss_out);
  if (st != IKL_OK) return st;
  if (grad && in->batch > 0 && !dtheta) return IKL_ERR_INVALID_ARGUMENT;
  if (close_epoch && !viol_accum) return IKL_ERR_INVALID_ARGUMENT;
  if (dual && (!close_epoch || !dual->lam_in || !dual->lam_out)) return IKL_ERR_INVALID_ARGUMENT;
  if (ex) {
    if (ex->world < 1 || ex->world > IKL_MAX_RANKS || ex->rank < 0 || ex->rank >= ex->world) return IKL_ERR_INVALID_ARGUMENT;
    for (int r = 0; r < ex->world; ++r)
      if (!ex->peers[r] || (reinterpret_cast<uintptr_t>(ex->peers[r]) & 15u)) return IKL_ERR_INVALID_ARGUMENT;
  }

  LagrangianParams p;
  Variant v;
  const int layout = prepare_lagrangian_params(spec, in, w->host, grad, dtheta, dt_sb, dt_sj, p, v);
  if (layout < 0) return (IklStatus)(-layout);
  v.mode = !grad ? kFwd : (sums ? (w->host ? kGradHostW : kGradDevW) : (w->host ? kGradOnlyHostW : kGradOnlyDevW));
  p.q_ee = q_ee_out;
  p.err_sq = err_sq_out;
  p.w_dev = w->device;
  p.partials = reinterpret_cast<double*>(workspace);
  p.counter = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(workspace) + kPartialsBytes);
  p.viol_out = viol_out;
  p.loss_out = loss_out;
  p.viol_accum = viol_accum;
  if (ex && ex->world > 1) {
    p.ex_world = ex->world;
    p.ex_rank = ex->rank;
    for (int r = 0; r < ex->world; ++r) p.ex_peer[r] = reinterpret_cast<ExchangeBuf*>(ex->peers[r]);
  }
  p.close_epoch = close_epoch ? 1 : 0;
  if (dual) {
    p.dual_lam_in = dual->lam_in;
    p.dual_lam_out = dual->lam_out;
    p.dual_lr = dual->multiplier_lr;
  }

  const int forced = forced_layout();
  cudaStream_t cs = (cudaStream_t)stream;
  bool found = false;
  cudaError_t e = cudaSuccess;
  // packed families: 3 links with the unit lengths folded (the reference's arm), 3 links with runtime lengths, 2 links
  if (layout == kRows) {
    if (v.links == 3 && v.unit) e = launch_family<kRows, 3, true>(v, p, cs, kMaxGrid, &found);
    else if (v.links == 3) e = launch_family<kRows, 3, false>(v, p, cs, kMaxGrid, &found);
    else e = launch_family<kRows, 2, false>(v, p, cs, kMaxGrid, &found);
  } else if (layout == kPlanes) {
    if (v.links == 3 && v.unit) e = launch_family<kPlanes, 3, true>(v, p, cs, kMaxGrid, &found);
    else if (v.links == 3) e = launch_family<kPlanes, 3, false>(v, p, cs, kMaxGrid, &found);
    else e = launch_family<kPlanes, 2, false>(v, p, cs, kMaxGrid, &found);
  }
  if (!found) {
    if (forced == kRows || forced == kPlanes) return IKL_ERR_UNSUPPORTED;
    if (spec->num_links == 3) e = launch_family<kStrided, 3, false>(v, p, cs, kMaxGrid, &found);
    else e = launch_family<kStrided, 2, false>(v, p, cs, kMaxGrid, &found);
    if (!found) return IKL_ERR_UNSUPPORTED;
  }
  return cuda_status(e);
}

}  // namespace
}  // namespace ikl

using namespace ikl;

extern "C" {

int32_t ikl_num_constraints(const IklSpec* spec) {
  if (!spec) return -1;
  return 1 + (spec->enforce_joint_limits ? spec->num_links : 0) + spec->num_links * spec->num_obstacles;
}

size_t ikl_workspace_bytes(int64_t batch) {
  (void)batch;          // the persistent grid is bounded, so the workspace does not grow with the batch
  return kWorkspaceBytes;
}

IklStatus ikl_lagrangian_forward(const IklSpec* spec, const IklBatch* in, const IklWeights* w, void* workspace, float* viol_out,
                                 float* loss_out, double* viol_accum, float* q_ee_out, float* err_sq_out, void* stream) {
  return run(spec, in, w, workspace, viol_out, loss_out, viol_accum, q_ee_out, err_sq_out, nullptr, 0, 0, false, stream);
}

IklStatus ikl_lagrangian_forward_backward(const IklSpec* spec, const IklBatch* in, const IklWeights* w, void* workspace, float* viol_out,
                                          float* loss_out, double* viol_accum, float* dtheta, int64_t dtheta_sb, int64_t dtheta_sj,
                                          void* stream) {
  return run(spec, in, w, workspace, viol_out, loss_out, viol_accum, nullptr, nullptr, dtheta, dtheta_sb, dtheta_sj, true, stream);
}

size_t ikl_exchange_bytes(void) { return sizeof(ExchangeBuf); }

IklStatus ikl_lagrangian_forward_allreduce(const IklSpec* spec, const IklBatch* in, const IklWeights* w, void* workspace, float* viol_out,
                                           float* loss_out, double* viol_accum, const IklExchange* ex, void* stream) {
  if (!ex) return IKL_ERR_INVALID_ARGUMENT;
  return run(spec, in, w, workspace, viol_out, loss_out, viol_accum, nullptr, nullptr, nullptr, 0, 0, false, stream, ex);
}

IklStatus ikl_lagrangian_forward_backward_allreduce(const IklSpec* spec, const IklBatch* in, const IklWeights* w, void* workspace,
                                                    float* viol_out, float* loss_out, double* viol_accum, float* dtheta,
                                                    int64_t dtheta_sb, int64_t dtheta_sj, const IklExchange* ex, void* stream) {
  if (!ex) return IKL_ERR_INVALID_ARGUMENT;
  return run(spec, in, w, workspace, viol_out, loss_out, viol_accum, nullptr, nullptr, dtheta, dtheta_sb, dtheta_sj, true, stream, ex);
}

IklStatus ikl_lagrangian_forward_dual_step(const IklSpec* spec, const IklBatch* in, const IklWeights* w, void* workspace, float* viol_out,
                                           float* loss_out, double* viol_accum, const IklExchange* ex, const IklDualStep* dual,
                                           void* stream) {
  return run(spec, in, w, workspace, viol_out, loss_out, viol_accum, nullptr, nullptr, nullptr, 0, 0, false, stream, ex, true, true, dual);
}

IklStatus ikl_lagrangian_forward_backward_dual_step(const IklSpec* spec, const IklBatch* in, const IklWeights* w, void* workspace,
                                                    float* viol_out, float* loss_out, double* viol_accum, float* dtheta,
                                                    int64_t dtheta_sb, int64_t dtheta_sj, const IklExchange* ex, const IklDualStep* dual,
                                                    void* stream) {
  return run(spec, in, w, workspace, viol_out, loss_out, viol_accum, nullptr, nullptr, dtheta, dtheta_sb, dtheta_sj, true, stream, ex,
             true, true, dual);
}

IklStatus ikl_lagrangian_backward(const IklSpec* spec, const IklBatch* in, const IklWeights* w, void* workspace, float* dtheta,
                                  int64_t dtheta_sb, int64_t dtheta_sj, void* stream) {
  if (!workspace) return IKL_ERR_INVALID_ARGUMENT;
  // gradient-only kernels where they exist (the staged layouts): no sums are formed at all.  Elsewhere the sums are a
  // by-product of the pass and land in the scratch tail of the workspace (after the completion counter).
  float* scratch = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kPartialsBytes + 64);
  static const bool with_sums = [] { const char* e = getenv("IKL_BACKWARD_SUMS"); return e && e[0] == '1'; }();
  return run(spec, in, w, workspace, scratch, scratch + kMaxK, nullptr, nullptr, nullptr, dtheta, dtheta_sb, dtheta_sj, true, stream,
             nullptr, with_sums);
}

int32_t ikl_abi_version(void) { return IKL_ABI_VERSION; }

void ikl_set_reserved_sms(int32_t n) { g_reserved_sms.store(n < 0 ? 0 : n, std::memory_order_relaxed); }

const char* ikl_strerror(IklStatus status) {
  switch (status) {
    case IKL_OK: return "ok";
    case IKL_ERR_INVALID_ARGUMENT: return "invalid argument";
    case IKL_ERR_UNSUPPORTED: return "unsupported configuration";
    case IKL_ERR_CUDA: return "CUDA error";
  }
  return "unknown status";
}

int32_t ikl_last_cuda_error(void) { return (int32_t)g_last_cuda_error; }
int64_t ikl_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }
const char* ikl_last_kernel_name(void) { return g_last_kernel; }

}  // extern "C"
