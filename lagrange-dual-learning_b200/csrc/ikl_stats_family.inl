// ikl_stats_family.inl -- included by ikl_stats_<layout>.cu after defining IKL_FAMILY_LAYOUT: the evaluation-statistics
// kernels for every (obstacle count, static / dynamic, joint limits on / off) combination.
#include "ikl_lagrangian.cuh"

namespace ikl {

template <>
cudaError_t launch_stats_family<IKL_FAMILY_LAYOUT>(const Variant& v, const LagrangianParams& p, cudaStream_t stream, int max_blocks,
                                                   bool* found) {
  constexpr int L = IKL_FAMILY_LAYOUT;
  *found = true;
#define IKL_CASE(NOBS, DYN, JL)                                                                        \
  if (v.nobs == NOBS && v.dyn == DYN && v.jl == JL) {                                                  \
    using C = Cfg<L, 3, NOBS, DYN, JL, kFwd, true>;                                                    \
    if constexpr (L == kPlanes || RowsTmaLayout<C>::kSupported) return launch_stats_cfg<C>(p, stream, max_blocks); \
  }
  IKL_CASE(0, false, false)
  IKL_CASE(0, false, true)
  IKL_CASE(1, false, true)
  IKL_CASE(4, false, true)
  IKL_CASE(1, true, true)
  IKL_CASE(2, true, true)
  IKL_CASE(3, true, true)
  IKL_CASE(4, true, true)
  IKL_CASE(2, false, true)
  IKL_CASE(3, false, true)
  IKL_CASE(1, false, false)
  IKL_CASE(2, false, false)
  IKL_CASE(3, false, false)
  IKL_CASE(4, false, false)
  IKL_CASE(1, true, false)
  IKL_CASE(2, true, false)
  IKL_CASE(3, true, false)
  IKL_CASE(4, true, false)
#undef IKL_CASE
  *found = false;
  return cudaSuccess;
}

}  // namespace ikl
