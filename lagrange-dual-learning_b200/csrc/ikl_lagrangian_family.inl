// ikl_lagrangian_family.inl -- included by each ikl_lagrangian_<family>.cu after defining
//   IKL_FAMILY_LAYOUT, IKL_FAMILY_J, IKL_FAMILY_UNIT and IKL_FAMILY_FULL (1: every combination,
//   0: only the combinations the five task configs and their dynamic siblings use).
// Splitting the instantiations over translation units keeps the build parallel.
#include <cstdio>

#include "ikl_lagrangian.cuh"

namespace ikl {
namespace {

template <int NOBS, bool DYN, bool JL>
cudaError_t launch_mode(const Variant& v, const LagrangianParams& p, cudaStream_t stream, int max_blocks) {
  constexpr int L = IKL_FAMILY_LAYOUT, J = IKL_FAMILY_J;
  constexpr bool UNIT = IKL_FAMILY_UNIT;
  static thread_local char name[96];
  static const char* const layout_names[] = {"strided", "rows", "planes"};
  const bool rows_tma = L == kRows && p.use_tma && RowsTmaLayout<Cfg<L, J, NOBS, DYN, JL, kFwd, UNIT>>::kSupported;
  const char* lname = (L == kPlanes && p.use_tma) ? "planes_tma" : (rows_tma ? "rows_tma" : layout_names[L]);
  static const char* const mode_names[] = {"fwd", "fwdbwd_hostw", "fwdbwd_devw", "grad_hostw", "grad_devw"};
  const bool grad_only = v.mode == kGradOnlyHostW || v.mode == kGradOnlyDevW;
  // gradient-only specialisations exist for the staged kernels only (launch_cfg redirects the rest to the full kernel)
  const bool staged = p.use_tma && (L == kPlanes || rows_tma);
  const int shown = (grad_only && !staged) ? (v.mode == kGradOnlyHostW ? kGradHostW : kGradDevW) : v.mode;
  snprintf(name, sizeof(name), "ikl_lagrangian<%s,J%d,obs%d,%s,%s,%s,%s>", lname, J, NOBS, DYN ? "dyn" : "stat",
           JL ? "jl" : "nojl", mode_names[shown], UNIT ? "unit" : "len");
  switch (v.mode) {
    case kFwd: return launch_cfg<Cfg<L, J, NOBS, DYN, JL, kFwd, UNIT>>(p, stream, max_blocks, name);
    case kGradHostW: return launch_cfg<Cfg<L, J, NOBS, DYN, JL, kGradHostW, UNIT>>(p, stream, max_blocks, name);
    case kGradDevW: return launch_cfg<Cfg<L, J, NOBS, DYN, JL, kGradDevW, UNIT>>(p, stream, max_blocks, name);
    case kGradOnlyHostW:
      if constexpr (L == kStrided) return launch_cfg<Cfg<L, J, NOBS, DYN, JL, kGradHostW, UNIT>>(p, stream, max_blocks, name);
      else return launch_cfg<Cfg<L, J, NOBS, DYN, JL, kGradOnlyHostW, UNIT>>(p, stream, max_blocks, name);
    case kGradOnlyDevW:
      if constexpr (L == kStrided) return launch_cfg<Cfg<L, J, NOBS, DYN, JL, kGradDevW, UNIT>>(p, stream, max_blocks, name);
      else return launch_cfg<Cfg<L, J, NOBS, DYN, JL, kGradOnlyDevW, UNIT>>(p, stream, max_blocks, name);
  }
  return cudaErrorInvalidValue;
}

}  // namespace

template <>
cudaError_t launch_family<IKL_FAMILY_LAYOUT, IKL_FAMILY_J, IKL_FAMILY_UNIT>(const Variant& v, const LagrangianParams& p,
                                                                             cudaStream_t stream, int max_blocks, bool* found) {
  *found = true;
#define IKL_CASE(NOBS, DYN, JL) \
  if (v.nobs == NOBS && v.dyn == DYN && v.jl == JL) return launch_mode<NOBS, DYN, JL>(v, p, stream, max_blocks);
  IKL_CASE(0, false, false)
  IKL_CASE(0, false, true)
  IKL_CASE(1, false, true)
  IKL_CASE(4, false, true)
  IKL_CASE(1, true, true)
  IKL_CASE(2, true, true)
  IKL_CASE(3, true, true)
  IKL_CASE(4, true, true)
#if IKL_FAMILY_FULL
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
#endif
#undef IKL_CASE
  *found = false;
  return cudaSuccess;
}

}  // namespace ikl
