// Evaluation-statistics kernels on the TMA-staged planes pipeline.
#define IKL_FAMILY_LAYOUT ::ikl::kPlanes
#include "ikl_stats_family.inl"
