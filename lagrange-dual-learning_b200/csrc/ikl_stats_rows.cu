// Evaluation-statistics kernels on the bulk-copy staged rows pipeline.
#define IKL_FAMILY_LAYOUT ::ikl::kRows
#include "ikl_stats_family.inl"
