// Instantiations of the fused Lagrangian kernels: layout=kStrided, J=2, unit link lengths=false.
#define IKL_FAMILY_LAYOUT ::ikl::kStrided
#define IKL_FAMILY_J 2
#define IKL_FAMILY_UNIT false
#define IKL_FAMILY_FULL 1
#include "ikl_lagrangian_family.inl"
