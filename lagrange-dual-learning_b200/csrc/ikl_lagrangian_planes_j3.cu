// Instantiations of the fused Lagrangian kernels: layout=kPlanes, J=3, unit link lengths=true.
#define IKL_FAMILY_LAYOUT ::ikl::kPlanes
#define IKL_FAMILY_J 3
#define IKL_FAMILY_UNIT true
#define IKL_FAMILY_FULL 1
#include "ikl_lagrangian_family.inl"
