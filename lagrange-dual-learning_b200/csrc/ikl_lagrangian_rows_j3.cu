// Instantiations of the fused Lagrangian kernels: layout=kRows, J=3, unit link lengths=true.
#define IKL_FAMILY_LAYOUT ::ikl::kRows
#define IKL_FAMILY_J 3
#define IKL_FAMILY_UNIT true
#define IKL_FAMILY_FULL 1
#include "ikl_lagrangian_family.inl"
