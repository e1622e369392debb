// Instantiations of the fused Lagrangian kernels: layout=kPlanes, J=3, unit link lengths=false (runtime lengths),
// for the combinations the five task configs and their dynamic siblings use.
#define IKL_FAMILY_LAYOUT ::ikl::kPlanes
#define IKL_FAMILY_J 3
#define IKL_FAMILY_UNIT false
#define IKL_FAMILY_FULL 0
#include "ikl_lagrangian_family.inl"
