// Instantiations of the fused Lagrangian kernels: layout=kRows, J=2, unit link lengths=false (runtime lengths),
// for the combinations the five task configs and their dynamic siblings use.
#define IKL_FAMILY_LAYOUT ::ikl::kRows
#define IKL_FAMILY_J 2
#define IKL_FAMILY_UNIT false
#define IKL_FAMILY_FULL 0
#include "ikl_lagrangian_family.inl"
