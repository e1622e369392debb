// ikl_common.h -- host-side plumbing shared by the translation units of libikl.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ikl.h"

namespace ikl {

// Thread-local record of the last CUDA error / dispatched kernel (ikl_last_cuda_error, ikl_last_kernel_name).
void set_last_cuda_error(cudaError_t e);
void set_last_kernel_name(const char* name);
void count_launch(int n = 1);

// SM count / of the current device, cached per device ordinal.
int sm_count();
// SMs the persistent Lagrangian grids leave free (ikl_set_reserved_sms), e.g. for a concurrent NCCL kernel.
int reserved_sms();

inline IklStatus cuda_status(cudaError_t e) {
  if (e == cudaSuccess) return IKL_OK;
  set_last_cuda_error(e);
  return IKL_ERR_CUDA;
}


}  // namespace ikl
