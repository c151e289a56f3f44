// Optimiser step + grid resampling launchers (optim_kernels.cu).
#pragma once
#include <cmath>

#include "common.cuh"

namespace tf {

int64_t adam_scratch_bytes(const int64_t* sizes, int n_leaves);
int adam_step(cudaStream_t st, const tensorf_adam_desc* d, const int64_t* sizes, float* const* params, const float* const* grads,
              float* const* mu, float* const* nu, const float* neg_lrs, float* grad_norm, void* scratch,
              int64_t scratch_bytes);

int64_t vm_resize_scratch_bytes(int C, int G_in, int G_out);
int vm_resize(cudaStream_t st, const float* vector_in, const float* matrix_in, int C, int G_in, int G_out, float* vector_out,
              float* matrix_out, void* scratch, int64_t scratch_bytes);

}  // namespace tf
