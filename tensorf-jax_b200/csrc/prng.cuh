// Device-side jax.random restatement + pixel-ray generation (prng_kernels.cu).
#pragma once
#include <algorithm>

#include "common.cuh"

namespace tf {

void threefry2x32_host(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* out2);
int prng_uniform(cudaStream_t st, uint32_t k0, uint32_t k1, int64_t first, int64_t n, float minval, float maxval, float* out);
int prng_gumbel(cudaStream_t st, uint32_t k0, uint32_t k1, int64_t n, float* out);
int pixel_rays(cudaStream_t st, const float* M_host, const float* origin_host, int W, int row0, int row1, uint32_t camera_index,
               float* origins, float* directions, uint32_t* camera_indices);
int pixel_rays_striped(cudaStream_t st, const float* M_host, const float* origin_host, int W, int H, int stripe, int rank, int world,
                       uint32_t camera_index, float* origins, float* directions, uint32_t* camera_indices, int64_t* n_rays);

int gather_rays(cudaStream_t st, const float* origins, const float* directions, const uint32_t* cams, const float* colors,
                int64_t n_table, const int64_t* idx, int64_t R, float* o_out, float* d_out, uint32_t* c_out, float* col_out,
                int* bad_count);
int rgba_over_white(cudaStream_t st, const float* rgba, int64_t n, float* rgb);

}  // namespace tf
