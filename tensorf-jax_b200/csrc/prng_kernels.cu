// Randomness and ray generation on the device.
//   - threefry2x32 (Random123, 20 rounds) with jax.random's counter layout for
//     random_bits / uniform / gumbel (jax 0.9.0.1, jax_threefry_partitionable=True): element i of a
//     flat array uses counter (hi32(i), lo32(i)) and bits = out0 ^ out1.  Replaces the host draw +
//     H2D copy of the (R,N) jitter of render.py:158-161 and the (N,) vectors of :375-379, :462-468.
//   - pixel rays of one camera (cameras.py:100-143) for a band of image rows: replaces the CPU-pinned
//     jit of cameras.py:124 and the per-frame H2D copy render_360.py makes.
#include <algorithm>

#include "prng.cuh"

namespace tf {

__host__ __device__ inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__host__ __device__ inline void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  const int rot[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  x0 += ks[0];
  x1 += ks[1];
#pragma unroll
  for (int blk = 0; blk < 5; ++blk) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      x0 += x1;
      x1 = rotl32(x1, rot[blk & 1][r]);
      x1 ^= x0;
    }
    x0 += ks[(blk + 1) % 3];
    x1 += ks[(blk + 2) % 3] + (uint32_t)(blk + 1);
  }
}

void threefry2x32_host(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* out2) {
  threefry2x32(k0, k1, x0, x1);
  out2[0] = x0;
  out2[1] = x1;
}

// jax.random.uniform(key, (n,), float32, minval, maxval): 23 mantissa bits | 1.0f, minus 1, scaled, clamped below.
__device__ __forceinline__ float uniform_from_bits(uint32_t bits, float lo, float hi) {
  const float f = __fsub_rn(__uint_as_float((bits >> 9) | 0x3F800000u), 1.0f);
  return fmaxf(lo, __fadd_rn(__fmul_rn(f, __fsub_rn(hi, lo)), lo));
}

template <bool GUMBEL>
__global__ void __launch_bounds__(256) k_prng(uint32_t k0, uint32_t k1, int64_t first, int64_t n, float lo, float hi, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t ctr = (uint64_t)(first + i);  // out[i] = element first + i of the whole draw
    uint32_t x0 = (uint32_t)(ctr >> 32), x1 = (uint32_t)ctr;
    threefry2x32(k0, k1, x0, x1);
    const float u = uniform_from_bits(x0 ^ x1, lo, hi);
    out[i] = GUMBEL ? -logf(-logf(u)) : u;
  }
}

int prng_uniform(cudaStream_t st, uint32_t k0, uint32_t k1, int64_t first, int64_t n, float minval, float maxval, float* out) {
  TF_CHECK_ARG(n >= 0 && first >= 0 && (n == 0 || out), "prng_uniform: bad arguments");
  if (n == 0) return 0;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(n, 256), (int64_t)kSMs * 16);
  k_prng<false><<<grid, 256, 0, st>>>(k0, k1, first, n, minval, maxval, out);
  TF_CHECK_LAUNCH();
  return 0;
}
int prng_gumbel(cudaStream_t st, uint32_t k0, uint32_t k1, int64_t n, float* out) {
  TF_CHECK_ARG(n >= 0 && (n == 0 || out), "prng_gumbel: bad arguments");
  if (n == 0) return 0;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(n, 256), (int64_t)kSMs * 16);
  k_prng<true><<<grid, 256, 0, st>>>(k0, k1, 0, n, 1.17549435e-38f /* finfo(float32).tiny */, 1.0f, out);
  TF_CHECK_LAUNCH();
  return 0;
}

// direction = M @ [u, v, 1] with M = R_world_camera @ K^-1 (cameras.py:107-113), normalised by
// (norm + 1e-8) (:116); origin = T_world_camera.translation() for every pixel.
struct PixelRayArgs {
  float M[9];
  float origin[3];
  int W, row0, row1;
  uint32_t cam;
  // striped variant: output row r (0 <= r < row1 - row0) is image row (r / stripe * world + rank) * stripe + r % stripe
  int stripe, rank, world;  // stripe == 0: contiguous band starting at row0
};
__global__ void __launch_bounds__(256) k_pixel_rays(const __grid_constant__ PixelRayArgs a, float* __restrict__ origins,
                                                    float* __restrict__ directions, uint32_t* __restrict__ cams) {
  const int64_t n = (int64_t)(a.row1 - a.row0) * a.W;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r_local = (int)(i / a.W);
  const int row = a.stripe > 0 ? (r_local / a.stripe * a.world + a.rank) * a.stripe + r_local % a.stripe : a.row0 + r_local;
  const float u = (float)(i % a.W), v = (float)row;
  float d[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) d[r] = __fadd_rn(__fadd_rn(__fmul_rn(a.M[3 * r], u), __fmul_rn(a.M[3 * r + 1], v)), a.M[3 * r + 2]);
  const float nrm = __fadd_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]))), 1e-8f);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    directions[3 * i + r] = __fdiv_rn(d[r], nrm);
    origins[3 * i + r] = a.origin[r];
  }
  if (cams) cams[i] = a.cam;
}

int pixel_rays(cudaStream_t st, const float* M_host, const float* origin_host, int W, int row0, int row1, uint32_t camera_index,
               float* origins, float* directions, uint32_t* camera_indices) {
  TF_CHECK_ARG(M_host && origin_host && origins && directions, "pixel_rays: null argument");
  TF_CHECK_ARG(W >= 1 && row0 >= 0 && row1 >= row0, "pixel_rays: bad image band W=%d rows [%d,%d)", W, row0, row1);
  const int64_t n = (int64_t)(row1 - row0) * W;
  if (n == 0) return 0;
  PixelRayArgs a;
  for (int i = 0; i < 9; ++i) a.M[i] = M_host[i];
  for (int i = 0; i < 3; ++i) a.origin[i] = origin_host[i];
  a.W = W; a.row0 = row0; a.row1 = row1; a.cam = camera_index;
  a.stripe = 0; a.rank = 0; a.world = 1;
  k_pixel_rays<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(a, origins, directions, camera_indices);
  TF_CHECK_LAUNCH();
  return 0;
}

int pixel_rays_striped(cudaStream_t st, const float* M_host, const float* origin_host, int W, int H, int stripe, int rank, int world,
                       uint32_t camera_index, float* origins, float* directions, uint32_t* camera_indices, int64_t* n_rays) {
  TF_CHECK_ARG(M_host && origin_host, "pixel_rays_striped: null argument");
  TF_CHECK_ARG(W >= 1 && H >= 1 && stripe >= 1 && world >= 1 && rank >= 0 && rank < world, "pixel_rays_striped: bad arguments");
  // rows of this rank: whole stripes k*world + rank, the last one possibly cut by the image height
  int rows = 0;
  for (int k = rank; k * stripe < H; k += world) rows += std::min(stripe, H - k * stripe);
  if (n_rays) *n_rays = (int64_t)rows * W;
  if (!origins && !directions) return 0;  // size query
  TF_CHECK_ARG(origins && directions, "pixel_rays_striped: null output");
  if (rows == 0) return 0;
  PixelRayArgs a;
  for (int i = 0; i < 9; ++i) a.M[i] = M_host[i];
  for (int i = 0; i < 3; ++i) a.origin[i] = origin_host[i];
  a.W = W; a.row0 = 0; a.row1 = rows; a.cam = camera_index;
  a.stripe = stripe; a.rank = rank; a.world = world;
  k_pixel_rays<<<(unsigned)ceil_div64((int64_t)rows * W, 256), 256, 0, st>>>(a, origins, directions, camera_indices);
  TF_CHECK_LAUNCH();
  return 0;
}

}  // namespace tf

// ---- training data path --------------------------------------------------
// The ray table (data.py:301-337: every pixel of every training view as origin, direction, camera id,
// colour = 40 B) stays resident on the device; a minibatch is a gather of R table rows by a shuffled index
// (training.py:318-323), so a step moves 4*R B of indices instead of 40*R B of rays and never leaves the stream.
namespace tf {

__global__ void __launch_bounds__(256) k_gather_rays(const float* __restrict__ origins, const float* __restrict__ directions,
                                                     const uint32_t* __restrict__ cams, const float* __restrict__ colors,
                                                     const int64_t* __restrict__ idx, int64_t R, int64_t n_table,
                                                     float* __restrict__ o_out, float* __restrict__ d_out,
                                                     uint32_t* __restrict__ c_out, float* __restrict__ col_out,
                                                     int* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  int64_t j = idx[i];
  if (j < 0 || j >= n_table) {  // never read out of the table; the host checks `bad` when it wants to
    if (bad) atomicAdd(bad, 1);
    j = 0;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    o_out[3 * i + a] = __ldg(origins + 3 * j + a);
    d_out[3 * i + a] = __ldg(directions + 3 * j + a);
    if (colors) col_out[3 * i + a] = __ldg(colors + 3 * j + a);
  }
  if (cams) c_out[i] = __ldg(cams + j);
}

int gather_rays(cudaStream_t st, const float* origins, const float* directions, const uint32_t* cams, const float* colors,
                int64_t n_table, const int64_t* idx, int64_t R, float* o_out, float* d_out, uint32_t* c_out, float* col_out,
                int* bad_count) {
  TF_CHECK_ARG(R >= 0 && n_table >= 1, "gather_rays: R=%lld n_table=%lld", (long long)R, (long long)n_table);
  if (bad_count) {
    TF_CHECK_CUDA(cudaMemsetAsync(bad_count, 0, sizeof(int), st));
    count_launch();
  }
  if (R == 0) return 0;  // empty minibatch: buffers may be null
  TF_CHECK_ARG(origins && directions && idx && o_out && d_out, "gather_rays: null argument");
  TF_CHECK_ARG((cams == nullptr) == (c_out == nullptr) && (colors == nullptr) == (col_out == nullptr),
               "gather_rays: optional table columns and outputs must be given together");
  k_gather_rays<<<(unsigned)ceil_div64(R, 256), 256, 0, st>>>(origins, directions, cams, colors, idx, R, n_table, o_out, d_out,
                                                               c_out, col_out, bad_count);
  TF_CHECK_LAUNCH();
  return 0;
}

// rgba (n,4) -> rgb over an opaque white background: rgb*a + (1-a)  (data.py:318-320)
__global__ void __launch_bounds__(256) k_rgba_over_white(const float* __restrict__ rgba, int64_t n, float* __restrict__ rgb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(reinterpret_cast<const float4*>(rgba) + i);
  const float bg = __fsub_rn(1.0f, p.w);
  rgb[3 * i + 0] = __fadd_rn(__fmul_rn(p.x, p.w), bg);
  rgb[3 * i + 1] = __fadd_rn(__fmul_rn(p.y, p.w), bg);
  rgb[3 * i + 2] = __fadd_rn(__fmul_rn(p.z, p.w), bg);
}

int rgba_over_white(cudaStream_t st, const float* rgba, int64_t n, float* rgb) {
  TF_CHECK_ARG(n >= 0 && (n == 0 || (rgba && rgb)), "rgba_over_white: bad arguments");
  TF_CHECK_ARG((reinterpret_cast<uintptr_t>(rgba) & 15) == 0, "rgba_over_white: rgba must be 16-byte aligned");
  if (n == 0) return 0;
  k_rgba_over_white<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(rgba, n, rgb);
  TF_CHECK_LAUNCH();
  return 0;
}

}  // namespace tf
