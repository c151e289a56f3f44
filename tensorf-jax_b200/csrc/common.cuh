// Shared host/device helpers for libtensorf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tensorf_b200.h"

namespace tf {

void set_error(const char* fmt, ...);  // thread-local message, api.cu

#define TF_CHECK_ARG(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      tf::set_error(__VA_ARGS__);          \
      return TENSORF_ERR_INVALID_ARGUMENT; \
    }                                      \
  } while (0)

#define TF_CHECK_CUDA(expr)                                                               \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      tf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return TENSORF_ERR_CUDA;                                                            \
    }                                                                                     \
  } while (0)

// Programmatic dependent launch.  A kernel launched through launch_pdl may be scheduled while the kernel before it in the
// stream drains (its set-up overlaps that tail, and the launch latency disappears); it calls pdl_wait() before it touches
// anything its predecessor wrote or reads - a no-op under a normal launch - and every CTA calls it, so that completion of
// the kernel implies completion of everything before it.  pdl_launch_dependents() in a kernel lets its successor's CTAs be
// scheduled as SMs free up instead of when the whole grid has exited.  Captured into CUDA graphs as programmatic edges.
int pdl_level();  // api.cu: TENSORF_PDL = 0 off, 1 the two fused MLP row kernels only, 2 (default) the per-ray kernels too
inline bool pdl_enabled() { return pdl_level() > 0; }
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename A>
inline cudaError_t launch_pdl(void (*kernel)(A), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const A& arg, bool pdl = true, int level = 1) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && pdl_level() >= level) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, arg);
}
#endif

// Launch errors are collected right after enqueue; no device synchronisation.
void count_launch();  // thread-local kernel-launch counter (tensorf_launch_count), api.cu
#define TF_CHECK_LAUNCH()               \
  do {                                  \
    tf::count_launch();                 \
    TF_CHECK_CUDA(cudaGetLastError());  \
  } while (0)

// Optional per-stage timing: when enabled (tensorf_profile_enable) every launcher brackets its
// stages with cudaEventRecord on the launch stream. Off by default (graph-capture safe).
struct StageTimer {
  StageTimer(cudaStream_t st, const char* name);
  ~StageTimer();
  cudaStream_t st_;
  int rec_;
};

#define TF_RETURN_IF_ERROR(expr) \
  do {                           \
    int rc_ = (expr);            \
    if (rc_ != 0) return rc_;    \
  } while (0)

constexpr int kSMs = 148;

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- packed ("texel-major") factor layout --------------------------------------------------
// lines  : [3][G][Cp]       at float offset 0
// planes : [3][G][G][Cp]    at float offset 3*G*Cp
// Cp = C rounded up to a multiple of 4 so that one texel is a whole number of float4.
__host__ __device__ inline int packed_cp(int C) { return (C + 3) & ~3; }
__host__ __device__ inline int64_t packed_line_floats(int C, int G) { return (int64_t)3 * G * packed_cp(C); }
__host__ __device__ inline int64_t packed_plane_floats(int C, int G) { return (int64_t)3 * G * G * packed_cp(C); }
__host__ __device__ inline int64_t packed_floats(int C, int G) { return packed_line_floats(C, G) + packed_plane_floats(C, G); }

#ifdef __CUDACC__
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Vector reduction to global memory (PTX ISA 8.1+, sm_90+): one 16-byte RED per texel fragment.
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
#endif

}  // namespace tf
