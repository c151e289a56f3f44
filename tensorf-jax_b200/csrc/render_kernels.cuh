// Argument blocks and host launchers of the per-ray kernels (render_kernels.cu) and of the
// TensorVM kernels (vm_kernels.cu).
#pragma once
#include "common.cuh"

namespace tf {

struct SceneArgs {
  const float* origins;     // (R,3)
  const float* directions;  // (R,3)
  const float* aabb;        // (2,3)
  const float* jitter;      // (N,) or (R,N)
  const float* base_ts;     // (N,) contracted
  const float* deltas;      // (N,) contracted
  int R, N, K, G, contracted;
};

struct DensityArgs : SceneArgs {
  const float* packed_d;
  const float* gumbel;  // (N,)
  int Cp, mode;
  float* z_out;        // (R,N)
  float* xs_out;       // (R,N,3) continuous grid coordinates of every sample (reused by the reverse kernels)
  int32_t* idx_out;    // (R,K) ascending
  float* pt_sel_out;   // (R,K)
  float* stats_out;    // (R,8): E_last, S, u, -, W0, W1, W2, -
  float* depth_out;    // (R,) depth modes
};

struct AppearanceArgs : SceneArgs {
  const float* packed_a;
  const int32_t* idx;  // (M,)
  const float* xs;     // (R,N,3) grid coordinates saved by the forward gather
  int C, Cp;
  int64_t M;
  float* feat;           // fwd: (M,3C) fp32, or with feat_slabs the fused MLP's operand format: per 128 rows and 8 columns one
                         // hi and one lo slab of 128 x 16 B (csrc/umma_tiles.cuh), rows M..ceil128(M) zero
  bool feat_slabs;
  const float* d_feat;   // bwd
  float* d_packed;       // bwd (+=)
  bool beside_mlp = false;  // bwd: launched beside the fused MLP's weight-gradient kernel (forked reverse pass): 96-register variant
};

struct CompositeArgs {
  const float* rgb_sel;  // (M,3)
  const float* pt_sel;   // (M,)
  float* stats;          // (R,8)
  const float* colors;   // (R,3) or null
  float* rgb_out;        // (R,3)
  float* go;             // (R,3) loss cotangent (when colors)
  float* loss;           // scalar (when colors)
  float loss_scale;
  int R, K;
};

struct RayBwdArgs : SceneArgs {
  const float* z;        // (R,N)
  const int32_t* idx;    // (R,K)
  const float* pt_sel;   // (R,K)
  const float* rgb_sel;  // (M,3)
  const float* stats;    // (R,8)
  const float* go;       // (R,3)
  float* d_rgb_sel;      // (M,3)
  float* amax;           // or null: receives max |d_rgb_sel| (bits, atomicMax; zeroed by the caller) - the fused MLP reverse scales by it
  float* dz;             // (R,N)
  // packed gradient accumulators zeroed by this kernel's threads before their ray work (the stores overlap the
  // latency-bound reverse scan instead of costing two memsets): float4 counts, either may be 0
  float4* zero0;
  int64_t zero0_n4;
  float4* zero1;
  int64_t zero1_n4;
  float* zleaf[8];  // small unaligned leaves (the MLP gradients), zeroed the same way
  int zleaf_n[8];
  int n_zleaf;
};

struct DensityBwdArgs : SceneArgs {
  const float* packed_d;
  const float* dz;  // (R,N)
  const float* xs;  // (R,N,3)
  float* d_packed;  // (+=)
  int Cp;
};

int launch_density_select(cudaStream_t st, const DensityArgs& A);
int launch_segment_probs(cudaStream_t st, const float* sigmas, const float* steps, float* p_exits, float* p_term, int R, int N);
int launch_topk_select(cudaStream_t st, const float* g, int R, int N, int K, int32_t* idx);
int launch_appearance(cudaStream_t st, const AppearanceArgs& A, bool bwd);
int launch_composite_fwd(cudaStream_t st, const CompositeArgs& A);
int launch_ray_bwd(cudaStream_t st, const RayBwdArgs& A);
int launch_density_scatter(cudaStream_t st, const DensityBwdArgs& A);
int launch_appearance_scatter(cudaStream_t st, const AppearanceArgs& A);

int vm_pack(cudaStream_t st, const float* vector, const float* matrix, float* packed, int C, int G);
int vm_unpack(cudaStream_t st, const float* packed, float* vector, float* matrix, int C, int G);
int vm_pack2(cudaStream_t st, const float* v0, const float* m0, float* p0, int C0, const float* v1, const float* m1, float* p1, int C1,
             int G);
int vm_unpack2(cudaStream_t st, const float* p0, float* v0, float* m0, int C0, const float* p1, float* v1, float* m1, int C1, int G);
int vm_interp_fwd(cudaStream_t st, const float* packed, const float* ijk, float* out, int C, int G, int64_t B, int feature_major);
int vm_interp_bwd(cudaStream_t st, const float* packed, const float* ijk, const float* d_out, float* d_packed, int C, int G,
                  int64_t B, int feature_major);

}  // namespace tf
