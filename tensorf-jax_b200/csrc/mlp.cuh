// FeatureMlp (networks.py:38-121) forward / reverse: shared declarations.
#pragma once
#include "common.cuh"

namespace tf {

struct MlpShape {
  int Ca;      // 3*ca input features
  int squash;  // 27
  int units;   // 128
  int Ff, Fv;  // fourier frequencies
  int enc;     // encoded dim (networks.py:77-82)
  int ncam;    // 0 = no embeddings
  int inference;  // forward only: skip residuals of the reverse pass (TENSORF_FLAG_INFERENCE)
};

inline MlpShape mlp_shape(const tensorf_render_desc& d) {
  MlpShape s;
  s.Ca = 3 * d.ca;
  s.squash = d.squash;
  s.units = d.units;
  s.inference = (d.flags & TENSORF_FLAG_INFERENCE) != 0;
  s.Ff = d.feat_freqs;
  s.Fv = d.view_freqs;
  s.enc = d.squash + 3 + 2 * d.feat_freqs * d.squash + 2 * d.view_freqs * 3;
  s.ncam = d.num_cameras;
  return s;
}

// Activation workspace of one MLP call (all fp32, row-major).
struct MlpWs {
  int ldf, ldx;  // row strides of f/df and x/dx: squash and enc rounded up to 16
  unsigned char* wpack;  // split-bf16 packed weights of the tensor-core path
  size_t wpack_bytes;
  float* f;    // (M, ldf)      Dense_0 output
  float* x;    // (M, ldx)      encoded input of Dense_1
  float* h1;   // (M, units)    relu(Dense_1)
  float* h2;   // (M, units)    relu(Dense_2), before FiLM
  float* dp2;  // (M, units)    d pre-activation of Dense_2
  float* dp1;  // (M, units)
  float* dx;   // (M, enc)
  float* df;   // (M, squash)
  uint32_t* bits1;  // (M, units/32) ReLU mask of layer 1, one bit per unit (tensor-core path); layer 2's follow (fused path)
  float* aux;       // fused path: [64] scalars (gradient scale), then the d(output) slab tiles (M, 16)
  float* wg_partial;  // fused path: per-CTA partial weight-gradient accumulators [min(tiles, 148)][320 + 3ca][128]
  bool wpack_ready = false;  // fused path: the caller has already run mlp_fused_pack on this workspace (on a side stream)
  bool pdl = false;  // fused path: launch the row kernels as programmatic dependents of the kernel before them in the stream
  bool beside_scatter = false;  // fused path: the weight-gradient kernel shares the GPU with a scatter kernel (forked reverse pass)
  const unsigned char* feat_slabs = nullptr;  // fused path: the feature rows already as slab tiles (render path: written by
                                              // k_appearance); null = `feat` is fp32 (M, 3ca) and is converted into `dx`
};
int64_t mlp_ws_floats(const MlpShape& s, int64_t M);
MlpWs mlp_ws_carve(const MlpShape& s, int64_t M, float* base);

struct MlpParams {
  const float *w0, *w1, *b1, *w2, *b2, *w3, *b3, *embed;
};
struct MlpGrads {
  float *w0, *w1, *b1, *w2, *b2, *w3, *b3, *embed;
  bool prezeroed = false;  // the caller already zeroed every leaf on the stream (render reverse: k_ray_bwd does it)
  bool amax_ready = false; // fused path: max |d_rgb| is already in ws.aux[0] (render reverse: k_ray_bwd computes it)
};

// rows_per_ray: viewdirs / camera_indices are indexed by row / rows_per_ray (render.py:487-496).
int mlp_simt_fwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const float* feat, const float* viewdirs,
                 const uint32_t* cams, int64_t M, int rows_per_ray, const MlpWs& ws, float* rgb);
int mlp_simt_bwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const float* feat, const float* viewdirs,
                 const uint32_t* cams, int64_t M, int rows_per_ray, const MlpWs& ws, const float* rgb, const float* d_rgb,
                 float* d_feat, const MlpGrads& g);
// tcgen05 tensor-core path (split-bf16 operands, fp32 accumulate), same contract.
int mlp_tc_fwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const float* feat, const float* viewdirs,
               const uint32_t* cams, int64_t M, int rows_per_ray, const MlpWs& ws, float* rgb);
int mlp_tc_bwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const float* feat, const float* viewdirs,
               const uint32_t* cams, int64_t M, int rows_per_ray, const MlpWs& ws, const float* rgb, const float* d_rgb,
               float* d_feat, const MlpGrads& g);
size_t mlp_tc_wpack_bytes(const MlpShape& s);
bool tc_make_a_tensor_map(void* tensor_map, const float* A, int64_t M, int K_valid, int64_t lda, int kc);  // CUtensorMap*, mlp_tc.cu
// per-row-tile fused kernels (mlp_fused.cu): lego-type networks only (mlp_fused_supported), same contract
bool mlp_fused_supported(const MlpShape& s);
size_t mlp_fused_wpack_bytes(const MlpShape& s);
int mlp_fused_fwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const float* feat, const float* viewdirs,
                  const uint32_t* cams, int64_t M, int rows_per_ray, const MlpWs& ws, float* rgb);
int mlp_fused_pack(cudaStream_t st, const MlpShape& s, const MlpParams& p, const MlpWs& ws);
bool mlp_fused_bwd_ready();
int mlp_fused_bwd_chain(cudaStream_t st, const MlpShape& s, const MlpParams& p, int64_t M, const MlpWs& ws, const float* rgb,
                        const float* d_rgb, float* d_feat, const MlpGrads& g);
int mlp_fused_bwd_wgrad(cudaStream_t st, const MlpShape& s, int64_t M, const MlpWs& ws, const MlpGrads& g);
int mlp_fused_bwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const float* feat, const float* viewdirs,
                  const uint32_t* cams, int64_t M, int rows_per_ray, const MlpWs& ws, const float* rgb, const float* d_rgb,
                  float* d_feat, const MlpGrads& g);

// pieces shared by both implementations (mlp_simt.cu)
int mlp_encode_fwd(cudaStream_t st, const MlpShape& s, const MlpWs& ws, const float* viewdirs, int64_t M, int rows_per_ray,
                   bool fast = false);
int mlp_encode_bwd(cudaStream_t st, const MlpShape& s, const MlpWs& ws, int64_t M, bool fast = false);
int mlp_out_fwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const MlpWs& ws, const uint32_t* cams, int64_t M,
                int rows_per_ray, float* rgb);
int mlp_out_bwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const MlpWs& ws, const uint32_t* cams, int64_t M,
                int rows_per_ray, const float* rgb, const float* d_rgb, const MlpGrads& g);
int mlp_colsum128(cudaStream_t st, const float* G, float* out, int64_t M);
int mlp_zero_grads(cudaStream_t st, const MlpShape& s, const MlpGrads& g);

inline bool mlp_use_tc(int mlp_impl) { return mlp_impl != TENSORF_MLP_SIMT_FP32; }

}  // namespace tf
