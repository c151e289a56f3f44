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
};

inline MlpShape mlp_shape(const tensorf_render_desc& d) {
  MlpShape s;
  s.Ca = 3 * d.ca;
  s.squash = d.squash;
  s.units = d.units;
  s.Ff = d.feat_freqs;
  s.Fv = d.view_freqs;
  s.enc = d.squash + 3 + 2 * d.feat_freqs * d.squash + 2 * d.view_freqs * 3;
  s.ncam = d.num_cameras;
  return s;
}

// Activation workspace of one MLP call (all fp32, row-major).
struct MlpWs {
  float* f;    // (M, squash)   Dense_0 output
  float* x;    // (M, enc)      encoded input of Dense_1
  float* h1;   // (M, units)    relu(Dense_1)
  float* h2;   // (M, units)    relu(Dense_2), before FiLM
  float* dp2;  // (M, units)    d pre-activation of Dense_2
  float* dp1;  // (M, units)
  float* dx;   // (M, enc)
  float* df;   // (M, squash)
};
int64_t mlp_ws_floats(const MlpShape& s, int64_t M);
MlpWs mlp_ws_carve(const MlpShape& s, int64_t M, float* base);

struct MlpParams {
  const float *w0, *w1, *b1, *w2, *b2, *w3, *b3, *embed;
};
struct MlpGrads {
  float *w0, *w1, *b1, *w2, *b2, *w3, *b3, *embed;
};

// rows_per_ray: viewdirs / camera_indices are indexed by row / rows_per_ray (render.py:487-496).
int mlp_simt_fwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const float* feat, const float* viewdirs,
                 const uint32_t* cams, int64_t M, int rows_per_ray, const MlpWs& ws, float* rgb);
int mlp_simt_bwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const float* feat, const float* viewdirs,
                 const uint32_t* cams, int64_t M, int rows_per_ray, const MlpWs& ws, const float* rgb, const float* d_rgb,
                 float* d_feat, const MlpGrads& g);

}  // namespace tf
