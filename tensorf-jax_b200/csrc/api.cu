// extern "C" entry points of libtensorf_b200.so (see include/tensorf_b200.h).
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "mlp.cuh"
#include "optim.cuh"
#include "prng.cuh"
#include "render_kernels.cuh"

namespace tf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static thread_local int64_t g_launches = 0;
void count_launch() { ++g_launches; }

// ---- stage profiler -------------------------------------------------------------------------
struct ProfRec {
  const char* name;
  cudaEvent_t e0, e1;
};
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  std::vector<ProfRec> recs;
  cudaEvent_t get() {
    if (used == pool.size()) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
      pool.push_back(e);
    }
    return pool[used++];
  }
};
static thread_local Profiler g_prof;

StageTimer::StageTimer(cudaStream_t st, const char* name) : st_(st), rec_(-1) {
  if (!g_prof.on) return;
  cudaEvent_t e0 = g_prof.get(), e1 = g_prof.get();
  if (!e0 || !e1) return;
  cudaEventRecord(e0, st);
  rec_ = (int)g_prof.recs.size();
  g_prof.recs.push_back(ProfRec{name, e0, e1});
}
StageTimer::~StageTimer() {
  if (rec_ >= 0) cudaEventRecord(g_prof.recs[rec_].e1, st_);
}

// One internal side stream (+ three fork / join events) per host thread and device, created on first use.
struct SideStream {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
};
static SideStream* side_stream() {
  static thread_local SideStream ss[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SideStream& s = ss[dev];
  if (s.stream == nullptr) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (auto& e : s.ev)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    s.device = dev;
  }
  return &s;
}

int pdl_level() {
  static const int lv = getenv("TENSORF_PDL") != nullptr ? atoi(getenv("TENSORF_PDL")) : 2;
  return lv;
}

static inline int64_t al(int64_t floats) { return round_up64(floats, 64); }  // 256-byte granules

struct RenderWs {
  float *packed_d, *packed_a, *gpacked_d, *gpacked_a;
  float *z, *dz, *xs, *pt_sel, *stats, *feat, *d_feat, *rgb_sel, *d_rgb_sel, *go;
  int32_t* idx;
  float* mlp_base;
  int64_t total_floats;
};

static RenderWs carve(const tensorf_render_desc& d, float* base) {
  RenderWs w{};
  const int64_t R = d.R, N = d.N, K = d.K, M = R * K;
  const MlpShape ms = mlp_shape(d);
  int64_t off = 0;
  auto take = [&](int64_t n) {
    float* p = base ? base + off : nullptr;
    off += al(n);
    return p;
  };
  w.packed_d = take(packed_floats(d.cd, d.G));
  w.packed_a = take(packed_floats(d.ca, d.G));
  w.gpacked_d = take(packed_floats(d.cd, d.G));
  w.gpacked_a = take(packed_floats(d.ca, d.G));
  w.z = take(R * N);
  w.dz = take(R * N);
  w.xs = take(R * N * 3);
  w.idx = reinterpret_cast<int32_t*>(take(M));
  w.pt_sel = take(M);
  w.stats = take(R * 8);
  w.go = take(R * 3);
  w.feat = take(round_up64(M, 128) * ms.Ca);  // fused MLP: fp16 slab tiles of 128 rows (same bytes per row as fp32)
  w.d_feat = take(M * ms.Ca);
  w.rgb_sel = take(M * 3);
  w.d_rgb_sel = take(M * 3);
  w.mlp_base = take(mlp_ws_floats(ms, M));
  w.total_floats = off;
  return w;
}

// TENSORF_FLAG_PACKED_FACTORS: the caller's factor buffers ARE the packed copies (no pack / unpack passes)
static bool packed_factors(const tensorf_render_desc& d) { return (d.flags & TENSORF_FLAG_PACKED_FACTORS) != 0; }
static int check_packed_leaf(const tensorf_render_desc& d, const float* vec, const float* mat, int C, const char* what) {
  TF_CHECK_ARG(vec != nullptr, "%s: packed factor buffer is NULL", what);
  TF_CHECK_ARG((reinterpret_cast<uintptr_t>(vec) & 15) == 0, "%s: packed factor buffer must be 16-byte aligned", what);
  TF_CHECK_ARG(mat == nullptr || mat == vec + packed_line_floats(C, d.G), "%s: with TENSORF_FLAG_PACKED_FACTORS the matrix pointer must be NULL or vector + 3*G*Cp", what);
  return 0;
}

static int check_desc(const tensorf_render_desc* d) {
  TF_CHECK_ARG(d != nullptr, "desc is NULL");
  TF_CHECK_ARG(d->R >= 0 && d->N >= 1 && d->G >= 2, "bad shape R=%d N=%d G=%d", d->R, d->N, d->G);
  TF_CHECK_ARG(d->cd >= 1 && d->ca >= 1, "bad channel dims cd=%d ca=%d", d->cd, d->ca);
  TF_CHECK_ARG(d->mode >= TENSORF_MODE_RGB && d->mode <= TENSORF_MODE_DIST_MEAN, "bad render mode %d", d->mode);
  if (d->mode == TENSORF_MODE_RGB) {
    TF_CHECK_ARG(d->K >= 1 && d->K <= d->N, "appearance_samples_per_ray=%d must be in [1, %d]", d->K, d->N);
    if (d->units != 128) {
      set_error("FeatureMlp units=%d unsupported (128 only)", d->units);
      return TENSORF_ERR_UNSUPPORTED;
    }
    TF_CHECK_ARG(d->squash >= 1 && d->squash <= 32, "feature_squash_dim=%d must be in [1,32]", d->squash);
    TF_CHECK_ARG(d->feat_freqs >= 0 && d->view_freqs >= 0 && d->num_cameras >= 0, "bad MLP config");
    TF_CHECK_ARG((int64_t)d->R * d->K < (int64_t)1 << 31, "R*K too large");
  }
  TF_CHECK_ARG((int64_t)d->R * d->N < (int64_t)1 << 31, "R*N too large");
  // the gather / scatter kernels index the packed factors with 32-bit float offsets
  TF_CHECK_ARG(packed_floats(d->cd, d->G) < ((int64_t)1 << 31) && packed_floats(d->ca, d->G) < ((int64_t)1 << 31),
               "factor grids too large (G=%d cd=%d ca=%d): >= 2^31 packed floats", d->G, d->cd, d->ca);
  return 0;
}

static int check_inputs(const tensorf_render_desc* d, const tensorf_render_inputs* in) {
  TF_CHECK_ARG(in != nullptr, "inputs is NULL");
  TF_CHECK_ARG(in->origins && in->directions && in->aabb && in->jitter, "origins/directions/aabb/jitter must be non-NULL");
  if (d->contracted) TF_CHECK_ARG(in->base_ts && in->deltas, "contracted scene needs base_ts and deltas");
  if (d->mode == TENSORF_MODE_RGB) {
    TF_CHECK_ARG(in->gumbel != nullptr, "RGB mode needs the gumbel vector");
    if (d->num_cameras > 0) TF_CHECK_ARG(in->camera_indices != nullptr, "camera embeddings need camera_indices");
  }
  return 0;
}

static void fill_scene(SceneArgs& a, const tensorf_render_desc& d, const tensorf_render_inputs& in) {
  a.origins = in.origins;
  a.directions = in.directions;
  a.aabb = in.aabb;
  a.jitter = in.jitter;
  a.base_ts = in.base_ts;
  a.deltas = in.deltas;
  a.R = d.R;
  a.N = d.N;
  a.K = d.K;
  a.G = d.G;
  a.contracted = d.contracted;
}

static MlpParams mlp_params(const tensorf_params& p) {
  return MlpParams{p.w0, p.w1, p.b1, p.w2, p.b2, p.w3, p.b3, p.embed};
}
static MlpGrads mlp_grads(const tensorf_params& p) { return MlpGrads{p.w0, p.w1, p.b1, p.w2, p.b2, p.w3, p.b3, p.embed}; }

// MLP implementation of a call: SIMT fp32, per-layer tcgen05 kernels, or the per-row-tile fused tcgen05 kernels
// (AUTO = fused when the network qualifies, else per-layer).
static int mlp_pick(const tensorf_render_desc& d, const MlpShape& ms, int* impl) {
  int m = d.mlp_impl;
  if (m == TENSORF_MLP_AUTO)
    m = (mlp_fused_supported(ms) && (ms.inference || mlp_fused_bwd_ready()) && !getenv("TENSORF_NO_FUSED_MLP")) ? TENSORF_MLP_FUSED
                                                                                                              : TENSORF_MLP_TCGEN05;
  if (m == TENSORF_MLP_FUSED && !mlp_fused_supported(ms)) {
    set_error("mlp_impl FUSED needs squash 27, units 128, 2 + 2 frequencies, no camera embeddings, 3*ca %% 16 == 0 <= 160");
    return TENSORF_ERR_UNSUPPORTED;
  }
  *impl = m;
  return 0;
}
static int mlp_fwd_any(int impl, cudaStream_t st, const MlpShape& s, const MlpParams& p, const float* feat, const float* viewdirs,
                       const uint32_t* cams, int64_t M, int rows_per_ray, const MlpWs& ws, float* rgb) {
  if (impl == TENSORF_MLP_FUSED) return mlp_fused_fwd(st, s, p, feat, viewdirs, cams, M, rows_per_ray, ws, rgb);
  return (impl == TENSORF_MLP_SIMT_FP32 ? mlp_simt_fwd : mlp_tc_fwd)(st, s, p, feat, viewdirs, cams, M, rows_per_ray, ws, rgb);
}
static int mlp_bwd_any(int impl, cudaStream_t st, const MlpShape& s, const MlpParams& p, const float* feat, const float* viewdirs,
                       const uint32_t* cams, int64_t M, int rows_per_ray, const MlpWs& ws, const float* rgb, const float* d_rgb,
                       float* d_feat, const MlpGrads& g) {
  if (impl == TENSORF_MLP_FUSED) return mlp_fused_bwd(st, s, p, feat, viewdirs, cams, M, rows_per_ray, ws, rgb, d_rgb, d_feat, g);
  return (impl == TENSORF_MLP_SIMT_FP32 ? mlp_simt_bwd : mlp_tc_bwd)(st, s, p, feat, viewdirs, cams, M, rows_per_ray, ws, rgb, d_rgb, d_feat, g);
}

static int check_mlp_params(const tensorf_render_desc& d, const tensorf_params* p, const char* what) {
  TF_CHECK_ARG(p != nullptr, "%s is NULL", what);
  TF_CHECK_ARG(p->w0 && p->w1 && p->b1 && p->w2 && p->b2 && p->w3 && p->b3, "%s: MLP leaves must be non-NULL", what);
  if (d.num_cameras > 0) TF_CHECK_ARG(p->embed != nullptr, "%s: embed must be non-NULL with camera embeddings", what);
  return 0;
}

int tc_rowgemm_test(cudaStream_t st, const float* A, int64_t M, int K, const float* W, int N, const float* bias, int relu,
                    const uint32_t* mask_bits, uint32_t* bits_out, float* C, void* scratch, size_t scratch_bytes, int nsplit);
int tc_redgemm_test(cudaStream_t st, const float* G, int Mg, const float* X, int Nx, int64_t rows, float* out);
int tc_trace_read(long long* host, int n);
int tc_umma_bench(cudaStream_t st, int N, int layout_type, int lbo, int sbo, int count, long long* out_dev);
int umma_probe(cudaStream_t st, const float* A, const float* B, float* D, int N, int K, int a_mode, int b_mode, int a_fmt, int b_fmt);

}  // namespace tf

using namespace tf;

extern "C" {

const char* tensorf_last_error(void) { return g_err; }
int tensorf_version(void) { return 100; }

int64_t tensorf_launch_count(void) { return g_launches; }

int tensorf_profile_enable(int enable) {
  g_prof.on = enable != 0;
  g_prof.used = 0;
  g_prof.recs.clear();
  return 0;
}

int tensorf_profile_read(int max_entries, char* names, float* total_ms, int* calls, int* count) {
  TF_CHECK_ARG(names && total_ms && calls && count && max_entries > 0, "NULL argument");
  int n = 0;
  for (auto& r : g_prof.recs) {
    TF_CHECK_CUDA(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    TF_CHECK_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    int j = 0;
    for (; j < n; ++j)
      if (strncmp(names + 32 * j, r.name, 31) == 0) break;
    if (j == n) {
      if (n == max_entries) continue;
      strncpy(names + 32 * n, r.name, 31);
      names[32 * n + 31] = 0;
      total_ms[n] = 0.f;
      calls[n] = 0;
      ++n;
    }
    total_ms[j] += ms;
    calls[j] += 1;
  }
  *count = n;
  g_prof.used = 0;
  g_prof.recs.clear();
  return 0;
}

int64_t tensorf_vm_packed_floats(int C, int G) { return packed_floats(C, G); }

int tensorf_vm_pack(tensorf_stream_t s, const float* vector, const float* matrix, float* packed, int C, int G) {
  TF_CHECK_ARG(vector && matrix && packed, "NULL buffer");
  TF_CHECK_ARG(C >= 1 && G >= 2, "bad C=%d G=%d", C, G);
  return vm_pack((cudaStream_t)s, vector, matrix, packed, C, G);
}
int tensorf_vm_unpack(tensorf_stream_t s, const float* packed, float* vector, float* matrix, int C, int G) {
  TF_CHECK_ARG(vector && matrix && packed, "NULL buffer");
  TF_CHECK_ARG(C >= 1 && G >= 2, "bad C=%d G=%d", C, G);
  return vm_unpack((cudaStream_t)s, packed, vector, matrix, C, G);
}

int tensorf_vm_interp_fwd(tensorf_stream_t s, const float* packed, const float* ijk, float* out, int C, int G, int64_t B,
                          int feature_major) {
  TF_CHECK_ARG(B >= 0 && C >= 1 && G >= 2, "bad shape C=%d G=%d B=%lld", C, G, (long long)B);
  TF_CHECK_ARG(packed_floats(C, G) < ((int64_t)1 << 31), "factor grid too large (C=%d G=%d): >= 2^31 packed floats", C, G);
  TF_CHECK_ARG(B == 0 || (packed && ijk && out), "NULL buffer");
  return vm_interp_fwd((cudaStream_t)s, packed, ijk, out, C, G, B, feature_major);
}
int tensorf_vm_interp_bwd(tensorf_stream_t s, const float* packed, const float* ijk, const float* d_out, float* d_packed,
                          int C, int G, int64_t B, int feature_major) {
  TF_CHECK_ARG(B >= 0 && C >= 1 && G >= 2, "bad shape C=%d G=%d B=%lld", C, G, (long long)B);
  TF_CHECK_ARG(packed_floats(C, G) < ((int64_t)1 << 31), "factor grid too large (C=%d G=%d): >= 2^31 packed floats", C, G);
  TF_CHECK_ARG(B == 0 || (packed && ijk && d_out && d_packed), "NULL buffer");
  return vm_interp_bwd((cudaStream_t)s, packed, ijk, d_out, d_packed, C, G, B, feature_major);
}

int tensorf_topk_select(tensorf_stream_t s, const float* g, int R, int N, int K, int32_t* idx) {
  TF_CHECK_ARG(R >= 0 && N >= 1 && K >= 1 && K <= N, "bad shape R=%d N=%d K=%d (need 1 <= K <= N)", R, N, K);
  TF_CHECK_ARG(R == 0 || (g && idx), "NULL buffer");
  return launch_topk_select((cudaStream_t)s, g, R, N, K, idx);
}

int tensorf_segment_probabilities(tensorf_stream_t s, const float* sigmas, const float* step_sizes, int R, int N,
                                  float* p_exits, float* p_terminates) {
  TF_CHECK_ARG(R >= 0 && N >= 1, "bad shape R=%d N=%d", R, N);
  TF_CHECK_ARG(R == 0 || (sigmas && step_sizes && p_exits && p_terminates), "NULL buffer");
  return launch_segment_probs((cudaStream_t)s, sigmas, step_sizes, p_exits, p_terminates, R, N);
}

int tensorf_tc_rowgemm_test(tensorf_stream_t s, const float* A, int64_t M, int K, const float* W, int N, const float* bias,
                            int relu, const uint32_t* mask_bits, uint32_t* bits_out, float* C, void* scratch,
                            int64_t scratch_bytes, int nsplit) {
  TF_CHECK_ARG(A && W && C && scratch && M >= 0 && K >= 1 && N >= 1, "bad argument");
  TF_CHECK_ARG(nsplit == 2 || nsplit == 3, "nsplit must be 2 or 3");
  TF_CHECK_ARG(!(mask_bits || bits_out) || N <= 256, "bit masks need N <= 256");
  return tc_rowgemm_test((cudaStream_t)s, A, M, K, W, N, bias, relu, mask_bits, bits_out, C, scratch, (size_t)scratch_bytes, nsplit);
}
int tensorf_tc_redgemm_test(tensorf_stream_t s, const float* G, int Mg, const float* X, int Nx, int64_t rows, float* out) {
  TF_CHECK_ARG(G && X && out && Mg >= 1 && Mg <= 128 && Nx >= 1 && Nx <= 512 && rows >= 0, "bad argument");
  return tc_redgemm_test((cudaStream_t)s, G, Mg, X, Nx, rows, out);
}

int tensorf_tc_trace_read(long long* host, int n) { return tc_trace_read(host, n); }
int tensorf_tc_umma_bench(tensorf_stream_t s, int N, int layout_type, int lbo, int sbo, int count, long long* out_dev) {
  TF_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0 && count > 0 && out_dev, "bad argument");
  return tc_umma_bench((cudaStream_t)s, N, layout_type, lbo, sbo, count, out_dev);
}

int tensorf_tc_umma_probe(tensorf_stream_t s, const float* A, const float* B, float* D, int N, int K, int a_mode, int b_mode,
                          int a_fmt, int b_fmt) {
  TF_CHECK_ARG(A && B && D, "NULL buffer");
  return umma_probe((cudaStream_t)s, A, B, D, N, K, a_mode, b_mode, a_fmt, b_fmt);
}

int64_t tensorf_mlp_workspace_bytes(const tensorf_render_desc* d, int64_t M) {
  if (!d || M < 0) return -1;
  return (int64_t)sizeof(float) * mlp_ws_floats(mlp_shape(*d), M);
}

int tensorf_mlp_workspace_layout(const tensorf_render_desc* d, int64_t M, int64_t* offsets) {
  TF_CHECK_ARG(d && M >= 0 && offsets, "bad arguments");
  const MlpShape ms = mlp_shape(*d);
  float* base = reinterpret_cast<float*>(uintptr_t(1) << 40);
  const MlpWs w = mlp_ws_carve(ms, M, base);
  const float* p[10] = {w.f, w.df, w.x, w.dx, w.h1, w.h2, w.dp2, w.dp1, reinterpret_cast<float*>(w.bits1),
                        reinterpret_cast<float*>(w.bits1) + round_up64(M, 128) * 4};
  for (int i = 0; i < 10; ++i) offsets[i] = p[i] - base;
  return 0;
}

int tensorf_mlp_fwd(tensorf_stream_t s, const tensorf_render_desc* d, const tensorf_params* p, const float* features,
                    const float* viewdirs, const uint32_t* camera_indices, int64_t M, int rows_per_ray, void* workspace,
                    float* rgb) {
  TF_CHECK_ARG(d != nullptr, "desc is NULL");
  if (d->units != 128) {
    set_error("FeatureMlp units=%d unsupported (128 only)", d->units);
    return TENSORF_ERR_UNSUPPORTED;
  }
  TF_CHECK_ARG(d->squash >= 1 && d->squash <= 32 && d->ca >= 1, "bad MLP config");
  TF_CHECK_ARG(M >= 0 && rows_per_ray >= 1 && M % rows_per_ray == 0, "M=%lld must be a multiple of rows_per_ray=%d",
               (long long)M, rows_per_ray);
  TF_RETURN_IF_ERROR(check_mlp_params(*d, p, "params"));
  TF_CHECK_ARG(M == 0 || (features && viewdirs && workspace && rgb), "NULL buffer");
  if (d->num_cameras > 0) TF_CHECK_ARG(M == 0 || camera_indices, "camera embeddings need camera_indices");
  MlpShape ms = mlp_shape(*d);
  MlpWs ws = mlp_ws_carve(ms, M, (float*)workspace);
  int impl;
  TF_RETURN_IF_ERROR(mlp_pick(*d, ms, &impl));
  return mlp_fwd_any(impl, (cudaStream_t)s, ms, mlp_params(*p), features, viewdirs,
                                                                camera_indices, M, rows_per_ray, ws, rgb);
}

int tensorf_mlp_bwd(tensorf_stream_t s, const tensorf_render_desc* d, const tensorf_params* p, const float* features,
                    const float* viewdirs, const uint32_t* camera_indices, int64_t M, int rows_per_ray, void* workspace,
                    const float* rgb, const float* d_rgb, float* d_features, const tensorf_params* grads) {
  TF_CHECK_ARG(d != nullptr, "desc is NULL");
  if (d->units != 128) {
    set_error("FeatureMlp units=%d unsupported (128 only)", d->units);
    return TENSORF_ERR_UNSUPPORTED;
  }
  TF_CHECK_ARG(M >= 0 && rows_per_ray >= 1 && M % rows_per_ray == 0, "M=%lld must be a multiple of rows_per_ray=%d",
               (long long)M, rows_per_ray);
  TF_RETURN_IF_ERROR(check_mlp_params(*d, p, "params"));
  TF_RETURN_IF_ERROR(check_mlp_params(*d, grads, "grads"));
  TF_CHECK_ARG(M == 0 || (features && viewdirs && workspace && rgb && d_rgb && d_features), "NULL buffer");
  MlpShape ms = mlp_shape(*d);
  MlpWs ws = mlp_ws_carve(ms, M, (float*)workspace);
  int impl;
  TF_RETURN_IF_ERROR(mlp_pick(*d, ms, &impl));
  return mlp_bwd_any(impl, (cudaStream_t)s, ms, mlp_params(*p), features, viewdirs,
                                                                camera_indices, M, rows_per_ray, ws, rgb, d_rgb, d_features,
                                                                mlp_grads(*grads));
}

int tensorf_render_workspace_bytes(const tensorf_render_desc* d, int64_t* bytes) {
  TF_RETURN_IF_ERROR(check_desc(d));
  TF_CHECK_ARG(bytes != nullptr, "bytes is NULL");
  tensorf_render_desc dd = *d;
  if (dd.mode != TENSORF_MODE_RGB) dd.K = 1;
  *bytes = carve(dd, nullptr).total_floats * (int64_t)sizeof(float);
  return 0;
}

int tensorf_render_workspace_view(const tensorf_render_desc* d, void* workspace, const char* name, void** ptr,
                                  int64_t* count) {
  TF_RETURN_IF_ERROR(check_desc(d));
  TF_CHECK_ARG(workspace && name && ptr && count, "NULL argument");
  RenderWs w = carve(*d, (float*)workspace);
  const int64_t R = d->R, N = d->N, M = R * d->K, Ca = 3 * d->ca;
  struct { const char* n; void* p; int64_t c; } views[] = {
      {"z", w.z, R * N},           {"dz", w.dz, R * N},          {"idx", w.idx, M},
      {"pt_sel", w.pt_sel, M},     {"stats", w.stats, R * 8},    {"feat", w.feat, M * Ca},
      {"d_feat", w.d_feat, M * Ca}, {"rgb_sel", w.rgb_sel, M * 3}, {"d_rgb_sel", w.d_rgb_sel, M * 3},
      {"go", w.go, R * 3},         {"packed_d", w.packed_d, packed_floats(d->cd, d->G)},
      {"packed_a", w.packed_a, packed_floats(d->ca, d->G)},
  };
  for (auto& v : views)
    if (strcmp(v.n, name) == 0) {
      *ptr = v.p;
      *count = v.c;
      return 0;
    }
  set_error("unknown workspace view '%s'", name);
  return TENSORF_ERR_INVALID_ARGUMENT;
}

int tensorf_render_depth(tensorf_stream_t s, const tensorf_render_desc* d, const tensorf_params* p,
                         const tensorf_render_inputs* in, void* workspace, float* depth) {
  TF_RETURN_IF_ERROR(check_desc(d));
  TF_CHECK_ARG(d->mode == TENSORF_MODE_DIST_MEDIAN || d->mode == TENSORF_MODE_DIST_MEAN, "render_depth needs a DIST_* mode");
  TF_RETURN_IF_ERROR(check_inputs(d, in));
  TF_CHECK_ARG(p != nullptr, "params is NULL");
  const bool packed = packed_factors(*d);
  if (packed) TF_RETURN_IF_ERROR(check_packed_leaf(*d, p->density_vector, p->density_matrix, d->cd, "params density"));
  else TF_CHECK_ARG(p->density_vector && p->density_matrix, "density factors must be non-NULL");
  TF_CHECK_ARG(workspace && (depth || d->R == 0), "NULL buffer");
  cudaStream_t st = (cudaStream_t)s;
  tensorf_render_desc dd = *d;
  dd.K = 1;
  RenderWs w = carve(dd, (float*)workspace);
  if (packed) w.packed_d = p->density_vector;
  else TF_RETURN_IF_ERROR(vm_pack(st, p->density_vector, p->density_matrix, w.packed_d, d->cd, d->G));
  DensityArgs a{};
  fill_scene(a, dd, *in);
  a.packed_d = w.packed_d;
  a.gumbel = nullptr;
  a.Cp = packed_cp(d->cd);
  a.mode = d->mode;
  a.z_out = w.z;
  a.depth_out = depth;
  return launch_density_select(st, a);
}

int tensorf_render_rgb_fwd(tensorf_stream_t s, const tensorf_render_desc* d, const tensorf_params* p,
                           const tensorf_render_inputs* in, void* workspace, float* rgb, float* loss) {
  TF_RETURN_IF_ERROR(check_desc(d));
  TF_CHECK_ARG(d->mode == TENSORF_MODE_RGB, "render_rgb_fwd needs mode RGB");
  TF_RETURN_IF_ERROR(check_inputs(d, in));
  TF_RETURN_IF_ERROR(check_mlp_params(*d, p, "params"));
  const bool packed = packed_factors(*d);
  if (packed) {
    TF_RETURN_IF_ERROR(check_packed_leaf(*d, p->density_vector, p->density_matrix, d->cd, "params density"));
    TF_RETURN_IF_ERROR(check_packed_leaf(*d, p->appearance_vector, p->appearance_matrix, d->ca, "params appearance"));
  } else {
    TF_CHECK_ARG(p->density_vector && p->density_matrix && p->appearance_vector && p->appearance_matrix,
                 "factor leaves must be non-NULL");
  }
  TF_CHECK_ARG(workspace && (rgb || d->R == 0), "NULL buffer");
  TF_CHECK_ARG(!in->colors || loss, "loss must be non-NULL when colors are given");
  cudaStream_t st = (cudaStream_t)s;
  RenderWs w = carve(*d, (float*)workspace);
  if (packed) {
    w.packed_d = p->density_vector;
    w.packed_a = p->appearance_vector;
  }
  const MlpShape ms = mlp_shape(*d);
  int mlp_impl;
  TF_RETURN_IF_ERROR(mlp_pick(*d, ms, &mlp_impl));
  const int64_t M = (int64_t)d->R * d->K;

  // Programmatic dependent launch of the fused MLP kernels (their set-up - 166 KB of weights into shared memory, tensor
  // memory, barriers - overlaps the tail of the kernel before them).  The weights are packed first thing, so that the only
  // output of its predecessor the forward kernel waits for are the feature tiles.
  if (in->colors) TF_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));  // (not between two kernels of the chain)
  MlpWs mws = mlp_ws_carve(ms, M, w.mlp_base);
  if (mlp_impl == TENSORF_MLP_FUSED && pdl_enabled() && M > 0) {
    StageTimer t_(st, "mlp_pack");
    TF_RETURN_IF_ERROR(mlp_fused_pack(st, ms, mlp_params(*p), mws));
    mws.wpack_ready = true;
    mws.pdl = true;
  }
  if (!packed) {
    StageTimer t_(st, "pack");
    TF_RETURN_IF_ERROR(vm_pack2(st, p->density_vector, p->density_matrix, w.packed_d, d->cd, p->appearance_vector,
                                p->appearance_matrix, w.packed_a, d->ca, d->G));
  }

  DensityArgs a{};
  fill_scene(a, *d, *in);
  a.packed_d = w.packed_d;
  a.gumbel = in->gumbel;
  a.Cp = packed_cp(d->cd);
  a.mode = d->mode;
  a.z_out = w.z;
  a.xs_out = w.xs;
  a.idx_out = w.idx;
  a.pt_sel_out = w.pt_sel;
  a.stats_out = w.stats;
  {
    StageTimer t_(st, "density_select");
    TF_RETURN_IF_ERROR(launch_density_select(st, a));
  }

  AppearanceArgs ap{};
  fill_scene(ap, *d, *in);
  ap.packed_a = w.packed_a;
  ap.idx = w.idx;
  ap.xs = w.xs;
  ap.C = d->ca;
  ap.Cp = packed_cp(d->ca);
  ap.M = M;
  ap.feat = w.feat;
  {
    StageTimer t_(st, "appearance_gather");
    ap.feat_slabs = mlp_impl == TENSORF_MLP_FUSED;  // the fused MLP kernels take the rows as two-term fp16 slab tiles
    TF_RETURN_IF_ERROR(launch_appearance(st, ap, false));
  }

  if (mlp_impl == TENSORF_MLP_FUSED) mws.feat_slabs = reinterpret_cast<const unsigned char*>(w.feat);  // written by k_appearance
  {
    StageTimer t_(st, "mlp_fwd");
    TF_RETURN_IF_ERROR(mlp_fwd_any(mlp_impl, st, ms, mlp_params(*p), w.feat, in->directions,
                                                                              in->camera_indices, M, d->K, mws, w.rgb_sel));
  }
  StageTimer t_(st, "composite");
  CompositeArgs c{};
  c.rgb_sel = w.rgb_sel;
  c.pt_sel = w.pt_sel;
  c.stats = w.stats;
  c.colors = in->colors;
  c.rgb_out = rgb;
  c.go = w.go;
  c.loss = loss;
  c.loss_scale = d->loss_scale;
  c.R = d->R;
  c.K = d->K;
  return launch_composite_fwd(st, c);
}

// phases: 0 = whole reverse pass (ray, density scatter, MLP, appearance scatter, one unpack);
// 1 = appearance half (ray_bwd, MLP reverse, appearance scatter, unpack appearance): every gradient leaf except the
//     density factors is final when it returns, so a sharded caller can start their exchange and overlap it with
// 2 = density half (density scatter, unpack density); needs phase 1 of the same step (dz lives in the workspace).
static int render_rgb_bwd_impl(tensorf_stream_t s, const tensorf_render_desc* d, const tensorf_params* p,
                               const tensorf_render_inputs* in, void* workspace, const float* d_rgb,
                               const tensorf_params* grads, int phase) {
  TF_RETURN_IF_ERROR(check_desc(d));
  TF_CHECK_ARG(phase >= 0 && phase <= 2, "render_rgb_bwd: phase %d outside [0,2]", phase);
  TF_CHECK_ARG(d->mode == TENSORF_MODE_RGB, "render_rgb_bwd needs mode RGB");
  TF_CHECK_ARG(!(d->flags & TENSORF_FLAG_INFERENCE), "render_rgb_bwd after a forward with TENSORF_FLAG_INFERENCE (residuals were not kept)");
  TF_RETURN_IF_ERROR(check_inputs(d, in));
  TF_RETURN_IF_ERROR(check_mlp_params(*d, p, "params"));
  TF_RETURN_IF_ERROR(check_mlp_params(*d, grads, "grads"));
  const bool packed = packed_factors(*d);
  if (packed) {
    TF_RETURN_IF_ERROR(check_packed_leaf(*d, p->density_vector, p->density_matrix, d->cd, "params density"));
    TF_RETURN_IF_ERROR(check_packed_leaf(*d, p->appearance_vector, p->appearance_matrix, d->ca, "params appearance"));
    TF_RETURN_IF_ERROR(check_packed_leaf(*d, grads->density_vector, grads->density_matrix, d->cd, "grads density"));
    TF_RETURN_IF_ERROR(check_packed_leaf(*d, grads->appearance_vector, grads->appearance_matrix, d->ca, "grads appearance"));
  } else {
    TF_CHECK_ARG(grads->density_vector && grads->density_matrix && grads->appearance_vector && grads->appearance_matrix,
                 "factor gradient leaves must be non-NULL");
  }
  TF_CHECK_ARG(workspace != nullptr, "workspace is NULL");
  TF_CHECK_ARG(d_rgb || in->colors, "either d_rgb or inputs->colors (fused loss) is required");
  cudaStream_t st = (cudaStream_t)s;
  RenderWs w = carve(*d, (float*)workspace);
  if (packed) {  // parameters are read, gradients accumulated, straight in the caller's packed buffers
    w.packed_d = p->density_vector;
    w.packed_a = p->appearance_vector;
    w.gpacked_d = grads->density_vector;
    w.gpacked_a = grads->appearance_vector;
  }
  const MlpShape ms = mlp_shape(*d);
  int mlp_impl;
  TF_RETURN_IF_ERROR(mlp_pick(*d, ms, &mlp_impl));
  const int64_t M = (int64_t)d->R * d->K;
  const bool do_app = phase != 2, do_den = phase != 1;

  if (do_app) {
    RayBwdArgs rb{};
    fill_scene(rb, *d, *in);
    rb.z = w.z;
    rb.idx = w.idx;
    rb.pt_sel = w.pt_sel;
    rb.rgb_sel = w.rgb_sel;
    rb.stats = w.stats;
    rb.go = d_rgb ? d_rgb : w.go;
    rb.d_rgb_sel = w.d_rgb_sel;
    rb.dz = w.dz;
    if (mlp_impl == TENSORF_MLP_FUSED) {  // the fused MLP reverse carries its gradients scaled by a power of two near 1 / max |d_rgb_sel|
      rb.amax = mlp_ws_carve(ms, M, w.mlp_base).aux;
      TF_CHECK_CUDA(cudaMemsetAsync(rb.amax, 0, sizeof(float), st));
    }
    // the packed gradient accumulators are zeroed by k_ray_bwd's threads (no memset launches)
    rb.zero0 = reinterpret_cast<float4*>(w.gpacked_a);
    rb.zero0_n4 = packed_floats(d->ca, d->G) / 4;
    if (do_den) {
      rb.zero1 = reinterpret_cast<float4*>(w.gpacked_d);
      rb.zero1_n4 = packed_floats(d->cd, d->G) / 4;
    }
    {  // ... and so are the MLP gradient leaves (mlp_*_bwd accumulates into them)
      const MlpGrads g = mlp_grads(*grads);
      const int U = ms.units;
      auto add = [&](float* ptr, int64_t n) {
        if (ptr && n > 0) {
          rb.zleaf[rb.n_zleaf] = ptr;
          rb.zleaf_n[rb.n_zleaf] = (int)n;
          ++rb.n_zleaf;
        }
      };
      add(g.w0, (int64_t)ms.Ca * ms.squash);
      add(g.w1, (int64_t)ms.enc * U);
      add(g.b1, U);
      add(g.w2, (int64_t)U * U);
      add(g.b2, U);
      add(g.w3, (int64_t)U * 3);
      add(g.b3, 3);
      if (ms.ncam) add(g.embed, (int64_t)ms.ncam * U);
    }
    StageTimer t_(st, "ray_bwd");
    TF_RETURN_IF_ERROR(launch_ray_bwd(st, rb));
  } else {  // phase 2: the density accumulator was left alone by phase 1
    StageTimer t_(st, "zero_grads");
    TF_CHECK_CUDA(cudaMemsetAsync(w.gpacked_d, 0, sizeof(float) * packed_floats(d->cd, d->G), st));
  }

  // Whole pass with the fused MLP (and no stage timers): two branches after k_ray_bwd that only meet again at the end.
  //   launch stream : activation-gradient chain -> weight gradients (HBM-bound, 6 warps per SM) -> reduce
  //   side stream   :                              density scatter -> appearance scatter (needs the chain's d_features)
  // The scatters are issue-bound at L2 rates and leave the DRAM pipe idle, the weight-gradient kernel streams 0.54 GB with
  // a handful of warps: they share the SMs.  Fork / join are event dependencies, so the schedule is CUDA-graph capturable.
  const bool fork = phase == 0 && mlp_impl == TENSORF_MLP_FUSED && !g_prof.on && !getenv("TENSORF_NO_FORK") && M > 0;
  SideStream* ss = fork ? side_stream() : nullptr;
  auto density_scatter = [&](cudaStream_t s_) -> int {
    DensityBwdArgs db{};
    fill_scene(db, *d, *in);
    db.packed_d = w.packed_d;
    db.dz = w.dz;
    db.xs = w.xs;
    db.d_packed = w.gpacked_d;
    db.Cp = packed_cp(d->cd);
    StageTimer t_(s_, "density_scatter");
    return launch_density_scatter(s_, db);
  };
  auto appearance_scatter = [&](cudaStream_t s_) -> int {
    AppearanceArgs ap{};
    fill_scene(ap, *d, *in);
    ap.packed_a = w.packed_a;
    ap.idx = w.idx;
    ap.xs = w.xs;
    ap.C = d->ca;
    ap.Cp = packed_cp(d->ca);
    ap.M = M;
    ap.d_feat = w.d_feat;
    ap.d_packed = w.gpacked_a;
    ap.beside_mlp = fork || (g_prof.on && phase == 0 && mlp_impl == TENSORF_MLP_FUSED && !getenv("TENSORF_NO_FORK"));
    StageTimer t_(s_, "appearance_scatter");
    return launch_appearance_scatter(s_, ap);
  };
  if (ss != nullptr) {
    MlpWs mws = mlp_ws_carve(ms, M, w.mlp_base);
    mws.feat_slabs = reinterpret_cast<const unsigned char*>(w.feat);
    MlpGrads mg = mlp_grads(*grads);
    mg.prezeroed = true;  // by k_ray_bwd above
    mg.amax_ready = true;
    mws.beside_scatter = true;
    mws.pdl = true;
    // Launch order = placement order: the MLP kernels are launched first so that their one CTA per SM is resident
    // everywhere (chain: 576 threads x 72 registers, weight gradients: 192 x 98) and the scatter CTAs fill what is left of
    // the register file (one / two density-scatter CTAs per SM instead of three; the scatters use no shared memory).
    // (Measured and dropped: a third branch for the appearance scatter, and the MLP branch on a high-priority stream so
    // that the weight-gradient CTAs are placed before the density scatter's pending ones - the branch ends 110 us earlier
    // but the join does not move: the overlapped region is bound by the SMs' total work, not by the dependencies.)
    static const int fork_mode = getenv("TENSORF_FORK_MODE") ? atoi(getenv("TENSORF_FORK_MODE")) : 2;
    // TENSORF_FORK_TIMES=1 (eager launches only): prints where each branch of the PREVIOUS call ended, relative to the fork
    static const bool fork_times = getenv("TENSORF_FORK_TIMES") != nullptr;
    static thread_local cudaEvent_t tev[6] = {nullptr};
    static thread_local int tcalls = 0;
    auto mark = [&](int i, cudaStream_t s_) {
      if (fork_times) cudaEventRecord(tev[i], s_);
    };
    if (fork_times) {
      if (tev[0] == nullptr) {
        for (auto& e : tev) cudaEventCreate(&e);
      } else if (cudaEventSynchronize(tev[5]) == cudaSuccess && (++tcalls % 8) == 0) {
        float t[6] = {0};
        for (int i = 1; i < 6; ++i) cudaEventElapsedTime(&t[i], tev[0], tev[i]);
        fprintf(stderr, "tensorf_b200 fork (us from k_ray_bwd's end): chain %.1f  wgrad+reduce %.1f | density scatter %.1f  appearance scatter %.1f | join %.1f\n",
                1e3f * t[1], 1e3f * t[2], 1e3f * t[3], 1e3f * t[4], 1e3f * t[5]);
      }
    }
    mark(0, st);
    if (fork_mode == 2) TF_CHECK_CUDA(cudaEventRecord(ss->ev[0], st));  // k_ray_bwd done: dz is final
    TF_RETURN_IF_ERROR(mlp_fused_bwd_chain(st, ms, mlp_params(*p), M, mws, w.rgb_sel, w.d_rgb_sel, w.d_feat, mg));
    mark(1, st);
    if (fork_mode == 2) {  // density scatter beside the chain kernel already
      TF_CHECK_CUDA(cudaStreamWaitEvent(ss->stream, ss->ev[0], 0));
      TF_RETURN_IF_ERROR(density_scatter(ss->stream));
      mark(3, ss->stream);
    }
    TF_CHECK_CUDA(cudaEventRecord(ss->ev[1], st));  // d_features are final
    TF_RETURN_IF_ERROR(mlp_fused_bwd_wgrad(st, ms, M, mws, mg));
    mark(2, st);
    TF_CHECK_CUDA(cudaStreamWaitEvent(ss->stream, ss->ev[1], 0));
    if (fork_mode != 2) {
      TF_RETURN_IF_ERROR(density_scatter(ss->stream));
      mark(3, ss->stream);
    }
    TF_RETURN_IF_ERROR(appearance_scatter(ss->stream));
    mark(4, ss->stream);
    TF_CHECK_CUDA(cudaEventRecord(ss->ev[2], ss->stream));
    TF_CHECK_CUDA(cudaStreamWaitEvent(st, ss->ev[2], 0));
    mark(5, st);
  } else {
    if (do_den && phase == 0) TF_RETURN_IF_ERROR(density_scatter(st));
    if (do_app) {
      MlpWs mws = mlp_ws_carve(ms, M, w.mlp_base);
      if (mlp_impl == TENSORF_MLP_FUSED) mws.feat_slabs = reinterpret_cast<const unsigned char*>(w.feat);  // written by k_appearance
      // with stage timers on, time the kernel variants of the forked pass (which the timers alone keep this call from using)
      mws.beside_scatter = g_prof.on && phase == 0 && !getenv("TENSORF_NO_FORK");
      {
        StageTimer t_(st, "mlp_bwd");
        MlpGrads mg = mlp_grads(*grads);
        mg.prezeroed = true;  // by k_ray_bwd above
        mg.amax_ready = mlp_impl == TENSORF_MLP_FUSED;
        TF_RETURN_IF_ERROR(mlp_bwd_any(mlp_impl, st, ms, mlp_params(*p), w.feat, in->directions, in->camera_indices, M, d->K, mws,
                                       w.rgb_sel, w.d_rgb_sel, w.d_feat, mg));
      }
      TF_RETURN_IF_ERROR(appearance_scatter(st));
    }
  }
  if (phase == 2) TF_RETURN_IF_ERROR(density_scatter(st));
  if (packed) return 0;  // the gradients are already where the caller wants them
  StageTimer t_(st, "unpack");
  if (phase == 0)
    return vm_unpack2(st, w.gpacked_d, grads->density_vector, grads->density_matrix, d->cd, w.gpacked_a,
                      grads->appearance_vector, grads->appearance_matrix, d->ca, d->G);
  if (phase == 1) return vm_unpack(st, w.gpacked_a, grads->appearance_vector, grads->appearance_matrix, d->ca, d->G);
  return vm_unpack(st, w.gpacked_d, grads->density_vector, grads->density_matrix, d->cd, d->G);
}

int tensorf_render_rgb_bwd(tensorf_stream_t s, const tensorf_render_desc* d, const tensorf_params* p,
                           const tensorf_render_inputs* in, void* workspace, const float* d_rgb,
                           const tensorf_params* grads) {
  return render_rgb_bwd_impl(s, d, p, in, workspace, d_rgb, grads, 0);
}
int tensorf_render_rgb_bwd_phase(tensorf_stream_t s, const tensorf_render_desc* d, const tensorf_params* p,
                                 const tensorf_render_inputs* in, void* workspace, const float* d_rgb,
                                 const tensorf_params* grads, int phase) {
  return render_rgb_bwd_impl(s, d, p, in, workspace, d_rgb, grads, phase);
}

int64_t tensorf_adam_scratch_bytes(const int64_t* sizes, int n_leaves) {
  if (!sizes || n_leaves < 0) return -1;
  return adam_scratch_bytes(sizes, n_leaves);
}
int tensorf_adam_step(tensorf_stream_t s, const tensorf_adam_desc* d, const int64_t* sizes, float* const* params,
                      const float* const* grads, float* const* mu, float* const* nu, const float* neg_lrs,
                      float* grad_norm, void* scratch, int64_t scratch_bytes) {
  return adam_step((cudaStream_t)s, d, sizes, params, grads, mu, nu, neg_lrs, grad_norm, scratch, scratch_bytes);
}

void tensorf_peer_shard(int64_t total, int rank, int world, int64_t* begin, int64_t* end) {
  if (!begin || !end) return;
  if (world < 1 || rank < 0 || rank >= world || total < 0) {
    *begin = *end = 0;
    return;
  }
  peer_shard(total, rank, world, begin, end);
}
int64_t tensorf_peer_adam_scratch_bytes(int64_t shard_floats) { return peer_adam_scratch_bytes(shard_floats); }
int tensorf_adam_step_peer(tensorf_stream_t s, const tensorf_peer_adam_desc* d, const int64_t* leaf_offsets,
                           const float* neg_lrs, const float* const* grad_peers, float* const* param_peers,
                           const float* grad_mc, float* param_mc, float* mu_shard, float* nu_shard,
                           float* const* norm_slot_peers, void* scratch, int64_t scratch_bytes) {
  return adam_step_peer((cudaStream_t)s, d, leaf_offsets, neg_lrs, grad_peers, param_peers, grad_mc, param_mc, mu_shard,
                        nu_shard, norm_slot_peers, scratch, scratch_bytes);
}
int tensorf_peer_set_max_ctas(int max_ctas) { return peer_set_max_ctas(max_ctas); }
int tensorf_peer_allreduce(tensorf_stream_t s, int rank, int world, int64_t total, float* const* peers, float* mc) {
  return peer_allreduce((cudaStream_t)s, rank, world, total, peers, mc, nullptr, nullptr, 0);
}
int tensorf_peer_allreduce_sync(tensorf_stream_t s, int rank, int world, int64_t total, float* const* peers, float* mc,
                                uint32_t* const* signal_peers, uint32_t* local_flags, uint32_t epoch) {
  if (!signal_peers || !local_flags) {
    set_error("peer_allreduce_sync: signal_peers and local_flags are required");
    return TENSORF_ERR_INVALID_ARGUMENT;
  }
  return peer_allreduce((cudaStream_t)s, rank, world, total, peers, mc, signal_peers, local_flags, epoch);
}
int tensorf_peer_grad_norm(tensorf_stream_t s, const float* norm_slots, int world, float* grad_norm) {
  return peer_grad_norm((cudaStream_t)s, norm_slots, world, grad_norm);
}

int64_t tensorf_vm_resize_scratch_bytes(int C, int G_in, int G_out) {
  if (C < 1 || G_in < 2 || G_out < 2) return -1;
  return vm_resize_scratch_bytes(C, G_in, G_out);
}
int tensorf_vm_resize(tensorf_stream_t s, const float* vector_in, const float* matrix_in, int C, int G_in, int G_out,
                      float* vector_out, float* matrix_out, void* scratch, int64_t scratch_bytes) {
  return vm_resize((cudaStream_t)s, vector_in, matrix_in, C, G_in, G_out, vector_out, matrix_out, scratch, scratch_bytes);
}

void tensorf_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* out2) {
  if (out2) threefry2x32_host(k0, k1, x0, x1, out2);
}
int tensorf_prng_uniform(tensorf_stream_t s, uint32_t k0, uint32_t k1, int64_t n, float minval, float maxval, float* out) {
  return prng_uniform((cudaStream_t)s, k0, k1, 0, n, minval, maxval, out);
}
int tensorf_prng_uniform_slice(tensorf_stream_t s, uint32_t k0, uint32_t k1, int64_t first, int64_t n, float minval, float maxval,
                               float* out) {
  return prng_uniform((cudaStream_t)s, k0, k1, first, n, minval, maxval, out);
}
int tensorf_prng_gumbel(tensorf_stream_t s, uint32_t k0, uint32_t k1, int64_t n, float* out) {
  return prng_gumbel((cudaStream_t)s, k0, k1, n, out);
}
int tensorf_pixel_rays_striped(tensorf_stream_t s, const float* M, const float* origin, int W, int H, int stripe, int rank, int world,
                               uint32_t camera_index, float* origins, float* directions, uint32_t* camera_indices, int64_t* n_rays) {
  return pixel_rays_striped((cudaStream_t)s, M, origin, W, H, stripe, rank, world, camera_index, origins, directions, camera_indices, n_rays);
}
int tensorf_pixel_rays(tensorf_stream_t s, const float* M, const float* origin, int W, int row0, int row1,
                       uint32_t camera_index, float* origins, float* directions, uint32_t* camera_indices) {
  return pixel_rays((cudaStream_t)s, M, origin, W, row0, row1, camera_index, origins, directions, camera_indices);
}

int tensorf_gather_rays(tensorf_stream_t s, const float* origins, const float* directions, const uint32_t* camera_indices,
                        const float* colors, int64_t n_table, const int64_t* idx, int64_t R, float* out_origins,
                        float* out_directions, uint32_t* out_camera_indices, float* out_colors, int* bad_count) {
  return gather_rays((cudaStream_t)s, origins, directions, camera_indices, colors, n_table, idx, R, out_origins, out_directions,
                     out_camera_indices, out_colors, bad_count);
}
int tensorf_rgba_over_white(tensorf_stream_t s, const float* rgba, int64_t n, float* rgb) {
  return rgba_over_white((cudaStream_t)s, rgba, n, rgb);
}

}  // extern "C"
