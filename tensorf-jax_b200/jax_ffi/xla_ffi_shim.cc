// XLA typed-FFI handlers over the C ABI of libtensorf_b200.so (include/tensorf_b200.h): one handler per entry point
// the JAX host code of tensorf-jax needs on the hot path (jax_ffi/tensorf_jax.py binds them):
//
//   TensorfRenderRgbFwd   render.py:105-279 (RGB)            tensorf_render_rgb_fwd
//   TensorfRenderRgbBwd   its reverse, training.py:153-156    tensorf_render_rgb_bwd
//   TensorfRenderDepth    render.py:248-276                   tensorf_render_depth
//   TensorfVmInterpFwd    tensor_vm.py:42-89                  tensorf_vm_pack + tensorf_vm_interp_fwd
//   TensorfVmInterpBwd    its reverse                         tensorf_vm_interp_bwd + tensorf_vm_unpack
//   TensorfAdamStep       training.py:158-243                 tensorf_adam_step
//
// NOT built in this image: the XLA FFI headers ship with jaxlib (`jax.ffi.include_dir()`) and JAX is not installable
// here; tests/test_jax_ffi_syntax.py type-checks this file against a minimal stand-in for xla/ffi/api/ffi.h.  Build
// where JAX is available:
//
//   g++ -O2 -fPIC -shared -std=c++17 -I"$(python -c 'import jax; print(jax.ffi.include_dir())')" \
//       -I../../include -I/usr/local/cuda/include xla_ffi_shim.cc -L../tensorf_b200 -ltensorf_b200 -lcudart \
//       -o libtensorf_b200_xla.so
//
// Contract: XLA owns every buffer; scratch / workspace memory is declared as an extra result so a
// handler never allocates; work is only enqueued on the stream XLA passes in; errors come back as ffi::Error
// (-> XlaRuntimeError), nothing aborts.
#include <cstdint>

#include <cuda_runtime_api.h>

#include "tensorf_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error Status(int rc) {
  if (rc == 0) return ffi::Error::Success();
  return ffi::Error(rc == TENSORF_ERR_INVALID_ARGUMENT ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    tensorf_last_error());
}
ffi::Error CudaStatus(cudaError_t e) {
  if (e == cudaSuccess) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, cudaGetErrorString(e));
}

tensorf_render_desc MakeDesc(int64_t R, int64_t N, int64_t K, int64_t G, int64_t cd, int64_t ca, int64_t mode,
                             int64_t contracted, int64_t feat_freqs, int64_t view_freqs, int64_t num_cameras,
                             int64_t inference) {
  tensorf_render_desc d{};
  d.R = (int32_t)R; d.N = (int32_t)N; d.K = (int32_t)K; d.G = (int32_t)G; d.cd = (int32_t)cd; d.ca = (int32_t)ca;
  d.mode = (int32_t)mode; d.contracted = (int32_t)contracted; d.squash = 27; d.units = 128;
  d.feat_freqs = (int32_t)feat_freqs; d.view_freqs = (int32_t)view_freqs; d.num_cameras = (int32_t)num_cameras;
  d.mlp_impl = TENSORF_MLP_AUTO; d.loss_scale = 0.f; d.flags = inference ? TENSORF_FLAG_INFERENCE : 0;
  return d;
}

using F32 = ffi::Buffer<ffi::F32>;
using U32 = ffi::Buffer<ffi::U32>;
using U8 = ffi::Buffer<ffi::U8>;

// ---- render_rays, RGB mode (render.py:105-279) ------------------------------------------------------------------
// Arguments: the MLP leaves and factors in the flatten order of LearnableParams, then the ray batch
// and the noise arrays JAX drew from the key.  Results: rgb (R,3) and the workspace (tensorf_render_workspace_bytes),
// which the reverse handler takes back as an argument (the residual of the custom_vjp).
ffi::Error RenderRgbFwd(cudaStream_t stream, F32 w0, F32 b1, F32 w1, F32 b2, F32 w2, F32 b3, F32 w3, F32 embed,
                        F32 app_vec, F32 app_mat, F32 den_vec, F32 den_mat, F32 aabb, F32 origins, F32 directions,
                        U32 cams, F32 jitter, F32 gumbel, F32 base_ts, F32 deltas, ffi::Result<F32> rgb,
                        ffi::Result<U8> workspace, int64_t N, int64_t K, int64_t contracted, int64_t feat_freqs,
                        int64_t view_freqs, int64_t num_cameras, int64_t inference) {
  const auto gd = den_mat.dimensions();  // (3, cd, G, G)
  const auto ad = app_mat.dimensions();
  tensorf_render_desc d = MakeDesc(origins.dimensions()[0], N, K, gd[2], gd[1], ad[1], TENSORF_MODE_RGB, contracted,
                                   feat_freqs, view_freqs, num_cameras, inference);
  tensorf_params p{den_vec.typed_data(), den_mat.typed_data(), app_vec.typed_data(), app_mat.typed_data(),
                   w0.typed_data(), w1.typed_data(), b1.typed_data(), w2.typed_data(), b2.typed_data(),
                   w3.typed_data(), b3.typed_data(), num_cameras ? embed.typed_data() : nullptr};
  tensorf_render_inputs in{origins.typed_data(), directions.typed_data(), cams.typed_data(), aabb.typed_data(),
                           jitter.typed_data(), gumbel.typed_data(), contracted ? base_ts.typed_data() : nullptr,
                           contracted ? deltas.typed_data() : nullptr, nullptr};
  return Status(tensorf_render_rgb_fwd(stream, &d, &p, &in, workspace->typed_data(), rgb->typed_data(), nullptr));
}

// Reverse of the above w.r.t. every leaf (training.py:153-156).  `workspace` is the forward's result: XLA passes it
// as an operand; the kernels only read their residuals from it and use the rest as scratch, so the handler declares
// it as an aliased result (input_output_aliases in tensorf_jax.py) and writes through `ws_out`.
ffi::Error RenderRgbBwd(cudaStream_t stream, F32 w0, F32 b1, F32 w1, F32 b2, F32 w2, F32 b3, F32 w3, F32 embed,
                        F32 app_vec, F32 app_mat, F32 den_vec, F32 den_mat, F32 aabb, F32 origins, F32 directions,
                        U32 cams, F32 jitter, F32 gumbel, F32 base_ts, F32 deltas, U8 workspace, F32 d_rgb,
                        ffi::Result<F32> g_w0, ffi::Result<F32> g_b1, ffi::Result<F32> g_w1, ffi::Result<F32> g_b2,
                        ffi::Result<F32> g_w2, ffi::Result<F32> g_b3, ffi::Result<F32> g_w3, ffi::Result<F32> g_embed,
                        ffi::Result<F32> g_app_vec, ffi::Result<F32> g_app_mat, ffi::Result<F32> g_den_vec,
                        ffi::Result<F32> g_den_mat, ffi::Result<U8> ws_out, int64_t N, int64_t K, int64_t contracted,
                        int64_t feat_freqs, int64_t view_freqs, int64_t num_cameras) {
  (void)workspace;  // same memory as ws_out
  const auto gd = den_mat.dimensions();
  const auto ad = app_mat.dimensions();
  tensorf_render_desc d = MakeDesc(origins.dimensions()[0], N, K, gd[2], gd[1], ad[1], TENSORF_MODE_RGB, contracted,
                                   feat_freqs, view_freqs, num_cameras, 0);
  tensorf_params p{den_vec.typed_data(), den_mat.typed_data(), app_vec.typed_data(), app_mat.typed_data(),
                   w0.typed_data(), w1.typed_data(), b1.typed_data(), w2.typed_data(), b2.typed_data(),
                   w3.typed_data(), b3.typed_data(), num_cameras ? embed.typed_data() : nullptr};
  tensorf_params g{g_den_vec->typed_data(), g_den_mat->typed_data(), g_app_vec->typed_data(), g_app_mat->typed_data(),
                   g_w0->typed_data(), g_w1->typed_data(), g_b1->typed_data(), g_w2->typed_data(), g_b2->typed_data(),
                   g_w3->typed_data(), g_b3->typed_data(), num_cameras ? g_embed->typed_data() : nullptr};
  tensorf_render_inputs in{origins.typed_data(), directions.typed_data(), cams.typed_data(), aabb.typed_data(),
                           jitter.typed_data(), gumbel.typed_data(), contracted ? base_ts.typed_data() : nullptr,
                           contracted ? deltas.typed_data() : nullptr, nullptr};
  if (!num_cameras) {  // the placeholder embedding's gradient is a defined zero
    ffi::Error e = CudaStatus(cudaMemsetAsync(g_embed->typed_data(), 0, g_embed->size_bytes(), stream));
    if (!e.success()) return e;
  }
  return Status(tensorf_render_rgb_bwd(stream, &d, &p, &in, ws_out->typed_data(), d_rgb.typed_data(), &g));
}

// ---- render_rays, DIST_MEDIAN / DIST_MEAN (render.py:248-276): only the density factors are read ------------------
ffi::Error RenderDepth(cudaStream_t stream, F32 den_vec, F32 den_mat, F32 aabb, F32 origins, F32 directions, F32 jitter,
                       F32 base_ts, F32 deltas, ffi::Result<F32> depth, ffi::Result<U8> workspace, int64_t N, int64_t mode,
                       int64_t contracted) {
  const auto gd = den_mat.dimensions();
  tensorf_render_desc d = MakeDesc(origins.dimensions()[0], N, 1, gd[2], gd[1], gd[1], mode, contracted, 0, 0, 0, 1);
  tensorf_params p{};
  p.density_vector = den_vec.typed_data();
  p.density_matrix = den_mat.typed_data();
  tensorf_render_inputs in{origins.typed_data(), directions.typed_data(), nullptr, aabb.typed_data(), jitter.typed_data(),
                           nullptr, contracted ? base_ts.typed_data() : nullptr, contracted ? deltas.typed_data() : nullptr,
                           nullptr};
  return Status(tensorf_render_depth(stream, &d, &p, &in, workspace->typed_data(), depth->typed_data()));
}

// ---- TensorVM.interpolate (tensor_vm.py:42-89): ijk (3,B) -> (3C,B) -------------------------------------------------
// `packed` (tensorf_vm_packed_floats) is the kernel-native copy of the factors; it is a result so that the reverse
// handler can take it back instead of packing again.
ffi::Error VmInterpFwd(cudaStream_t stream, F32 vector, F32 matrix, F32 ijk, ffi::Result<F32> out, ffi::Result<F32> packed) {
  const auto vd = vector.dimensions();  // (3, C, G)
  const int C = (int)vd[1], G = (int)vd[2];
  const int64_t B = (int64_t)ijk.element_count() / 3;
  int rc = tensorf_vm_pack(stream, vector.typed_data(), matrix.typed_data(), packed->typed_data(), C, G);
  if (rc != 0) return Status(rc);
  return Status(tensorf_vm_interp_fwd(stream, packed->typed_data(), ijk.typed_data(), out->typed_data(), C, G, B, 0));
}
ffi::Error VmInterpBwd(cudaStream_t stream, F32 packed, F32 ijk, F32 d_out, ffi::Result<F32> d_vector, ffi::Result<F32> d_matrix,
                       ffi::Result<F32> d_packed) {
  const auto vd = d_vector->dimensions();
  const int C = (int)vd[1], G = (int)vd[2];
  const int64_t B = (int64_t)ijk.element_count() / 3;
  ffi::Error e = CudaStatus(cudaMemsetAsync(d_packed->typed_data(), 0, d_packed->size_bytes(), stream));
  if (!e.success()) return e;
  int rc = tensorf_vm_interp_bwd(stream, packed.typed_data(), ijk.typed_data(), d_out.typed_data(), d_packed->typed_data(), C, G, B, 0);
  if (rc != 0) return Status(rc);
  return Status(tensorf_vm_unpack(stream, d_packed->typed_data(), d_vector->typed_data(), d_matrix->typed_data(), C, G));
}

// ---- optimiser step (training.py:158-243) ---------------------------------------------------------------------------
// Operands: n parameter leaves, n gradients, n first moments, n second moments.  Results: n new parameters, n new
// first moments, n new second moments (each aliased to its operand: the step is in place, training.py:101 donates the
// state), grad_norm (scalar) and the scratch buffer (tensorf_adam_scratch_bytes).  neg_lrs: -learning rate per leaf.
ffi::Error AdamStep(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, ffi::Span<const float> neg_lrs, float b1,
                    float b2, float eps, float eps_root, float bias_correction1, float bias_correction2, float lr_decay) {
  const size_t n = neg_lrs.size();
  if (n == 0 || n > TENSORF_ADAM_MAX_LEAVES || args.size() != 4 * n || rets.size() != 3 * n + 2)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "TensorfAdamStep: expected 4n operands and 3n+2 results for n leaves");
  int64_t sizes[TENSORF_ADAM_MAX_LEAVES];
  float *params[TENSORF_ADAM_MAX_LEAVES], *mu[TENSORF_ADAM_MAX_LEAVES], *nu[TENSORF_ADAM_MAX_LEAVES];
  const float* grads[TENSORF_ADAM_MAX_LEAVES];
  float lrs[TENSORF_ADAM_MAX_LEAVES];
  for (size_t i = 0; i < n; ++i) {
    auto g = args.get<F32>(n + i);
    auto p = rets.get<F32>(i);
    auto m = rets.get<F32>(n + i);
    auto v = rets.get<F32>(2 * n + i);
    if (!g.has_value() || !p.has_value() || !m.has_value() || !v.has_value())
      return ffi::Error(ffi::ErrorCode::kInvalidArgument, "TensorfAdamStep: every leaf must be an f32 buffer");
    sizes[i] = (int64_t)(*p)->element_count();
    params[i] = (*p)->typed_data();
    mu[i] = (*m)->typed_data();
    nu[i] = (*v)->typed_data();
    grads[i] = g->typed_data();
    lrs[i] = neg_lrs[i];
  }
  auto norm = rets.get<F32>(3 * n);
  auto scratch = rets.get<U8>(3 * n + 1);
  if (!norm.has_value() || !scratch.has_value()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "TensorfAdamStep: bad results");
  tensorf_adam_desc d{};
  d.n_leaves = (int32_t)n;
  d.b1 = b1; d.b2 = b2; d.eps = eps; d.eps_root = eps_root;
  d.bias_correction1 = bias_correction1; d.bias_correction2 = bias_correction2; d.lr_decay = lr_decay;
  return Status(tensorf_adam_step(stream, &d, sizes, params, grads, mu, nu, lrs, (*norm)->typed_data(), (*scratch)->typed_data(),
                                  (int64_t)(*scratch)->size_bytes()));
}

}  // namespace

#define TENSORF_LEAF_ARGS                                                                            \
  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>() /* MLP */ \
  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()                                             /* factors */
#define TENSORF_RAY_ARGS .Arg<F32>().Arg<F32>().Arg<F32>().Arg<U32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>() /* aabb, rays, noise */

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    TensorfRenderRgbFwd, RenderRgbFwd,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>() TENSORF_LEAF_ARGS TENSORF_RAY_ARGS
        .Ret<F32>().Ret<U8>()
        .Attr<int64_t>("density_samples_per_ray").Attr<int64_t>("appearance_samples_per_ray")
        .Attr<int64_t>("scene_contraction").Attr<int64_t>("feature_n_freqs").Attr<int64_t>("viewdir_n_freqs")
        .Attr<int64_t>("num_cameras").Attr<int64_t>("inference"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    TensorfRenderRgbBwd, RenderRgbBwd,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>() TENSORF_LEAF_ARGS TENSORF_RAY_ARGS
        .Arg<U8>().Arg<F32>()  // workspace, d_rgb
        .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()
        .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<U8>()
        .Attr<int64_t>("density_samples_per_ray").Attr<int64_t>("appearance_samples_per_ray")
        .Attr<int64_t>("scene_contraction").Attr<int64_t>("feature_n_freqs").Attr<int64_t>("viewdir_n_freqs")
        .Attr<int64_t>("num_cameras"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    TensorfRenderDepth, RenderDepth,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
        .Ret<F32>().Ret<U8>()
        .Attr<int64_t>("density_samples_per_ray").Attr<int64_t>("mode").Attr<int64_t>("scene_contraction"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(TensorfVmInterpFwd, VmInterpFwd,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    TensorfVmInterpBwd, VmInterpBwd,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F32>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    TensorfAdamStep, AdamStep,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .RemainingArgs().RemainingRets()
        .Attr<ffi::Span<const float>>("neg_lrs").Attr<float>("b1").Attr<float>("b2").Attr<float>("eps").Attr<float>("eps_root")
        .Attr<float>("bias_correction1").Attr<float>("bias_correction2").Attr<float>("lr_decay"));
