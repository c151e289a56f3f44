// XLA typed-FFI handlers over the C ABI of libtensorf_b200.so (include/tensorf_b200.h).
//
// NOT compiled in this image: the XLA FFI headers ship with jaxlib (`jax.ffi.include_dir()`),
// and JAX is not installable here.  Build where JAX is available:
//
//   g++ -O2 -fPIC -shared -std=c++17 -I"$(python -c 'import jax; print(jax.ffi.include_dir())')" \
//       -I../../include xla_ffi_shim.cc -L../tensorf_b200 -ltensorf_b200 -o libtensorf_b200_xla.so
//
// Contract: XLA owns every buffer; scratch is declared as an extra result so the
// handler never allocates; work is only enqueued on the stream XLA passes in; errors come back
// as ffi::Error (-> XlaRuntimeError), nothing aborts.
#include <cstdint>

#include "tensorf_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error Status(int rc) {
  if (rc == 0) return ffi::Error::Success();
  return ffi::Error(rc == TENSORF_ERR_INVALID_ARGUMENT ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    tensorf_last_error());
}

tensorf_render_desc MakeDesc(int64_t R, int64_t N, int64_t K, int64_t G, int64_t cd, int64_t ca, int64_t mode,
                             int64_t contracted, int64_t feat_freqs, int64_t view_freqs, int64_t num_cameras,
                             float loss_scale) {
  tensorf_render_desc d{};
  d.R = (int32_t)R; d.N = (int32_t)N; d.K = (int32_t)K; d.G = (int32_t)G; d.cd = (int32_t)cd; d.ca = (int32_t)ca;
  d.mode = (int32_t)mode; d.contracted = (int32_t)contracted; d.squash = 27; d.units = 128;
  d.feat_freqs = (int32_t)feat_freqs; d.view_freqs = (int32_t)view_freqs; d.num_cameras = (int32_t)num_cameras;
  d.mlp_impl = TENSORF_MLP_AUTO; d.loss_scale = loss_scale;
  return d;
}

using F32 = ffi::Buffer<ffi::F32>;
using U32 = ffi::Buffer<ffi::U32>;
using U8 = ffi::Buffer<ffi::U8>;

// render_rays forward (render.py:105-279, RGB mode). Arguments in the flatten order of
// LearnableParams followed by the ray batch and the host-drawn noise vectors.
ffi::Error RenderRgbFwd(cudaStream_t stream, F32 w0, F32 b1, F32 w1, F32 b2, F32 w2, F32 b3, F32 w3, F32 embed,
                        F32 app_vec, F32 app_mat, F32 den_vec, F32 den_mat, F32 aabb, F32 origins, F32 directions,
                        U32 cams, F32 jitter, F32 gumbel, F32 base_ts, F32 deltas, ffi::Result<F32> rgb,
                        ffi::Result<U8> workspace, int64_t N, int64_t K, int64_t contracted, int64_t feat_freqs,
                        int64_t view_freqs, int64_t num_cameras) {
  const auto gd = den_mat.dimensions();  // (3, cd, G, G)
  const auto ad = app_mat.dimensions();
  tensorf_render_desc d = MakeDesc(origins.dimensions()[0], N, K, gd[2], gd[1], ad[1], TENSORF_MODE_RGB, contracted,
                                   feat_freqs, view_freqs, num_cameras, 0.f);
  tensorf_params p{den_vec.typed_data(), den_mat.typed_data(), app_vec.typed_data(), app_mat.typed_data(),
                   w0.typed_data(), w1.typed_data(), b1.typed_data(), w2.typed_data(), b2.typed_data(),
                   w3.typed_data(), b3.typed_data(), num_cameras ? embed.typed_data() : nullptr};
  tensorf_render_inputs in{origins.typed_data(), directions.typed_data(), cams.typed_data(), aabb.typed_data(),
                           jitter.typed_data(), gumbel.typed_data(), contracted ? base_ts.typed_data() : nullptr,
                           contracted ? deltas.typed_data() : nullptr, nullptr};
  return Status(tensorf_render_rgb_fwd(stream, &d, &p, &in, workspace->typed_data(), rgb->typed_data(), nullptr));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    TensorfRenderRgbFwd, RenderRgbFwd,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()   // MLP leaves
        .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()                                               // factors
        .Arg<F32>().Arg<F32>().Arg<F32>().Arg<U32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()   // aabb, rays, noise
        .Ret<F32>().Ret<U8>()
        .Attr<int64_t>("density_samples_per_ray").Attr<int64_t>("appearance_samples_per_ray")
        .Attr<int64_t>("scene_contraction").Attr<int64_t>("feature_n_freqs").Attr<int64_t>("viewdir_n_freqs")
        .Attr<int64_t>("num_cameras"));
// The reverse handler (tensorf_render_rgb_bwd: workspace + d_rgb in, one gradient buffer per leaf
// out), the depth handler (tensorf_render_depth) and TensorVM.interpolate (tensorf_vm_pack +
// tensorf_vm_interp_fwd/bwd) follow the same pattern, one handler per C-ABI entry point.
