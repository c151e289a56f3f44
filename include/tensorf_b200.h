/*
 * tensorf_b200.h — C ABI of the B200-native tensorf-jax hot path (libtensorf_b200.so).
 *
 * The reference (brentyi/tensorf-jax) is pure Python/JAX and has no FFI of its own; the
 * entry points below are what a `jax.ffi` / XLA custom-call binding for its per-ray hot path
 * binds (see INTEGRATION.md for the reference-side stub).  Each entry point cites the
 * reference interface (file:line under /root/reference) it replaces.
 *
 * Conventions
 *  - Plain pointers and sizes only.  All pointers are DEVICE pointers unless stated; all
 *    floating point is fp32, indices int32, camera ids uint32; arrays are dense row-major in
 *    the reference's own layouts (factors channel-first: vector (3,C,G), matrix (3,C,G,G),
 *    tensor_vm.py:129-138; flax Dense kernels (in,out), networks.py:57-117).
 *  - The caller (XLA) owns every buffer.  Entry points never allocate, free or retain device
 *    memory; scratch is a caller-provided workspace whose size `tensorf_render_workspace_bytes`
 *    reports.  Result and workspace buffers may arrive uninitialised: gradient accumulators
 *    are zeroed on-stream by the entry point.
 *  - Work is only enqueued on the given stream (no device synchronisation, no host-blocking
 *    waits), so calls are CUDA-graph capturable and re-entrant across devices/threads.
 *  - Return value: 0 on success, negative `tensorf_status` on failure with a thread-local
 *    message available from `tensorf_last_error()`.  Nothing aborts.
 *  - Randomness stays in the host framework: the shared jitter/Gumbel vectors that
 *    render.py:158-161, :375-379 and :461-469 draw with jax.random enter as input arrays.
 */
#ifndef TENSORF_B200_H_
#define TENSORF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tensorf_stream_t; /* cudaStream_t */

enum tensorf_status {
  TENSORF_OK = 0,
  TENSORF_ERR_INVALID_ARGUMENT = -1,
  TENSORF_ERR_CUDA = -2,
  TENSORF_ERR_UNSUPPORTED = -3
};

/* render.py:18-23 (RenderMode) */
enum tensorf_render_mode { TENSORF_MODE_RGB = 0, TENSORF_MODE_DIST_MEDIAN = 1, TENSORF_MODE_DIST_MEAN = 2 };

/* TENSORF_FLAG_INFERENCE: forward only (render_360.py, render_rays_batched): residuals that only the reverse pass
 * reads (ReLU masks, second hidden layer) are not written; tensorf_render_rgb_bwd / tensorf_mlp_bwd then fail. */
/* TENSORF_FLAG_PACKED_FACTORS: the factor leaves are ALREADY in the kernel-native texel-major layout (tensorf_vm_pack;
 * tensorf_vm_packed_floats floats per TensorVM: lines [3][G][Cp] then planes [3][G][G][Cp]): `*_vector` points to that
 * buffer and `*_matrix` is NULL or `*_vector + 3*G*Cp`.  The same holds for the gradient struct, whose packed buffers are
 * zeroed and accumulated in place.  A training loop that keeps parameters, gradients and Adam moments packed (every
 * optimiser operation is elementwise) skips the pack and unpack passes of each step: 2 x 12.7 MB at 128^3, 2 x 69 MB at
 * 300^3; tensorf_vm_unpack converts back at the boundary (checkpoints, grid resampling). */
enum tensorf_render_flags { TENSORF_FLAG_INFERENCE = 1, TENSORF_FLAG_PACKED_FACTORS = 2 };

/* MLP arithmetic: exact fp32 on CUDA cores; tcgen05 tensor cores with split-bf16 operands, one kernel per layer;
 * or the per-row-tile fused tcgen05 kernels (split-fp16 operands, activations stay in tensor memory between the
 * layers; networks with feature_squash_dim 27, units 128, 2 + 2 frequencies and no camera embeddings only -
 * TENSORF_ERR_UNSUPPORTED otherwise).  AUTO picks the fused kernels when the network qualifies. */
enum tensorf_mlp_impl { TENSORF_MLP_AUTO = 0, TENSORF_MLP_SIMT_FP32 = 1, TENSORF_MLP_TCGEN05 = 2, TENSORF_MLP_FUSED = 3 };

/* Static configuration of one render call: render.py:26-36 (RenderConfig), the static fields
 * of networks.py:38-43 (FeatureMlp) and the array shapes render.py:105-113 receives. */
typedef struct tensorf_render_desc {
  int32_t R;           /* rays in this call */
  int32_t N;           /* density_samples_per_ray */
  int32_t K;           /* appearance_samples_per_ray, 1 <= K <= N */
  int32_t G;           /* grid_dim */
  int32_t cd;          /* density per-axis channels (density feature dim = 3*cd) */
  int32_t ca;          /* appearance per-axis channels */
  int32_t mode;        /* tensorf_render_mode */
  int32_t contracted;  /* LearnableParams.scene_contraction */
  int32_t squash;      /* feature_squash_dim (27) */
  int32_t units;       /* 128 */
  int32_t feat_freqs;  /* feature_n_freqs */
  int32_t view_freqs;  /* viewdir_n_freqs */
  int32_t num_cameras; /* 0 = no camera embeddings */
  int32_t mlp_impl;    /* tensorf_mlp_impl */
  float loss_scale;    /* 1/(3*R_global): training.py:140 is a mean over the WHOLE batch */
  int32_t flags;       /* tensorf_render_flags */
} tensorf_render_desc;

/* Leaves of render.py:39-46 (LearnableParams), reference layouts. `embed` may be NULL when
 * num_cameras == 0.  For gradients the same struct is used with writable buffers. */
typedef struct tensorf_params {
  float* density_vector;    /* (3, cd, G) */
  float* density_matrix;    /* (3, cd, G, G) */
  float* appearance_vector; /* (3, ca, G) */
  float* appearance_matrix; /* (3, ca, G, G) */
  float* w0;                /* Dense_0.kernel (3*ca, squash), no bias */
  float* w1;                /* Dense_1.kernel (enc, units) */
  float* b1;                /* Dense_1.bias (units) */
  float* w2;                /* Dense_2.kernel (units, units) */
  float* b2;                /* Dense_2.bias (units) */
  float* w3;                /* Dense_3.kernel (units, 3) */
  float* b3;                /* Dense_3.bias (3) */
  float* embed;             /* Embed_0.embedding (num_cameras, units) or NULL */
} tensorf_params;

/* cameras.py:10-20 (Rays3D) + the arrays render_rays derives from its key and config. */
typedef struct tensorf_render_inputs {
  const float* origins;           /* (R,3) */
  const float* directions;        /* (R,3) */
  const uint32_t* camera_indices; /* (R,) ; may be NULL when num_cameras == 0 */
  const float* aabb;              /* (2,3) */
  const float* jitter;            /* bounded: (N,) shared by all rays; contracted: (R,N) */
  const float* gumbel;            /* (N,) shared by all rays; RGB mode only */
  const float* base_ts;           /* contracted only: (N,) constant schedule, render.py:130-151 */
  const float* deltas;            /* contracted only: (N,) step sizes, render.py:154-155 */
  const float* colors;            /* (R,3) or NULL; when given the MSE loss is fused */
} tensorf_render_inputs;

const char* tensorf_last_error(void);
int tensorf_version(void);

/* ---- measurement hooks (bench.py; off by default) ---------------------------------------- */
/* Kernels + memsets this thread has enqueued through the library so far. */
int64_t tensorf_launch_count(void);
/* When enabled, every entry point brackets its stages (pack, density_select, mlp_fwd, ...) with
 * cudaEventRecord on the launch stream. Not CUDA-graph-capture safe while enabled. */
int tensorf_profile_enable(int enable);
/* HOST buffers: names[max_entries][32], total_ms[max_entries], calls[max_entries]. Blocks until
 * the recorded events have completed, returns per-stage totals since the last read and resets. */
int tensorf_profile_read(int max_entries, char* names, float* total_ms, int* calls, int* count);

/* ---- factor layout (kernel-native texel-major copy of the channel-first factors) -------- */
/* Number of floats of the packed copy of one TensorVM: [3][G][Cp] lines then [3][G][G][Cp]
 * planes, Cp = C rounded up to a multiple of 4. */
int64_t tensorf_vm_packed_floats(int C, int G);
/* tensor_vm.py:129-138 layout -> packed. */
int tensorf_vm_pack(tensorf_stream_t s, const float* vector, const float* matrix, float* packed, int C, int G);
/* packed gradient -> (3,C,G) / (3,C,G,G) gradient buffers (overwrites). */
int tensorf_vm_unpack(tensorf_stream_t s, const float* packed, float* vector, float* matrix, int C, int G);

/* ---- tensor_vm.py:42-89 TensorVM.interpolate ---------------------------------------------- */
/* ijk (3,B) in [-1,1] -> out (3C,B) (feature_major=0, the reference's layout) or (B,3C)
 * (feature_major=1, the (M,3ca) row layout render.py:484-486 builds). */
int tensorf_vm_interp_fwd(tensorf_stream_t s, const float* packed, const float* ijk, float* out, int C, int G, int64_t B,
                          int feature_major);
/* Reverse mode w.r.t. the factors: d_out has out's layout; d_packed (+=) must be zeroed by
 * the caller (or hold a running sum). */
int tensorf_vm_interp_bwd(tensorf_stream_t s, const float* packed, const float* ijk, const float* d_out, float* d_packed,
                          int C, int G, int64_t B, int feature_major);

/* ---- render.py:461-469 selection stage (test entry point) -------------------------------- */
/* idx (R,K) = indices of the K largest g per row, ties -> lower index, emitted in ASCENDING
 * index order (the order is not observable through render_rays). g (R,N). */
int tensorf_topk_select(tensorf_stream_t s, const float* g, int R, int N, int K, int32_t* idx);

/* ---- render.py:300-347 compute_segment_probabilities ---------------------------------------- */
/* sigmas, step_sizes (R,N) -> p_exits = exp(cumsum(-sigma*step)), p_terminates (R,N). */
int tensorf_segment_probabilities(tensorf_stream_t s, const float* sigmas, const float* step_sizes, int R, int N,
                                  float* p_exits, float* p_terminates);

/* ---- tensor-core GEMM building blocks of the MLP (test entry points) ------------------------ */
/* C (M,N) = [mask bit] relu?(A (M,K) @ W (K,N) + bias); mask_bits / bits_out: ceil16(N)/32 (rounded up)
 * uint32 words per row, bit n of row m (bits_out = C > 0); split-bf16 operands on tcgen05 (nsplit = 2:
 * hi+lo, 3 products, ~2^-16; nsplit = 3: hi+mid+lo, 6 products, fp32-exact operands), fp32
 * accumulation in TMEM. scratch: 6*ceil16(N)*ceil64(K)+ bytes (128 KiB * ceil(K/64) * ceil(N/256) suffices). */
int tensorf_tc_rowgemm_test(tensorf_stream_t s, const float* A, int64_t M, int K, const float* W, int N, const float* bias,
                            int relu, const uint32_t* mask_bits, uint32_t* bits_out, float* C, void* scratch,
                            int64_t scratch_bytes, int nsplit);
/* Debug: with TENSORF_TC_TRACE=1 CTA 0 of the row GEMM records clock64() per role/chunk; copies 8 x 1024
 * int64 slots to HOST memory (synchronises the device). */
int tensorf_tc_trace_read(long long* host, int n);
/* Microbenchmark: SM cycles for `count` back-to-back tcgen05.mma (M=128, K=16, bf16, given N) with the given
 * shared-memory descriptor layout type / LBO / SBO; result in out_dev[0] (device int64). */
int tensorf_tc_umma_bench(tensorf_stream_t s, int N, int layout_type, int lbo, int sbo, int count, long long* out_dev);
/* out (Nx, Mg) += X^T (Nx,rows) @ G (rows,Mg): the weight-gradient contraction over rows. Mg <= 128, Nx <= 512. */
int tensorf_tc_redgemm_test(tensorf_stream_t s, const float* G, int Mg, const float* X, int Nx, int64_t rows, float* out);
/* Test entry point of the fused MLP kernels' operand conventions (csrc/umma_tiles.cuh): D[128 x N] = A[128 x K] * B[N x K]^T
 * with two-term 16-bit split operands.  a_mode 0 = shared memory K-major, 1 = shared memory MN-major, 2 = tensor memory;
 * b_mode 0 = K-major, 1 = MN-major; a_fmt / b_fmt 0 = fp16, 1 = bf16.  N % 16 == 0 <= 128, K % 16 == 0 <= 160. */
int tensorf_tc_umma_probe(tensorf_stream_t s, const float* A, const float* B, float* D, int N, int K, int a_mode, int b_mode,
                          int a_fmt, int b_fmt);

/* ---- networks.py:46-121 FeatureMlp.__call__ ----------------------------------------------- */
/* features (M, 3*ca); viewdirs (M/rows_per_ray, 3); camera_indices (M/rows_per_ray); rgb (M,3).
 * workspace: tensorf_mlp_workspace_bytes. */
int64_t tensorf_mlp_workspace_bytes(const tensorf_render_desc* d, int64_t M);
/* Test entry point: float offsets into the MLP workspace of the saved activations, in the order f, df, x, dx, h1, h2,
 * dp2, dp1, relu-mask words of layer 1, of layer 2.  With TENSORF_MLP_FUSED x / h1 / h2 / dp2 / dp1 hold two-term
 * 16-bit slab tiles per 128 rows (csrc/umma_tiles.cuh) and f / df hold [tile][4-column group][row][4] floats. */
int tensorf_mlp_workspace_layout(const tensorf_render_desc* d, int64_t M, int64_t* offsets);
int tensorf_mlp_fwd(tensorf_stream_t s, const tensorf_render_desc* d, const tensorf_params* p, const float* features,
                    const float* viewdirs, const uint32_t* camera_indices, int64_t M, int rows_per_ray, void* workspace,
                    float* rgb);
/* rgb = the forward output, d_rgb (M,3) -> d_features (M,3*ca) and the MLP leaves of `grads`
 * (overwritten). Must follow tensorf_mlp_fwd on the same workspace. */
int tensorf_mlp_bwd(tensorf_stream_t s, const tensorf_render_desc* d, const tensorf_params* p, const float* features,
                    const float* viewdirs, const uint32_t* camera_indices, int64_t M, int rows_per_ray, void* workspace,
                    const float* rgb, const float* d_rgb, float* d_features, const tensorf_params* grads);

/* ---- render.py:105-279 render_rays --------------------------------------------------------- */
int tensorf_render_workspace_bytes(const tensorf_render_desc* d, int64_t* bytes);
/* mode RGB: rgb (R,3).  If inputs->colors != NULL, *loss (device scalar, overwritten) receives
 * loss_scale * sum((rgb-colors)^2) (training.py:140) and the loss cotangent is kept in the
 * workspace for tensorf_render_rgb_bwd(d_rgb = NULL). */
int tensorf_render_rgb_fwd(tensorf_stream_t s, const tensorf_render_desc* d, const tensorf_params* p,
                           const tensorf_render_inputs* in, void* workspace, float* rgb, float* loss);
/* Reverse mode of render_rays w.r.t. every leaf of LearnableParams (what
 * training.py:153-156 `jax.value_and_grad` asks for).  d_rgb (R,3) or NULL (= use the fused
 * loss cotangent).  All leaves of `grads` are overwritten.  Must follow the forward call on the
 * same workspace. */
int tensorf_render_rgb_bwd(tensorf_stream_t s, const tensorf_render_desc* d, const tensorf_params* p,
                           const tensorf_render_inputs* in, void* workspace, const float* d_rgb,
                           const tensorf_params* grads);
/* The same reverse pass in two halves, for sharded training: phase 1 = ray reverse, MLP reverse,
 * appearance scatter (afterwards every leaf of `grads` except the density factors is final, so their exchange can
 * start and overlap with) phase 2 = density scatter (density_vector / density_matrix of `grads`).  phase 0 = the whole
 * pass, identical to tensorf_render_rgb_bwd.  Phase 2 must follow phase 1 of the same step on the same workspace. */
int tensorf_render_rgb_bwd_phase(tensorf_stream_t s, const tensorf_render_desc* d, const tensorf_params* p,
                                 const tensorf_render_inputs* in, void* workspace, const float* d_rgb,
                                 const tensorf_params* grads, int phase);
/* modes DIST_MEDIAN / DIST_MEAN (render.py:248-276): depth (R,). Only the density factors of
 * `p` are read. */
int tensorf_render_depth(tensorf_stream_t s, const tensorf_render_desc* d, const tensorf_params* p,
                         const tensorf_render_inputs* in, void* workspace, float* depth);

/* Debug/test views into the workspace after tensorf_render_rgb_fwd (device pointers). */
int tensorf_render_workspace_view(const tensorf_render_desc* d, void* workspace, const char* name, void** ptr,
                                  int64_t* count);

/* ---- training.py:158-243 optimiser step ------------------------------------ */
/* optax.chain(scale_by_adam(b1, b2, eps, eps_root), masked(scale(-lr_mlp)), masked(scale(-lr_tensor)))
 * (training.py:213-243), the learning-rate decay factor of training.py:176-186 and
 * optax.apply_updates (training.py:199-201), over all leaves in one launch; optax.global_norm(grads)
 * (training.py:194) comes out of the same pass.  The step count stays on the host: the caller passes
 * the bias corrections 1 - b^t (t = count + 1) and lr_decay = exponential_decay(...)(resetted_step). */
#define TENSORF_ADAM_MAX_LEAVES 16
typedef struct tensorf_adam_desc {
  int32_t n_leaves;
  int32_t reserved;
  float b1, b2, eps, eps_root;
  float bias_correction1; /* 1 - b1^t */
  float bias_correction2; /* 1 - b2^t */
  float lr_decay;         /* lr_decay_coeff, training.py:176-181 */
  float reserved2;
} tensorf_adam_desc;
/* HOST arrays of length n_leaves: sizes (elements), params/grads/mu/nu (device pointers), neg_lrs (-lr of
 * the leaf's group).  params, mu, nu are updated in place (training_step donates its state, training.py:101).
 * grad_norm: device scalar (overwritten) or NULL.  scratch: tensorf_adam_scratch_bytes. */
int64_t tensorf_adam_scratch_bytes(const int64_t* sizes, int n_leaves);
int tensorf_adam_step(tensorf_stream_t s, const tensorf_adam_desc* d, const int64_t* sizes, float* const* params,
                      const float* const* grads, float* const* mu, float* const* nu, const float* neg_lrs,
                      float* grad_norm, void* scratch, int64_t scratch_bytes);

/* ---- SURVEY 8e fused follow-up: gradient reduce-scatter + Adam + parameter all-gather in ONE kernel over
 * peer memory (NVLink 5 / NVSwitch) ---------------------------------------------------------------------
 * Replaces `ncclAllReduce(grads)` followed by `tensorf_adam_step` on every rank (training.py:153-156 -> :158-243
 * when the ray batch is sharded).  All leaves live in one FLAT fp32 buffer per rank (leaf i =
 * elements [leaf_offsets[i], leaf_offsets[i+1]); elements from leaf_offsets[n_leaves] to `total` are padding),
 * allocated symmetrically so that every rank holds a peer mapping of every other rank's buffer.  Rank r owns the
 * element range [shard_begin, shard_end) (multiples of 4): it sums the `world` gradient copies of its range in
 * rank order (P2P loads, or one `multimem.ld_reduce` through the NVSwitch when grad_mc != NULL), applies the Adam
 * step of tensorf_adam_step with ITS shard of the moments (mu, nu: shard_end - shard_begin floats, local), and
 * stores the new parameters into every rank's parameter buffer (P2P stores, or one `multimem.st` when
 * param_mc != NULL).  The sum of squares of the reduced gradient over the shard goes to slot [rank] of every
 * rank's `norm_slots` (world floats); tensorf_peer_grad_norm turns the slots into optax.global_norm(grads).
 * The CALLER orders the ranks: a cross-rank barrier on the stream before this call (all gradients written) and
 * one after it (all parameter / slot stores landed) - torch symmetric-memory `barrier()`, or any equivalent.
 * With world == 1 (pointers to the rank's own buffers, no barrier) the result is bit-identical to
 * tensorf_adam_step. */
#define TENSORF_PEER_MAX_WORLD 16
#define TENSORF_PEER_MAX_LEAVES 32 /* leaves + alignment gaps (a gap is a leaf with neg_lr 0 whose gradient stays 0) */
typedef struct tensorf_peer_adam_desc {
  tensorf_adam_desc adam;     /* n_leaves = leaves of the flat buffer, <= TENSORF_PEER_MAX_LEAVES */
  int32_t rank, world;
  int64_t total;              /* floats in the flat buffers, multiple of 4 */
  int64_t shard_begin, shard_end; /* this rank's range, multiples of 4, within [0,total] */
} tensorf_peer_adam_desc;
/* Equal contiguous shards in units of 4 floats: [begin,end) of `rank`. */
void tensorf_peer_shard(int64_t total, int rank, int world, int64_t* begin, int64_t* end);
int64_t tensorf_peer_adam_scratch_bytes(int64_t shard_floats);
/* HOST arrays: leaf_offsets [n_leaves+1], neg_lrs [n_leaves], grad_peers / param_peers / norm_slot_peers [world]
 * (device pointers valid in THIS process: own buffer at [rank], peer mappings elsewhere). */
int tensorf_adam_step_peer(tensorf_stream_t s, const tensorf_peer_adam_desc* d, const int64_t* leaf_offsets,
                           const float* neg_lrs, const float* const* grad_peers, float* const* param_peers,
                           const float* grad_mc, float* param_mc, float* mu_shard, float* nu_shard,
                           float* const* norm_slot_peers, void* scratch, int64_t scratch_bytes);
/* The exchange alone (sum all-reduce, in place) over the same symmetric buffers, for callers that keep their own
 * optimiser: rank r sums elements tensorf_peer_shard(total, r, world) of the `world` buffers in rank order (or through
 * the multicast address `mc` when non-NULL) and stores the sums into every rank's buffer.  HOST array peers[world];
 * total a multiple of 4.  Same barrier contract as tensorf_adam_step_peer.  Runs on the caller's stream, so it
 * is ordered with the reverse pass without a second stream. */
int tensorf_peer_allreduce(tensorf_stream_t s, int rank, int world, int64_t total, float* const* peers, float* mc);
/* Upper bound on the CTAs of the exchange kernels launched from this thread from now on (0 = default, 4 per SM): an exchange
 * running on a side stream beside a compute kernel (the early bucket beside the density scatter) should trickle through a
 * few SMs instead of competing for issue slots on all of them. */
int tensorf_peer_set_max_ctas(int max_ctas);
/* The same exchange with the two cross-rank barriers INSIDE the kernel (no barrier launches around it):
 * signal_peers[world] (HOST array) = every rank's signal pad, 32 uint32 in symmetric memory, zeroed once before the
 * first call ([0,16) "buffer complete", [16,32) "stores landed", slot = signalling rank); local_flags = 4 uint32 of this
 * rank's own device memory, zeroed once; epoch = 1, 2, 3, ... the call number, identical on every rank, or 0 = the
 * kernel keeps the call number itself in local_flags[2] (then every call on every rank must pass 0).  Block 0
 * signals / awaits "complete" with st.release.sys / ld.acquire.sys and opens a gate for the grid; the last block to
 * finish its stores signals / awaits "landed", so stream completion of the kernel means every buffer holds every sum.
 * Every rank must make the call (even with an empty shard).  With epoch = 0 the launch has no per-call argument and is
 * CUDA-graph capturable (every replay advances the device-side counter). */
int tensorf_peer_allreduce_sync(tensorf_stream_t s, int rank, int world, int64_t total, float* const* peers, float* mc,
                                uint32_t* const* signal_peers, uint32_t* local_flags, uint32_t epoch);
/* grad_norm[0] = sqrt(sum of the `world` slots, in rank order, fp64 accumulation): identical on every rank. */
int tensorf_peer_grad_norm(tensorf_stream_t s, const float* norm_slots, int world, float* grad_norm);

/* ---- tensor_vm.py:183-223 TensorVMSingle.resize ------------------------------ */
/* vector (3,C,G_in) -> (3,C,G_out), matrix (3,C,G_in,G_in) -> (3,C,G_out,G_out): jax.image.scale_and_translate,
 * "linear" kernel, scale (G_out-1)/(G_in-1), translation -(scale/2 - 0.5) (align corners), antialiased when
 * shrinking.  Applied to parameters and to both Adam moments by training.py:245-276. */
int64_t tensorf_vm_resize_scratch_bytes(int C, int G_in, int G_out);
int tensorf_vm_resize(tensorf_stream_t s, const float* vector_in, const float* matrix_in, int C, int G_in, int G_out,
                      float* vector_out, float* matrix_out, void* scratch, int64_t scratch_bytes);

/* ---- jax.random on the device + pixel rays ------------------------------------ */
/* Threefry-2x32 (20 rounds), HOST function: out2 = cipher(key (k0,k1), counter (x0,x1)); the Random123
 * known-answer vectors check it (tests/test_prng.py). */
void tensorf_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* out2);
/* jax.random.uniform(key, (n,), float32, minval, maxval) / jax.random.gumbel(key, (n,)) for a raw threefry key
 * (render.py:158-161, :375-379, :462-468): element i uses counter (hi32(i), lo32(i)), bits = out0 ^ out1
 * (jax_threefry_partitionable layout), so any (R,N) array is the flat array reshaped. */
int tensorf_prng_uniform(tensorf_stream_t s, uint32_t k0, uint32_t k1, int64_t n, float minval, float maxval, float* out);
int tensorf_prng_gumbel(tensorf_stream_t s, uint32_t k0, uint32_t k1, int64_t n, float* out);
/* Elements [first, first + n) of the same uniform draw: a rank that holds rows [a, b) of a sharded (R,N) jitter
 * (render.py:158-161 with the ray batch split over ranks) draws first = a*N, n = (b-a)*N and gets exactly its slice of
 * what the single-device reference draws for the whole batch. */
int tensorf_prng_uniform_slice(tensorf_stream_t s, uint32_t k0, uint32_t k1, int64_t first, int64_t n, float minval, float maxval,
                               float* out);
/* cameras.py:100-143 pixel_rays_wrt_world for image rows [row0,row1): M (HOST, 3x3 row-major) =
 * R_world_camera @ K^-1, origin (HOST, 3) = T_world_camera.translation(); direction = M @ [u,v,1] / (norm + 1e-8).
 * origins, directions ((row1-row0)*W, 3); camera_indices ((row1-row0)*W) or NULL. */
int tensorf_pixel_rays(tensorf_stream_t s, const float* M, const float* origin, int W, int row0, int row1,
                       uint32_t camera_index, float* origins, float* directions, uint32_t* camera_indices);

/* The same for the rows a rank owns when a frame is dealt to `world` ranks in interleaved stripes of `stripe` rows (stripe k
 * = rows [k*stripe, (k+1)*stripe) belongs to rank k % world; render_360-style frames sharded with no collective): the
 * rank's rows are written contiguously, in image order, by ONE launch.  *n_rays (HOST, may be NULL) receives rows * W; with
 * origins == directions == NULL the call only reports it. */
int tensorf_pixel_rays_striped(tensorf_stream_t s, const float* M, const float* origin, int W, int H, int stripe, int rank, int world,
                               uint32_t camera_index, float* origins, float* directions, uint32_t* camera_indices, int64_t* n_rays);

/* ---- training data path -------------------------------------------------------- */
/* data.py:301-337 keeps every training ray in memory; training.py:318-323 draws shuffled minibatches.  Here the
 * table (origins, directions (n_table,3); camera_indices (n_table) or NULL; colors (n_table,3) or NULL) lives on
 * the device and a minibatch is a gather by idx (R,) int64.  Indices outside [0,n_table) read row 0 and are
 * counted in *bad_count (device int, overwritten; may be NULL). */
int tensorf_gather_rays(tensorf_stream_t s, const float* origins, const float* directions, const uint32_t* camera_indices,
                        const float* colors, int64_t n_table, const int64_t* idx, int64_t R, float* out_origins,
                        float* out_directions, uint32_t* out_camera_indices, float* out_colors, int* bad_count);
/* data.py:318-320: rgb = rgba[:3]*a + (1-a) (opaque white background); rgba (n,4) 16-byte aligned -> rgb (n,3). */
int tensorf_rgba_over_white(tensorf_stream_t s, const float* rgba, int64_t n, float* rgb);

#ifdef __cplusplus
}
#endif
#endif /* TENSORF_B200_H_ */
