// Optimiser step and grid resampling: the two callers right after the reverse pass
// (SURVEY §8f rows 1-2).
//   - k_adam: optax.scale_by_adam(b1,b2,eps,eps_root) -> masked scale(-lr) -> lr_decay -> apply_updates
//     (training.py:176-201, :213-243) over every leaf of LearnableParams in ONE launch, with the
//     grads' global L2 norm (training.py:194, optax.global_norm) reduced in the same pass.
//     HBM-bound: 16 B read + 12 B written per parameter.
//   - k_resize_*: TensorVMSingle.resize (tensor_vm.py:183-223): jax.image.scale_and_translate with
//     the "linear" (triangle) kernel, align-corners scale/translation, antialiased when shrinking.
#include <algorithm>

#include "optim.cuh"

namespace tf {

// ---------------------------------------------------------------------------------------------
// Adam
// ---------------------------------------------------------------------------------------------
constexpr int kAdamThreads = 256;
constexpr int kAdamChunk = 4096;  // elements per CTA: 4 float4 per thread

struct AdamLeaf {
  float* p;
  const float* g;
  float* mu;
  float* nu;
  int64_t n;
  float neg_lr;
  int block_begin;  // first CTA of this leaf
};
struct AdamArgs {
  AdamLeaf leaf[TENSORF_ADAM_MAX_LEAVES];
  int n_leaves;
  float b1, b2, one_minus_b1, one_minus_b2, eps, eps_root, bc1, bc2, lr_decay;
  float* partial;          // [gridDim.x] per-CTA sum of g^2
  unsigned int* ticket;    // arrival counter (zeroed before launch)
  float* grad_norm;        // device scalar or nullptr
};

__global__ void __launch_bounds__(kAdamThreads) k_adam(const __grid_constant__ AdamArgs a) {
  __shared__ float s_red[kAdamThreads / 32];
  __shared__ bool s_last;
  int li = 0;
#pragma unroll 1
  while (li + 1 < a.n_leaves && (int)blockIdx.x >= a.leaf[li + 1].block_begin) ++li;
  const AdamLeaf& L = a.leaf[li];
  const int64_t base = (int64_t)(blockIdx.x - L.block_begin) * kAdamChunk;
  const int64_t n_here = min((int64_t)kAdamChunk, L.n - base);
  const bool vec = ((reinterpret_cast<uintptr_t>(L.p) | reinterpret_cast<uintptr_t>(L.g) | reinterpret_cast<uintptr_t>(L.mu) |
                     reinterpret_cast<uintptr_t>(L.nu)) & 15) == 0;
  float ss = 0.f;
  if (vec && n_here == kAdamChunk) {
    float4 p4[4], g4[4], m4[4], v4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // all 16 loads in flight before any arithmetic
      const int64_t e = base + ((int64_t)i * kAdamThreads + threadIdx.x) * 4;
      g4[i] = __ldcs(reinterpret_cast<const float4*>(L.g + e));
      p4[i] = *reinterpret_cast<const float4*>(L.p + e);
      m4[i] = *reinterpret_cast<const float4*>(L.mu + e);
      v4[i] = *reinterpret_cast<const float4*>(L.nu + e);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t e = base + ((int64_t)i * kAdamThreads + threadIdx.x) * 4;
      ss += g4[i].x * g4[i].x + g4[i].y * g4[i].y + g4[i].z * g4[i].z + g4[i].w * g4[i].w;
      adam_one(p4[i].x, g4[i].x, m4[i].x, v4[i].x, a, L.neg_lr);
      adam_one(p4[i].y, g4[i].y, m4[i].y, v4[i].y, a, L.neg_lr);
      adam_one(p4[i].z, g4[i].z, m4[i].z, v4[i].z, a, L.neg_lr);
      adam_one(p4[i].w, g4[i].w, m4[i].w, v4[i].w, a, L.neg_lr);
      *reinterpret_cast<float4*>(L.p + e) = p4[i];
      *reinterpret_cast<float4*>(L.mu + e) = m4[i];
      *reinterpret_cast<float4*>(L.nu + e) = v4[i];
    }
  } else {
    for (int64_t i = threadIdx.x; i < n_here; i += kAdamThreads) {
      const int64_t e = base + i;
      float p = L.p[e], g = L.g[e], m = L.mu[e], v = L.nu[e];
      ss += g * g;
      adam_one(p, g, m, v, a, L.neg_lr);
      L.p[e] = p;
      L.mu[e] = m;
      L.nu[e] = v;
    }
  }
  // ---- global_norm(grads): per-CTA partial, then the last CTA to arrive sums them in index order
  // (deterministic for a given leaf list) ----
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kAdamThreads / 32; ++w) t += s_red[w];
    a.partial[blockIdx.x] = t;
    __threadfence();
    s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last || a.grad_norm == nullptr) return;
  __threadfence();
  double acc = 0.0;  // fp64 accumulation of the (few thousand) partials
  for (unsigned i = threadIdx.x; i < gridDim.x; i += kAdamThreads) acc += (double)__ldcg(a.partial + i);
  __shared__ double s_acc[kAdamThreads];
  s_acc[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kAdamThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_acc[threadIdx.x] += s_acc[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) a.grad_norm[0] = (float)sqrt(s_acc[0]);
}

int64_t adam_scratch_bytes(const int64_t* sizes, int n_leaves) {
  int64_t blocks = 0;
  for (int i = 0; i < n_leaves; ++i) blocks += ceil_div64(sizes[i], kAdamChunk);
  return 16 + blocks * (int64_t)sizeof(float);
}

int adam_step(cudaStream_t st, const tensorf_adam_desc* d, const int64_t* sizes, float* const* params, const float* const* grads,
              float* const* mu, float* const* nu, const float* neg_lrs, float* grad_norm, void* scratch,
              int64_t scratch_bytes) {
  TF_CHECK_ARG(d && sizes && params && grads && mu && nu && neg_lrs, "adam: null argument");
  TF_CHECK_ARG(d->n_leaves >= 1 && d->n_leaves <= TENSORF_ADAM_MAX_LEAVES, "adam: n_leaves=%d outside [1,%d]", d->n_leaves,
               TENSORF_ADAM_MAX_LEAVES);
  TF_CHECK_ARG(d->bias_correction1 > 0.f && d->bias_correction2 > 0.f, "adam: bias corrections must be > 0 (step count >= 1)");
  AdamArgs a{};
  int64_t blocks = 0;
  int nl = 0;
  for (int i = 0; i < d->n_leaves; ++i) {
    TF_CHECK_ARG(sizes[i] >= 0, "adam: leaf %d has negative size", i);
    if (sizes[i] == 0) continue;
    TF_CHECK_ARG(params[i] && grads[i] && mu[i] && nu[i], "adam: leaf %d has a null buffer", i);
    AdamLeaf& L = a.leaf[nl++];
    L.p = params[i]; L.g = grads[i]; L.mu = mu[i]; L.nu = nu[i];
    L.n = sizes[i];
    L.neg_lr = neg_lrs[i];
    L.block_begin = (int)blocks;
    blocks += ceil_div64(sizes[i], kAdamChunk);
  }
  if (nl == 0) {
    if (grad_norm) {
      TF_CHECK_CUDA(cudaMemsetAsync(grad_norm, 0, sizeof(float), st));
      count_launch();
    }
    return 0;
  }
  TF_CHECK_ARG(blocks < (int64_t)1 << 31, "adam: too many elements");
  TF_CHECK_ARG(scratch && scratch_bytes >= 16 + blocks * (int64_t)sizeof(float),
               "adam: scratch too small (%lld < %lld bytes; see tensorf_adam_scratch_bytes)", (long long)scratch_bytes,
               (long long)(16 + blocks * (int64_t)sizeof(float)));
  a.n_leaves = nl;
  a.b1 = d->b1; a.b2 = d->b2;
  a.one_minus_b1 = 1.0f - d->b1; a.one_minus_b2 = 1.0f - d->b2;  // fp32, as optax computes (1 - decay)
  a.eps = d->eps; a.eps_root = d->eps_root;
  a.bc1 = d->bias_correction1; a.bc2 = d->bias_correction2;
  a.lr_decay = d->lr_decay;
  a.ticket = reinterpret_cast<unsigned int*>(scratch);
  a.partial = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scratch) + 16);
  a.grad_norm = grad_norm;
  StageTimer t(st, "adam");
  TF_CHECK_CUDA(cudaMemsetAsync(a.ticket, 0, 16, st));
  count_launch();
  k_adam<<<(unsigned)blocks, kAdamThreads, 0, st>>>(a);
  TF_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// resize
// ---------------------------------------------------------------------------------------------
// Per output index o: the (few) input taps with non-zero weight, jax/_src/image/scale.py
// compute_weight_mat restated in fp32 with the same operation order:
//   inv = 1/scale ; ks = max(inv, 1) (antialias) ; f = (o+0.5)*inv - translation*inv - 0.5
//   w_i = max(0, 1 - |f - i| / ks) ; w_i /= sum_i w_i (when |sum| > 1000 eps) ; 0 if f outside [-0.5, in-0.5]
// scale = (out-1)/(in-1), translation = -(scale/2 - 0.5)   (tensor_vm.py:204-213)
struct ResizeTaps {
  int* lo;     // [out]
  int* cnt;    // [out]
  float* w;    // [out][max_taps]
  int max_taps;
};

__global__ void k_resize_taps(int in, int out, ResizeTaps T) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= out) return;
  const float scale = __fdiv_rn(__fsub_rn((float)out, 1.0f), __fsub_rn((float)in, 1.0f));
  const float translation = -__fsub_rn(__fdiv_rn(scale, 2.0f), 0.5f);
  const float inv = __fdiv_rn(1.0f, scale);
  const float ks = fmaxf(inv, 1.0f);
  const float f = __fsub_rn(__fsub_rn(__fmul_rn(__fadd_rn((float)o, 0.5f), inv), __fmul_rn(translation, inv)), 0.5f);
  int lo = max(0, (int)ceilf(f - ks) - 1), hi = min(in - 1, (int)floorf(f + ks) + 1);
  // shrink to the non-zero support
  float tot = 0.f;
  int first = -1, last = -2;
  for (int i = lo; i <= hi; ++i) {
    const float w = fmaxf(0.f, __fsub_rn(1.0f, __fdiv_rn(fabsf(__fsub_rn(f, (float)i)), ks)));
    if (w != 0.f) {
      if (first < 0) first = i;
      last = i;
    }
    tot = __fadd_rn(tot, w);
  }
  const bool inside = f >= -0.5f && f <= __fsub_rn((float)in, 0.5f);
  const bool ok = fabsf(tot) > 1000.f * 1.1920928955078125e-07f;
  int cnt = (inside && ok && first >= 0) ? last - first + 1 : 0;
  cnt = min(cnt, T.max_taps);
  T.lo[o] = cnt ? first : 0;
  T.cnt[o] = cnt;
  for (int t = 0; t < T.max_taps; ++t) {
    float w = 0.f;
    if (t < cnt) {
      const int i = first + t;
      w = fmaxf(0.f, __fsub_rn(1.0f, __fdiv_rn(fabsf(__fsub_rn(f, (float)i)), ks)));
      w = __fdiv_rn(w, tot);
    }
    T.w[(int64_t)o * T.max_taps + t] = w;
  }
}

// Resample along the LAST axis: src (rows, in) -> dst (rows, out).
__global__ void __launch_bounds__(256) k_resize_last(const float* __restrict__ src, float* __restrict__ dst, int64_t rows,
                                                     int in, int out, ResizeTaps T) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * out) return;
  const int64_t r = e / out;
  const int o = (int)(e % out);
  const int lo = T.lo[o], cnt = T.cnt[o];
  const float* s = src + r * in + lo;
  float acc = 0.f;
  for (int t = 0; t < cnt; ++t) acc = fmaf(T.w[(int64_t)o * T.max_taps + t], __ldg(s + t), acc);
  dst[e] = acc;
}

// Resample along the MIDDLE axis: src (slabs, in, width) -> dst (slabs, out, width); coalesced along width.
__global__ void __launch_bounds__(256) k_resize_mid(const float* __restrict__ src, float* __restrict__ dst, int64_t slabs,
                                                    int in, int out, int width, ResizeTaps T) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= slabs * out * width) return;
  const int x = (int)(e % width);
  const int o = (int)((e / width) % out);
  const int64_t sl = e / ((int64_t)width * out);
  const int lo = T.lo[o], cnt = T.cnt[o];
  const float* s = src + (sl * in + lo) * width + x;
  float acc = 0.f;
  for (int t = 0; t < cnt; ++t) acc = fmaf(T.w[(int64_t)o * T.max_taps + t], __ldg(s + (int64_t)t * width), acc);
  dst[e] = acc;
}

static int resize_max_taps(int in, int out) {
  const double inv = (double)(in - 1) / (double)(out - 1);
  return 2 * (int)std::ceil(std::max(1.0, inv)) + 2;
}
static int64_t resize_taps_bytes(int in, int out) {
  return round_up64((int64_t)out * (2 * 4 + resize_max_taps(in, out) * 4), 256);
}

int64_t vm_resize_scratch_bytes(int C, int G_in, int G_out) {
  if (G_in == G_out) return 0;
  return resize_taps_bytes(G_in, G_out) + (int64_t)3 * C * G_out * G_in * (int64_t)sizeof(float);
}

int vm_resize(cudaStream_t st, const float* vector_in, const float* matrix_in, int C, int G_in, int G_out, float* vector_out,
              float* matrix_out, void* scratch, int64_t scratch_bytes) {
  TF_CHECK_ARG(vector_in && matrix_in && vector_out && matrix_out, "vm_resize: null buffer");
  TF_CHECK_ARG(C >= 1 && G_in >= 2 && G_out >= 2, "vm_resize: C=%d G_in=%d G_out=%d unsupported (grid dims >= 2)", C, G_in, G_out);
  StageTimer t(st, "resize");
  if (G_in == G_out) {  // no spatial dim changes: scale_and_translate is the identity (tensor_vm.py:204-206)
    TF_CHECK_CUDA(cudaMemcpyAsync(vector_out, vector_in, (size_t)3 * C * G_in * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TF_CHECK_CUDA(cudaMemcpyAsync(matrix_out, matrix_in, (size_t)3 * C * G_in * G_in * sizeof(float), cudaMemcpyDeviceToDevice, st));
    count_launch();
    count_launch();
    return 0;
  }
  TF_CHECK_ARG(scratch && scratch_bytes >= vm_resize_scratch_bytes(C, G_in, G_out), "vm_resize: scratch too small (%lld < %lld)",
               (long long)scratch_bytes, (long long)vm_resize_scratch_bytes(C, G_in, G_out));
  ResizeTaps T;
  T.max_taps = resize_max_taps(G_in, G_out);
  unsigned char* sp = reinterpret_cast<unsigned char*>(scratch);
  T.lo = reinterpret_cast<int*>(sp);
  T.cnt = T.lo + G_out;
  T.w = reinterpret_cast<float*>(T.cnt + G_out);
  float* tmp = reinterpret_cast<float*>(sp + resize_taps_bytes(G_in, G_out));
  k_resize_taps<<<(G_out + 127) / 128, 128, 0, st>>>(G_in, G_out, T);
  TF_CHECK_LAUNCH();
  const int64_t PC = (int64_t)3 * C;
  // lines: (3C, G_in) -> (3C, G_out)
  k_resize_last<<<(unsigned)ceil_div64(PC * G_out, 256), 256, 0, st>>>(vector_in, vector_out, PC, G_in, G_out, T);
  TF_CHECK_LAUNCH();
  // planes: first spatial axis, then the second (the einsum contracts both; fp32 either way)
  k_resize_mid<<<(unsigned)ceil_div64(PC * G_out * G_in, 256), 256, 0, st>>>(matrix_in, tmp, PC, G_in, G_out, G_in, T);
  TF_CHECK_LAUNCH();
  k_resize_last<<<(unsigned)ceil_div64(PC * G_out * G_out, 256), 256, 0, st>>>(tmp, matrix_out, PC * G_out, G_in, G_out, T);
  TF_CHECK_LAUNCH();
  return 0;
}

}  // namespace tf
