// Per-ray kernels of render.py:105-279 / :437-549 and their reverse mode.
//
//   k_density_select   one warp per ray: sampling (AABB or L-inf contraction) -> fused VM
//                      density gather (no (3cd,R,N) feature tensor) -> softplus -> prefix sum
//                      / exp -> Gumbel top-k (radix select, ties -> lower index) -> depth modes
//   k_appearance_gather  VM appearance lookup at the selected samples, rows (M, 3ca)
//   k_composite_fwd    weighted RGB sum x unbias + white background (+ fused MSE loss)
//   k_ray_bwd          reverse of composite + unbias + segment probabilities (suffix scan)
//   k_density_scatter / k_appearance_scatter  re-gather + vector RED into packed gradients
#include <stdlib.h>

#include <algorithm>

#include <cuda_fp16.h>

#include "render_kernels.cuh"
#include "vm.cuh"

#ifndef TF_DS_MINB
#define TF_DS_MINB 8  // measured on B200: 8 CTAs/SM (64 regs) 0.113 ms; 6: 0.119; default: 0.125; 1: 0.156
#endif

namespace tf {

// ---------------------------------------------------------------------------------------------
// Warp-level radix select of the K largest keys, ties -> lower index.
// keys: shared [N]; hist: shared int[256]. On return thr = key of the K-th largest element,
// n_eq = how many elements equal to thr to take (in ascending index order).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_radix_select(const uint32_t* keys, int N, int K, int* hist, int lane, uint32_t& thr,
                                                  int& n_eq) {
  uint32_t prefix = 0, mask = 0;
  int krem = K;
#pragma unroll 1
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = lane; i < 256; i += 32) hist[i] = 0;
    __syncwarp();
    for (int s = lane; s < N; s += 32) {
      uint32_t k = keys[s];
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1);
    }
    __syncwarp();
    int c[8], tot = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c[j] = hist[8 * lane + j];
      tot += c[j];
    }
    int incl = tot;  // suffix sum over lanes: bins of higher lanes hold larger digits
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += t;
    }
    int a = incl - tot, digit = -1, nk = 0;
#pragma unroll
    for (int j = 7; j >= 0; --j) {
      if (digit < 0 && a < krem && a + c[j] >= krem) {
        digit = 8 * lane + j;
        nk = krem - a;
      }
      a += c[j];
    }
    unsigned found = __ballot_sync(0xffffffffu, digit >= 0);
    int src = __ffs(found) - 1;
    digit = __shfl_sync(0xffffffffu, digit, src);
    krem = __shfl_sync(0xffffffffu, nk, src);
    prefix |= (uint32_t)digit << shift;
    mask |= 0xffu << shift;
    __syncwarp();
  }
  thr = prefix;
  n_eq = krem;
}

// Emit the selected indices in ascending order. F(pos, s).
template <typename F>
__device__ __forceinline__ void warp_emit_selected(const uint32_t* keys, int N, uint32_t thr, int n_eq, int lane, F emit) {
  int pos = 0, eq_seen = 0;
  const unsigned lt = (1u << lane) - 1u;
  for (int base = 0; base < N; base += 32) {
    int s = base + lane;
    uint32_t k = (s < N) ? keys[s] : 0u;
    bool gt = (s < N) && (k > thr);
    bool eq = (s < N) && (k == thr);
    unsigned eqb = __ballot_sync(0xffffffffu, eq);
    bool take = gt || (eq && (eq_seen + __popc(eqb & lt) < n_eq));
    unsigned tb = __ballot_sync(0xffffffffu, take);
    if (take) emit(pos + __popc(tb & lt), s);
    pos += __popc(tb);
    eq_seen += __popc(eqb);
  }
}

// ---------------------------------------------------------------------------------------------
// Segment probabilities for one 32-sample tile (render.py:300-347). Called with identical
// arguments by the forward and the reverse kernel so both see bit-identical E / pt.
// ---------------------------------------------------------------------------------------------
struct SegTile {
  float a, E, pt;
};
__device__ __forceinline__ SegTile seg_tile_a(float a, int lane, float& carry_c, float& carry_E) {
  SegTile o;
  o.a = a;
  float c = __fadd_rn(carry_c, warp_incl_scan(o.a, lane));
  o.E = expf(c);  // p_exits
  float Eprev = __shfl_up_sync(0xffffffffu, o.E, 1);
  if (lane == 0) Eprev = carry_E;
  o.pt = __fmul_rn(__fsub_rn(1.0f, expf(o.a)), Eprev);  // p_terminates
  carry_c = __shfl_sync(0xffffffffu, c, 31);
  carry_E = __shfl_sync(0xffffffffu, o.E, 31);
  return o;
}
__device__ __forceinline__ SegTile seg_tile(float z, float delta, bool valid, int lane, float& carry_c, float& carry_E) {
  float sigma = softplus_f(z);
  // neg_scaled_sigmas = -sigmas * step_sizes
  return seg_tile_a(valid ? __fmul_rn(-sigma, delta) : 0.0f, lane, carry_c, carry_E);
}

// render.py:300-347 compute_segment_probabilities as a standalone op: one warp per row.
__global__ void __launch_bounds__(128) k_segment_probs(const float* __restrict__ sigmas, const float* __restrict__ steps,
                                                       float* __restrict__ p_exits, float* __restrict__ p_term, int R, int N) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= R) return;
  float carry_c = 0.f, carry_E = 1.f;
  for (int base = 0; base < N; base += 32) {
    int s = base + lane;
    bool valid = s < N;
    float a = valid ? __fmul_rn(-sigmas[(int64_t)r * N + s], steps[(int64_t)r * N + s]) : 0.0f;
    SegTile sg = seg_tile_a(a, lane, carry_c, carry_E);
    if (valid) {
      p_exits[(int64_t)r * N + s] = sg.E;
      p_term[(int64_t)r * N + s] = sg.pt;
    }
  }
}

int launch_segment_probs(cudaStream_t st, const float* sigmas, const float* steps, float* p_exits, float* p_term, int R,
                         int N) {
  if (R == 0) return 0;
  k_segment_probs<<<(R + 3) / 4, 128, 0, st>>>(sigmas, steps, p_exits, p_term, R, N);
  TF_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// k_density_select
// ---------------------------------------------------------------------------------------------
template <int LPS>
__device__ __forceinline__ float density_sample_sum(const float* __restrict__ packed, const VmTaps& taps, int G, int Cp,
                                                    int sub) {
  const int nvec = Cp >> 2;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int v = sub; v < nvec; v += LPS) {
#pragma unroll
    for (int P = 0; P < 3; ++P) {
      PairAddr pa = pair_addr(taps, P, G, Cp, v);
      float4 lin, bil;
      pair_values(packed, pa, lin, bil);
      acc.x = fmaf(lin.x, bil.x, acc.x);
      acc.y = fmaf(lin.y, bil.y, acc.y);
      acc.z = fmaf(lin.z, bil.z, acc.z);
      acc.w = fmaf(lin.w, bil.w, acc.w);
    }
  }
  return (acc.x + acc.y) + (acc.z + acc.w);
}

template <int LPS>
__global__ void __launch_bounds__(128, TF_DS_MINB) k_density_select(DensityArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Npad = round_up(A.N, 32);
  // per warp: float vals[Npad]; uint32 keys[Npad]; int hist[256]
  unsigned char* wbase = smem_raw + (size_t)warp * ((size_t)Npad * 8 + 1024);
  float* vals = reinterpret_cast<float*>(wbase);
  uint32_t* keys = reinterpret_cast<uint32_t*>(wbase + (size_t)Npad * 4);
  int* hist = reinterpret_cast<int*>(wbase + (size_t)Npad * 8);

  const int r = blockIdx.x * 4 + warp;
  if (r >= A.R) return;

  SceneParams sc;
  load_scene(sc, A.aabb, A.N, A.G, A.contracted, A.jitter, A.base_ts, A.deltas);
  RayParams ray;
  load_ray(ray, sc, A.origins, A.directions, A.aabb, r);

  // ---- phase 1: density gather, LPS lanes per sample --------------------------------------
  constexpr int SPI = 32 / LPS;
  const int sub = lane % LPS, sl = lane / LPS;
  for (int sbase = 0; sbase < A.N; sbase += SPI) {
    int s = sbase + sl;
    float acc = 0.f;
    if (s < A.N) {
      float delta;
      float t = sample_t(ray, sc, r, s, delta);
      float x[3];
      sample_grid_coords(ray, sc, t, x);
      if (sub == 0 && A.xs_out) {
        float* xo = A.xs_out + ((int64_t)r * A.N + s) * 3;
        xo[0] = x[0];
        xo[1] = x[1];
        xo[2] = x[2];
      }
      VmTaps taps;
      make_vm_taps(taps, x, A.G);
      acc = density_sample_sum<LPS>(A.packed_d, taps, A.G, A.Cp, sub);
    }
#pragma unroll
    for (int o = LPS >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (s < A.N && sub == 0) vals[s] = acc + 10.0f;  // render.py:207
  }
  __syncwarp();

  // ---- phase 2: segment probabilities, Gumbel keys, depth accumulators ---------------------
  float carry_c = 0.f, carry_E = 1.f;
  float mean_acc = 0.f, med_acc = 0.f;
  bool carry_mask = false;
  for (int base = 0; base < A.N; base += 32) {
    int s = base + lane;
    bool valid = s < A.N;
    float z = valid ? vals[s] : 0.f;
    float delta = 0.f, t = 0.f;
    if (valid) t = sample_t(ray, sc, r, s, delta);
    SegTile sg = seg_tile(z, delta, valid, lane, carry_c, carry_E);
    bool mask = __fsub_rn(1.0f, sg.E) > 0.5f;  // render.py:253-258
    bool pmask = __shfl_up_sync(0xffffffffu, (int)mask, 1) != 0;
    if (lane == 0) pmask = carry_mask;
    carry_mask = __shfl_sync(0xffffffffu, (int)mask, 31) != 0;
    if (valid) {
      A.z_out[(int64_t)r * A.N + s] = z;
      vals[s] = sg.pt;
      if (A.mode == TENSORF_MODE_RGB) keys[s] = ordered_key(__fadd_rn(A.gumbel[s], logf(sg.pt)));
      mean_acc = fmaf(sg.pt, t, mean_acc);         // render.py:271-276
      if (s >= 1 && mask != pmask) med_acc += t;  // render.py:259-266 (where-semantics)
    }
  }
  const float E_last = carry_E;
  __syncwarp();

  if (A.mode != TENSORF_MODE_RGB) {
    float dlt;
    float t_last = sample_t(ray, sc, r, A.N - 1, dlt);
    float out;
    if (A.mode == TENSORF_MODE_DIST_MEAN) {
      out = warp_sum(mean_acc) + E_last * t_last;
    } else {
      out = warp_sum(med_acc);
      if (!carry_mask) out += INFINITY;  // padded (True, inf) element
    }
    if (lane == 0) A.depth_out[r] = out;
    return;
  }

  // ---- phase 3: Gumbel top-k (render.py:461-469) -------------------------------------------
  uint32_t thr;
  int n_eq;
  warp_radix_select(keys, A.N, A.K, hist, lane, thr, n_eq);
  float S = 0.f;
  int32_t* idx_row = A.idx_out + (int64_t)r * A.K;
  float* pt_row = A.pt_sel_out + (int64_t)r * A.K;
  warp_emit_selected(keys, A.N, thr, n_eq, lane, [&](int pos, int s) {
    float pt = vals[s];
    idx_row[pos] = s;
    pt_row[pos] = pt;
    S += pt;
  });
  S = warp_sum(S);
  if (lane == 0) {
    A.stats_out[(int64_t)r * 8 + 0] = E_last;
    A.stats_out[(int64_t)r * 8 + 1] = S;
  }
}

int launch_density_select(cudaStream_t st, const DensityArgs& A) {
  if (A.R == 0) return 0;
  const int nvec = A.Cp / 4;
  // Lanes per sample: every lane of a sample recomputes the sample position and the taps, so few lanes with up to
  // four float4 channel groups each win (measured on B200: cd=16: 1 lane 0.093 ms, 2: 0.104, 4: 0.115;
  // cd=32: 2 lanes 0.303 ms, 1: 0.307, 4: 0.326, 8: 0.389).
  int lps = 1;
  while (lps < 8 && nvec > 4 * lps) lps <<= 1;
  if (const char* e = getenv("TENSORF_DS_LPS")) lps = atoi(e);  // 1, 2, 4 or 8
  const int Npad = round_up(A.N, 32);
  size_t smem = 4 * ((size_t)Npad * 8 + 1024);
  TF_CHECK_ARG(smem <= 200 * 1024, "density_samples_per_ray=%d too large for the per-ray kernel", A.N);
  dim3 grid((A.R + 3) / 4), block(128);
#define TF_LAUNCH_DS(L)                                                                                      \
  do {                                                                                                       \
    if (smem > 48 * 1024)                                                                                    \
      TF_CHECK_CUDA(cudaFuncSetAttribute(k_density_select<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_density_select<L><<<grid, block, smem, st>>>(A);                                                       \
  } while (0)
  switch (lps) {
    case 8: TF_LAUNCH_DS(8); break;
    case 4: TF_LAUNCH_DS(4); break;
    case 2: TF_LAUNCH_DS(2); break;
    default: TF_LAUNCH_DS(1); break;
  }
#undef TF_LAUNCH_DS
  TF_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// standalone selection stage (test entry point)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_topk_select(const float* __restrict__ g, int R, int N, int K, int32_t* __restrict__ idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Npad = round_up(N, 32);
  unsigned char* wbase = smem_raw + (size_t)warp * ((size_t)Npad * 4 + 1024);
  uint32_t* keys = reinterpret_cast<uint32_t*>(wbase);
  int* hist = reinterpret_cast<int*>(wbase + (size_t)Npad * 4);
  const int r = blockIdx.x * 4 + warp;
  if (r >= R) return;
  for (int s = lane; s < N; s += 32) keys[s] = ordered_key(g[(int64_t)r * N + s]);
  __syncwarp();
  uint32_t thr;
  int n_eq;
  warp_radix_select(keys, N, K, hist, lane, thr, n_eq);
  int32_t* row = idx + (int64_t)r * K;
  warp_emit_selected(keys, N, thr, n_eq, lane, [&](int pos, int s) { row[pos] = s; });
}

int launch_topk_select(cudaStream_t st, const float* g, int R, int N, int K, int32_t* idx) {
  if (R == 0) return 0;
  size_t smem = 4 * ((size_t)round_up(N, 32) * 4 + 1024);
  TF_CHECK_ARG(smem <= 200 * 1024, "N=%d too large for the selection kernel", N);
  if (smem > 48 * 1024)
    TF_CHECK_CUDA(cudaFuncSetAttribute(k_topk_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_topk_select<<<(R + 3) / 4, 128, smem, st>>>(g, R, N, K, idx);
  TF_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// appearance gather / scatter at the selected samples (render.py:472-486)
// work item = (row m, float4 channel group v)
// ---------------------------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(256) k_appearance(AppearanceArgs A) {
  const int nvec = A.Cp >> 2;
  int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t m = item / nvec;
  int v = (int)(item % nvec);
  if (m >= A.M) return;
  int r = (int)(m / A.K);
  int s = A.idx[m];
  const float* xp = A.xs + ((int64_t)r * A.N + s) * 3;  // coordinates computed once by k_density_select
  float x[3] = {xp[0], xp[1], xp[2]};
  VmTaps taps;
  make_vm_taps(taps, x, A.G);
  const int Ca = 3 * A.C;
#pragma unroll
  for (int P = 0; P < 3; ++P) {
    PairAddr pa = pair_addr(taps, P, A.G, A.Cp, v);
    float4 lin, bil;
    pair_values(A.packed_a, pa, lin, bil);
    int c0 = 4 * v;
    if (!BWD) {
      float4 f = f4_mul(lin, bil);
      float* dst = A.feat + m * Ca + P * A.C + c0;
      if ((A.C & 3) == 0) {
        *reinterpret_cast<float4*>(dst) = f;
      } else {
        float fv[4] = {f.x, f.y, f.z, f.w};
        for (int j = 0; j < 4; ++j)
          if (c0 + j < A.C) dst[j] = fv[j];
      }
    } else {
      const float* src = A.d_feat + m * Ca + P * A.C + c0;
      float4 g;
      if ((A.C & 3) == 0) {
        g = *reinterpret_cast<const float4*>(src);
      } else {
        float gv[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < 4; ++j)
          if (c0 + j < A.C) gv[j] = src[j];
        g = make_float4(gv[0], gv[1], gv[2], gv[3]);
      }
      float4 gl = f4_mul(g, bil), gb = f4_mul(g, lin);
      red_add_v4(A.d_packed + pa.l0, f4_scale(gl, pa.wl0));
      red_add_v4(A.d_packed + pa.l1, f4_scale(gl, pa.wl1));
      red_add_v4(A.d_packed + pa.m00, f4_scale(gb, pa.w00));
      red_add_v4(A.d_packed + pa.m01, f4_scale(gb, pa.w01));
      red_add_v4(A.d_packed + pa.m10, f4_scale(gb, pa.w10));
      red_add_v4(A.d_packed + pa.m11, f4_scale(gb, pa.w11));
    }
  }
}

// The same lookup for the fused MLP kernels (csrc/mlp_fused.cu): rows are written as two-term fp16 slab tiles (per 128 rows
// and 8 columns one hi and one lo slab of 128 x 16 B, csrc/umma_tiles.cuh) instead of fp32 rows - the same number of bytes,
// but the MLP's loader copies them straight into its operand ring.  The gather keeps the mapping of k_appearance (item =
// (row, float4 channel group): the 12 lanes of a row read one contiguous 192-byte texel per tap); a block owns 32 rows,
// stages their hi / lo halves in shared memory in slab order and writes every slab segment (32 rows x 16 B = 512 B) with
// coalesced 16-byte stores - scattering 8-byte fragments straight to global memory cost 3-4x the L2 write requests.
// Needs C % 8 == 0.  Rows M..ceil128(M) are written as zeros (they must be finite).
constexpr int kSlabRows = 32, kSlabPitch = kSlabRows * 16 + 16;  // +16: column groups land on different banks
__global__ void __launch_bounds__(384, 3) k_appearance_slab(AppearanceArgs A) {  // blockDim = 32 rows x C/4 items: one item per thread
  // (this kernel itself is launched normally: as a programmatic dependent its 3 CTAs per SM sit on the SMs of
  //  k_density_select's last wave while they wait - measured 11 us slower per step)
  pdl_launch_dependents();  // the fused MLP kernel may load its weights while this grid's last wave drains
  extern __shared__ __align__(16) unsigned char sm_slab[];
  const int nvec = A.C >> 2, Ca = 3 * A.C, nslab = (Ca >> 3) * 2;
  const int64_t m0 = (int64_t)blockIdx.x * kSlabRows;
  for (int item = threadIdx.x; item < kSlabRows * nvec; item += blockDim.x) {
    const int rl = item / nvec, v = item % nvec;
    const int64_t m = m0 + rl;
    uint2 hi[3], lo[3];
#pragma unroll
    for (int P = 0; P < 3; ++P) hi[P] = lo[P] = make_uint2(0u, 0u);
    if (m < A.M) {
      const int r = (int)(m / A.K);
      const int s = A.idx[m];
      const float* xp = A.xs + ((int64_t)r * A.N + s) * 3;
      float x[3] = {xp[0], xp[1], xp[2]};
      VmTaps taps;
      make_vm_taps(taps, x, A.G);
#pragma unroll
      for (int P = 0; P < 3; ++P) {
        PairAddr pa = pair_addr(taps, P, A.G, A.Cp, v);
        float4 lin, bil;
        pair_values(A.packed_a, pa, lin, bil);
        const float4 f = f4_mul(lin, bil);
        const __half2 h01 = __floats2half2_rn(f.x, f.y), h23 = __floats2half2_rn(f.z, f.w);
        const float2 r01 = __half22float2(h01), r23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(f.x - r01.x, f.y - r01.y), l23 = __floats2half2_rn(f.z - r23.x, f.w - r23.y);
        hi[P] = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        lo[P] = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
      }
    }
#pragma unroll
    for (int P = 0; P < 3; ++P) {
      const int cg = (P * A.C + 4 * v) >> 3;
      unsigned char* d = sm_slab + (size_t)(2 * cg) * kSlabPitch + rl * 16 + (v & 1) * 8;
      *reinterpret_cast<uint2*>(d) = hi[P];
      *reinterpret_cast<uint2*>(d + kSlabPitch) = lo[P];
    }
  }
  __syncthreads();
  unsigned char* tile = reinterpret_cast<unsigned char*>(A.feat) + (m0 >> 7) * ((int64_t)Ca * 512) + (m0 & 127) * 16;
  for (int i = threadIdx.x; i < nslab * kSlabRows; i += blockDim.x) {
    const int sl = i / kSlabRows, rl = i % kSlabRows;
    *reinterpret_cast<uint4*>(tile + (int64_t)sl * 2048 + rl * 16) = *reinterpret_cast<const uint4*>(sm_slab + (size_t)sl * kSlabPitch + rl * 16);
  }
}

int launch_appearance(cudaStream_t st, const AppearanceArgs& A, bool bwd) {
  if (A.M == 0) return 0;
  if (!bwd && A.feat_slabs) {
    TF_CHECK_ARG(A.C % 8 == 0 && A.Cp == A.C, "appearance slab rows need ca %% 8 == 0");
    const size_t smem = (size_t)(3 * A.C / 8) * 2 * kSlabPitch;
    TF_CHECK_ARG(smem <= 96 * 1024, "appearance slab rows: too many channels");
    if (smem > 48 * 1024) TF_CHECK_CUDA(cudaFuncSetAttribute(k_appearance_slab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TF_CHECK_ARG(kSlabRows * (A.C / 4) <= 384, "appearance slab rows: ca=%d too wide", A.C);
    k_appearance_slab<<<(unsigned)(round_up64(A.M, 128) / kSlabRows), kSlabRows * (A.C / 4), smem, st>>>(A);
    TF_CHECK_LAUNCH();
    return 0;
  }
  int64_t items = A.M * (A.Cp / 4);
  unsigned grid = (unsigned)ceil_div64(items, 256);
  if (bwd)
    k_appearance<true><<<grid, 256, 0, st>>>(A);
  else
    k_appearance<false><<<grid, 256, 0, st>>>(A);
  TF_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// composite (render.py:233-246, :529-546) + fused MSE (training.py:140)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_composite_fwd(CompositeArgs A) {
  pdl_launch_dependents();
  pdl_wait();  // rgb of the MLP kernel before this one
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = blockIdx.x * 4 + warp;
  __shared__ float loss_part[4];
  float sq = 0.f;
  if (r < A.R) {
    float W[3] = {0.f, 0.f, 0.f};
    for (int k = lane; k < A.K; k += 32) {
      int64_t m = (int64_t)r * A.K + k;
      float pt = A.pt_sel[m];
      W[0] = fmaf(A.rgb_sel[3 * m + 0], pt, W[0]);
      W[1] = fmaf(A.rgb_sel[3 * m + 1], pt, W[1]);
      W[2] = fmaf(A.rgb_sel[3 * m + 2], pt, W[2]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) W[c] = warp_sum(W[c]);
    float* st = A.stats + (int64_t)r * 8;
    float E_last = st[0], S = st[1];
    float u = __fdiv_rn(__fadd_rn(__fsub_rn(1.0f, E_last), 1e-8f), __fadd_rn(S, 1e-8f));
    if (lane < 3) {
      float w = lane == 0 ? W[0] : (lane == 1 ? W[1] : W[2]);
      float rgb = __fadd_rn(__fmul_rn(w, u), E_last);
      A.rgb_out[3 * r + lane] = rgb;
      st[4 + lane] = w;
      if (A.colors) {
        float diff = rgb - A.colors[3 * r + lane];
        A.go[3 * r + lane] = 2.0f * diff * A.loss_scale;
        sq = diff * diff;
      }
    }
    if (lane == 3) st[2] = u;
  }
  if (A.colors) {
    sq = warp_sum(sq);
    if (lane == 0) loss_part[warp] = sq;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(A.loss, (loss_part[0] + loss_part[1] + loss_part[2] + loss_part[3]) * A.loss_scale);
  }
}

int launch_composite_fwd(cudaStream_t st, const CompositeArgs& A) {
  if (A.R == 0) return 0;
  TF_CHECK_CUDA(launch_pdl(k_composite_fwd, dim3((unsigned)((A.R + 3) / 4)), dim3(128), 0, st, A, true, 2));
  TF_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// k_ray_bwd: reverse of composite/unbias/segment probabilities
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_ray_bwd(RayBwdArgs A) {
  pdl_launch_dependents();  // the reverse MLP kernel may set itself up while this grid drains (it waits before reading)
  pdl_wait();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Npad = round_up(A.N, 32);
  float* gE_s = reinterpret_cast<float*>(smem_raw) + (size_t)warp * 2 * Npad;  // gpt_s * E_s
  float* q_s = gE_s + Npad;                                                    // gpt_s, then gpt_s * pt_s
  {  // zero the scatter kernels' accumulators (fire-and-forget stores, issued before the ray work)
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = gtid; i < A.zero0_n4; i += nthr) A.zero0[i] = z4;
    for (int64_t i = gtid; i < A.zero1_n4; i += nthr) A.zero1[i] = z4;
    for (int j = 0; j < A.n_zleaf; ++j)
      for (int64_t i = gtid; i < A.zleaf_n[j]; i += nthr) A.zleaf[j][i] = 0.f;
  }
  const int r = blockIdx.x * 4 + warp;
  if (r >= A.R) return;

  SceneParams sc;
  load_scene(sc, A.aabb, A.N, A.G, A.contracted, A.jitter, A.base_ts, A.deltas);
  RayParams ray;
  load_ray(ray, sc, A.origins, A.directions, A.aabb, r);

  const float* st = A.stats + (int64_t)r * 8;
  const float E_last = st[0], S = st[1];
  const float Wc[3] = {st[4], st[5], st[6]};
  const float go[3] = {A.go[3 * r + 0], A.go[3 * r + 1], A.go[3 * r + 2]};
  const float Au = __fadd_rn(__fsub_rn(1.0f, E_last), 1e-8f), Bu = __fadd_rn(S, 1e-8f);
  const float u = __fdiv_rn(Au, Bu);
  const float AoB2 = Au / (Bu * Bu);

  for (int s = lane; s < Npad; s += 32) q_s[s] = 0.f;
  __syncwarp();
  float dmax = 0.f;
  for (int k = lane; k < A.K; k += 32) {
    int64_t m = (int64_t)r * A.K + k;
    float pt = A.pt_sel[m];
    float gpt = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float ck = A.rgb_sel[3 * m + c];
      const float dsel = go[c] * u * pt;
      A.d_rgb_sel[3 * m + c] = dsel;
      dmax = fmaxf(dmax, fabsf(dsel));
      gpt += go[c] * (u * ck - Wc[c] * AoB2);
    }
    q_s[A.idx[m]] = gpt;
  }
  if (A.amax != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    if (lane == 0 && dmax > 0.f) atomicMax(reinterpret_cast<unsigned int*>(A.amax), __float_as_uint(dmax));
  }
  float gE = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) gE += go[c] * (1.0f - Wc[c] / Bu);
  __syncwarp();

  // forward recompute of E_s, pt_s (bit-identical to k_density_select)
  float carry_c = 0.f, carry_E = 1.f;
  for (int base = 0; base < A.N; base += 32) {
    int s = base + lane;
    bool valid = s < A.N;
    float z = valid ? A.z[(int64_t)r * A.N + s] : 0.f;
    float delta = 0.f;
    if (valid) (void)sample_t(ray, sc, r, s, delta);
    SegTile sg = seg_tile(z, delta, valid, lane, carry_c, carry_E);
    if (valid) {
      float gpt = q_s[s];
      gE_s[s] = gpt * sg.E;
      q_s[s] = gpt * sg.pt;
    }
  }
  __syncwarp();

  // reverse pass: dL/da_j = -gpt_j E_j + sum_{s>j} gpt_s pt_s + gE E_last
  float carry_suf = 0.f;
  const float tail = gE * E_last;
  for (int base = Npad - 32; base >= 0; base -= 32) {
    int s = base + lane;
    bool valid = s < A.N;
    float q = valid ? q_s[s] : 0.f;
    float incl = q;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float t = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += t;
    }
    float excl = incl - q + carry_suf;
    carry_suf += __shfl_sync(0xffffffffu, incl, 0);
    if (valid) {
      float z = A.z[(int64_t)r * A.N + s];
      float delta;
      (void)sample_t(ray, sc, r, s, delta);
      float da = -gE_s[s] + excl + tail;
      float sigma = softplus_f(z);
      A.dz[(int64_t)r * A.N + s] = da * (-delta) * expf(z - sigma);
    }
  }
}

int launch_ray_bwd(cudaStream_t st, const RayBwdArgs& A) {
  if (A.R == 0) {  // no rays: the accumulators still have to read as zero
    if (A.zero0_n4) TF_CHECK_CUDA(cudaMemsetAsync(A.zero0, 0, sizeof(float4) * A.zero0_n4, st));
    if (A.zero1_n4) TF_CHECK_CUDA(cudaMemsetAsync(A.zero1, 0, sizeof(float4) * A.zero1_n4, st));
    for (int j = 0; j < A.n_zleaf; ++j) TF_CHECK_CUDA(cudaMemsetAsync(A.zleaf[j], 0, sizeof(float) * A.zleaf_n[j], st));
    return 0;
  }
  size_t smem = 4 * (size_t)2 * round_up(A.N, 32) * sizeof(float);
  if (smem > 48 * 1024) TF_CHECK_CUDA(cudaFuncSetAttribute(k_ray_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TF_CHECK_CUDA(launch_pdl(k_ray_bwd, dim3((unsigned)((A.R + 3) / 4)), dim3(128), smem, st, A, true, 2));
  TF_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Sliding-window scatter-add (reverse of the VM gather) for samples that are ordered along a ray.
//
// Work item = (ray, segment of consecutive samples, VM pair P, block of LPS float4 channel
// groups); LPS lanes per item, one float4 of channels per lane.  Consecutive samples of a ray
// move less than a texel per step, so instead of 6 REDs per sample the lane keeps the current
// 2-texel line window and 2x2 plane window of gradient sums in registers and issues a vector
// RED only when a texel leaves the window: one RED per (ray segment, texel visit).
//   APP=false: density   — walks s in [seg), upstream = scalar dz[r,s] for every channel
//   APP=true : appearance — walks the ascending selected samples idx[r,k], upstream = d_feat row
// ---------------------------------------------------------------------------------------------
struct WalkArgs : SceneArgs {
  const float* packed;
  float* d_packed;
  const float* dz;      // density
  const float* xs;      // (R,N,3) grid coordinates
  const int32_t* idx;   // appearance
  const float* d_feat;  // appearance (M, 3C)
  int C, Cp, seg_len, segs, count;  // count = N (density) or K (appearance)
  bool slim;                         // appearance: the 96-register variant (runs beside the fused MLP's weight-gradient kernel)
};

// Register budgets: density 80 (768 threads per SM); appearance 128 (114 used; 512 threads per SM) when the kernel has the
// GPU to itself, and SLIM = 96 for the forked reverse pass, where 512 of its threads then fit beside the weight-gradient
// kernel's CTA (192 threads x 48 registers) instead of 256: alone the slim variant is 12 us slower (0.094 -> 0.106 ms),
// beside that kernel the step is 15 us faster.  (Density at 72 or 76 registers spills and loses.)
// CTA size (launch_walk): 64 threads for the density walk, 128 for the appearance walk - small CTAs fill the register
// file left over by the MLP kernels at a finer grain and shorten the tail: 256 -> 64 / 128 threads took another 18 us off
// the step (the density walk alone: 0.152 -> 0.144 ms).
template <int LPS, bool APP, bool SLIM = false>
__global__ void __maxnreg__(APP ? (SLIM ? 96 : 128) : 80) k_scatter_walk(WalkArgs A) {  // measured: density 3 CTAs/SM, appearance 2
  const int nvec = A.Cp >> 2;
  const int vblocks = (nvec + LPS - 1) / LPS;
  const int sub = threadIdx.x % LPS;
  const int per_ray = A.segs * 3 * vblocks;
  int64_t item = (int64_t)blockIdx.x * (blockDim.x / LPS) + threadIdx.x / LPS;
  if (item >= (int64_t)A.R * per_ray) return;
  const int r = (int)(item / per_ray);
  int rem = (int)(item % per_ray);
  const int seg = rem / (3 * vblocks);
  rem -= seg * 3 * vblocks;
  const int P = rem / vblocks;
  const int v = (rem % vblocks) * LPS + sub;
  if (v >= nvec) return;
  const int j_begin = seg * A.seg_len, j_end = min(A.count, j_begin + A.seg_len);

  const int G = A.G, Cp = A.Cp;
  const int lbase = P * G * Cp + 4 * v;  // 32-bit float offsets (entry points reject >= 2^31 packed floats)
  const int mbase = 3 * G * Cp + P * G * G * Cp + 4 * v;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int EMPTY = INT_MIN;
  int lw = EMPTY;           // line window covers texels lw, lw+1
  float4 la0 = zero4, la1 = zero4;
  int pa = EMPTY, pb = 0;   // plane window covers rows pa, pa+1 x cols pb, pb+1
  float4 m00 = zero4, m01 = zero4, m10 = zero4, m11 = zero4;
  auto flush_l = [&](int i, const float4& acc) { red_add_v4(A.d_packed + (lbase + i * Cp), acc); };
  auto flush_m = [&](int a, int b, const float4& acc) { red_add_v4(A.d_packed + (mbase + (a * G + b) * Cp), acc); };

  // One-sample software pipeline over the per-sample operands: the cotangent (appearance: the d_feat row, the only
  // operand that comes from DRAM, 90 MB per step read once; density: dz), the saved grid coordinates, and for the
  // appearance walk the selected sample index two steps ahead (the coordinate address depends on it).
  auto load_g = [&](int j, float4& g_out) {
    if (APP) {
      const float* src = A.d_feat + ((int64_t)r * A.K + j) * (3 * A.C) + P * A.C + 4 * v;
      if ((A.C & 3) == 0) {
        g_out = __ldcs(reinterpret_cast<const float4*>(src));
      } else {
        float gv[4] = {0.f, 0.f, 0.f, 0.f};
        for (int q = 0; q < 4; ++q)
          if (4 * v + q < A.C) gv[q] = src[q];
        g_out = make_float4(gv[0], gv[1], gv[2], gv[3]);
      }
    } else {
      const float gz = __ldg(A.dz + (int64_t)r * A.N + j);
      g_out = make_float4(gz, gz, gz, gz);
    }
  };
  auto load_s = [&](int j) -> int { return APP ? __ldg(A.idx + (int64_t)r * A.K + j) : j; };
  auto load_x = [&](int sidx, float& x0, float& x1, float& x2) {
    const float* xp = A.xs + ((int64_t)r * A.N + sidx) * 3;  // coordinates saved by the forward gather
    x0 = __ldg(xp);
    x1 = __ldg(xp + 1);
    x2 = __ldg(xp + 2);
  };
  // (density walk: 80 registers at 3 CTAs/SM leave no room for the pipeline registers - it spilled and lost 12 %;
  //  its operands are L2 hits, so it loads them in place)
  float4 g_nxt = zero4;
  float xn0 = 0.f, xn1 = 0.f, xn2 = 0.f;
  int s_nxt2 = 0;  // sample index of step j+2
  if (APP && j_begin < j_end) {
    load_g(j_begin, g_nxt);
    load_x(load_s(j_begin), xn0, xn1, xn2);
    if (j_begin + 1 < j_end) s_nxt2 = load_s(j_begin + 1);
  }
  for (int j = j_begin; j < j_end; ++j) {
    float4 g;
    float x[3];
    if (APP) {
      g = g_nxt;
      x[0] = xn0; x[1] = xn1; x[2] = xn2;
      if (j + 1 < j_end) {
        load_g(j + 1, g_nxt);
        load_x(s_nxt2, xn0, xn1, xn2);
        if (j + 2 < j_end) s_nxt2 = load_s(j + 2);
      }
    } else {
      load_g(j, g);
      load_x(j, x[0], x[1], x[2]);
    }
    // axis roles of pair P (tensor_vm.py:50-52), selected without indexing local arrays
    const float xl = P == 0 ? x[0] : (P == 1 ? x[2] : x[1]);
    const float xa = P == 0 ? x[1] : (P == 1 ? x[0] : x[2]);
    const float xb = P == 0 ? x[2] : (P == 1 ? x[1] : x[0]);
    const Tap L = make_tap(xl, G), Ta = make_tap(xa, G), Tb = make_tap(xb, G);

    // re-gather (bit-identical to the forward)
    const float4 l0 = ld4(A.packed + (lbase + L.i0 * Cp)), l1 = ld4(A.packed + (lbase + L.i1 * Cp));
    const float4 v00 = ld4(A.packed + (mbase + (Ta.i0 * G + Tb.i0) * Cp));
    const float4 v01 = ld4(A.packed + (mbase + (Ta.i0 * G + Tb.i1) * Cp));
    const float4 v10 = ld4(A.packed + (mbase + (Ta.i1 * G + Tb.i0) * Cp));
    const float4 v11 = ld4(A.packed + (mbase + (Ta.i1 * G + Tb.i1) * Cp));
    float4 lin = f4_fma(l1, L.w1, f4_scale(l0, L.w0));
    float4 bil = f4_scale(v00, __fmul_rn(Ta.w0, Tb.w0));
    bil = f4_fma(v01, __fmul_rn(Ta.w0, Tb.w1), bil);
    bil = f4_fma(v10, __fmul_rn(Ta.w1, Tb.w0), bil);
    bil = f4_fma(v11, __fmul_rn(Ta.w1, Tb.w1), bil);
    const float4 gl = f4_mul(g, bil), gb = f4_mul(g, lin);

    // ---- line window ----
    const int wb = min(L.i0, G - 2);
    if (wb != lw) {
      if (lw != EMPTY) {
        const int d = wb - lw;
        if (d == 1) {
          flush_l(lw, la0);
          la0 = la1;
          la1 = zero4;
        } else if (d == -1) {
          flush_l(lw + 1, la1);
          la1 = la0;
          la0 = zero4;
        } else {
          flush_l(lw, la0);
          flush_l(lw + 1, la1);
          la0 = la1 = zero4;
        }
      }
      lw = wb;
    }
    {
      const float s0 = (L.i0 == wb ? L.w0 : 0.f) + (L.i1 == wb ? L.w1 : 0.f);
      const float s1 = (L.i0 == wb + 1 ? L.w0 : 0.f) + (L.i1 == wb + 1 ? L.w1 : 0.f);
      la0 = f4_fma(gl, s0, la0);
      la1 = f4_fma(gl, s1, la1);
    }
    // ---- plane window ----
    const int ab = min(Ta.i0, G - 2), bb = min(Tb.i0, G - 2);
    if (ab != pa || bb != pb) {
      if (pa != EMPTY) {
        const int da = ab - pa, db = bb - pb;
        if (da > 1 || da < -1 || db > 1 || db < -1) {
          flush_m(pa, pb, m00);
          flush_m(pa, pb + 1, m01);
          flush_m(pa + 1, pb, m10);
          flush_m(pa + 1, pb + 1, m11);
          m00 = m01 = m10 = m11 = zero4;
        } else {
          if (da == 1) {
            flush_m(pa, pb, m00);
            flush_m(pa, pb + 1, m01);
            m00 = m10;
            m01 = m11;
            m10 = m11 = zero4;
          } else if (da == -1) {
            flush_m(pa + 1, pb, m10);
            flush_m(pa + 1, pb + 1, m11);
            m10 = m00;
            m11 = m01;
            m00 = m01 = zero4;
          }
          // rows now are ab, ab+1; the row that just entered the window is still all zero
          if (db == 1) {
            if (da != -1) flush_m(ab, pb, m00);
            if (da != 1) flush_m(ab + 1, pb, m10);
            m00 = m01;
            m10 = m11;
            m01 = m11 = zero4;
          } else if (db == -1) {
            if (da != -1) flush_m(ab, pb + 1, m01);
            if (da != 1) flush_m(ab + 1, pb + 1, m11);
            m01 = m00;
            m11 = m10;
            m00 = m10 = zero4;
          }
        }
      }
      pa = ab;
      pb = bb;
    }
    {
      const float r0 = (Ta.i0 == ab ? Ta.w0 : 0.f) + (Ta.i1 == ab ? Ta.w1 : 0.f);
      const float r1 = (Ta.i0 == ab + 1 ? Ta.w0 : 0.f) + (Ta.i1 == ab + 1 ? Ta.w1 : 0.f);
      const float c0 = (Tb.i0 == bb ? Tb.w0 : 0.f) + (Tb.i1 == bb ? Tb.w1 : 0.f);
      const float c1 = (Tb.i0 == bb + 1 ? Tb.w0 : 0.f) + (Tb.i1 == bb + 1 ? Tb.w1 : 0.f);
      m00 = f4_fma(gb, r0 * c0, m00);
      m01 = f4_fma(gb, r0 * c1, m01);
      m10 = f4_fma(gb, r1 * c0, m10);
      m11 = f4_fma(gb, r1 * c1, m11);
    }
  }
  if (lw != EMPTY) {
    flush_l(lw, la0);
    flush_l(lw + 1, la1);
  }
  if (pa != EMPTY) {
    flush_m(pa, pb, m00);
    flush_m(pa, pb + 1, m01);
    flush_m(pa + 1, pb, m10);
    flush_m(pa + 1, pb + 1, m11);
  }
}

// (Measured and dropped: a density walk with TWO float4 channel groups per thread, so that the scalar work of a sample -
// coordinates, taps, window bookkeeping, ~250 of the 555 instructions per item - is paid once per 32 bytes: 128 registers,
// 2 CTAs per SM, 0.151 -> 0.220 ms.  The walk needs its 24 resident warps per SM more than it needs fewer instructions.)
static int pick_lps(int nvec) {
  for (int p = 8; p >= 1; p >>= 1)
    if (nvec % p == 0) return p;
  return 1;
}

template <bool APP>
static int launch_walk(cudaStream_t st, WalkArgs A) {
  if (A.R == 0 || A.count == 0) return 0;
  const int nvec = A.Cp / 4;
  int lps = pick_lps(nvec);
  if (lps == 1 && nvec > 2) lps = 4;  // odd channel counts: round the block up, idle lanes exit
  lps = min(lps, 4);                  // 64-byte texel fragments per item keep more items in flight
  A.seg_len = APP ? 128 : 16;  // measured on B200: density 16 (lego 0.151 ms; 8: 0.158, 32: 0.164); appearance 128 (dozer K=99: 0.089 ms vs 0.107 at 64; lego K=38 unchanged)
  if (const char* e = getenv(APP ? "TENSORF_SEG_LEN_APP" : "TENSORF_SEG_LEN")) A.seg_len = std::max(8, atoi(e));
  A.segs = (A.count + A.seg_len - 1) / A.seg_len;
  const int vblocks = (nvec + lps - 1) / lps;
  int64_t items = (int64_t)A.R * A.segs * 3 * vblocks;
  // tuning switches: TENSORF_DSW_BLOCK / TENSORF_ASW_BLOCK (CTA size of the density / appearance walk), TENSORF_ASW_SLIM=0
  static const int den_block = getenv("TENSORF_DSW_BLOCK") ? atoi(getenv("TENSORF_DSW_BLOCK")) : 64;
  static const int app_block = getenv("TENSORF_ASW_BLOCK") ? atoi(getenv("TENSORF_ASW_BLOCK")) : 128;
  const int block = APP ? app_block : den_block;
  unsigned grid = (unsigned)ceil_div64(items, block / lps);
  static const int slim_env = getenv("TENSORF_ASW_SLIM") ? atoi(getenv("TENSORF_ASW_SLIM")) : 1;
  if (APP && A.slim && slim_env && lps == 4) {
    k_scatter_walk<4, APP, APP><<<grid, block, 0, st>>>(A);
  } else {
    switch (lps) {
      case 4: k_scatter_walk<4, APP><<<grid, block, 0, st>>>(A); break;
      case 2: k_scatter_walk<2, APP><<<grid, block, 0, st>>>(A); break;
      default: k_scatter_walk<1, APP><<<grid, block, 0, st>>>(A); break;
    }
  }
  TF_CHECK_LAUNCH();
  return 0;
}

int launch_density_scatter(cudaStream_t st, const DensityBwdArgs& D) {
  WalkArgs A{};
  static_cast<SceneArgs&>(A) = static_cast<const SceneArgs&>(D);
  A.packed = D.packed_d;
  A.d_packed = D.d_packed;
  A.dz = D.dz;
  A.xs = D.xs;
  A.C = D.Cp;
  A.Cp = D.Cp;
  A.count = D.N;
  return launch_walk<false>(st, A);
}

int launch_appearance_scatter(cudaStream_t st, const AppearanceArgs& D) {
  WalkArgs A{};
  static_cast<SceneArgs&>(A) = static_cast<const SceneArgs&>(D);
  A.packed = D.packed_a;
  A.d_packed = D.d_packed;
  A.idx = D.idx;
  A.xs = D.xs;
  A.d_feat = D.d_feat;
  A.C = D.C;
  A.Cp = D.Cp;
  A.count = D.K;
  A.slim = D.beside_mlp;
  return launch_walk<true>(st, A);
}

}  // namespace tf
