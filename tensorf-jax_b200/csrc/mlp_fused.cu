// FeatureMlp (networks.py:38-121) as per-row-tile fused tcgen05 kernels: activations never leave the SM between the
// layers of a pass.  See umma_tiles.cuh for the operand format (two-term 16-bit splits in "slab" tiles).
//
//   k_umma_probe : one 128 x N x K product with every operand source / orientation / format the fused kernels use
//                  (tests/test_gpu_fused.py pins the descriptor conventions on the hardware).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <cuda.h>

#include "mlp.cuh"
#include "umma_tiles.cuh"

namespace tf {

using namespace tc;

// ---------------------------------------------------------------------------------------------------------------
// k_umma_probe: D[128 x N] = A[128 x K] * B[N x K]^T, two-term split operands, three products per k-step.
//   a_mode 0: A tile rows = m, columns = k (K-major)      1: tile rows = k, columns = m (MN-major)     2: A in tensor memory
//   b_mode 0: B tile rows = n, columns = k (K-major)      1: tile rows = k, columns = n (MN-major)
// ---------------------------------------------------------------------------------------------------------------
template <int FMT>
__device__ void probe_fill(unsigned char* tile, int tile_rows, const float* src, int R, int Ccols, bool transposed, int tid, int nthr) {
  // logical matrix src[R][Ccols]; tile rows index r (or c when transposed)
  const int trows = transposed ? Ccols : R, tcols = transposed ? R : Ccols;
  (void)tile_rows;
  for (int item = tid; item < trows * (tcols / 8); item += nthr) {
    const int tr = item % trows, c8 = item / trows;
    float x[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int tc_ = c8 * 8 + q;
      x[q] = transposed ? src[tc_ * Ccols + tr] : src[tr * Ccols + tc_];
    }
    uint4 hi, lo;
    split8_fmt<FMT>(x, hi, lo);
    *reinterpret_cast<uint4*>(tile + slab_offset(trows, 0, tr, c8 * 8)) = hi;
    *reinterpret_cast<uint4*>(tile + slab_offset(trows, 1, tr, c8 * 8)) = lo;
  }
}

__global__ void __launch_bounds__(128, 1) k_umma_probe(const float* A, const float* B, float* D, int N, int K, int a_mode, int b_mode,
                                                      int a_fmt, int b_fmt) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 64);
  unsigned char* sA = smem + 1024;
  unsigned char* sB = sA + slab_tile_bytes(128, 160);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  const uint32_t tm_d = tm + 256;

  if (a_mode == 2) {  // thread = row: packed pairs, hi columns [0, K/2), lo columns [K/2, K)
    for (int c0 = 0; c0 < K; c0 += 16) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float x0 = A[tid * K + c0 + 2 * q], x1 = A[tid * K + c0 + 2 * q + 1];
        if (a_fmt == FMT_F16) split_pair<FMT_F16>(x0, x1, hi[q], lo[q]);
        else split_pair<FMT_BF16>(x0, x1, hi[q], lo[q]);
      }
      const uint32_t t = tm + ((uint32_t)(warp * 32) << 16);
      tmem_st8(t + c0 / 2, hi);
      tmem_st8(t + K / 2 + c0 / 2, lo);
    }
    tmem_st_wait();
  } else if (a_fmt == FMT_F16) {
    probe_fill<FMT_F16>(sA, 0, A, 128, K, a_mode == 1, tid, 128);
  } else {
    probe_fill<FMT_BF16>(sA, 0, A, 128, K, a_mode == 1, tid, 128);
  }
  if (b_fmt == FMT_F16) probe_fill<FMT_F16>(sB, 0, B, N, K, b_mode == 1, tid, 128);
  else probe_fill<FMT_BF16>(sB, 0, B, N, K, b_mode == 1, tid, 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();

  if (warp == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(128, N, a_fmt, b_fmt, a_mode == 1, b_mode == 1);
    // A: K-major tile has 128 rows, MN-major tile has K rows
    const uint32_t a_rows = a_mode == 1 ? (uint32_t)K : 128u, b_rows = b_mode == 1 ? (uint32_t)K : (uint32_t)N;
    const uint32_t a_lo0 = desc_lo(smem_u32(sA), a_mode == 1 ? 128u : 2u * slab_bytes(a_rows));
    const uint32_t a_hi = desc_hi(a_mode == 1 ? 2u * slab_bytes(a_rows) : 128u);
    const uint32_t b_lo0 = desc_lo(smem_u32(sB), b_mode == 1 ? 128u : 2u * slab_bytes(b_rows));
    const uint32_t b_hi = desc_hi(b_mode == 1 ? 2u * slab_bytes(b_rows) : 128u);
    const uint32_t a_term = slab_bytes(a_rows) >> 4, b_term = slab_bytes(b_rows) >> 4;
    // one k-step = 16 k: K-major -> two column groups (4 slabs), MN-major -> 16 rows
    const uint32_t a_step = (a_mode == 1 ? 256u : 4u * slab_bytes(a_rows)) >> 4;
    const uint32_t b_step = (b_mode == 1 ? 256u : 4u * slab_bytes(b_rows)) >> 4;
    if (elect_one()) {
      for (int ks = 0; ks < K / 16; ++ks) {
        if (a_mode == 2)
          umma_ts_split2(tm_d, tm + ks * 8, K / 2, b_lo0 + ks * b_step, b_term, b_hi, idesc, ks == 0);
        else
          umma_ss_split2(tm_d, a_lo0 + ks * a_step, a_term, a_hi, b_lo0 + ks * b_step, b_term, b_hi, idesc, ks == 0);
      }
      umma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int n0 = 0; n0 < N; n0 += 16) {
    float v[16];
    tmem_ld16(tm_d + ((uint32_t)(warp * 32) << 16) + n0, v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 16; ++q) D[(warp * 32 + lane) * N + n0 + q] = v[q];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tm, 512);
  }
}

int umma_probe(cudaStream_t st, const float* A, const float* B, float* D, int N, int K, int a_mode, int b_mode, int a_fmt, int b_fmt) {
  TF_CHECK_ARG(N >= 16 && N <= 128 && N % 16 == 0 && K >= 16 && K <= 160 && K % 16 == 0, "umma_probe: N=%d K=%d unsupported", N, K);
  TF_CHECK_ARG(a_mode >= 0 && a_mode <= 2 && b_mode >= 0 && b_mode <= 1 && (a_fmt | b_fmt) >= 0 && (a_fmt | b_fmt) <= 1, "umma_probe: bad mode");
  const size_t smem = 1024 + 2 * (size_t)slab_tile_bytes(128, 160);
  TF_CHECK_CUDA(cudaFuncSetAttribute(k_umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_umma_probe<<<1, 128, smem, st>>>(A, B, D, N, K, a_mode, b_mode, a_fmt, b_fmt);
  TF_CHECK_LAUNCH();
  return 0;
}


// ===============================================================================================================
// Fused FeatureMlp for the lego-type network (networks.py:38-121 with feature_squash_dim 27, units 128,
// feature_n_freqs = viewdir_n_freqs = 2, no camera embeddings; 3*ca a multiple of 16, <= 160).
//
// Encoded input in kernel order ("x'", 160 columns): column 80*s + 5*i + v holds, for source dimension d = 15*s + i
// (d < 27: squashed feature f_d, d >= 27: view direction component d - 27), v = 0: the value, 1: sin, 2: sin 2x,
// 3: cos, 4: cos 2x.  Columns 75..79 and 155..159 are padding; column 75 is the constant 1 (its Dense_1 weight row
// is zero; in the weight-gradient product it yields the bias gradient as one more column sum).  fz_perm maps a
// kernel column to the reference's column of networks.py:68-76 ([f, v, enc(f), enc(v)]).
// ===============================================================================================================
// Residual / gradient slab tiles are written once and read back a kernel or more later: streaming stores (evict-first in
// the L2), so that they do not displace the factor and gradient planes the gather / scatter kernels keep there.
#define TF_SLAB_ST(ptr, val) __stcs((ptr), (val))
namespace fz {
constexpr int kU = 128, kSq = 27, kQDims = 8, kPer = 5, kXQ = 40, kX = 160, kOnesCol = 150;
constexpr int kWorkWarps = 16, kMmaWarp = 16, kLoadWarp = 17, kThreads = 576;
constexpr int kRawSlots = 3, kRawBytes = 128 * 32 * 4, kHeader = 8192;  // feature ring: pieces of 32 columns (4 column groups x hi, lo slab)
constexpr uint32_t kAccX = 0, kAccY = 128, kAop = 256, kAopLo = 80, kAccD0 = 416;  // tensor-memory columns (Dense_0: 2 x 32 at kAccD0)
constexpr uint32_t kB1Bytes = 2u * kX * kU * 2u, kB2Bytes = 2u * kU * kU * 2u;
// kernel column 40*q + 5*i + v of x' (source dimension d = 8*q + i; d = 30, 31 are padding, column 150 is the constant 1)
// -> column of the reference's [f, v, enc(f), enc(v)] (networks.py:68-76), or -1
__host__ __device__ inline int perm(int kp) {
  const int d = kQDims * (kp / kXQ) + (kp % kXQ) / kPer, v = kp % kPer;
  if (d >= 30) return -1;
  return v == 0 ? d : 30 + 4 * d + (v - 1);
}
// Dense_0 output columns are in the reference's order (quarter q of the workers reads columns 8q..8q+7; 27..31 are zero)
__host__ __device__ inline int sq_nat(int np) { return np < kSq ? np : -1; }
__host__ __device__ inline uint32_t b0_bytes(int K0) { return (uint32_t)K0 * 128u; }  // 32 rows x K0 columns x 2 terms x 2 B
// x' k-chunk c (16 columns) is complete after this many warp arrivals: chunks 2 and 7 straddle two quarters
__host__ __device__ inline uint32_t x_chunk_warps(int c) { return (c == 2 || c == 7) ? 8u : 4u; }
}  // namespace fz

bool mlp_fused_supported(const MlpShape& s) {
  return s.squash == fz::kSq && s.units == fz::kU && s.Ff == 2 && s.Fv == 2 && s.ncam == 0 && s.Ca % 16 == 0 && s.Ca >= 16 && s.Ca <= 160;
}
size_t mlp_fused_wpack_bytes(const MlpShape& s) { return (size_t)fz::b0_bytes(s.Ca) + fz::kB1Bytes + fz::kB2Bytes; }

// Weights -> fp16 two-term slab tiles [B0 | B1 | B2]: tile rows = output unit n, columns = input k (K-major B of the
// forward; the same bytes are the MN-major B of the reverse products, umma_tiles.cuh).
struct FusedPackArgs {
  const float *w0, *w1, *w2, *b1;
  int Ca;
  unsigned char* out;
};
__global__ void __launch_bounds__(256) k_fused_pack(FusedPackArgs a) {
  const int K0 = a.Ca;
  const int n0 = 32 * (K0 / 8), n1 = fz::kU * (fz::kX / 8), n2 = fz::kU * (fz::kU / 8);
  int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n0 + n1 + n2) return;
  float x[8];
  unsigned char* tile;
  int rows, n, c8;
  if (item < n0) {
    rows = 32; n = item % 32; c8 = item / 32; tile = a.out;
#pragma unroll
    for (int q = 0; q < 8; ++q) x[q] = fz::sq_nat(n) >= 0 ? a.w0[(c8 * 8 + q) * fz::kSq + fz::sq_nat(n)] : 0.f;
  } else if (item < n0 + n1) {
    item -= n0;
    rows = fz::kU; n = item % fz::kU; c8 = item / fz::kU; tile = a.out + fz::b0_bytes(K0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int nat = fz::perm(c8 * 8 + q);
      // the row of x's constant-1 column carries Dense_1's bias: the tensor core adds it
      x[q] = nat >= 0 ? a.w1[nat * fz::kU + n] : (c8 * 8 + q == fz::kOnesCol ? a.b1[n] : 0.f);
    }
  } else {
    item -= n0 + n1;
    rows = fz::kU; n = item % fz::kU; c8 = item / fz::kU; tile = a.out + fz::b0_bytes(K0) + fz::kB1Bytes;
#pragma unroll
    for (int q = 0; q < 8; ++q) x[q] = a.w2[(c8 * 8 + q) * fz::kU + n];
  }
  uint4 hi, lo;
  split8_fmt<FMT_F16>(x, hi, lo);
  *reinterpret_cast<uint4*>(tile + slab_offset(rows, 0, n, c8 * 8)) = hi;
  *reinterpret_cast<uint4*>(tile + slab_offset(rows, 1, n, c8 * 8)) = lo;
}

struct FusedFwdArgs {
  const unsigned char* wpack;
  const float *b1, *b2, *w3, *b3;
  const float* viewdirs;
  int rows_per_ray;
  int64_t M;
  int K0;  // 3*ca
  float* rgb;
  // residuals of the reverse pass (TRAIN): slab tiles per 128-row tile, ReLU masks, Dense_0 output
  const unsigned char* feats;  // feature rows as slab tiles
  unsigned char *xs, *h1s, *h2s;
  float* fs;
  uint32_t *bits1, *bits2;
};

// 16 fp32 values of one row -> fp16 (hi, lo) words; hi[0..3] / lo[0..3] = columns 0-7, [4..7] = columns 8-15
__device__ __forceinline__ void split16(const float* y, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int q = 0; q < 8; ++q) split_pair<FMT_F16>(y[2 * q], y[2 * q + 1], hi[q], lo[q]);
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// "this warp's part of an A-operand chunk is in tensor memory": make the stores visible to the MMA warp, then arrive
__device__ __forceinline__ void chunk_arrive(uint64_t* bar, int lane) {
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
// 8 columns (one slab column group cg) of one row: the next layer's A operand in tensor memory (hi at packed column
// 4*cg, lo at 80 + 4*cg of the operand region) and, for the reverse pass, the slab tile in global memory
template <bool TRAIN>
__device__ __forceinline__ void emit8(const float* y, uint32_t aop_lane, int cg, unsigned char* slab_row) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) split_pair<FMT_F16>(y[2 * q], y[2 * q + 1], hi[q], lo[q]);
  tmem_st4(aop_lane + 4u * cg, hi);
  tmem_st4(aop_lane + fz::kAopLo + 4u * cg, lo);
  if (TRAIN) {
    uint4* p = reinterpret_cast<uint4*>(slab_row + (size_t)(2 * cg) * 2048);
    TF_SLAB_ST(&p[0], make_uint4(hi[0], hi[1], hi[2], hi[3]));
    TF_SLAB_ST(&p[128], make_uint4(lo[0], lo[1], lo[2], lo[3]));
  }
}
// sin / cos of a squashed feature or view-direction component (networks.py:13-35: sin(2^j x), sin(2^j x + pi/2), j < 2):
// special-function unit after an explicit reduction to [-pi, pi] (absolute error ~5e-7), second level by double angle
__device__ __forceinline__ void fast_sincos(float x, float& sn, float& cs) {
  const float k = rintf(x * 0.15915494309189535f);
  const float r = fmaf(k, -1.7484555e-7f, fmaf(k, -6.2831854820251465f, x));  // x - k * 2pi in two terms
  sn = __sinf(r);
  cs = __cosf(r);
}

// Dense_0 accumulator -> Fourier-encoded input x' (this quarter's 40 columns)
template <bool TRAIN>
__device__ __forceinline__ void fused_epi0(const FusedFwdArgs& g, int Q, uint32_t acc_lane, uint32_t aop_lane, int64_t m, int64_t tile, int row,
                                           uint64_t* xready, int lane) {
  float f[8];
  tmem_ld8(acc_lane + 8u * Q, f);
  tmem_ld_wait();
  if (TRAIN) {
    float4* fp = reinterpret_cast<float4*>(g.fs + (size_t)tile * 4096 + (size_t)(2 * Q) * 512) + row;
    fp[0] = make_float4(f[0], f[1], f[2], f[3]);
    fp[128] = make_float4(f[4], f[5], f[6], f[7]);
  }
  if (Q == 3) {  // dimensions 27..29 are the view direction, 30..31 padding
    f[3] = f[4] = f[5] = f[6] = f[7] = 0.f;
    if (m < g.M) {
      const float* v = g.viewdirs + (m / g.rows_per_ray) * 3;
      f[3] = __ldg(v); f[4] = __ldg(v + 1); f[5] = __ldg(v + 2);
    }
  }
  unsigned char* slab_row = TRAIN ? g.xs + (size_t)tile * (fz::kX * 512) + (size_t)row * 16 : nullptr;
  float y[40];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float sn, cs;
    fast_sincos(f[i], sn, cs);
    y[5 * i] = f[i];
    y[5 * i + 1] = sn;
    y[5 * i + 2] = 2.0f * sn * cs;
    y[5 * i + 3] = cs;
    y[5 * i + 4] = fmaf(-2.0f * sn, sn, 1.0f);
  }
  if (Q == 3) {
#pragma unroll
    for (int e = 30; e < 40; ++e) y[e] = e == 30 ? 1.0f : 0.0f;  // column 150: the constant 1 (bias-gradient column of x')
  }
  // 8-column units: [40Q + 8u, +8) = slab column group 5Q + u of chunk (5Q + u) / 2
#pragma unroll
  for (int u = 0; u < 5; ++u) {
    const int cg = 5 * Q + u;
    emit8<TRAIN>(y + 8 * u, aop_lane, cg, slab_row);
    if ((cg & 1) || u == 4) chunk_arrive(&xready[cg >> 1], lane);  // the chunk's last unit of this quarter
  }
}

// hidden-layer accumulator -> bias, ReLU, mask bits -> next layer's A operand (this quarter's 32 columns = 2 k-chunks).
// LAST: no next layer; the output layer (networks.py:114-120) is accumulated into o3 instead.
template <bool TRAIN, bool LAST>
__device__ __forceinline__ void fused_epi_hidden(int Q, uint32_t acc_lane, uint32_t aop_lane, const float* s_bias /* null: already in the product */, const float* s_w3, unsigned char* slab_row,
                                                 uint32_t* bits_row, uint64_t* kready, int lane, float* o3) {
  uint32_t bw = 0u;
  float y[32];
  tmem_ld16(acc_lane + 32 * Q, y);
  tmem_ld16(acc_lane + 32 * Q + 16, y + 16);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int n0 = 32 * Q + 16 * c;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float v = fmaxf(LAST ? y[16 * c + q] + s_bias[n0 + q] : y[16 * c + q], 0.f);  // Dense_1's bias rides in W1' (ones column of x')
      y[16 * c + q] = v;
      if (TRAIN && v > 0.f) bw |= 1u << (16 * c + q);
      if (LAST) {
        o3[0] = fmaf(v, s_w3[3 * (n0 + q) + 0], o3[0]);
        o3[1] = fmaf(v, s_w3[3 * (n0 + q) + 1], o3[1]);
        o3[2] = fmaf(v, s_w3[3 * (n0 + q) + 2], o3[2]);
      }
    }
    if (!LAST) {
      emit8<TRAIN>(y + 16 * c, aop_lane, 4 * Q + 2 * c, slab_row);
      emit8<TRAIN>(y + 16 * c + 8, aop_lane, 4 * Q + 2 * c + 1, slab_row);
      chunk_arrive(&kready[2 * Q + c], lane);
    } else if (TRAIN) {
      uint32_t hi[8], lo[8];
      split16(y + 16 * c, hi, lo);
      uint4* p = reinterpret_cast<uint4*>(slab_row + (size_t)(4 * (2 * Q + c)) * 2048);
      TF_SLAB_ST(&p[0], make_uint4(hi[0], hi[1], hi[2], hi[3]));
      TF_SLAB_ST(&p[128], make_uint4(lo[0], lo[1], lo[2], lo[3]));
      TF_SLAB_ST(&p[256], make_uint4(hi[4], hi[5], hi[6], hi[7]));
      TF_SLAB_ST(&p[384], make_uint4(lo[4], lo[5], lo[6], lo[7]));
    }
  }
  if (TRAIN) bits_row[Q] = bw;
}

// The column quarter Q is a RUN-TIME value on purpose: with four template instances the kernel was 133 KB of SASS and 43 % of
// its warp stalls were instruction fetches (ncu: stall_no_inst) - sixteen warps walking four copies of the same unrolled
// epilogues do not fit the instruction cache.
template <bool TRAIN>
__device__ __forceinline__ void fused_fwd_worker(const FusedFwdArgs& g, int Q, uint32_t tmem, int quad, int lane, uint64_t* xready, uint64_t* kready,
                                                 uint64_t* accfull, const float* s_b1, const float* s_b2, const float* s_w3, float* s_part) {
  const int64_t ntiles = (g.M + 127) / 128;
  const int row = quad * 32 + lane;
  const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
  const uint32_t accX = lane_base + fz::kAccX, accY = lane_base + fz::kAccY, aop = lane_base + fz::kAop;
  // Dense_0 accumulator of tile number `it` -> x' (the MMA warp runs Dense_0 one tile ahead)
  auto encode = [&](int64_t t, uint32_t it) {
    mbar_wait(&accfull[0], it & 1, 10);
    tc_fence_after();
    fused_epi0<TRAIN>(g, Q, lane_base + fz::kAccD0 + 32u * (it & 1), aop, t * 128 + row, t, row, xready, lane);
  };
  const int64_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  // Software pipeline over the CTA's tiles, written so that every epilogue appears ONCE in the code (instruction cache):
  // step k finishes Dense_1 of tile k-1, encodes tile k, then finishes tile k-1.
  for (int64_t k = 0; k <= my_tiles; ++k) {
    const int64_t t = blockIdx.x + (k - 1) * (int64_t)gridDim.x;  // the tile being finished (k >= 1)
    const int64_t m = t * 128 + row;
    const uint32_t ph = (uint32_t)(k - 1) & 1u;
    float o3[3] = {0.f, 0.f, 0.f};
    if (k > 0) {
      // Dense_1 -> h1
      mbar_wait(&accfull[1], ph, 11);
      tc_fence_after();
      fused_epi_hidden<TRAIN, false>(Q, accX, aop, s_b1, s_w3, TRAIN ? g.h1s + (size_t)t * (fz::kU * 512) + (size_t)row * 16 : nullptr,
                                     TRAIN ? g.bits1 + m * 4 : nullptr, kready, lane, nullptr);
      // Dense_2 complete: h1 in the operand region is dead.  The NEXT tile's x' goes in first, so that its Dense_1 runs on
      // the tensor pipe while this tile's output layer is evaluated on the CUDA cores.
      mbar_wait(&accfull[2], ph, 12);
      tc_fence_after();
    }
    if (k < my_tiles) encode(blockIdx.x + k * (int64_t)gridDim.x, (uint32_t)k);
    if (k > 0) {
      // Dense_2 -> h2 -> Dense_3 + sigmoid
      fused_epi_hidden<TRAIN, true>(Q, accY, aop, s_b2, s_w3, TRAIN ? g.h2s + (size_t)t * (fz::kU * 512) + (size_t)row * 16 : nullptr,
                                    TRAIN ? g.bits2 + m * 4 : nullptr, kready, lane, o3);
      tc_fence_before();
      if (Q != 0) {
        float* sp = s_part + (Q - 1) * 384 + 3 * row;
        sp[0] = o3[0]; sp[1] = o3[1]; sp[2] = o3[2];
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");  // also: every quarter has read accumulator Y before the next Dense_2 can start
      if (Q == 0 && m < g.M) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          g.rgb[3 * m + c] = 1.0f / (1.0f + expf(-(o3[c] + s_part[3 * row + c] + s_part[384 + 3 * row + c] + s_part[768 + 3 * row + c] + s_w3[384 + c])));
      }
    }
  }
}

template <bool TRAIN>
__global__ void __launch_bounds__(fz::kThreads, 1) k_mlp_fused_fwd(FusedFwdArgs g) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* wfull = reinterpret_cast<uint64_t*>(smem);
  uint64_t* rfull = wfull + 1;               // [kRawSlots] feature piece landed (bulk copy tx bytes)
  uint64_t* rempty = rfull + fz::kRawSlots;  // [kRawSlots] Dense_0 MMAs have read the piece (tcgen05.commit)
  uint64_t* kready = rempty + fz::kRawSlots; // [8]  h1 chunks (Dense_2's A operand): one quarter each
  uint64_t* xready = kready + 8;             // [10] x' chunks (Dense_1's A operand): one or two quarters each
  uint64_t* accfull = xready + 10;           // [3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 3);
  float* s_b1 = reinterpret_cast<float*>(smem + 512);
  float* s_b2 = s_b1 + 128;
  float* s_w3 = s_b2 + 128;   // [384 + 3]
  float* s_part = s_w3 + 388; // [3][384] partial output-layer sums of quarters 1..3 (ends at byte 7696 < kHeader)
  const uint32_t wbytes = fz::b0_bytes(g.K0) + fz::kB1Bytes + fz::kB2Bytes;
  unsigned char* sW = smem + fz::kHeader;
  unsigned char* raw0 = smem + ((fz::kHeader + wbytes + 1023u) & ~1023u);
  const int nk0 = g.K0 / 16, npiece = (g.K0 + 31) / 32;
  const int64_t ntiles = (g.M + 127) / 128;
  const int64_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (tid == 0) {
    mbar_init(wfull, 1);
    for (int r = 0; r < fz::kRawSlots; ++r) {
      mbar_init(&rfull[r], 1);
      mbar_init(&rempty[r], 1);
    }
    for (int c = 0; c < 8; ++c) mbar_init(&kready[c], 4);
    for (int c = 0; c < 10; ++c) mbar_init(&xready[c], fz::x_chunk_warps(c));
    for (int l = 0; l < 3; ++l) mbar_init(&accfull[l], 1);
    fence_barrier_init();
  }
  for (int n = tid; n < 128; n += fz::kThreads) {
    s_b1[n] = g.b1[n];
    s_b2[n] = g.b2[n];
  }
  for (int n = tid; n < 387; n += fz::kThreads) s_w3[n] = n < 384 ? g.w3[n] : g.b3[n - 384];
  if (warp == fz::kMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < fz::kWorkWarps) {
    // ================= row workers: thread = row; quarter q (warps 4q..4q+3) owns a quarter of every layer's columns ========
    fused_fwd_worker<TRAIN>(g, warp >> 2, tmem, warp & 3, lane, xready, kready, accfull, s_b1, s_b2, s_w3, s_part);
  } else if (warp == fz::kLoadWarp) {
    // ================= loader: weights once, then the feature tile of every row tile in 16 KB pieces (32 columns) =============
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull, wbytes);
      for (uint32_t off = 0; off < wbytes; off += 16384u)
        bulk_copy_g2s(sW + off, g.wpack + off, min(16384u, wbytes - off), wfull);
      pdl_wait();  // the feature tiles are the predecessor's output (the weights were packed earlier in the stream)
      uint32_t sl = 0, ph = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        const unsigned char* tile = g.feats + (size_t)(blockIdx.x + it * gridDim.x) * ((size_t)g.K0 * 512);
        for (int p = 0; p < npiece; ++p) {
          const uint32_t bytes = (uint32_t)min(32, g.K0 - 32 * p) * 512u;
          mbar_wait(&rempty[sl], ph ^ 1, 30);
          mbar_arrive_expect_tx(&rfull[sl], bytes);
          bulk_copy_g2s(raw0 + (size_t)sl * fz::kRawBytes, tile + (size_t)p * fz::kRawBytes, bytes, &rfull[sl]);
          if (++sl == fz::kRawSlots) { sl = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ================= MMA issuer (warp-uniform loop, one elected lane issues) =================
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc0 = make_idesc(128, 32, FMT_F16, FMT_F16, 0, 0), idesc1 = make_idesc(128, 128, FMT_F16, FMT_F16, 0, 0);
    const uint32_t sw = smem_u32(sW), sr = smem_u32(raw0);
    const uint32_t b0_lo = desc_lo(sw, 2u * slab_bytes(32)), b1_lo = desc_lo(sw + fz::b0_bytes(g.K0), 2u * slab_bytes(128)),
                   b2_lo = desc_lo(sw + fz::b0_bytes(g.K0) + fz::kB1Bytes, 2u * slab_bytes(128));
    const uint32_t b_hi = desc_hi(128u);  // K-major slabs: 8-row group stride 128 B (A tiles of 128 rows use the same high word)
    const uint32_t t32 = slab_bytes(32) >> 4, s32 = (4u * slab_bytes(32)) >> 4, t128 = slab_bytes(128) >> 4, s128 = (4u * slab_bytes(128)) >> 4;
    const uint32_t aop = tm + fz::kAop, accX = tm + fz::kAccX, accY = tm + fz::kAccY;
    uint32_t kph = 0, xph = 0, sl = 0, rph = 0;
    // Dense_0 of tile `it`: A = the feature pieces in the shared-memory ring (K-major slab tiles), one 32-column accumulator per
    // tile parity.  It is issued one tile AHEAD (between Dense_1 and Dense_2 of the previous tile), so the features of the
    // next tile are multiplied while the workers are busy with this one.
    auto dense0 = [&](int64_t it) {
      const uint32_t acc = tm + fz::kAccD0 + 32u * (uint32_t)(it & 1);
      for (int p = 0; p < npiece; ++p) {
        mbar_wait(&rfull[sl], rph, 41);
        tc_fence_after();
        const uint32_t a_lo = desc_lo(sr + sl * fz::kRawBytes, 2u * slab_bytes(128));
        if (elect_one()) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = 2 * p + h;
            if (c < nk0) umma_ss_split2(acc, a_lo + h * s128, t128, b_hi, b0_lo + c * s32, t32, b_hi, idesc0, c == 0);
          }
          umma_commit(&rempty[sl]);
        }
        __syncwarp();
        if (++sl == fz::kRawSlots) { sl = 0; rph ^= 1; }
      }
      if (elect_one()) umma_commit(&accfull[0]);
      __syncwarp();
    };
    mbar_wait(wfull, 0, 40);
    dense0(0);
    for (int64_t it = 0; it < my_tiles; ++it) {
      for (int i = 0; i < 10; ++i) {  // Dense_1: chunks roughly in the order the four quarters complete them
        const int c = (0x7294618350ull >> (4 * i)) & 15;  // 0,5,3,8,1,6,4,9,2,7
        mbar_wait(&xready[c], (xph >> c) & 1u, 42);
        xph ^= 1u << c;
        tc_fence_after();
        if (elect_one()) umma_ts_split2(accX, aop + 8u * c, fz::kAopLo, b1_lo + c * s128, t128, b_hi, idesc1, i == 0);
        __syncwarp();
      }
      if (elect_one()) umma_commit(&accfull[1]);
      __syncwarp();
      if (it + 1 < my_tiles) dense0(it + 1);
      for (int i = 0; i < 8; ++i) {  // Dense_2: quarter q completes chunk 2q, then 2q + 1
        const int c = (i & 3) * 2 + (i >> 2);
        mbar_wait(&kready[c], (kph >> c) & 1u, 43);
        kph ^= 1u << c;
        tc_fence_after();
        if (elect_one()) umma_ts_split2(accY, aop + 8u * c, fz::kAopLo, b2_lo + c * s128, t128, b_hi, idesc1, i == 0);
        __syncwarp();
      }
      if (elect_one()) umma_commit(&accfull[2]);
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == fz::kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// fp32 feature rows (M, K0) -> slab tiles (standalone tensorf_mlp_fwd; the render path's k_appearance writes slabs itself)
__global__ void __launch_bounds__(256) k_feat_to_slab(const float* feat, int64_t M, int K0, unsigned char* out) {
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (row, column group)
  const int ncg = K0 / 8;
  const int64_t m = item / ncg;
  const int cg = (int)(item % ncg);
  if (m >= round_up64(M, 128)) return;
  float x[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) x[q] = m < M ? feat[m * K0 + cg * 8 + q] : 0.f;
  uint4 hi, lo;
  split8_fmt<FMT_F16>(x, hi, lo);
  unsigned char* tile = out + (m >> 7) * ((int64_t)K0 * 512) + (m & 127) * 16;
  *reinterpret_cast<uint4*>(tile + (int64_t)(2 * cg) * 2048) = hi;
  *reinterpret_cast<uint4*>(tile + (int64_t)(2 * cg + 1) * 2048) = lo;
}

int mlp_fused_pack(cudaStream_t st, const MlpShape& s, const MlpParams& p, const MlpWs& ws) {
  TF_CHECK_ARG(mlp_fused_supported(s), "fused MLP: unsupported network shape");
  TF_CHECK_ARG(ws.wpack_bytes >= mlp_fused_wpack_bytes(s), "fused MLP: weight scratch too small");
  FusedPackArgs a{p.w0, p.w1, p.w2, p.b1, s.Ca, ws.wpack};
  const int items = 32 * (s.Ca / 8) + fz::kU * (fz::kX / 8) + fz::kU * (fz::kU / 8);
  k_fused_pack<<<(items + 255) / 256, 256, 0, st>>>(a);
  TF_CHECK_LAUNCH();
  return 0;
}

int mlp_fused_fwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const float* feat, const float* viewdirs, const uint32_t* cams,
                  int64_t M, int rows_per_ray, const MlpWs& ws, float* rgb) {
  (void)cams;
  if (M == 0) return 0;
  if (!ws.wpack_ready) TF_RETURN_IF_ERROR(mlp_fused_pack(st, s, p, ws));
  const unsigned char* slabs = ws.feat_slabs;
  if (slabs == nullptr) {  // fp32 rows from the caller: convert once (the reverse pass reads the same tiles)
    unsigned char* out = reinterpret_cast<unsigned char*>(ws.dx);
    const int64_t items = round_up64(M, 128) * (s.Ca / 8);
    k_feat_to_slab<<<(unsigned)ceil_div64(items, 256), 256, 0, st>>>(feat, M, s.Ca, out);
    TF_CHECK_LAUNCH();
    slabs = out;
  }
  FusedFwdArgs g{};
  g.wpack = ws.wpack; g.b1 = p.b1; g.b2 = p.b2; g.w3 = p.w3; g.b3 = p.b3;
  g.viewdirs = viewdirs; g.rows_per_ray = rows_per_ray; g.M = M; g.K0 = s.Ca; g.rgb = rgb; g.feats = slabs;
  const bool train = !s.inference;
  if (train) {
    g.xs = reinterpret_cast<unsigned char*>(ws.x);
    g.h1s = reinterpret_cast<unsigned char*>(ws.h1);
    g.h2s = reinterpret_cast<unsigned char*>(ws.h2);
    g.fs = ws.f;
    g.bits1 = ws.bits1;
    g.bits2 = ws.bits1 + round_up64(M, 128) * 4;
  }
  const size_t smem = ((fz::kHeader + mlp_fused_wpack_bytes(s) + 1023) & ~(size_t)1023) + (size_t)fz::kRawSlots * fz::kRawBytes;
  TF_CHECK_ARG(smem <= 227 * 1024, "fused MLP: shared memory budget exceeded");
  const unsigned grid = (unsigned)std::min<int64_t>((M + 127) / 128, kSMs);
  if (train) {
    TF_CHECK_CUDA(cudaFuncSetAttribute(k_mlp_fused_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TF_CHECK_CUDA(launch_pdl(k_mlp_fused_fwd<true>, dim3(grid), dim3(fz::kThreads), smem, st, g, ws.pdl));
  } else {
    TF_CHECK_CUDA(cudaFuncSetAttribute(k_mlp_fused_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TF_CHECK_CUDA(launch_pdl(k_mlp_fused_fwd<false>, dim3(grid), dim3(fz::kThreads), smem, st, g, ws.pdl));
  }
  TF_CHECK_LAUNCH();
  return 0;
}

// ===============================================================================================================
// Reverse pass, kernel 1 of 2: the activation-gradient chain of one 128-row tile,
//   d(out) -> dp2 = (d(out) W3^T) * relu'(h2) -> dp1 = (dp2 W2^T) * relu'(h1) -> dx' = dp1 W1'^T -> df (Fourier reverse)
//   -> d_features = df W0^T,
// with every intermediate in tensor memory (A operands) and the forward's packed weights as MN-major B operands.
// All gradients are carried multiplied by a power of two S (|d_rgb|_max * S in [1, 2)) so that two fp16 terms keep
// ~22 significant bits; outputs are rescaled.  dp2 / dp1 / df / d(out) are also written as slab tiles for kernel 2.
// ===============================================================================================================
namespace fz {
constexpr uint32_t kBAccA = 0, kBAccB = 160, kBAop = 320;  // reverse: accumulators up to 160 columns wide
constexpr int kBwdMmaWarp = 16, kBwdLoadWarp = 17, kBwdThreads = 576;  // warps 0-15 workers (four column quarters)
}
__device__ __forceinline__ void grad_scale(const float* amax_slot, float& S, float& invS) {
  const uint32_t E = (__float_as_uint(*amax_slot) >> 23) & 0xFFu;
  // |d_rgb|_max * S in [2^7, 2^8): typical gradient magnitudes sit around 1, far above the fp16 subnormal floor of the
  // lo terms (6e-8), with 2^8 of headroom to the fp16 maximum for growth through the three layers
  const bool ok = E >= 8u && E <= 246u;
  S = ok ? __uint_as_float((261u - E) << 23) : 1.0f;
  invS = ok ? __uint_as_float((E - 7u) << 23) : 1.0f;
}
__global__ void __launch_bounds__(256) k_fused_amax(const float* x, int64_t n, float* slot) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(slot), __float_as_uint(m));
}

struct FusedBwdArgs {
  const unsigned char* wpack;
  const float* w3;
  const float *rgb, *d_rgb;
  const float* fs;
  const uint32_t *bits1, *bits2;
  const float* amax;
  int64_t M;
  int K0;
  unsigned char *dp2s, *dp1s, *dfs, *douts;
  float* d_feat;
  float* db3;
};

__device__ __forceinline__ void store_chunk_slab(unsigned char* slab_row, int c, const uint32_t* hi, const uint32_t* lo) {
  uint4* p = reinterpret_cast<uint4*>(slab_row + (size_t)(4 * c) * 2048);
  TF_SLAB_ST(&p[0], make_uint4(hi[0], hi[1], hi[2], hi[3]));
  TF_SLAB_ST(&p[128], make_uint4(lo[0], lo[1], lo[2], lo[3]));
  TF_SLAB_ST(&p[256], make_uint4(hi[4], hi[5], hi[6], hi[7]));
  TF_SLAB_ST(&p[384], make_uint4(lo[4], lo[5], lo[6], lo[7]));
}

__device__ __forceinline__ void fused_bwd_worker(const FusedBwdArgs& g, int Q, uint32_t tmem, int quad, int lane, uint64_t* kready, uint64_t* dfready,
                                                 uint64_t* accfull, uint64_t* s3done, const float* s_w3) {
  const int64_t ntiles = (g.M + 127) / 128;
  const int row = quad * 32 + lane;
  const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
  const uint32_t aop = lane_base + fz::kBAop;
  pdl_wait();  // d_rgb, max|d_rgb| and the zeroed db3 are the predecessor's (k_ray_bwd's) output
  float S, invS;
  grad_scale(g.amax, S, invS);
  const int nk0 = g.K0 / 16;
  // ---- output layer reverse (networks.py:114-120) and relu'(h2): this quarter's 32 columns of dp2 of tile t ----
  auto out_reverse = [&](int64_t t) {
    const int64_t m = t * 128 + row;
    float dout[3] = {0.f, 0.f, 0.f};
    if (m < g.M) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float y = g.rgb[3 * m + c];
        dout[c] = S * g.d_rgb[3 * m + c] * y * (1.0f - y);
      }
    }
    const uint32_t bw = g.bits2[m * 4 + Q];
    unsigned char* slab_row = g.dp2s + (size_t)t * (fz::kU * 512) + (size_t)row * 16;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int n0 = 32 * Q + 16 * c;
      float y[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float v = fmaf(dout[2], s_w3[3 * (n0 + q) + 2], fmaf(dout[1], s_w3[3 * (n0 + q) + 1], dout[0] * s_w3[3 * (n0 + q)]));
        y[q] = ((bw >> (16 * c + q)) & 1u) ? v : 0.f;
      }
      emit8<true>(y, aop, 4 * Q + 2 * c, slab_row);
      emit8<true>(y + 8, aop, 4 * Q + 2 * c + 1, slab_row);
      chunk_arrive(&kready[2 * Q + c], lane);
    }
    if (Q == 0) {
      float y[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) y[q] = q < 3 ? dout[q < 3 ? q : 0] : 0.f;
      uint32_t hi[8], lo[8];
      split16(y, hi, lo);
      store_chunk_slab(g.douts + (size_t)t * (16 * 512) + (size_t)row * 16, 0, hi, lo);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float sum = warp_sum(dout[c]);
        if (lane == 0) atomicAdd(g.db3 + c, sum * invS);
      }
    }
  };
  const int64_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  // Software pipeline over the CTA's tiles with every stage appearing once in the code: step k finishes dp1 and df of tile
  // k-1, writes tile k's dp2, then stores tile k-1's d_features.
  for (int64_t k = 0; k <= my_tiles; ++k) {
    const int64_t t = blockIdx.x + (k - 1) * (int64_t)gridDim.x;  // the tile being finished (k >= 1)
    const int64_t m = t * 128 + row;
    const uint32_t ph = (uint32_t)(k - 1) & 1u;
    // accumulators alternate with the tile parity: dh1 and d_features in P, dx' in Q_ (see the MMA warp)
    const uint32_t accP = lane_base + (ph ? fz::kBAccB : fz::kBAccA), accQ = lane_base + (ph ? fz::kBAccA : fz::kBAccB);
    if (k > 0) {
      // ---- dp1 = (dp2 W2^T) * relu'(h1) ----
      mbar_wait(&accfull[0], ph, 50);
      tc_fence_after();
      {
        const uint32_t bw = g.bits1[m * 4 + Q];
        unsigned char* slab_row = g.dp1s + (size_t)t * (fz::kU * 512) + (size_t)row * 16;
        float y[32];
        tmem_ld16(accP + 32 * Q, y);
        tmem_ld16(accP + 32 * Q + 16, y + 16);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int q = 0; q < 16; ++q) y[16 * c + q] = ((bw >> (16 * c + q)) & 1u) ? y[16 * c + q] : 0.f;
          emit8<true>(y + 16 * c, aop, 4 * Q + 2 * c, slab_row);
          emit8<true>(y + 16 * c + 8, aop, 4 * Q + 2 * c + 1, slab_row);
          chunk_arrive(&kready[2 * Q + c], lane);
        }
      }
      // ---- Fourier reverse (networks.py:13-35, :68-76): df_d = dx_d + cos(f) dx_sin1 + 2 cos(2f) dx_sin2 - sin(f) dx_cos1 - 2 sin(2f) dx_cos2 ----
      mbar_wait(&accfull[1], ph, 51);
      tc_fence_after();
      {
        float fv[8];
        const float4* fp = reinterpret_cast<const float4*>(g.fs + (size_t)t * 4096 + (size_t)(2 * Q) * 512) + row;
        const float4 f0 = fp[0], f1 = fp[128];
        fv[0] = f0.x; fv[1] = f0.y; fv[2] = f0.z; fv[3] = f0.w; fv[4] = f1.x; fv[5] = f1.y; fv[6] = f1.z; fv[7] = f1.w;
        float dx[40];
#pragma unroll
        for (int c = 0; c < 5; ++c) tmem_ld8(accQ + 40 * Q + 8 * c, dx + 8 * c);
        tmem_ld_wait();
        float y[8];
        const int ND = Q < 3 ? 8 : 3;  // feature dimensions of this quarter (the view direction needs no gradient)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < ND) {
            float sn, cs;
            fast_sincos(fv[i], sn, cs);
            const float sn2 = 2.0f * sn * cs, cs2 = fmaf(-2.0f * sn, sn, 1.0f);
            y[i] = dx[5 * i] + cs * dx[5 * i + 1] + 2.0f * cs2 * dx[5 * i + 2] - sn * dx[5 * i + 3] - 2.0f * sn2 * dx[5 * i + 4];
          } else {
            y[i] = 0.f;
          }
        }
        emit8<true>(y, aop, Q, g.dfs + (size_t)t * (32 * 512) + (size_t)row * 16);
        chunk_arrive(&dfready[Q >> 1], lane);
      }
      // ---- d_features = df W0^T complete: the operand region is free, so the NEXT tile's dp2 goes in first and its dh1 product
      // runs on the tensor pipe while this tile's d_features are written out ----
      mbar_wait(&accfull[2], ph, 52);
      tc_fence_after();
    }
    if (k < my_tiles) out_reverse(blockIdx.x + k * (int64_t)gridDim.x);
    if (k > 0) {
      {  // accumulator fragments (16 rows x 256 bit) -> whole 32-byte sectors
        const int lrow = lane >> 2, lc = (lane & 3) * 2;
        const int c_beg = (nk0 * Q) / 4, c_end = (nk0 * (Q + 1)) / 4;
        const uint32_t accP0 = tmem + (ph ? fz::kBAccB : fz::kBAccA);
        for (int c = c_beg; c < c_end; ++c) {
          float v[2][8];
          tmem_ld_16x256b_x2(accP0 + ((uint32_t)(quad * 32) << 16) + 16u * c, v[0]);
          tmem_ld_16x256b_x2(accP0 + ((uint32_t)(quad * 32 + 16) << 16) + 16u * c, v[1]);
          tmem_ld_wait();
#pragma unroll
          for (int blk = 0; blk < 2; ++blk)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int64_t mr = t * 128 + quad * 32 + blk * 16 + h * 8 + lrow;
              if (mr < g.M) {
                float* dst = g.d_feat + mr * g.K0 + 16 * c + lc;
                *reinterpret_cast<float2*>(dst) = make_float2(v[blk][h * 2] * invS, v[blk][h * 2 + 1] * invS);
                *reinterpret_cast<float2*>(dst + 8) = make_float2(v[blk][4 + h * 2] * invS, v[blk][4 + h * 2 + 1] * invS);
              }
            }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s3done);  // this warp has read accumulator P: the next tile's dx' product may overwrite it
    }
  }
}

__global__ void __maxnreg__(72) k_mlp_fused_bwd(FusedBwdArgs g) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* wfull = reinterpret_cast<uint64_t*>(smem);
  uint64_t* kready = wfull + 1;     // [8] dp2 / dp1 chunks (one quarter each)
  uint64_t* dfready = kready + 8;   // [2] df chunks (two quarters each)
  uint64_t* accfull = dfready + 2;  // [3]
  uint64_t* s3done = accfull + 3;   // all 16 worker warps have read the d_features accumulator of the tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s3done + 1);
  float* s_w3 = reinterpret_cast<float*>(smem + 1024);  // [384]
  const uint32_t wbytes = fz::b0_bytes(g.K0) + fz::kB1Bytes + fz::kB2Bytes;
  unsigned char* sW = smem + fz::kHeader;
  const int64_t ntiles = (g.M + 127) / 128;
  const int64_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  if (tid == 0) {
    mbar_init(wfull, 1);
    for (int c = 0; c < 8; ++c) mbar_init(&kready[c], 4);
    for (int c = 0; c < 2; ++c) mbar_init(&dfready[c], 8);
    for (int l = 0; l < 3; ++l) mbar_init(&accfull[l], 1);
    mbar_init(s3done, 16);
    fence_barrier_init();
  }
  for (int n = tid; n < 384; n += fz::kBwdThreads) s_w3[n] = g.w3[n];
  if (warp == fz::kBwdMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 16) {
    fused_bwd_worker(g, warp >> 2, tmem, warp & 3, lane, kready, dfready, accfull, s3done, s_w3);
  } else if (warp == fz::kBwdLoadWarp) {
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull, wbytes);
      for (uint32_t off = 0; off < wbytes; off += 16384u) bulk_copy_g2s(sW + off, g.wpack + off, min(16384u, wbytes - off), wfull);
    }
  } else {
    // ================= MMA issuer: the weight tiles of the forward, read as MN-major B (tile rows = contraction index) ======
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t sw = smem_u32(sW);
    const uint32_t b0_lo = desc_lo(sw, 128u), b1_lo = desc_lo(sw + fz::b0_bytes(g.K0), 128u),
                   b2_lo = desc_lo(sw + fz::b0_bytes(g.K0) + fz::kB1Bytes, 128u);
    const uint32_t bh32 = desc_hi(2u * slab_bytes(32)), bh128 = desc_hi(2u * slab_bytes(128));
    const uint32_t t32 = slab_bytes(32) >> 4, t128 = slab_bytes(128) >> 4;
    const uint32_t id2 = make_idesc(128, fz::kU, FMT_F16, FMT_F16, 0, 1), id1 = make_idesc(128, fz::kX, FMT_F16, FMT_F16, 0, 1),
                   id0 = make_idesc(128, g.K0, FMT_F16, FMT_F16, 0, 1);
    const uint32_t aop = tm + fz::kBAop;
    uint32_t kph = 0, dph = 0;
    mbar_wait(wfull, 0, 60);
    for (int64_t it = 0; it < my_tiles; ++it) {
      // accumulators alternate with the tile parity: the workers feed tile it+1's dp2 while they still read tile it's
      // d_features accumulator (P), so dh1 of the next tile goes to the other one
      const uint32_t accP = tm + ((it & 1) ? fz::kBAccB : fz::kBAccA), accQ = tm + ((it & 1) ? fz::kBAccA : fz::kBAccB);
      for (int i = 0; i < 8; ++i) {  // dh1 = dp2 W2^T (quarter q completes chunk 2q, then 2q + 1)
        const int c = (i & 3) * 2 + (i >> 2);
        mbar_wait(&kready[c], (kph >> c) & 1u, 61);
        kph ^= 1u << c;
        tc_fence_after();
        if (elect_one()) umma_ts_split2(accP, aop + 8u * c, fz::kAopLo, b2_lo + 16u * c, t128, bh128, id2, i == 0);
        __syncwarp();
      }
      if (elect_one()) umma_commit(&accfull[0]);
      __syncwarp();
      if (it > 0) {  // accQ is the previous tile's d_features accumulator: every worker warp must have read it
        mbar_wait(s3done, (uint32_t)(it - 1) & 1u, 64);
        tc_fence_after();
      }
      for (int i = 0; i < 8; ++i) {  // dx' = dp1 W1'^T
        const int c = (i & 3) * 2 + (i >> 2);
        mbar_wait(&kready[c], (kph >> c) & 1u, 62);
        kph ^= 1u << c;
        tc_fence_after();
        if (elect_one()) umma_ts_split2(accQ, aop + 8u * c, fz::kAopLo, b1_lo + 16u * c, t128, bh128, id1, i == 0);
        __syncwarp();
      }
      if (elect_one()) umma_commit(&accfull[1]);
      __syncwarp();
      for (int c = 0; c < 2; ++c) {  // d_features = df W0^T
        mbar_wait(&dfready[c], (dph >> c) & 1u, 63);
        dph ^= 1u << c;
        tc_fence_after();
        if (elect_one()) umma_ts_split2(accP, aop + 8u * c, fz::kAopLo, b0_lo + 16u * c, t32, bh32, id0, c == 0);
        __syncwarp();
      }
      if (elect_one()) umma_commit(&accfull[2]);
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == fz::kBwdMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ===============================================================================================================
// Reverse pass, kernel 2 of 2: every weight / bias gradient of the MLP as products over the rows, K = 128 rows per tile,
// both operands MN-major views of slab tiles streamed from global memory in 16 KB pieces (cp.async.bulk):
//   dW3^T[n][c]  += h2^T  d(out)          dW2^T[n][k] += dp2^T [h1 | 1]
//   dW1'^T[n][k'] += dp1^T x'  (x' has a ones column: db1)          dW0'^T[n'][k] += df^T features
// The four accumulators (16 + 144 + 160 + 3*ca columns of tensor memory) stay resident over all tiles of the CTA and
// are added to the gradient leaves once at the end.
// ===============================================================================================================
namespace fz {
constexpr int kWgThreads = 192;  // warps 0-3 epilogue, 4 MMA, 5 loader
constexpr int kWgSlots = 5;
constexpr uint32_t kPiece = 16384, kATile = 65536;
constexpr uint32_t kWgOnes = 1024, kWgA = kWgOnes + 8192, kWgB = kWgA + 2 * kATile, kWgSmem = kWgB + kWgSlots * kPiece;
constexpr uint32_t kD3 = 0, kD2 = 16, kD1 = 160, kD0 = 320;
}
struct FusedWgradArgs {
  const unsigned char *h2s, *douts, *dp2s, *h1s, *dp1s, *xs, *dfs, *feats;
  const float* amax;
  int64_t M;
  int K0;
  float *dw0, *dw1, *db1, *dw2, *db2, *dw3;
  float* partial;  // [gridDim.x][320 + K0][128]
  int group;       // B pieces per MMA: 2 (N = 64) when the kernel runs alone, 1 beside a scatter kernel
};

// Sum of the per-CTA partial accumulators of k_mlp_fused_wgrad -> the gradient leaves (overwritten), un-scaled, with the
// kernel column orders mapped back to the reference's (fz::perm, fz::sq_nat).  One thread per (accumulator column, row).
struct WgradReduceArgs {
  const float* partial;
  const float* amax;
  int ncta, K0;
  float *dw0, *dw1, *db1, *dw2, *db2, *dw3;
};
__global__ void __launch_bounds__(1024) k_wgrad_reduce(WgradReduceArgs a) {
  // block = (128 rows, 8 slices of the CTA range): 8x the loads in flight of a single pass over the partials (the sum is
  // latency-bound: ~148 strided 512-byte reads per column), combined through shared memory in a fixed order (deterministic)
  __shared__ float s_red[8][128];
  const int col = blockIdx.x, n = threadIdx.x, j = threadIdx.y;
  const int ncols = fz::kD0 + a.K0;
  const bool live = !(col >= (int)fz::kD0 && n >= 32);
  const float* p = a.partial + (size_t)col * 128 + n;
  const size_t stride = (size_t)ncols * 128;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (live) {
    int c = j;
    for (; c + 24 < a.ncta; c += 32) {
      s0 += __ldcg(p + (size_t)c * stride);
      s1 += __ldcg(p + (size_t)(c + 8) * stride);
      s2 += __ldcg(p + (size_t)(c + 16) * stride);
      s3 += __ldcg(p + (size_t)(c + 24) * stride);
    }
    for (; c < a.ncta; c += 8) s0 += __ldcg(p + (size_t)c * stride);
  }
  s_red[j][n] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (j != 0 || !live) return;
  float tot = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) tot += s_red[q][n];
  float S, invS;
  grad_scale(a.amax, S, invS);
  const float x = tot * invS;
  if (col < (int)fz::kD2) {
    if (col < 3) a.dw3[n * 3 + col] = x;
  } else if (col < (int)fz::kD1) {
    const int k = col - fz::kD2;
    if (k < fz::kU) a.dw2[k * fz::kU + n] = x;
    else if (k == fz::kU) a.db2[n] = x;
  } else if (col < (int)fz::kD0) {
    const int kp = col - fz::kD1, nat = fz::perm(kp);
    if (nat >= 0) a.dw1[nat * fz::kU + n] = x;
    else if (kp == fz::kOnesCol) a.db1[n] = x;
  } else {
    const int nat0 = fz::sq_nat(n);
    if (nat0 >= 0) a.dw0[(col - fz::kD0) * fz::kSq + nat0] = x;
  }
}

// 64-register cap (48 used, no spills; 98 without it): 192 x 48 registers leave room for two 96-register CTAs of the
// appearance scatter on the same SM (forked reverse pass).
__global__ void __maxnreg__(64) k_mlp_fused_wgrad(FusedWgradArgs g) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem);  // [2]
  uint64_t* aempty = afull + 2;                          // [2]
  uint64_t* bfull = aempty + 2;                          // [kWgSlots]
  uint64_t* bempty = bfull + fz::kWgSlots;               // [kWgSlots]
  uint64_t* done = bempty + fz::kWgSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  unsigned char* sOnes = smem + fz::kWgOnes;
  unsigned char* sA = smem + fz::kWgA;
  unsigned char* sB = smem + fz::kWgB;
  const int64_t ntiles = (g.M + 127) / 128;
  const int64_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int K0 = g.K0;
  // pieces of the B operand per product: (bytes, columns)
  const int npf = (K0 + 31) / 32;  // feature pieces; the last one may be half (16 columns)

  if (tid == 0) {
    for (int a = 0; a < 2; ++a) {
      mbar_init(&afull[a], 1);
      mbar_init(&aempty[a], 1);
    }
    for (int b = 0; b < fz::kWgSlots; ++b) {
      mbar_init(&bfull[b], 1);
      mbar_init(&bempty[b], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  // constant B piece [128 rows x 16 columns]: column 0 = 1 (hi term), everything else 0
  for (int i = tid; i < 8192 / 16; i += fz::kWgThreads)
    reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(i < 128 ? 0x00003C00u : 0u, 0u, 0u, 0u);
  fence_proxy_async();
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 5) {
    // ================= loader =================
    if (lane == 0) {
      uint32_t ab = 0, aph = 0, sl = 0, bph = 0;
      // read-once streams: evict-first, so that they do not push the gradient planes of the scatter kernel running beside
      // this one out of the L2 (-12 us per step)
      const uint64_t pol = l2_policy_evict_first();
      auto copy = [&](void* dst, const void* src, uint32_t bytes, uint64_t* bar) { bulk_copy_g2s_hint(dst, src, bytes, bar, pol); };
      auto load_a = [&](const unsigned char* src, uint32_t bytes) {
        mbar_wait(&aempty[ab], aph ^ 1, 70);
        mbar_arrive_expect_tx(&afull[ab], bytes);
        for (uint32_t off = 0; off < bytes; off += fz::kPiece) copy(sA + ab * fz::kATile + off, src + off, fz::kPiece, &afull[ab]);
        if (++ab == 2) { ab = 0; aph ^= 1; }
      };
      auto load_b = [&](const unsigned char* src, uint32_t bytes) {
        mbar_wait(&bempty[sl], bph ^ 1, 71);
        mbar_arrive_expect_tx(&bfull[sl], bytes);
        copy(sB + sl * fz::kPiece, src, bytes, &bfull[sl]);
        if (++sl == fz::kWgSlots) { sl = 0; bph ^= 1; }
      };
      for (int64_t it = 0; it < my_tiles; ++it) {
        const size_t t = (size_t)(blockIdx.x + it * gridDim.x);
        load_a(g.h2s + t * fz::kATile, fz::kATile);
        load_b(g.douts + t * 8192, 8192);
        load_a(g.dp2s + t * fz::kATile, fz::kATile);
        for (int j = 0; j < 4; ++j) load_b(g.h1s + t * fz::kATile + j * fz::kPiece, fz::kPiece);
        load_a(g.dp1s + t * fz::kATile, fz::kATile);
        for (int j = 0; j < 5; ++j) load_b(g.xs + t * (fz::kX * 512) + j * fz::kPiece, fz::kPiece);
        load_a(g.dfs + t * fz::kPiece, fz::kPiece);
        for (int j = 0; j < npf; ++j) load_b(g.feats + t * ((size_t)K0 * 512) + j * fz::kPiece, (uint32_t)min(32, K0 - 32 * j) * 512u);
      }
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t a_hi = desc_hi(2u * slab_bytes(128)), term = slab_bytes(128) >> 4;
    const uint32_t id32 = make_idesc(128, 32, FMT_F16, FMT_F16, 1, 1), id16 = make_idesc(128, 16, FMT_F16, FMT_F16, 1, 1);
    const uint32_t sa = smem_u32(sA), sb = smem_u32(sB), so = smem_u32(sOnes);
    uint32_t ab = 0, aph = 0, sl = 0, bph = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const bool first = it == 0;
      auto wait_a = [&]() {
        mbar_wait(&afull[ab], aph, 80);
        tc_fence_after();
      };
      auto piece = [&](uint32_t b_addr, uint32_t dcol, uint32_t idesc, uint64_t* release) {
        const uint32_t a_lo = desc_lo(sa + ab * fz::kATile, 128u), b_lo = desc_lo(b_addr, 128u);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) umma_ss_split2(tm + dcol, a_lo + 16u * ks, term, a_hi, b_lo + 16u * ks, term, a_hi, idesc, first && ks == 0);
          if (release) umma_commit(release);
        }
        __syncwarp();
      };
      auto ring_piece = [&](uint32_t dcol, uint32_t idesc) {
        mbar_wait(&bfull[sl], bph, 81);
        tc_fence_after();
        piece(sb + sl * fz::kPiece, dcol, idesc, &bempty[sl]);
        if (++sl == fz::kWgSlots) { sl = 0; bph ^= 1; }
      };
      // `cols` columns of B in 32-column pieces (the last one may be shorter).  Two pieces in adjacent slots go into ONE
      // instruction stream of N = 64: every MMA re-reads its 128 x 16 A operand from shared memory whatever N is, and at
      // N = 32 the kernel was bound by exactly those reads (ncu: tensor-core shared-memory wavefronts at 59 % of peak
      // with the HMMA pipe never idle).
      auto ring_cols = [&](uint32_t dcol, int cols) {
        for (int c = 0; c < cols;) {
          const int left = cols - c;
          if (g.group > 1 && left > 32 && sl + 1 < (uint32_t)fz::kWgSlots) {
            const int n = min(64, left);
            mbar_wait(&bfull[sl], bph, 81);
            mbar_wait(&bfull[sl + 1], bph, 82);
            tc_fence_after();
            const uint32_t a_lo = desc_lo(sa + ab * fz::kATile, 128u), b_lo = desc_lo(sb + sl * fz::kPiece, 128u);
            const uint32_t idesc = make_idesc(128, n, FMT_F16, FMT_F16, 1, 1);
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) umma_ss_split2(tm + dcol + c, a_lo + 16u * ks, term, a_hi, b_lo + 16u * ks, term, a_hi, idesc, first && ks == 0);
              umma_commit(&bempty[sl]);
              umma_commit(&bempty[sl + 1]);
            }
            __syncwarp();
            sl += 2;
            if (sl == (uint32_t)fz::kWgSlots) { sl = 0; bph ^= 1; }
            c += n;
          } else {
            ring_piece(dcol + c, left >= 32 ? id32 : make_idesc(128, left, FMT_F16, FMT_F16, 1, 1));
            c += min(32, left);
          }
        }
      };
      auto release_a = [&]() {
        if (elect_one()) umma_commit(&aempty[ab]);
        __syncwarp();
        if (++ab == 2) { ab = 0; aph ^= 1; }
      };
      wait_a();  // h2^T d(out)
      ring_piece(fz::kD3, id16);
      release_a();
      wait_a();  // dp2^T [h1 | 1]
      ring_cols(fz::kD2, 128);
      piece(so, fz::kD2 + 128u, id16, nullptr);
      release_a();
      wait_a();  // dp1^T x'
      ring_cols(fz::kD1, fz::kX);
      release_a();
      wait_a();  // df^T features (rows >= 32 of this accumulator are meaningless and never read)
      ring_cols(fz::kD0, K0);
      release_a();
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
  } else {
    // ================= epilogue: this CTA's partial sums -> scratch [cta][column][128 rows], plain coalesced stores ==========
    // (148 CTAs adding 59 k values each straight into the leaves with REDs took 40 us: the L2 serialises them per address;
    // k_wgrad_reduce sums the partials instead - ~35 MB that are still in L2 - and is deterministic)
    mbar_wait(done, 0, 90);
    tc_fence_after();
    const int n = warp * 32 + lane;  // accumulator row = output unit of the layer
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    const int ncols = fz::kD0 + (warp == 0 ? K0 : 0);  // rows >= 32 of the Dense_0 accumulator are meaningless
    float* part = g.partial + (size_t)blockIdx.x * (size_t)(fz::kD0 + K0) * 128 + n;
    for (int c0 = 0; c0 < ncols; c0 += 16) {
      float v[16];
      tmem_ld16(tl + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 16; ++q) part[(size_t)(c0 + q) * 128] = v[q];
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

bool mlp_fused_bwd_ready() { return true; }

// The reverse pass in its two halves, so that a caller can run the (HBM-bound, 6-warp) weight-gradient kernel on one stream
// while the (issue-bound) scatter kernels that only need d_features run on another (render_rgb_bwd_impl):
//   mlp_fused_bwd_chain: max|d_rgb| (unless the caller has it) + k_mlp_fused_bwd  -> d_features, slab tiles of dp2 / dp1 / df / d(out), db3
//   mlp_fused_bwd_wgrad: k_mlp_fused_wgrad + k_wgrad_reduce                       -> dW0, dW1, db1, dW2, db2, dW3
int mlp_fused_bwd_chain(cudaStream_t st, const MlpShape& s, const MlpParams& p, int64_t M, const MlpWs& ws, const float* rgb,
                        const float* d_rgb, float* d_feat, const MlpGrads& gr) {
  TF_CHECK_ARG(!s.inference, "mlp reverse pass after a forward with TENSORF_FLAG_INFERENCE (residuals were not kept)");
  TF_CHECK_ARG(mlp_fused_supported(s), "fused MLP: unsupported network shape");
  if (!gr.prezeroed) TF_RETURN_IF_ERROR(mlp_zero_grads(st, s, gr));
  if (M == 0) return 0;
  const int64_t Mp = round_up64(M, 128);
  float* amax = ws.aux;
  if (!gr.amax_ready) {
    TF_CHECK_CUDA(cudaMemsetAsync(amax, 0, sizeof(float), st));
    k_fused_amax<<<kSMs, 256, 0, st>>>(d_rgb, 3 * M, amax);
    TF_CHECK_LAUNCH();
  }
  FusedBwdArgs b{};
  b.wpack = ws.wpack; b.w3 = p.w3; b.rgb = rgb; b.d_rgb = d_rgb; b.fs = ws.f; b.bits1 = ws.bits1; b.bits2 = ws.bits1 + Mp * 4;
  b.amax = amax; b.M = M; b.K0 = s.Ca;
  b.dp2s = reinterpret_cast<unsigned char*>(ws.dp2); b.dp1s = reinterpret_cast<unsigned char*>(ws.dp1);
  b.dfs = reinterpret_cast<unsigned char*>(ws.df); b.douts = reinterpret_cast<unsigned char*>(ws.aux + 64);
  b.d_feat = d_feat; b.db3 = gr.b3;
  const size_t smem_b = fz::kHeader + mlp_fused_wpack_bytes(s);
  const unsigned grid = (unsigned)std::min<int64_t>((M + 127) / 128, kSMs);
  TF_CHECK_CUDA(cudaFuncSetAttribute(k_mlp_fused_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
  TF_CHECK_CUDA(launch_pdl(k_mlp_fused_bwd, dim3(grid), dim3(fz::kBwdThreads), smem_b, st, b, ws.pdl));
  TF_CHECK_LAUNCH();
  return 0;
}

int mlp_fused_bwd_wgrad(cudaStream_t st, const MlpShape& s, int64_t M, const MlpWs& ws, const MlpGrads& gr) {
  if (M == 0) return 0;
  const unsigned grid = (unsigned)std::min<int64_t>((M + 127) / 128, kSMs);
  float* amax = ws.aux;
  FusedWgradArgs w{};
  w.h2s = reinterpret_cast<unsigned char*>(ws.h2); w.douts = reinterpret_cast<unsigned char*>(ws.aux + 64);
  w.dp2s = reinterpret_cast<unsigned char*>(ws.dp2); w.h1s = reinterpret_cast<unsigned char*>(ws.h1);
  w.dp1s = reinterpret_cast<unsigned char*>(ws.dp1); w.xs = reinterpret_cast<unsigned char*>(ws.x);
  w.dfs = reinterpret_cast<unsigned char*>(ws.df);
  w.feats = ws.feat_slabs ? ws.feat_slabs : reinterpret_cast<const unsigned char*>(ws.dx);
  w.amax = amax; w.M = M; w.K0 = s.Ca;
  w.dw0 = gr.w0; w.dw1 = gr.w1; w.db1 = gr.b1; w.dw2 = gr.w2; w.db2 = gr.b2; w.dw3 = gr.w3;
  w.partial = ws.wg_partial;
  // Alone, the kernel is 20 us faster with two B pieces per MMA (N = 64).  Beside the appearance scatter of the forked
  // reverse pass the whole step is 19 us SLOWER with it (measured in both orders): the scatter is the critical branch
  // and latency-bound on its DRAM reads, and a weight-gradient kernel that pulls its 538 MB faster starves it.
  static const int grp_env = getenv("TENSORF_WG_GROUP") ? atoi(getenv("TENSORF_WG_GROUP")) : 0;
  w.group = grp_env ? grp_env : (ws.beside_scatter ? 1 : 2);
  TF_CHECK_CUDA(cudaFuncSetAttribute(k_mlp_fused_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fz::kWgSmem));
  k_mlp_fused_wgrad<<<grid, fz::kWgThreads, fz::kWgSmem, st>>>(w);
  TF_CHECK_LAUNCH();
  WgradReduceArgs ra{ws.wg_partial, amax, (int)grid, s.Ca, gr.w0, gr.w1, gr.b1, gr.w2, gr.b2, gr.w3};
  k_wgrad_reduce<<<fz::kD0 + s.Ca, dim3(128, 8), 0, st>>>(ra);
  TF_CHECK_LAUNCH();
  return 0;
}

int mlp_fused_bwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const float* feat, const float* viewdirs, const uint32_t* cams,
                  int64_t M, int rows_per_ray, const MlpWs& ws, const float* rgb, const float* d_rgb, float* d_feat, const MlpGrads& gr) {
  (void)feat; (void)viewdirs; (void)cams; (void)rows_per_ray;
  TF_RETURN_IF_ERROR(mlp_fused_bwd_chain(st, s, p, M, ws, rgb, d_rgb, d_feat, gr));
  return mlp_fused_bwd_wgrad(st, s, M, ws, gr);
}

}  // namespace tf
