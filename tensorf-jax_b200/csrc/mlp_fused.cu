// FeatureMlp (networks.py:38-121) as per-row-tile fused tcgen05 kernels: activations never leave the SM between the
// layers of a pass.  See umma_tiles.cuh for the operand format (two-term 16-bit splits in "slab" tiles).
//
//   k_umma_probe : one 128 x N x K product with every operand source / orientation / format the fused kernels use
//                  (tests/test_gpu_fused.py pins the descriptor conventions on the hardware).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

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

}  // namespace tf
