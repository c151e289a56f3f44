// Split 16-bit operand tiles for tcgen05.mma and the instruction / descriptor variants the fused MLP kernels
// need on top of tc_prims.cuh: mixed fp16 / bf16 operands, MN-major shared-memory operands, A from tensor memory.
//
// "Slab layout" of a [ROWS x C] matrix whose fp32 values are split into two 16-bit terms (x ~ hi + lo):
//     byte offset(term, r, c) = ((c >> 3) * 2 + term) * ROWS * 16  +  r * 16  +  (c & 7) * 2
// i.e. per group of 8 columns one hi slab and one lo slab of ROWS x 16 bytes.  The same bytes are a canonical
// no-swizzle UMMA operand in BOTH orientations (cute mma_traits_sm100.hpp, make_umma_desc):
//   K-major  (rows = M or N index, columns = K): 8-row group stride SBO = 128 B, 8-column group stride LBO = 2*ROWS*16 B
//   MN-major (columns = M or N index, rows = K): 8-row (K) group stride LBO = 128 B, 8-column (MN) group stride SBO = 2*ROWS*16 B
// so a row tile of activations serves as the A operand of the forward / dX GEMMs (K = features) and as an operand
// of the weight-gradient GEMMs (K = rows) without a transpose, and one packed weight matrix serves the forward
// (K-major B) and the reverse (MN-major B) products.  A K-chunk of 16 columns is 4 contiguous slabs.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "tc_prims.cuh"

namespace tf {
namespace tc {

enum : int { FMT_F16 = 0, FMT_BF16 = 1 };

__host__ __device__ inline uint32_t slab_bytes(int rows) { return (uint32_t)rows * 16u; }
__host__ __device__ inline uint32_t slab_tile_bytes(int rows, int cols) { return (uint32_t)(cols >> 3) * 2u * slab_bytes(rows); }
__host__ __device__ inline uint32_t slab_offset(int rows, int term, int r, int c) {
  return (uint32_t)(((c >> 3) * 2 + term) * rows + r) * 16u + (uint32_t)(c & 7) * 2u;
}

// InstrDescriptor (cute/arch/mma_sm100_desc.hpp): c_format F32 [4,6); a_format [7,10), b_format [10,13) (0 = F16, 1 = BF16);
// a_major bit 15, b_major bit 16 (0 = K, 1 = MN); n_dim = N >> 3 [17,23); m_dim = M >> 4 [24,29).
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_fmt, int b_fmt, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Low / high words of a no-swizzle shared-memory descriptor (start >> 4 | LBO >> 4 << 16 ; SBO >> 4 | version 1 << 14).
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }

// D[tmem] (+)= A[smem] * B[smem]; descriptors as 32-bit words (uniform datapath, see tc_prims.cuh umma_bf16_lo).
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      :
      : "r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: A is [128 lanes x 16 k] of 16-bit values packed two per 32-bit column
// (lane = row, column = k / 2, low half = even k), K-major only.
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// registers -> tensor memory: lane i of the warp writes 8 / 16 consecutive 32-bit columns of TMEM lane (taddr.lane + i)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
          taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- fp32 pair -> packed 16-bit (hi, lo) terms ------------------------------------------------------------------
// hi = round(x), lo = round(x - hi): fp16 keeps 22 significant bits (absolute floor 2^-25), bf16 16 bits.
template <int FMT>
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  if (FMT == FMT_F16) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - __low2float(h), x1 - __high2float(h));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  }
}
// 8 consecutive columns of one row -> one 16-byte hi chunk and one 16-byte lo chunk
template <int FMT>
__device__ __forceinline__ void split8_fmt(const float* x, uint4& hi, uint4& lo) {
  split_pair<FMT>(x[0], x[1], hi.x, lo.x);
  split_pair<FMT>(x[2], x[3], hi.y, lo.y);
  split_pair<FMT>(x[4], x[5], hi.z, lo.z);
  split_pair<FMT>(x[6], x[7], hi.w, lo.w);
}

// The three products of a two-term split, smallest first: a_lo*b_hi, a_hi*b_lo, a_hi*b_hi.
// a_term / b_term: distance between the hi and lo operand in descriptor units (16 B) for shared-memory operands,
// in columns for a tensor-memory A.
__device__ __forceinline__ void umma_ss_split2(uint32_t d, uint32_t a_lo, uint32_t a_term, uint32_t a_hi, uint32_t b_lo, uint32_t b_term,
                                               uint32_t b_hi, uint32_t idesc, bool first) {
  umma_ss(d, a_lo + a_term, a_hi, b_lo, b_hi, idesc, !first);
  umma_ss(d, a_lo, a_hi, b_lo + b_term, b_hi, idesc, 1);
  umma_ss(d, a_lo, a_hi, b_lo, b_hi, idesc, 1);
}
__device__ __forceinline__ void umma_ts_split2(uint32_t d, uint32_t a_tmem, uint32_t a_term_cols, uint32_t b_lo, uint32_t b_term, uint32_t b_hi,
                                               uint32_t idesc, bool first) {
  umma_ts(d, a_tmem + a_term_cols, b_lo, b_hi, idesc, !first);
  umma_ts(d, a_tmem, b_lo + b_term, b_hi, idesc, 1);
  umma_ts(d, a_tmem, b_lo, b_hi, idesc, 1);
}

}  // namespace tc
}  // namespace tf
