// FeatureMlp (networks.py:38-121) on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// fp32 parity (1e-4) on bf16 tensor cores: every operand x is split into bf16 terms hi + mid (+ lo) and a
// product is accumulated term by term in fp32 (TMEM): 3 tcgen05.mma per k-step for the two-term split of
// the reverse GEMMs (~2^-16 per product), 6 for the three-term split of the forward GEMMs (fp32-exact
// operands; the composite reverse amplifies forward error).
//
//   k_tc_rowgemm  : C[M x N] = epilogue(A[M x K] * W[K x N])     forward layers, dX chain
//       persistent (1 CTA/SM), warp-specialised: one thread issues 2-D TMA loads (tensor map, 128-byte
//       swizzle) of fp32 A chunks into a raw ring; 3 producer groups x 4 warps convert shared -> shared
//       into the canonical no-swizzle K-major UMMA layout; weights pre-split once per call and resident in
//       shared memory (or streamed by cp.async.bulk); one MMA warp (elect.sync leader, uniform-register
//       descriptors); 8 epilogue warps in two sets, one TMEM accumulator each (tcgen05.ld.16x256b ->
//       bias / ReLU / 1-bit masks / fused output layer / fused Fourier encode -> 32-byte-sector global stores).
//   k_tc_redgemm2 : dW^T[128 x N] += G^T[128 x rows] * X[rows x N]   weight (and bias) gradients: rows split
//       across CTAs, raw fp32 row chunks by cp.async.bulk, 2 converter groups transpose + split, RED-add
//       epilogue.  k_tc_redgemm is the register-staged fallback for unaligned operands.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <cuda.h>

#include "mlp.cuh"
#include "tc_prims.cuh"

namespace tf {

using namespace tc;

// ---------------------------------------------------------------------------------------------
// Operand splitting. NSPLIT=2: x ~ hi+lo (16 mantissa bits), products hi*hi + hi*lo + lo*hi.
// NSPLIT=3: x = hi+mid+lo (24 bits, fp32-exact operands), six products down to 2^-24.
// ---------------------------------------------------------------------------------------------
template <int NSPLIT>
__device__ __forceinline__ void split_store(const float* x, unsigned char* tile0, uint32_t tile_stride, uint32_t off) {
  uint32_t h[4], m[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    float r0 = x[2 * i] - __low2float(hh), r1 = x[2 * i + 1] - __high2float(hh);
    __nv_bfloat162 mm = __floats2bfloat162_rn(r0, r1);
    h[i] = *reinterpret_cast<uint32_t*>(&hh);
    m[i] = *reinterpret_cast<uint32_t*>(&mm);
    if (NSPLIT == 3) {
      r0 -= __low2float(mm);
      r1 -= __high2float(mm);
      __nv_bfloat162 ll = __floats2bfloat162_rn(r0, r1);
      l[i] = *reinterpret_cast<uint32_t*>(&ll);
    }
  }
  *reinterpret_cast<uint4*>(tile0 + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(tile0 + tile_stride + off) = make_uint4(m[0], m[1], m[2], m[3]);
  if (NSPLIT == 3) *reinterpret_cast<uint4*>(tile0 + 2 * tile_stride + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// Issue the split products of one k-step, smallest terms first. a[i], b[i]: descriptors of split i.
template <int NSPLIT>
__device__ __forceinline__ void umma_split(uint32_t d, const uint64_t* a, const uint64_t* b, uint32_t idesc, bool first) {
  if (NSPLIT == 2) {
    umma_bf16(d, a[1], b[0], idesc, !first);
    umma_bf16(d, a[0], b[1], idesc, 1);
    umma_bf16(d, a[0], b[0], idesc, 1);
  } else {
    umma_bf16(d, a[2], b[0], idesc, !first);
    umma_bf16(d, a[0], b[2], idesc, 1);
    umma_bf16(d, a[1], b[1], idesc, 1);
    umma_bf16(d, a[1], b[0], idesc, 1);
    umma_bf16(d, a[0], b[1], idesc, 1);
    umma_bf16(d, a[0], b[0], idesc, 1);
  }
}

template <int NSPLIT>
__device__ __forceinline__ void umma_split_lo(uint32_t d, uint32_t a0, uint32_t a_split, uint32_t b0, uint32_t b_split,
                                              uint32_t desc_hi, uint32_t idesc, bool first) {
  if (NSPLIT == 2) {
    umma_bf16_lo(d, a0 + a_split, b0, desc_hi, idesc, !first);
    umma_bf16_lo(d, a0, b0 + b_split, desc_hi, idesc, 1);
    umma_bf16_lo(d, a0, b0, desc_hi, idesc, 1);
  } else {
    umma_bf16_lo(d, a0 + 2 * a_split, b0, desc_hi, idesc, !first);
    umma_bf16_lo(d, a0, b0 + 2 * b_split, desc_hi, idesc, 1);
    umma_bf16_lo(d, a0 + a_split, b0 + b_split, desc_hi, idesc, 1);
    umma_bf16_lo(d, a0 + a_split, b0, desc_hi, idesc, 1);
    umma_bf16_lo(d, a0, b0 + b_split, desc_hi, idesc, 1);
    umma_bf16_lo(d, a0, b0, desc_hi, idesc, 1);
  }
}

// ---------------------------------------------------------------------------------------------
// weight packing: W (fp32, any strides) -> per KC-wide K chunk NSPLIT tiles [N_pad x KC]
// ---------------------------------------------------------------------------------------------
constexpr int kRC = 32;  // row chunk of the reduction GEMM
template <int NSPLIT>
struct RowCfg {
  static constexpr int KC = 32;  // K chunk per pipeline stage
};

struct PackArgs {
  const float* W;
  int64_t sk, sn;  // W(k, n) = W[k*sk + n*sn]
  int K_valid, N_valid, K_pad, N_pad;
  unsigned char* out;
};

template <int NSPLIT>
__global__ void __launch_bounds__(256) k_pack_weights(PackArgs a) {
  constexpr int KC = RowCfg<NSPLIT>::KC;
  int item = blockIdx.x * blockDim.x + threadIdx.x;  // (n, k8)
  const int nchunks = (a.K_pad + KC - 1) / KC;
  if (item >= a.N_pad * nchunks * (KC / 8)) return;
  const int n = item % a.N_pad;
  const int k8g = item / a.N_pad;  // k8 index over padded chunks
  const int c = k8g / (KC / 8), j = k8g % (KC / 8);
  float x[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    int k = c * KC + j * 8 + q;
    x[q] = (k < a.K_valid && n < a.N_valid) ? a.W[k * a.sk + n * a.sn] : 0.f;
  }
  const uint32_t tb = tile_bytes(a.N_pad, KC);
  split_store<NSPLIT>(x, a.out + (size_t)c * NSPLIT * tb, tb, tile_offset(n, j * 8, KC));
}

template <int NSPLIT>
static size_t packed_weight_bytes(int K_pad, int N_pad) {
  constexpr int KC = RowCfg<NSPLIT>::KC;
  return (size_t)((K_pad + KC - 1) / KC) * NSPLIT * tile_bytes(N_pad, KC);
}

// Debug timeline (TENSORF_TC_TRACE=1): CTA 0 records clock64() per role and chunk.
__device__ long long g_trace[8192];
#define TF_TRACE(slot, idx) do { if (g.trace && blockIdx.x == 0 && (idx) < 1024) g_trace[(slot) * 1024 + (idx)] = clock64(); } while (0)

// ---------------------------------------------------------------------------------------------
// k_tc_rowgemm
// ---------------------------------------------------------------------------------------------
struct RowGemmArgs {
  const float* A;
  int64_t lda, M;
  int K_valid, K_pad;         // K_pad multiple of 16
  const unsigned char* Bp;    // packed weights
  int N_pad;                  // multiple of 16, <= 256
  float* C;
  int64_t ldc;
  int N_store;                // columns written (<= N_pad)
  const float* bias;          // [N_store] or null
  const uint32_t* bits_in;    // keep C(m,n) only where bit n of row m is set (ceil(N_pad/32) words per row), or null
  uint32_t* bits_out;         // receives bit(m,n) = C(m,n) > 0 (the ReLU mask for the reverse pass), or null
  int relu;
  int stages;
  int trace;                  // debug: record the CTA-0 timeline into g_trace
  int resident;               // 1: all weight chunks live in smem for the whole kernel (loaded once)
  int header_bytes;           // kRowFixed
  // fused output layer (networks.py:103-120): FiLM + Dense(3) + sigmoid on the epilogue's rows (N_pad == 128)
  const float* w3;            // (128,3) or null
  const float* b3;            // (3)
  const float* embed;         // (ncam,128) or null
  const uint32_t* cams;       // camera index per ray
  int rows_per_ray;
  float* rgb_out;             // (M,3)
  float* enc_x;               // EPI_ENC (Dense_0): the Fourier-encoded MLP input x (M, enc_ldx) is written by the epilogue
  const float* enc_vd;        //   viewdirs (M / rows_per_ray, 3)
  int enc_ldx, enc_squash, enc_Ff, enc_Fv;
  int groups;                 // active producer groups; stages % groups == 0 (see launch_rowgemm)
  int raw_slots;              // > 0: A chunks arrive by TMA (tensor map) into a raw fp32 ring of this many slots
  uint32_t tmem_cols;         // power of two >= 2*N_pad
};

// warps 0-7 epilogue (two per TMEM lane quadrant, each takes half of the columns), 8 MMA issuer,
// 9 weight loader, 10 TMA loader of the raw A ring, then kGroups producer groups of 4 warps.
// A operand path: ONE thread issues a 2-D TMA load (tensor map over the fp32 activation matrix, box =
// 128 rows x 32 columns, SWIZZLE_128B so the column-wise reads below are bank-conflict free) per K-chunk
// into a ring of raw fp32 slots; the producer groups only convert shared -> shared (split into bf16
// core matrices) and never wait on global memory.  Fallback (unaligned A): the groups load through
// registers; fence.proxy.async then drains the issuing thread's outstanding loads, so overlap only
// comes from every group owning whole chunks (chunk seq -> group seq % groups).
constexpr int kGroups = 3;
constexpr int kGroupThreads = 128;
constexpr int kEpiWarps = 8;
constexpr int kMmaWarp = kEpiWarps, kLoadWarp = kEpiWarps + 1, kRawWarp = kEpiWarps + 2, kProdWarp0 = kEpiWarps + 3;
constexpr int kRowThreads = 32 * kProdWarp0 + kGroups * kGroupThreads;
constexpr int kRawSlotsMax = 8;
constexpr int kRowFixed = 4096;  // barriers [0,1K), bias [1K,2K), W3/b3 [2K,4K)
enum : int { EPI_BITS_IN = 1, EPI_BITS_OUT = 2, EPI_OUT3 = 4, EPI_ENC = 8, EPI_FILM = 16 };  // FILM: OUT3 with camera embeddings

template <int NSPLIT, int EPI>
__global__ void __launch_bounds__(kRowThreads, 1) k_tc_rowgemm(RowGemmArgs g, const __grid_constant__ CUtensorMap tmapA) {
  constexpr int KC = RowCfg<NSPLIT>::KC;
  constexpr int ITEMS = 128 * (KC / 8) / kGroupThreads;  // 16-byte k-chunks per producer thread and stage
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = g.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [S]
  uint64_t* empty = full + S;                           // [S]
  uint64_t* tfull = empty + S;                          // [2]
  uint64_t* tempty = tfull + 2;                         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_bias = reinterpret_cast<float*>(smem + 1024);   // [256] bias staged once per CTA
  float* s_w3 = reinterpret_cast<float*>(smem + 2048);     // [128*3 + 3] output layer weights + bias
  const uint32_t a_tile = tile_bytes(128, KC), b_tile = tile_bytes(g.N_pad, KC);
  const int nchunks_w = (g.K_pad + KC - 1) / KC;
  // resident weights: [header][all weight chunks][A stages]; streamed: [header][stages of A+B]
  const uint32_t stage_bytes = g.resident ? NSPLIT * a_tile : NSPLIT * (a_tile + b_tile);
  unsigned char* wres = smem + g.header_bytes;
  unsigned char* stage0 = wres + (g.resident ? (size_t)nchunks_w * NSPLIT * b_tile : 0);
  uint64_t* wfull = tempty + 2 + 1;  // after tmem_slot (8-byte aligned slot)
  uint64_t* rfull = wfull + 1;             // [raw_slots] raw fp32 chunk landed (TMA tx bytes)
  uint64_t* rempty = rfull + kRawSlotsMax;  // [raw_slots] raw chunk read by its producer group (4 warps)
  constexpr uint32_t kRawBytes = 128 * KC * 4;
  unsigned char* raw0 = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(stage0 + (size_t)S * stage_bytes) + 1023) & ~(uintptr_t)1023);  // swizzle atom alignment
  const int T = g.raw_slots;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], kGroupThreads + (g.resident ? 0 : 1));  // producer group (+ the weight loader's expect_tx arrive)
      mbar_init(&empty[s], 1);                                    // tcgen05.commit
    }
    mbar_init(wfull, 1);
    for (int r = 0; r < T; ++r) {
      mbar_init(&rfull[r], 1);
      mbar_init(&rempty[r], kGroupThreads / 32);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);                // tcgen05.commit
      mbar_init(&tempty[a], 32 * kEpiWarps / 2);  // the 4 epilogue warps that own this accumulator
    }
    fence_barrier_init();
  }
  for (int n = tid; n < 256; n += kRowThreads) s_bias[n] = (g.bias && n < g.N_store) ? g.bias[n] : 0.f;
  if (EPI & EPI_OUT3)
    for (int n = tid; n < 387; n += kRowThreads) s_w3[n] = n < 384 ? g.w3[n] : g.b3[n - 384];
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, g.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t ntiles = (g.M + 127) / 128;
  const int nchunks = (g.K_pad + KC - 1) / KC;
  const int64_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int64_t total = my_tiles * nchunks;  // chunks this CTA processes, in (tile, k-chunk) order

  if (warp >= kProdWarp0) {
    // ================= A producers: fp32 rows -> split bf16 core matrices =================
    // 8 consecutive lanes fill one 128-byte core matrix (8 rows x 16 B).
    const int grp = (warp - kProdWarp0) >> 2, pw = (warp - kProdWarp0) & 3;
    const bool vec_ok = (g.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
    // chunk sequence number seq -> (tile, k-chunk), operand stage and raw slot are tracked incrementally: 64-bit
    // divisions by run-time values cost ~25 instructions each and there were six of them per chunk.
    // S and T are multiples of the group count, so stage / slot indices advance by `groups` and wrap exactly.
    const int ngrp = g.groups, ntot = (int)total;
    int c_idx = grp % nchunks;                                   // k-chunk of the tile
    int64_t t = blockIdx.x + (int64_t)(grp / nchunks) * gridDim.x;  // tile
    uint32_t st = (uint32_t)grp, ph = 0, sl = (uint32_t)grp, rph = 0;
    const uint32_t raw_s = smem_u32(raw0);
    for (int seq = grp; seq < ntot && grp < ngrp; seq += ngrp) {
      const int k0 = c_idx * KC;
      if (pw == 0 && lane == 0) TF_TRACE(0, seq);
      float x[ITEMS][8];
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const int q = it * 4 + pw;                         // 32-lane group: (row block, k8 quad)
        const int rb = q / (KC / 32), kh = q % (KC / 32);
        const int row = rb * 8 + (lane & 7);
        const int k = k0 + (kh * 4 + (lane >> 3)) * 8;
        const int64_t m = t * 128 + row;
        if (T > 0) continue;  // TMA path: filled from the raw ring below
        if (m < g.M && k < g.K_pad) {
          const float* src = g.A + m * g.lda + k;
          if (vec_ok && k + 8 <= g.K_valid) {
            float4 v0 = __ldg(reinterpret_cast<const float4*>(src)), v1 = __ldg(reinterpret_cast<const float4*>(src + 4));
            x[it][0] = v0.x; x[it][1] = v0.y; x[it][2] = v0.z; x[it][3] = v0.w;
            x[it][4] = v1.x; x[it][5] = v1.y; x[it][6] = v1.z; x[it][7] = v1.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) x[it][e] = (k + e < g.K_valid) ? src[e] : 0.f;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) x[it][e] = 0.f;
        }
      }
      if (T > 0) {
        // raw slot: [128 rows][128 B], 16-byte unit u of row r stored at unit u ^ (r & 7) (SWIZZLE_128B);
        // 8 consecutive lanes read the same logical unit of 8 consecutive rows -> 8 distinct units
        const uint32_t raw = raw_s + sl * kRawBytes;
        mbar_wait_backoff(&rfull[sl], rph, 300 + seq);
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
          const int q = it * 4 + pw;
          const int rb = q / (KC / 32), kh = q % (KC / 32);
          const int row = rb * 8 + (lane & 7);
          const int u = (kh * 4 + (lane >> 3)) * 2;
          const float4 v0 = lds128(raw + row * 128 + ((u ^ (row & 7)) << 4));
          const float4 v1 = lds128(raw + row * 128 + (((u + 1) ^ (row & 7)) << 4));
          x[it][0] = v0.x; x[it][1] = v0.y; x[it][2] = v0.z; x[it][3] = v0.w;
          x[it][4] = v1.x; x[it][5] = v1.y; x[it][6] = v1.z; x[it][7] = v1.w;
        }
      }
      mbar_wait_backoff(&empty[st], ph ^ 1, 100 + seq);
      if (pw == 0 && lane == 0) TF_TRACE(1, seq);
      unsigned char* sA = stage0 + (size_t)st * stage_bytes;
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const int q = it * 4 + pw;
        const int rb = q / (KC / 32), kh = q % (KC / 32);
        const int row = rb * 8 + (lane & 7);
        const int k8 = kh * 4 + (lane >> 3);
        split_store<NSPLIT>(x[it], sA, a_tile, tile_offset(row, k8 * 8, KC));
      }
      if (pw == 0 && lane == 0) TF_TRACE(2, seq);
      if (T > 0) {
        // Release the raw slot only now: the converted values have been consumed (true data dependence on the
        // shared-memory loads).  Releasing right after ISSUING the loads let the TMA refill overwrite the slot
        // before some loads had read it (measured: ~0.2 % of rows wrong, run to run).
        __syncwarp();
        if (lane == 0) mbar_arrive(&rempty[sl]);
      }
      fence_proxy_async();
      mbar_arrive(&full[st]);
      if (pw == 0 && lane == 0) TF_TRACE(3, seq);
      // advance to this group's next chunk
      c_idx += ngrp;
      while (c_idx >= nchunks) {
        c_idx -= nchunks;
        t += gridDim.x;
      }
      st += (uint32_t)ngrp;
      if (st >= (uint32_t)S) {
        st -= (uint32_t)S;
        ph ^= 1;
      }
      if (T > 0) {
        sl += (uint32_t)ngrp;
        if (sl >= (uint32_t)T) {
          sl -= (uint32_t)T;
          rph ^= 1;
        }
      }
    }
  } else if (warp == kLoadWarp) {
    // ================= weight loader: one bulk copy per chunk =================
    if (lane == 0 && g.resident) {
      // the weights are the same for every tile: load all chunks once
      mbar_arrive_expect_tx(wfull, (uint32_t)nchunks * NSPLIT * b_tile);
      for (int c = 0; c < nchunks; ++c)
        bulk_copy_g2s(wres + (size_t)c * NSPLIT * b_tile, g.Bp + (size_t)c * NSPLIT * b_tile, NSPLIT * b_tile, wfull);
    } else if (lane == 0) {
      uint32_t st = 0, ph = 0;
      int c = 0;
      for (int seq = 0; seq < (int)total; ++seq) {
        mbar_wait_backoff(&empty[st], ph ^ 1, 200 + seq);
        unsigned char* sB = stage0 + (size_t)st * stage_bytes + NSPLIT * a_tile;
        mbar_arrive_expect_tx(&full[st], NSPLIT * b_tile);
        bulk_copy_g2s(sB, g.Bp + (size_t)c * NSPLIT * b_tile, NSPLIT * b_tile, &full[st]);
        if (++c == nchunks) c = 0;
        if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == kRawWarp) {
    // ================= TMA loader of the raw A ring: one 2-D box (128 rows x KC columns) per chunk ==========
    if (lane == 0 && T > 0) {
      uint32_t sl = 0, ph = 0;
      int c = 0;
      int64_t t = blockIdx.x;
      for (int seq = 0; seq < (int)total; ++seq) {
        mbar_wait_backoff(&rempty[sl], ph ^ 1, 400 + seq);
        mbar_arrive_expect_tx(&rfull[sl], kRawBytes);  // the whole box counts, zero-filled parts included
        tma_load_2d(raw0 + (size_t)sl * kRawBytes, &tmapA, c * KC, (int)(t * 128), &rfull[sl]);
        if (++c == nchunks) {
          c = 0;
          t += gridDim.x;
        }
        if (++sl == (uint32_t)T) { sl = 0; ph ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    // The whole warp runs this loop (warp-uniform control flow, values derived from kernel
    // parameters and loop counters only) so the compiler keeps descriptors in uniform registers;
    // just the tcgen05 instructions themselves are predicated to one elected lane.
    {
      const uint32_t idesc = make_idesc_bf16(128, g.N_pad);
      const uint32_t desc_hi = ((KC * 16) >> 4) | (1u << 14);          // SBO = KC*16 B, version 1
      const uint32_t lo_const = (128u >> 4) << 16;                       // LBO = 128 B
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t a_lo0 = ((smem_base + (uint32_t)g.header_bytes + (g.resident ? (uint32_t)nchunks_w * NSPLIT * b_tile : 0u)) >> 4) | lo_const;
      const uint32_t b_lo0 = g.resident ? (((smem_base + (uint32_t)g.header_bytes) >> 4) | lo_const) : a_lo0 + ((NSPLIT * a_tile) >> 4);
      const uint32_t a_split = a_tile >> 4, b_split = b_tile >> 4, st_units = stage_bytes >> 4;
      const uint32_t w_chunk = (NSPLIT * b_tile) >> 4;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t acc = 0, acc_ph = 0;
      uint32_t st = 0, ph = 0;
      if (g.resident) mbar_wait(wfull, 0, 250);
      for (int64_t ti = 0; ti < my_tiles; ++ti) {
        mbar_wait(&tempty[acc], acc_ph ^ 1, 300);
        tc_fence_after();
        const uint32_t d = tmem_u + acc * (uint32_t)g.N_pad;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&full[st], ph, 400);
          tc_fence_after();
          const uint32_t a_st = a_lo0 + st * st_units;
          const uint32_t b_st = g.resident ? b_lo0 + (uint32_t)c * w_chunk : b_lo0 + st * st_units;
          const int nks = min(KC / 16, (g.K_pad - c * KC) / 16);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks)
              if (ks < nks) umma_split_lo<NSPLIT>(d, a_st + ks * 16, a_split, b_st + ks * 16, b_split, desc_hi, idesc, (c | ks) == 0);
            umma_commit(&empty[st]);  // smem stage free once these MMAs have read it
          }
          __syncwarp();
          if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(&tfull[acc]);  // accumulator complete
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else {
    // ================= epilogue: TMEM -> registers -> bias / relu / mask -> global ====================
    // Two warps share a TMEM lane quadrant and split the columns. tcgen05.ld.16x256b hands every group
    // of 4 lanes 32 contiguous bytes of one row (the mma C-fragment layout), so the registers are stored
    // straight to global memory as whole 32-byte sectors: no staging rows in shared memory and no
    // per-thread bulk stores (issuing 256 small cp.async.bulk per tile cost ~3k cycles of an ~8k epilogue).
    // A thread owns 4 rows of the tile (rr = 2*blk + h -> quad*32 + blk*16 + h*8 + lane/4) and columns
    // 2*(lane%4) + {0,1} of every 8-column group.  Row-wise results (ReLU bit masks, the fused output
    // layer) are combined across the 4 lanes of a row with two xor-shuffles.
    // ReLU masks travel as one bit per element (bits_out / bits_in), not as fp32 activations.
    // The 8 warps form two sets of 4 (one warp per TMEM lane quadrant); set s owns the CTA's tiles s, s+2, ...
    // and TMEM accumulator s, and handles all columns of its tiles.  Two tiles are in the epilogue at once, so
    // one set's per-tile overhead (accumulator wait, row set-up) overlaps the other set's column loop.
    const int quad = warp & 3, eset = warp >> 2;
    const uint32_t acc = (uint32_t)eset;
    uint32_t acc_ph = 0;
    const int lrow = lane >> 2, lq = lane & 3, lc = lq * 2;
    const int c_begin = 0, c_cols = g.N_pad;
    const bool vec_ok = (g.ldc % 2 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 7) == 0);
    // every column of the tile is stored and rows are 8-byte aligned: unconditional 8-byte stores
    const bool fast_store = vec_ok && g.N_store == g.N_pad;
    const float relu_floor = g.relu ? 0.f : -INFINITY;
    const int words = (g.N_pad + 31) / 32;
    for (int64_t t = blockIdx.x + (int64_t)eset * gridDim.x; t < ntiles; t += 2 * (int64_t)gridDim.x) {
      const int64_t mbase = t * 128 + quad * 32 + lrow;  // row rr = mbase + roff(rr)
#define TF_ROFF(rr) ((((rr) >> 1) * 16) + (((rr) & 1) * 8))
      int em_off[4];  // FiLM conditioner row of each row's camera (networks.py:103-111), -1 = none
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        em_off[rr] = -1;
        if ((EPI & EPI_FILM) && mbase + TF_ROFF(rr) < g.M)
          em_off[rr] = (int)g.cams[(mbase + TF_ROFF(rr)) / g.rows_per_ray] * 128;
      }
      int ray[4];  // EPI_ENC: viewdir row of each row's ray
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) ray[rr] = (EPI & EPI_ENC) ? (int)(min(mbase + TF_ROFF(rr), g.M - 1) / g.rows_per_ray) : 0;
      // this thread's first column of each of its rows; rows outside the matrix are clamped (their stores are
      // predicated off by row_ok), so the pointers advance unconditionally
      float* crow[4];
      uint32_t row_ok = 0;
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int64_t mr = mbase + TF_ROFF(rr);
        if (g.C != nullptr && mr < g.M) row_ok |= 1u << rr;
        crow[rr] = g.C + min(mr, g.M - 1) * g.ldc + c_begin + lc;
      }
      float o3[4][3];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) o3[rr][0] = o3[rr][1] = o3[rr][2] = 0.f;
      if (tid == 0) TF_TRACE(4, (t - blockIdx.x) / gridDim.x);
      mbar_wait(&tfull[acc], acc_ph, 500);
      if (tid == 0) TF_TRACE(6, (t - blockIdx.x) / gridDim.x);
      tc_fence_after();
      const uint32_t tcol = tmem_base + acc * (uint32_t)g.N_pad;
      for (int j = 0; j * 16 < c_cols; ++j) {
        const int n0 = c_begin + j * 16;
        float v[2][8];
        tmem_ld_16x256b_x2(tcol + ((uint32_t)(quad * 32) << 16) + n0, v[0]);
        tmem_ld_16x256b_x2(tcol + ((uint32_t)(quad * 32 + 16) << 16) + n0, v[1]);
        tmem_ld_wait();
        const float2 b0 = *reinterpret_cast<const float2*>(s_bias + n0 + lc);
        const float2 b1 = *reinterpret_cast<const float2*>(s_bias + n0 + 8 + lc);
        const float bq[4] = {b0.x, b0.y, b1.x, b1.y};
        uint32_t ow[4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int blk = rr >> 1, h = rr & 1;
          const int64_t mr = mbase + TF_ROFF(rr);
          uint32_t mw = 0u;
          if (EPI & EPI_BITS_IN) mw = mr < g.M ? (__ldg(g.bits_in + mr * words + (n0 >> 5)) >> (n0 & 16)) : 0u;
          ow[rr] = 0u;
#pragma unroll
          for (int e = 0; e < 4; ++e) {               // e = 2*(8-column group) + (column within the pair)
            const int q = (e >> 1) * 8 + lc + (e & 1);  // column within the 16-column chunk
            const float y0 = v[blk][(e >> 1) * 4 + h * 2 + (e & 1)] + bq[e];
            float y = fmaxf(y0, relu_floor);  // relu_floor = 0 or -inf: no per-element branch
            if ((EPI & EPI_BITS_IN) && !((mw >> q) & 1u)) y = 0.f;
            if ((EPI & EPI_BITS_OUT) && y > 0.f) ow[rr] |= 1u << q;
            v[blk][(e >> 1) * 4 + h * 2 + (e & 1)] = y;
            if (EPI & EPI_ENC) {
              // networks.py:13-35, :68-76: x = [f, v, enc(f), enc(v)]; this thread owns source dimension n of row mr
              // (n < squash: the accumulator; the next 3 padding columns stand in for the view direction)
              const int n = n0 + q;
              if (mr < g.M && n < g.enc_squash + 3) {
                const int D = g.enc_squash + 3;
                float val = y;
                int F = g.enc_Ff, off = D + n * 2 * g.enc_Ff;
                if (n >= g.enc_squash) {
                  val = __ldg(g.enc_vd + (int64_t)ray[rr] * 3 + (n - g.enc_squash));
                  F = g.enc_Fv;
                  off = D + g.enc_squash * 2 * g.enc_Ff + (n - g.enc_squash) * 2 * g.enc_Fv;
                }
                float* xr = g.enc_x + mr * g.enc_ldx;
                xr[n] = val;
                // Even frequency levels are evaluated directly (sincosf shares one range reduction; cos(a) stands in
                // for the reference's sin(a + pi/2), which differs by the rounding of a + pi/2 only); odd levels come
                // from the double-angle identities, ~3 ulp.  Half the transcendental work of 2F sinf calls.
                float scale = 1.0f;
                for (int jf = 0; jf < F; jf += 2) {
                  float sn, cs;
                  sincosf(val * scale, &sn, &cs);  // val * 2^jf is exact
                  const float sn2 = 2.0f * sn * cs, cs2 = fmaf(-2.0f * sn, sn, 1.0f);
                  if (jf + 1 < F && ((reinterpret_cast<uintptr_t>(xr + off + jf) | reinterpret_cast<uintptr_t>(xr + off + F + jf)) & 7) == 0) {
                    *reinterpret_cast<float2*>(xr + off + jf) = make_float2(sn, sn2);
                    *reinterpret_cast<float2*>(xr + off + F + jf) = make_float2(cs, cs2);
                  } else {
                    xr[off + jf] = sn;
                    xr[off + F + jf] = cs;
                    if (jf + 1 < F) {
                      xr[off + jf + 1] = sn2;
                      xr[off + F + jf + 1] = cs2;
                    }
                  }
                  scale *= 4.0f;
                }
              }
            }
            if (EPI & EPI_OUT3) {
              const int n = n0 + q;
              float yf = y;
              if ((EPI & EPI_FILM) && em_off[rr] >= 0 && n >= 64) yf = __fadd_rn(__fmul_rn(__ldg(g.embed + em_off[rr] + n - 64), y), __ldg(g.embed + em_off[rr] + n));
              o3[rr][0] = fmaf(yf, s_w3[3 * n + 0], o3[rr][0]);
              o3[rr][1] = fmaf(yf, s_w3[3 * n + 1], o3[rr][1]);
              o3[rr][2] = fmaf(yf, s_w3[3 * n + 2], o3[rr][2]);
            }
          }
        }
        if (EPI & EPI_BITS_OUT) {
          // 16 mask bits per (row, chunk): OR over the row's 4 lanes (two rows per 32-bit word), then lane
          // lq writes row rr == lq's half-word
          uint32_t p01 = ow[0] | (ow[1] << 16), p23 = ow[2] | (ow[3] << 16);
          p01 |= __shfl_xor_sync(0xffffffffu, p01, 1);
          p23 |= __shfl_xor_sync(0xffffffffu, p23, 1);
          p01 |= __shfl_xor_sync(0xffffffffu, p01, 2);
          p23 |= __shfl_xor_sync(0xffffffffu, p23, 2);
          const uint32_t mine = (lq & 2) ? p23 : p01;
          const int64_t mq = mbase + TF_ROFF(lq);
          if (mq < g.M) reinterpret_cast<uint16_t*>(g.bits_out + mq * words)[n0 >> 4] = (uint16_t)(mine >> ((lq & 1) * 16));
        }
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int blk = rr >> 1, h = rr & 1;
          if (!((row_ok >> rr) & 1u)) {
          } else if (fast_store) {
            *reinterpret_cast<float2*>(crow[rr]) = make_float2(v[blk][h * 2], v[blk][h * 2 + 1]);
            *reinterpret_cast<float2*>(crow[rr] + 8) = make_float2(v[blk][4 + h * 2], v[blk][4 + h * 2 + 1]);
          } else {
#pragma unroll
            for (int gq = 0; gq < 2; ++gq) {
              const int n = n0 + gq * 8 + lc;
              float* dst = crow[rr] + gq * 8;
              const float y0 = v[blk][gq * 4 + h * 2], y1 = v[blk][gq * 4 + h * 2 + 1];
              if (vec_ok && n + 1 < g.N_store) {
                *reinterpret_cast<float2*>(dst) = make_float2(y0, y1);
              } else {
                if (n < g.N_store) dst[0] = y0;
                if (n + 1 < g.N_store) dst[1] = y1;
              }
            }
          }
          crow[rr] += 16;
        }
      }
      if (tid == 0) TF_TRACE(5, (t - blockIdx.x) / gridDim.x);
      if (EPI & EPI_OUT3) {
        // sum the row's partial outputs over its 4 lanes, apply the sigmoid, store
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            o3[rr][c] += __shfl_xor_sync(0xffffffffu, o3[rr][c], 1);
            o3[rr][c] += __shfl_xor_sync(0xffffffffu, o3[rr][c], 2);
          }
        float mine3[3];  // lane lq keeps row rr == lq
#pragma unroll
        for (int c = 0; c < 3; ++c) mine3[c] = lq == 0 ? o3[0][c] : lq == 1 ? o3[1][c] : lq == 2 ? o3[2][c] : o3[3][c];
        const int64_t mm = mbase + TF_ROFF(lq);
        if (mm < g.M) {
#pragma unroll
          for (int c = 0; c < 3; ++c) g.rgb_out[3 * mm + c] = 1.0f / (1.0f + expf(-(mine3[c] + s_w3[384 + c])));
        }
      }
      if (tid == 0) TF_TRACE(7, (t - blockIdx.x) / gridDim.x);
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc_ph ^= 1;
    }
#undef TF_ROFF
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, g.tmem_cols);
  }
}

static uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// Tensor map over the fp32 activation matrix A (M rows, K_valid columns, row stride lda): box = 128 rows x
// KC columns (one A chunk), 128-byte swizzle, zero fill outside the matrix (row and column tails).
bool tc_make_a_tensor_map(void* tm_, const float* A, int64_t M, int K_valid, int64_t lda, int kc);
static bool make_a_tensor_map(CUtensorMap* tm, const float* A, int64_t M, int K_valid, int64_t lda, int kc) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || (reinterpret_cast<uintptr_t>(A) & 15) != 0 || (lda * 4) % 16 != 0 || K_valid < 1 || M < 1 || kc * 4 != 128) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)K_valid, (cuuint64_t)M};
  const cuuint64_t strides[1] = {(cuuint64_t)lda * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kc, 128};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(A), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool tc_make_a_tensor_map(void* tm_, const float* A, int64_t M, int K_valid, int64_t lda, int kc) {
  return make_a_tensor_map(reinterpret_cast<CUtensorMap*>(tm_), A, M, K_valid, lda, kc);
}

template <int NSPLIT, int EPI>
static int launch_rowgemm_epi(cudaStream_t st, const RowGemmArgs& g, const CUtensorMap& tm, unsigned grid, size_t smem) {
  TF_CHECK_CUDA(cudaFuncSetAttribute(k_tc_rowgemm<NSPLIT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_tc_rowgemm<NSPLIT, EPI><<<grid, kRowThreads, smem, st>>>(g, tm);
  TF_CHECK_LAUNCH();
  return 0;
}

template <int NSPLIT>
static int launch_rowgemm(cudaStream_t st, RowGemmArgs g) {
  constexpr int KC = RowCfg<NSPLIT>::KC;
  if (g.M == 0) return 0;
  TF_CHECK_ARG(g.N_pad % 16 == 0 && g.N_pad >= 16 && g.N_pad <= 256, "tc rowgemm: N_pad=%d unsupported", g.N_pad);
  TF_CHECK_ARG(g.K_pad % 16 == 0 && g.K_pad >= 16, "tc rowgemm: K_pad=%d unsupported", g.K_pad);
  const size_t stage = (size_t)NSPLIT * (tile_bytes(128, KC) + tile_bytes(g.N_pad, KC));
  g.header_bytes = kRowFixed;
  // weights resident in smem for the whole kernel when they fit next to >= 2 A-only stages
  const size_t wres = (size_t)((g.K_pad + KC - 1) / KC) * NSPLIT * tile_bytes(g.N_pad, KC);
  const size_t a_stage = (size_t)NSPLIT * tile_bytes(128, KC);
  constexpr size_t kRaw = (size_t)128 * KC * 4;
  // raw A ring by TMA when A is 16-byte aligned with 16-byte row pitch
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  const bool tma = !getenv("TENSORF_TC_NO_TMA") && make_a_tensor_map(&tm, g.A, g.M, g.K_valid, g.lda, KC);
  g.resident = 0;
  size_t stage_sz = stage, base = g.header_bytes;
  if (!getenv("TENSORF_TC_STREAM_W") && g.header_bytes + wres + (tma ? kGroups * a_stage + 3 * kRaw + 1024 : 2 * a_stage) <= 227 * 1024) {
    g.resident = 1;
    stage_sz = a_stage;
    base += wres;
  }
  size_t avail = 227 * 1024 - base;
  g.raw_slots = 0;
  if (tma) {
    // the raw ring hides the global latency: give it up to 6 slots, keep at least one operand stage per group
    avail -= 1024;  // alignment slack of the ring (1024-byte swizzle atoms)
    const size_t min_stages = (size_t)kGroups * stage_sz;
    if (avail >= min_stages + kGroups * kRaw) {
      // A producer group may only wait on a raw slot whose previous occupant it consumed itself (a parity wait
      // cannot tell "previous fill not landed yet" from "my fill landed"): slot count = multiple of the group count
      g.raw_slots = (int)std::min<size_t>(2 * kGroups, (avail - min_stages) / kRaw) / kGroups * kGroups;
      if (const char* e = getenv("TENSORF_TC_RAW_SLOTS")) g.raw_slots = std::max(kGroups, std::min(g.raw_slots, atoi(e)) / kGroups * kGroups);
      avail -= (size_t)g.raw_slots * kRaw;
    } else {
      avail += 1024;
    }
  }
  int stages = (int)std::min<size_t>(g.raw_slots ? 6 : 8, avail / stage_sz);
  TF_CHECK_ARG(stages >= 2, "tc rowgemm: tile too large for shared memory");
  if (const char* e = getenv("TENSORF_TC_STAGES")) stages = std::max(2, std::min(stages, atoi(e)));
  g.trace = getenv("TENSORF_TC_TRACE") != nullptr;
  // A stage must always be filled by the same producer group, in order: mbarrier parity waits
  // cannot tell "two phases ahead" from "not yet", so the stage count is a multiple of the
  // number of active groups (chunk seq -> group seq % groups, stage seq % stages).
  g.groups = std::min(kGroups, stages);
  stages = stages / g.groups * g.groups;
  g.stages = stages;
  g.tmem_cols = pow2_cols(2 * g.N_pad);
  const size_t smem = base + stages * stage_sz + (g.raw_slots ? 1024 + (size_t)g.raw_slots * kRaw : 0);
  const int64_t ntiles = (g.M + 127) / 128;
  const unsigned grid = (unsigned)std::min<int64_t>(ntiles, kSMs);
  if (getenv("TENSORF_TC_DEBUG"))
    fprintf(stderr, "rowgemm<%d> M=%lld K=%d/%d N=%d resident=%d stages=%d groups=%d raw_slots=%d smem=%zu\n", NSPLIT, (long long)g.M,
            g.K_valid, g.K_pad, g.N_pad, g.resident, g.stages, g.groups, g.raw_slots, smem);
  const int epi = (g.bits_in ? EPI_BITS_IN : 0) | (g.bits_out ? EPI_BITS_OUT : 0) | (g.rgb_out ? EPI_OUT3 : 0) | (g.enc_x ? EPI_ENC : 0) |
                  (g.rgb_out && g.embed ? EPI_FILM : 0);
  switch (epi) {
    case 0: return launch_rowgemm_epi<NSPLIT, 0>(st, g, tm, grid, smem);
    case EPI_ENC: return launch_rowgemm_epi<NSPLIT, EPI_ENC>(st, g, tm, grid, smem);
    case EPI_BITS_IN: return launch_rowgemm_epi<NSPLIT, EPI_BITS_IN>(st, g, tm, grid, smem);
    case EPI_BITS_OUT: return launch_rowgemm_epi<NSPLIT, EPI_BITS_OUT>(st, g, tm, grid, smem);
    case EPI_BITS_IN | EPI_BITS_OUT: return launch_rowgemm_epi<NSPLIT, EPI_BITS_IN | EPI_BITS_OUT>(st, g, tm, grid, smem);
    case EPI_OUT3: return launch_rowgemm_epi<NSPLIT, EPI_OUT3>(st, g, tm, grid, smem);
    case EPI_OUT3 | EPI_FILM: return launch_rowgemm_epi<NSPLIT, EPI_OUT3 | EPI_FILM>(st, g, tm, grid, smem);
    default: set_error("tc rowgemm: unsupported epilogue combination %d", epi); return TENSORF_ERR_UNSUPPORTED;
  }
}

// ---------------------------------------------------------------------------------------------
// k_tc_redgemm: T[m][n] = sum_rows G[row][m] * X[row][n];  out[n*ldo + m] += T[m][n]
// ---------------------------------------------------------------------------------------------
struct RedGemmArgs {
  const float* G;  // (rows, ldg), m < Mg valid columns (<= 128)
  int64_t ldg;
  int Mg;
  const float* X;  // (rows, ldx), n < Nx valid columns
  int64_t ldx;
  int Nx, N_pad;   // N_pad multiple of 16, <= 512
  int64_t rows;
  int64_t rows_per_cta;  // multiple of kRC
  float* out;
  int64_t ldo;
  int trace;
  float* colsum;   // optional: colsum[m] += sum_rows G[row][m] via a virtual all-ones column n == Nx of X (N_pad > Nx)
  int stages, groups;
  uint32_t tmem_cols;
};

// warps 0-3 epilogue, 4 MMA issuer, then kGroups producer groups of 4 warps (chunk c -> group c % kGroups)
constexpr int kRedGroups = 4;
constexpr int kRedGroupThreads = 96;
constexpr int kRedThreads = 160 + kRedGroups * kRedGroupThreads;

__global__ void __launch_bounds__(kRedThreads, 1) k_tc_redgemm(RedGemmArgs g) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = g.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + S;
  uint64_t* tfull = empty + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  const uint32_t a_tile = tile_bytes(128, kRC), b_tile = tile_bytes(g.N_pad, kRC);
  const uint32_t stage_bytes = 2 * a_tile + 2 * b_tile;
  unsigned char* stage0 = smem + 1024;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], kRedGroupThreads);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, g.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t r_begin = (int64_t)blockIdx.x * g.rows_per_cta;
  const int64_t r_end = min(g.rows, r_begin + g.rows_per_cta);
  const int nchunks = (int)((r_end - r_begin + kRC - 1) / kRC);

  if (warp >= 5) {
    // ===== producers: item = 8 rows x 4 columns (8 coalesced float4 loads), transposed in
    // registers into four 16-byte k-chunks (one per column) =====
    const int grp = (warp - 5) / 3, pt = (tid - 160) % kRedGroupThreads;
    const bool vecA = (g.ldg % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.G) & 15) == 0);
    const bool vecB = (g.ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.X) & 15) == 0);
    const int a_items = 32 * (kRC / 8), b_items = (g.N_pad / 4) * (kRC / 8);
    for (int c = grp; c < nchunks && grp < g.groups; c += g.groups) {
      const uint32_t st = (uint32_t)(c % S), ph = (uint32_t)((c / S) & 1);
      unsigned char* sA = stage0 + (size_t)st * stage_bytes;
      unsigned char* sB = sA + 2 * a_tile;
      const int64_t r0 = r_begin + (int64_t)c * kRC;
      bool waited = false;
      if (pt == 0) TF_TRACE(0, c);
      for (int base = 0; base < a_items + b_items; base += 2 * kRedGroupThreads) {
        float x[2][8][4];
        int col4[2], jj[2];
        bool isA[2], live[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int item = base + u * kRedGroupThreads + pt;
          live[u] = item < a_items + b_items;
          isA[u] = item < a_items;
          const int e = isA[u] ? item : item - a_items;
          const int nc4 = isA[u] ? 32 : g.N_pad / 4;
          col4[u] = e % nc4;
          jj[u] = e / nc4;
          const float* src = isA[u] ? g.G : g.X;
          const int64_t ld = isA[u] ? g.ldg : g.ldx;
          const int valid = isA[u] ? g.Mg : g.Nx;
          const bool vec = isA[u] ? vecA : vecB;
          const int c0 = col4[u] * 4;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int64_t r = r0 + jj[u] * 8 + q;
            if (live[u] && r < r_end && c0 < valid) {
              if (vec && c0 + 4 <= valid) {
                float4 v = __ldg(reinterpret_cast<const float4*>(src + r * ld + c0));
                x[u][q][0] = v.x; x[u][q][1] = v.y; x[u][q][2] = v.z; x[u][q][3] = v.w;
              } else {
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4)
                  x[u][q][e4] = (c0 + e4 < valid) ? src[r * ld + c0 + e4] : ((!isA[u] && g.colsum && c0 + e4 == g.Nx) ? 1.0f : 0.f);
              }
            } else if (live[u] && r < r_end && !isA[u] && g.colsum && c0 <= g.Nx && g.Nx < c0 + 4) {
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) x[u][q][e4] = (c0 + e4 == g.Nx) ? 1.0f : 0.f;  // ones column -> bias gradient
            } else {
              x[u][q][0] = x[u][q][1] = x[u][q][2] = x[u][q][3] = 0.f;
            }
          }
        }
        if (!waited) {
          if (pt == 0) TF_TRACE(1, c);
          mbar_wait_backoff(&empty[st], ph ^ 1);
          waited = true;
          if (pt == 0) TF_TRACE(2, c);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (!live[u]) continue;
          unsigned char* tile = isA[u] ? sA : sB;
          const uint32_t tb = isA[u] ? a_tile : b_tile;
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            float col[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) col[q] = x[u][q][e4];
            split_store<2>(col, tile, tb, tile_offset(col4[u] * 4 + e4, jj[u] * 8, kRC));
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(&full[st]);
      if (pt == 0) TF_TRACE(3, c);
    }
  } else if (warp == 4) {
    {  // warp-uniform issue loop (see k_tc_rowgemm): descriptors stay in uniform registers
      const uint32_t desc_hi = ((kRC * 16) >> 4) | (1u << 14);
      const uint32_t lo_const = (128u >> 4) << 16;
      const uint32_t a_lo0 = ((smem_u32(smem) + 1024u) >> 4) | lo_const;
      const uint32_t a_split = a_tile >> 4, b_split = b_tile >> 4, st_units = stage_bytes >> 4, sbo_units = (kRC * 16) >> 4;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t st = 0, ph = 0;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t a_st = a_lo0 + st * st_units;
        const uint32_t b_st = a_st + 2 * a_split;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kRC / 16; ++ks) {
            for (int n0 = 0; n0 < g.N_pad; n0 += 256) {
              const int nn = min(256, g.N_pad - n0);
              const uint32_t idesc = make_idesc_bf16(128, nn);
              // rows n0.. of the B tile start (n0/8) core-matrix rows further
              umma_split_lo<2>(tmem_u + n0, a_st + ks * 16, a_split, b_st + (uint32_t)(n0 >> 3) * sbo_units + ks * 16, b_split,
                               desc_hi, idesc, (c | ks) == 0);
            }
          }
          umma_commit(&empty[st]);
        }
        __syncwarp();
        if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(tfull);
      __syncwarp();
    }
  } else {
    // ===== epilogue: RED-add the 128 x N tile into out[n*ldo + m] =====
    if (nchunks > 0) {
      mbar_wait(tfull, 0);
      if (tid == 0) TF_TRACE(6, 0);
      tc_fence_after();
      const int m = warp * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
      const int nblk = g.N_pad / 16;
      for (int i = 0; i < nblk; ++i) {
        const int n0 = ((i + (int)blockIdx.x) % nblk) * 16;  // CTAs start at different columns
        float v[16];
        tmem_ld16(trow + n0, v);
        tmem_ld_wait();
        if (m < g.Mg) {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            if (n0 + q < g.Nx) atomicAdd(g.out + (int64_t)(n0 + q) * g.ldo + m, v[q]);
            else if (g.colsum && n0 + q == g.Nx) atomicAdd(g.colsum + m, v[q]);
          }
        }
      }
    }
  }
  if (tid == 0) TF_TRACE(7, 0);
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, g.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// k_tc_redgemm2: same contraction, asynchronous loads. A 32-row chunk of G and of X is one
// contiguous block in global memory (ld % 4 == 0), so ONE thread streams raw fp32 chunks into a
// shared-memory staging ring with cp.async.bulk (deep prefetch, no registers, no LSU), and the
// converter warps transpose + split them smem -> smem into the UMMA operand tiles.
//   warps 0-3 epilogue, 4 MMA issuer, 5 loader, 6.. converters: kConvGroups groups that own alternate
//   chunks, so one group's barrier/fence latency overlaps the other group's conversion (stage and slot
//   counts are multiples of the group count: a stage is always filled by the same group, in order).
// ---------------------------------------------------------------------------------------------
constexpr int kConvGroups = 2;
constexpr int kConvGroupWarps = 8;
constexpr int kConvWarps = kConvGroups * kConvGroupWarps;
constexpr int kRed2Threads = 192 + 32 * kConvWarps;

// RC = rows per chunk (the MMA K extent of one operand stage): 32, or 16 when the wider staging of 32 rows does not fit
// (dozer: X has 400 columns).
template <int RC>
__global__ void __launch_bounds__(kRed2Threads, 1) k_tc_redgemm2(RedGemmArgs g, int T /*staging slots*/) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = g.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [S] operand stage converted
  uint64_t* empty = full + S;                           // [S] operand stage consumed (tcgen05.commit)
  uint64_t* sfull = empty + S;                          // [T] staging slot loaded (tx bytes)
  uint64_t* sempty = sfull + T;                         // [T] staging slot read by all converter warps
  uint64_t* tfull = sempty + T;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  const uint32_t a_tile = tile_bytes(128, RC), b_tile = tile_bytes(g.N_pad, RC);
  const uint32_t stage_bytes = 2 * a_tile + 2 * b_tile;
  const uint32_t slot_floats = (uint32_t)RC * (uint32_t)(g.ldg + g.ldx);
  unsigned char* stage0 = smem + 1024;
  float* slot0 = reinterpret_cast<float*>(stage0 + (size_t)S * stage_bytes);

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], kConvGroupWarps);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < T; ++s) {
      mbar_init(&sfull[s], 1);
      mbar_init(&sempty[s], kConvGroupWarps);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, g.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t r_begin = (int64_t)blockIdx.x * g.rows_per_cta;
  const int64_t r_end = min(g.rows, r_begin + g.rows_per_cta);
  const int nchunks = (int)((r_end - r_begin + RC - 1) / RC);

  if (warp == 5) {
    // ===== loader: two bulk copies per chunk (raw fp32 rows of G and X) =====
    if (lane == 0) {
      for (int c = 0; c < nchunks; ++c) {
        const uint32_t sl = (uint32_t)(c % T), ph = (uint32_t)((c / T) & 1);
        mbar_wait_backoff(&sempty[sl], ph ^ 1);
        TF_TRACE(4, c);
        const int64_t r0 = r_begin + (int64_t)c * RC;
        const uint32_t nrows = (uint32_t)min((int64_t)RC, r_end - r0);
        float* dst = slot0 + (size_t)sl * slot_floats;
        const uint32_t bg = nrows * (uint32_t)g.ldg * 4, bx = nrows * (uint32_t)g.ldx * 4;
        mbar_arrive_expect_tx(&sfull[sl], bg + bx);
        bulk_copy_g2s(dst, g.G + r0 * g.ldg, bg, &sfull[sl]);
        bulk_copy_g2s(dst + (size_t)RC * g.ldg, g.X + r0 * g.ldx, bx, &sfull[sl]);
      }
    }
  } else if (warp >= 6) {
    // ===== converters: item = (operand, column, group of 8 rows) -> one 16-byte k-chunk per split =====
    const int grp = (warp - 6) / kConvGroupWarps;
    const int ct = tid - 192 - grp * 32 * kConvGroupWarps;
    const int a_items = 128 * (RC / 8), b_items = g.N_pad * (RC / 8);
    uint32_t sl = (uint32_t)grp, sph = 0, st = (uint32_t)grp, ph = 0;  // advanced incrementally (no run-time divisions)
    for (int c = grp; c < nchunks; c += kConvGroups) {
      const int64_t r0 = r_begin + (int64_t)c * RC;
      const int nrows = (int)min((int64_t)RC, r_end - r0);
      const float* sG = slot0 + (size_t)sl * slot_floats;
      const float* sX = sG + (size_t)RC * g.ldg;
      unsigned char* sA = stage0 + (size_t)st * stage_bytes;
      unsigned char* sB = sA + 2 * a_tile;
      if (ct == 0) TF_TRACE(0, c);
      mbar_wait(&sfull[sl], sph);
      if (ct == 0) TF_TRACE(1, c);
      mbar_wait(&empty[st], ph ^ 1);
      if (ct == 0) TF_TRACE(2, c);
      const bool full_chunk = nrows == RC;
      for (int item = ct; item < a_items + b_items; item += 32 * kConvGroupWarps) {
        const bool isA = item < a_items;
        int col, j;
        if (isA) {  // 128 columns: shifts
          col = item & 127;
          j = item >> 7;
        } else {    // N_pad columns, RC/8 <= 4 row groups: compares instead of a division
          const int e = item - a_items;
          j = (e >= g.N_pad) + (RC > 16 ? (e >= 2 * g.N_pad) + (e >= 3 * g.N_pad) : 0);
          col = e - j * g.N_pad;
        }
        const float* src = (isA ? sG : sX) + col;
        const int ld = isA ? (int)g.ldg : (int)g.ldx;
        const int valid = isA ? g.Mg : g.Nx;
        float x[8];
        if (col < valid) {
          if (full_chunk) {
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = src[(j * 8 + q) * ld];
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = (j * 8 + q < nrows) ? src[(j * 8 + q) * ld] : 0.f;
          }
        } else {
          const bool ones = !isA && g.colsum && col == g.Nx;  // ones column -> bias gradient
#pragma unroll
          for (int q = 0; q < 8; ++q) x[q] = (ones && j * 8 + q < nrows) ? 1.0f : 0.f;
        }
        split_store<2>(x, isA ? sA : sB, isA ? a_tile : b_tile, tile_offset(col, j * 8, RC));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&full[st]);     // operand stage ready for the tensor core
        mbar_arrive(&sempty[sl]);   // staging slot may be overwritten
      }
      if (ct == 0) TF_TRACE(3, c);
      sl += kConvGroups;
      if (sl >= (uint32_t)T) { sl -= (uint32_t)T; sph ^= 1; }
      st += kConvGroups;
      if (st >= (uint32_t)S) { st -= (uint32_t)S; ph ^= 1; }
    }
  } else if (warp == 4) {
    {  // warp-uniform issue loop (see k_tc_rowgemm): descriptors stay in uniform registers
      const uint32_t desc_hi = ((RC * 16) >> 4) | (1u << 14);
      const uint32_t lo_const = (128u >> 4) << 16;
      const uint32_t a_lo0 = ((smem_u32(smem) + 1024u) >> 4) | lo_const;
      const uint32_t a_split = a_tile >> 4, b_split = b_tile >> 4, st_units = stage_bytes >> 4, sbo_units = (RC * 16) >> 4;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t st = 0, ph = 0;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&full[st], ph);
        if (lane == 0) TF_TRACE(5, c);
        tc_fence_after();
        const uint32_t a_st = a_lo0 + st * st_units;
        const uint32_t b_st = a_st + 2 * a_split;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < RC / 16; ++ks) {
            for (int n0 = 0; n0 < g.N_pad; n0 += 256) {
              const int nn = min(256, g.N_pad - n0);
              const uint32_t idesc = make_idesc_bf16(128, nn);
              // rows n0.. of the B tile start (n0/8) core-matrix rows further
              umma_split_lo<2>(tmem_u + n0, a_st + ks * 16, a_split, b_st + (uint32_t)(n0 >> 3) * sbo_units + ks * 16, b_split,
                               desc_hi, idesc, (c | ks) == 0);
            }
          }
          umma_commit(&empty[st]);
        }
        __syncwarp();
        if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(tfull);
      __syncwarp();
    }
  } else {
    // ===== epilogue: RED-add the 128 x N tile into out[n*ldo + m] =====
    if (nchunks > 0) {
      mbar_wait(tfull, 0);
      if (tid == 0) TF_TRACE(6, 0);
      tc_fence_after();
      const int m = warp * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
      const int nblk = g.N_pad / 16;
      for (int i = 0; i < nblk; ++i) {
        const int n0 = ((i + (int)blockIdx.x) % nblk) * 16;  // CTAs start at different columns
        float v[16];
        tmem_ld16(trow + n0, v);
        tmem_ld_wait();
        if (m < g.Mg) {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            if (n0 + q < g.Nx) atomicAdd(g.out + (int64_t)(n0 + q) * g.ldo + m, v[q]);
            else if (g.colsum && n0 + q == g.Nx) atomicAdd(g.colsum + m, v[q]);
          }
        }
      }
    }
  }
  if (tid == 0) TF_TRACE(7, 0);
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, g.tmem_cols);
  }
}

static int launch_redgemm(cudaStream_t st, RedGemmArgs g) {
  if (g.rows == 0) return 0;
  TF_CHECK_ARG(g.N_pad % 16 == 0 && g.N_pad <= 512 && g.Mg <= 128, "tc redgemm: shape unsupported (Mg=%d N_pad=%d)", g.Mg, g.N_pad);
  g.tmem_cols = pow2_cols(g.N_pad);
  g.trace = getenv("TENSORF_TC_TRACE") != nullptr;
  int64_t ctas = std::min<int64_t>(kSMs, std::max<int64_t>(1, g.rows / 256));
  g.rows_per_cta = round_up64(ceil_div64(g.rows, ctas), kRC);
  ctas = ceil_div64(g.rows, g.rows_per_cta);
  const size_t stage = 2 * (size_t)tile_bytes(128, kRC) + 2 * (size_t)tile_bytes(g.N_pad, kRC);
  // asynchronous-load variant: rows of G and X must be 16-byte multiples, and 2 operand stages plus
  // >= 2 staging slots must fit in shared memory (with 32 rows per chunk, else with 16)
  const bool aligned = g.ldg % 4 == 0 && g.ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(g.G) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(g.X) & 15) == 0;
  if (aligned && !getenv("TENSORF_TC_RED_REGS")) {
    for (int rc = 32; rc >= 16; rc >>= 1) {
      const size_t stg = 2 * (size_t)tile_bytes(128, rc) + 2 * (size_t)tile_bytes(g.N_pad, rc);
      const size_t slot = (size_t)rc * (size_t)(g.ldg + g.ldx) * 4;
      if (1024 + 2 * stg + 2 * slot > 227 * 1024) continue;
      g.stages = kConvGroups;
      g.groups = kConvGroups;
      int T = (int)std::min<size_t>(6, (227 * 1024 - 1024 - 2 * stg) / slot);
      T = T / kConvGroups * kConvGroups;
      const size_t smem = 1024 + g.stages * stg + (size_t)T * slot;
      // rows per CTA: whole chunks
      int64_t nct = std::min<int64_t>(kSMs, std::max<int64_t>(1, g.rows / 256));
      g.rows_per_cta = round_up64(ceil_div64(g.rows, nct), rc);
      nct = ceil_div64(g.rows, g.rows_per_cta);
      if (rc == 32) {
        TF_CHECK_CUDA(cudaFuncSetAttribute(k_tc_redgemm2<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_tc_redgemm2<32><<<(unsigned)nct, kRed2Threads, smem, st>>>(g, T);
      } else {
        TF_CHECK_CUDA(cudaFuncSetAttribute(k_tc_redgemm2<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_tc_redgemm2<16><<<(unsigned)nct, kRed2Threads, smem, st>>>(g, T);
      }
      TF_CHECK_LAUNCH();
      return 0;
    }
  }
  int stages = (int)std::min<size_t>(8, (227 * 1024 - 1024) / stage);
  TF_CHECK_ARG(stages >= 2, "tc redgemm: tile too large for shared memory");
  g.groups = std::min(kRedGroups, stages);
  stages = stages / g.groups * g.groups;  // same-group-per-stage rule, see launch_rowgemm
  g.stages = stages;
  const size_t smem = 1024 + stages * stage;
  TF_CHECK_CUDA(cudaFuncSetAttribute(k_tc_redgemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_tc_redgemm<<<(unsigned)ctas, kRedThreads, smem, st>>>(g);
  TF_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// FeatureMlp forward / reverse on the tensor-core GEMMs
// ---------------------------------------------------------------------------------------------
// All weight packings of one MLP call go out in ONE launch (blockIdx.y = job).
struct PackJobs {
  PackArgs a[8];
  int nsplit[8];
  int n;
};
__global__ void __launch_bounds__(256) k_pack_all(PackJobs jobs) {
  const PackArgs a = jobs.a[blockIdx.y];
  const int nsplit = jobs.nsplit[blockIdx.y];
  constexpr int KC = RowCfg<2>::KC;
  int item = blockIdx.x * blockDim.x + threadIdx.x;
  const int nchunks = (a.K_pad + KC - 1) / KC;
  if (item >= a.N_pad * nchunks * (KC / 8)) return;
  const int n = item % a.N_pad;
  const int k8g = item / a.N_pad;
  const int c = k8g / (KC / 8), j = k8g % (KC / 8);
  float x[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    int k = c * KC + j * 8 + q;
    x[q] = (k < a.K_valid && n < a.N_valid) ? a.W[k * a.sk + n * a.sn] : 0.f;
  }
  const uint32_t tb = tile_bytes(a.N_pad, KC);
  if (nsplit == 3)
    split_store<3>(x, a.out + (size_t)c * 3 * tb, tb, tile_offset(n, j * 8, KC));
  else
    split_store<2>(x, a.out + (size_t)c * 2 * tb, tb, tile_offset(n, j * 8, KC));
}
static int launch_pack_jobs(cudaStream_t st, const PackJobs& jobs) {
  if (jobs.n == 0) return 0;
  constexpr int KC = RowCfg<2>::KC;
  int max_items = 0;
  for (int i = 0; i < jobs.n; ++i)
    max_items = std::max(max_items, jobs.a[i].N_pad * ((jobs.a[i].K_pad + KC - 1) / KC) * (KC / 8));
  k_pack_all<<<dim3((max_items + 255) / 256, jobs.n), 256, 0, st>>>(jobs);
  TF_CHECK_LAUNCH();
  return 0;
}

struct RowEpilogue {  // optional fused pieces of a row GEMM
  float* enc_x = nullptr;  // Dense_0: write the encoded input x = [f, v, enc(f), enc(v)] from the epilogue
  const float* enc_vd = nullptr;
  int enc_ldx = 0, enc_squash = 0, enc_Ff = 0, enc_Fv = 0;
  const uint32_t* bits_in = nullptr;
  uint32_t* bits_out = nullptr;
  const float *w3 = nullptr, *b3 = nullptr, *embed = nullptr;
  const uint32_t* cams = nullptr;
  int rows_per_ray = 1;
  float* rgb_out = nullptr;
};

// C[:, n0:n0+nn] = epilogue(A @ W[:, n0:n0+nn]) for column tiles of at most 256 (UMMA N limit).
// `collect` != null: only record the weight-packing jobs; else: only launch the GEMMs (weights
// already packed into `scratch`).
template <int NSPLIT>
static int rowgemm_tiled(cudaStream_t st, const float* A, int64_t lda, int64_t M, int K_valid, const float* W, int64_t sk,
                         int64_t sn, int N_valid, float* C, int64_t ldc, int N_store_total, const float* bias, int relu,
                         const RowEpilogue& ep, unsigned char* scratch, PackJobs* collect) {
  const int K_pad = round_up(K_valid, 16);
  const int N_total = round_up(N_store_total, 16);
  for (int n0 = 0; n0 < N_total; n0 += 256) {
    const int nn = std::min(256, N_total - n0);
    if (collect) {
      TF_CHECK_ARG(collect->n < 8, "too many weight tiles");
      collect->a[collect->n] = PackArgs{W + n0 * sn, sk, sn, K_valid, std::max(0, N_valid - n0), K_pad, nn, scratch};
      collect->nsplit[collect->n++] = NSPLIT;
    } else {
      RowGemmArgs g{};
      g.A = A; g.lda = lda; g.M = M; g.K_valid = K_valid; g.K_pad = K_pad; g.Bp = scratch; g.N_pad = nn;
      g.C = C ? C + n0 : nullptr; g.ldc = ldc; g.N_store = std::min(nn, N_store_total - n0);
      g.bias = bias ? bias + n0 : nullptr; g.relu = relu;
      TF_CHECK_ARG(!(ep.bits_in || ep.bits_out || ep.rgb_out) || N_total <= 256, "fused epilogues need a single column tile");
      TF_CHECK_ARG(!ep.rgb_out || nn == 128, "fused output layer needs N == 128");
      g.bits_in = ep.bits_in; g.bits_out = ep.bits_out;
      g.w3 = ep.w3; g.b3 = ep.b3; g.embed = ep.embed; g.cams = ep.cams; g.rows_per_ray = ep.rows_per_ray; g.rgb_out = ep.rgb_out;
      TF_CHECK_ARG(!ep.enc_x || (N_total <= 256 && ep.enc_squash + 3 <= nn && ep.rows_per_ray >= 1 && !ep.rgb_out && !ep.bits_in && !ep.bits_out),
                   "fused encode needs a single column tile with 3 spare columns");
      g.enc_x = ep.enc_x; g.enc_vd = ep.enc_vd; g.enc_ldx = ep.enc_ldx; g.enc_squash = ep.enc_squash; g.enc_Ff = ep.enc_Ff; g.enc_Fv = ep.enc_Fv;
      TF_RETURN_IF_ERROR(launch_rowgemm<NSPLIT>(st, g));
    }
    scratch += packed_weight_bytes<NSPLIT>(K_pad, nn);
  }
  return 0;
}

// sized for the larger (3-way) split so either precision fits
static size_t rowgemm_scratch(int K_valid, int N_total) {
  const int K_pad = round_up(K_valid, 16);
  N_total = round_up(N_total, 16);
  size_t b = 0;
  for (int n0 = 0; n0 < N_total; n0 += 256)
    b += std::max(packed_weight_bytes<2>(K_pad, std::min(256, N_total - n0)), packed_weight_bytes<3>(K_pad, std::min(256, N_total - n0)));
  return b;
}

size_t mlp_tc_wpack_bytes(const MlpShape& s) {
  const int ldf = round_up(s.squash, 16), ldx = round_up(s.enc, 16), U = s.units;
  return rowgemm_scratch(s.Ca, ldf) + rowgemm_scratch(s.enc, U) + rowgemm_scratch(U, U)      // forward
         + rowgemm_scratch(U, U) + rowgemm_scratch(U, ldx) + rowgemm_scratch(s.squash, s.Ca);  // reverse (dX)
}

// The forward feeds the composite reverse, which cancels (u*c_k - W*A/B^2): it needs fp32-exact
// operands (3-way split). The reverse GEMMs feed gradients directly: the 2-way split suffices.
constexpr int kFwdSplit = 3, kBwdSplit = 2;

// The six row GEMMs of one MLP call. pass 0 = record weight packing jobs, 1 = forward GEMMs,
// 2 = reverse GEMMs (interleaved with the other reverse kernels by the caller through `stage`).
struct TcPlan {
  const MlpShape& s;
  const MlpParams& p;
  const MlpWs& ws;
  const float* feat;
  int64_t M;
  unsigned char* sc[6];
  TcPlan(const MlpShape& s_, const MlpParams& p_, const MlpWs& ws_, const float* feat_, int64_t M_)
      : s(s_), p(p_), ws(ws_), feat(feat_), M(M_) {
    const int U = s.units;
    size_t off[6] = {rowgemm_scratch(s.Ca, ws.ldf), rowgemm_scratch(s.enc, U), rowgemm_scratch(U, U),
                     rowgemm_scratch(U, U), rowgemm_scratch(U, ws.ldx), rowgemm_scratch(s.squash, s.Ca)};
    unsigned char* q = ws.wpack;
    for (int i = 0; i < 6; ++i) {
      sc[i] = q;
      q += off[i];
    }
  }
  int gemm(cudaStream_t st, int i, PackJobs* collect, const RowEpilogue& ep, float* d_feat) const {
    const int U = s.units;
    switch (i) {
      case 0:  // Dense_0: f = feat @ W0 (no bias); padding columns of f come out as exact zeros
        return rowgemm_tiled<kFwdSplit>(st, feat, s.Ca, M, s.Ca, p.w0, s.squash, 1, s.squash, ws.f, ws.ldf, ws.ldf, nullptr, 0,
                                        ep, sc[0], collect);
      case 1:  // Dense_1 + relu (+ ReLU bitmask for the reverse pass)
        return rowgemm_tiled<kFwdSplit>(st, ws.x, ws.ldx, M, s.enc, p.w1, U, 1, U, ws.h1, U, U, p.b1, 1, ep, sc[1], collect);
      case 2:  // Dense_2 + relu (+ fused FiLM / Dense_3 / sigmoid)
        // (forward-only calls do not keep h2: the fused output layer consumes it in registers)
        return rowgemm_tiled<kFwdSplit>(st, ws.h1, U, M, U, p.w2, U, 1, U, (s.inference && ep.rgb_out) ? nullptr : ws.h2, U, U, p.b2, 1,
                                        ep, sc[2], collect);
      case 3:  // dp1 = (dp2 @ W2^T) * relu'(h1)
        return rowgemm_tiled<kBwdSplit>(st, ws.dp2, U, M, U, p.w2, 1, U, U, ws.dp1, U, U, nullptr, 0, ep, sc[3], collect);
      case 4:  // dx = dp1 @ W1^T
        return rowgemm_tiled<kBwdSplit>(st, ws.dp1, U, M, U, p.w1, 1, U, s.enc, ws.dx, ws.ldx, ws.ldx, nullptr, 0, ep, sc[4],
                                        collect);
      default:  // d_feat = df @ W0^T
        return rowgemm_tiled<kBwdSplit>(st, ws.df, ws.ldf, M, s.squash, p.w0, 1, s.squash, s.Ca, d_feat, s.Ca, s.Ca, nullptr, 0,
                                        ep, sc[5], collect);
    }
  }
};

int mlp_tc_fwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const float* feat, const float* viewdirs,
               const uint32_t* cams, int64_t M, int rows_per_ray, const MlpWs& ws, float* rgb) {
  if (M == 0) return 0;
  TcPlan plan(s, p, ws, feat, M);
  // pack the weights of all six GEMMs (forward and reverse) in one launch
  PackJobs jobs{};
  RowEpilogue none;
  for (int i = 0; i < 6; ++i) TF_RETURN_IF_ERROR(plan.gemm(st, i, &jobs, none, nullptr));
  TF_RETURN_IF_ERROR(launch_pack_jobs(st, jobs));
  // Dense_0 with the Fourier encode fused into its epilogue: x = [f, v, enc(f), enc(v)] is written straight from the
  // accumulator fragments (the three padding columns after f stand in for the view direction), so there is no encode
  // kernel and f is not re-read.  Needs squash + 3 <= ceil16(squash) columns; otherwise the standalone kernel runs.
  if (s.squash + 3 <= ws.ldf && getenv("TENSORF_TC_ENC_FUSION")) {  // off by default: measured slower than the standalone kernel so far
    RowEpilogue e0;
    e0.enc_x = ws.x; e0.enc_vd = viewdirs; e0.enc_ldx = ws.ldx; e0.enc_squash = s.squash; e0.enc_Ff = s.Ff; e0.enc_Fv = s.Fv;
    e0.rows_per_ray = rows_per_ray;
    TF_RETURN_IF_ERROR(plan.gemm(st, 0, nullptr, e0, nullptr));
  } else {
    TF_RETURN_IF_ERROR(plan.gemm(st, 0, nullptr, none, nullptr));
    TF_RETURN_IF_ERROR(mlp_encode_fwd(st, s, ws, viewdirs, M, rows_per_ray, true));
  }
  RowEpilogue e1;
  e1.bits_out = s.inference ? nullptr : ws.bits1;
  TF_RETURN_IF_ERROR(plan.gemm(st, 1, nullptr, e1, nullptr));
  RowEpilogue e2;
  e2.w3 = p.w3; e2.b3 = p.b3; e2.embed = s.ncam ? p.embed : nullptr; e2.cams = cams; e2.rows_per_ray = rows_per_ray;
  e2.rgb_out = rgb;
  return plan.gemm(st, 2, nullptr, e2, nullptr);
}

int mlp_tc_bwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const float* feat, const float* viewdirs,
               const uint32_t* cams, int64_t M, int rows_per_ray, const MlpWs& ws, const float* rgb, const float* d_rgb,
               float* d_feat, const MlpGrads& gr) {
  const int U = s.units;
  (void)viewdirs;
  TF_CHECK_ARG(!s.inference, "mlp reverse pass after a forward with TENSORF_FLAG_INFERENCE (residuals were not kept)");
  if (!gr.prezeroed) TF_RETURN_IF_ERROR(mlp_zero_grads(st, s, gr));
  if (M == 0) return 0;
  TcPlan plan(s, p, ws, feat, M);  // weights were packed by mlp_tc_fwd on the same workspace
  RowEpilogue none;
  TF_RETURN_IF_ERROR(mlp_out_bwd(st, s, p, ws, cams, M, rows_per_ray, rgb, d_rgb, gr));
  RedGemmArgs r{};
  // dW2[k][n] = sum_rows h1[row][k] * dp2[row][n] ; db2 through the ones column
  r = RedGemmArgs{};
  r.G = ws.dp2; r.ldg = U; r.Mg = U; r.X = ws.h1; r.ldx = U; r.Nx = U; r.N_pad = round_up(U + 1, 16); r.rows = M; r.out = gr.w2;
  r.ldo = U; r.colsum = gr.b2;
  TF_RETURN_IF_ERROR(launch_redgemm(st, r));
  RowEpilogue e3;
  e3.bits_in = ws.bits1;
  TF_RETURN_IF_ERROR(plan.gemm(st, 3, nullptr, e3, nullptr));
  // dW1[k][n] = sum_rows x[row][k] * dp1[row][n] ; db1
  r = RedGemmArgs{};
  r.G = ws.dp1; r.ldg = U; r.Mg = U; r.X = ws.x; r.ldx = ws.ldx; r.Nx = s.enc; r.N_pad = round_up(s.enc + 1, 16); r.rows = M;
  r.out = gr.w1; r.ldo = U; r.colsum = gr.b1;
  TF_RETURN_IF_ERROR(launch_redgemm(st, r));
  TF_RETURN_IF_ERROR(plan.gemm(st, 4, nullptr, none, nullptr));
  TF_RETURN_IF_ERROR(mlp_encode_bwd(st, s, ws, M, true));
  // dW0[k][n] = sum_rows feat[row][k] * df[row][n]
  r = RedGemmArgs{};
  r.G = ws.df; r.ldg = ws.ldf; r.Mg = s.squash; r.X = feat; r.ldx = s.Ca; r.Nx = s.Ca; r.N_pad = round_up(s.Ca, 16); r.rows = M;
  r.out = gr.w0; r.ldo = s.squash;
  TF_RETURN_IF_ERROR(launch_redgemm(st, r));
  return plan.gemm(st, 5, nullptr, none, d_feat);
}

// ---------------------------------------------------------------------------------------------
// microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16) for a given N and smem layout type
// (operand contents are irrelevant). out[0] = cycles for `count` MMAs on one SM.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_umma_bench(int N, int layout_type, uint32_t lbo, uint32_t sbo, int count, int distinct,
                                                      long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 64);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (distinct < 0 && threadIdx.x < 32) {  // warp-uniform issue loop, MMA predicated by elect.sync
    distinct = -distinct;
    const uint32_t base = smem_u32(smem + 1024);
    const uint32_t idesc = make_idesc_bf16(128, N);
    long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const uint32_t off = (uint32_t)(i % distinct) * 4096;
      uint64_t da = (uint64_t)(((base + off) >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
                    ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);
      uint64_t db = (uint64_t)(((base + 65536 + off) >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
                    ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);
      if (elect_one()) umma_bf16(tm, da, db, idesc, i != 0);
    }
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    if (threadIdx.x == 0) out[0] = clock64() - t0;
  } else if (distinct > 0 && threadIdx.x == 0) {
    const uint32_t base = smem_u32(smem + 1024);
    const uint32_t idesc = make_idesc_bf16(128, N);
    long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const uint32_t off = (uint32_t)(i % distinct) * 4096;  // different operand tiles, like a K loop
      uint64_t da = (uint64_t)(((base + off) >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
                    ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);
      uint64_t db = (uint64_t)(((base + 65536 + off) >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
                    ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);
      umma_bf16(tm, da, db, idesc, i != 0);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    out[0] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tm, 256);
  }
}

int tc_umma_bench(cudaStream_t st, int N, int layout_type, int lbo, int sbo, int count, long long* out_dev) {
  const int distinct = getenv("TENSORF_UMMA_UNIFORM") ? -8 : 8;
  const size_t smem = 1024 + 2 * 65536 + 16384;
  TF_CHECK_CUDA(cudaFuncSetAttribute(k_umma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_umma_bench<<<1, 128, smem, st>>>(N, layout_type, (uint32_t)lbo, (uint32_t)sbo, count, distinct, out_dev);
  TF_CHECK_LAUNCH();
  return 0;
}

int tc_trace_read(long long* host, int n) {
  TF_CHECK_CUDA(cudaDeviceSynchronize());
  TF_CHECK_CUDA(cudaMemcpyFromSymbol(host, g_trace, sizeof(long long) * (size_t)std::min(n, 8192)));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// test entry points (tests/test_gpu_tc.py): plain GEMMs through the tensor-core kernels
// ---------------------------------------------------------------------------------------------
int tc_rowgemm_test(cudaStream_t st, const float* A, int64_t M, int K, const float* W, int N, const float* bias, int relu,
                    const uint32_t* mask_bits, uint32_t* bits_out, float* C, void* scratch, size_t scratch_bytes, int nsplit) {
  TF_CHECK_ARG(rowgemm_scratch(K, N) <= scratch_bytes, "scratch too small");
  RowEpilogue ep;
  ep.bits_in = mask_bits;
  ep.bits_out = bits_out;
  PackJobs jobs{};
  if (nsplit == 3) {
    TF_RETURN_IF_ERROR(rowgemm_tiled<3>(st, A, K, M, K, W, N, 1, N, C, N, N, bias, relu, ep, (unsigned char*)scratch, &jobs));
    TF_RETURN_IF_ERROR(launch_pack_jobs(st, jobs));
    return rowgemm_tiled<3>(st, A, K, M, K, W, N, 1, N, C, N, N, bias, relu, ep, (unsigned char*)scratch, nullptr);
  }
  TF_RETURN_IF_ERROR(rowgemm_tiled<2>(st, A, K, M, K, W, N, 1, N, C, N, N, bias, relu, ep, (unsigned char*)scratch, &jobs));
  TF_RETURN_IF_ERROR(launch_pack_jobs(st, jobs));
  return rowgemm_tiled<2>(st, A, K, M, K, W, N, 1, N, C, N, N, bias, relu, ep, (unsigned char*)scratch, nullptr);
}

int tc_redgemm_test(cudaStream_t st, const float* G, int Mg, const float* X, int Nx, int64_t rows, float* out) {
  RedGemmArgs g{};
  g.G = G; g.ldg = Mg; g.Mg = Mg; g.X = X; g.ldx = Nx; g.Nx = Nx; g.N_pad = round_up(Nx, 16); g.rows = rows;
  g.out = out; g.ldo = Mg;
  return launch_redgemm(st, g);
}

}  // namespace tf
