// FeatureMlp (networks.py:38-121) on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// fp32 parity (1e-4) on bf16 tensor cores: every operand x is split into hi = bf16(x) and
// lo = bf16(x - hi); a product is accumulated as hi*hi + hi*lo + lo*hi in fp32 (TMEM), i.e.
// three tcgen05.mma per k-step ("bf16x3", ~2^-16 relative per product).
//
//   k_tc_rowgemm : C[M x N] = epilogue(A[M x K] * W[K x N])     forward layers and dX
//       persistent, warp-specialised: 4 producer warps (fp32 -> hi/lo bf16 into the canonical
//       no-swizzle K-major UMMA layout), 1 bulk-copy thread streaming pre-packed weights
//       (cp.async.bulk), 1 MMA-issuing thread, 4 epilogue warps (tcgen05.ld -> bias/relu/mask
//       -> global). smem stage ring + double-buffered TMEM accumulators.
//   k_tc_redgemm : dW^T[128 x N] += G^T[128 x rows] * X[rows x N]   weight gradients (split over
//       rows across CTAs, RED-add epilogue).
#include <algorithm>

#include "mlp.cuh"
#include "tc_prims.cuh"

namespace tf {

using namespace tc;

// ---------------------------------------------------------------------------------------------
// weight packing: W (fp32, any strides) -> per 64-wide K chunk [hi tile N_pad x 64][lo tile]
// ---------------------------------------------------------------------------------------------
constexpr int kKC = 64;   // K chunk of the row GEMM
constexpr int kRC = 32;   // row chunk of the reduction GEMM

struct PackArgs {
  const float* W;
  int64_t sk, sn;  // W(k, n) = W[k*sk + n*sn]
  int K_valid, N_valid, K_pad, N_pad;
  unsigned char* out;
};

__global__ void __launch_bounds__(256) k_pack_weights(PackArgs a) {
  const int k8s = (a.K_pad + 7) / 8;
  int item = blockIdx.x * blockDim.x + threadIdx.x;  // (n, k8)
  const int nchunks = (a.K_pad + kKC - 1) / kKC;
  if (item >= a.N_pad * nchunks * (kKC / 8)) return;
  (void)k8s;
  const int n = item % a.N_pad;
  const int k8g = item / a.N_pad;  // global k8 index over padded chunks
  const int c = k8g / (kKC / 8), j = k8g % (kKC / 8);
  float x[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    int k = c * kKC + j * 8 + q;
    x[q] = (k < a.K_valid && n < a.N_valid) ? a.W[k * a.sk + n * a.sn] : 0.f;
  }
  uint4 hi, lo;
  split8(x, hi, lo);
  const uint32_t tb = tile_bytes(a.N_pad, kKC);
  unsigned char* chunk = a.out + (size_t)c * 2 * tb;
  const uint32_t off = tile_offset(n, j * 8, kKC);
  *reinterpret_cast<uint4*>(chunk + off) = hi;
  *reinterpret_cast<uint4*>(chunk + tb + off) = lo;
}

static size_t packed_weight_bytes(int K_pad, int N_pad) {
  return (size_t)((K_pad + kKC - 1) / kKC) * 2 * tile_bytes(N_pad, kKC);
}

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
  const float* mask;          // keep where mask(m,n) > 0, or null
  int64_t ldmask;
  int relu;
  int stages;
  uint32_t tmem_cols;         // power of two >= 2*N_pad
};

constexpr int kRowThreads = 320;  // warps 0-3 epilogue, 4 MMA, 5 weight loader, 6-9 producers

__global__ void __launch_bounds__(kRowThreads, 1) k_tc_rowgemm(RowGemmArgs g) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = g.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [S]
  uint64_t* empty = full + S;                           // [S]
  uint64_t* tfull = empty + S;                          // [2]
  uint64_t* tempty = tfull + 2;                         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const uint32_t a_tile = tile_bytes(128, kKC), b_tile = tile_bytes(g.N_pad, kKC);
  const uint32_t stage_bytes = 2 * a_tile + 2 * b_tile;
  unsigned char* stage0 = smem + 1024;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 128 + 1);  // 128 producer threads + the weight loader's expect_tx arrive
      mbar_init(&empty[s], 1);       // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);       // tcgen05.commit
      mbar_init(&tempty[a], 128);    // epilogue threads
    }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, g.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t ntiles = (g.M + 127) / 128;
  const int nchunks = (g.K_pad + kKC - 1) / kKC;

  if (warp >= 6) {
    // ================= A producers: fp32 rows -> hi/lo bf16 core matrices =================
    const int pt = tid - 192, pw = pt >> 5;
    const bool vec_ok = (g.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
    uint32_t it_s = 0, ph = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int64_t m0 = t * 128;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&empty[it_s], ph ^ 1);
        unsigned char* sA = stage0 + (size_t)it_s * stage_bytes;
        const int k0 = c * kKC;
#pragma unroll 2
        for (int it = 0; it < 8; ++it) {
          const int rb = pw + 4 * (it >> 1);
          const int k8 = (lane >> 3) + 4 * (it & 1);
          const int row = rb * 8 + (lane & 7);
          const int k = k0 + k8 * 8;
          if (k >= g.K_pad) continue;
          float x[8];
          const int64_t m = m0 + row;
          if (m < g.M) {
            const float* src = g.A + m * g.lda + k;
            if (vec_ok && k + 8 <= g.K_valid) {
              float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
              x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
              x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) x[q] = (k + q < g.K_valid) ? src[q] : 0.f;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = 0.f;
          }
          uint4 hi, lo;
          split8(x, hi, lo);
          const uint32_t off = tile_offset(row, k8 * 8, kKC);
          *reinterpret_cast<uint4*>(sA + off) = hi;
          *reinterpret_cast<uint4*>(sA + a_tile + off) = lo;
        }
        fence_proxy_async();
        mbar_arrive(&full[it_s]);
        if (++it_s == (uint32_t)S) { it_s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 5) {
    // ================= weight loader: one bulk copy per chunk =================
    if (lane == 0) {
      uint32_t it_s = 0, ph = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&empty[it_s], ph ^ 1);
          unsigned char* sB = stage0 + (size_t)it_s * stage_bytes + 2 * a_tile;
          mbar_arrive_expect_tx(&full[it_s], 2 * b_tile);
          bulk_copy_g2s(sB, g.Bp + (size_t)c * 2 * b_tile, 2 * b_tile, &full[it_s]);
          if (++it_s == (uint32_t)S) { it_s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, g.N_pad);
      const uint32_t sbo = kKC * 16;  // next 8 rows
      uint32_t it_s = 0, ph = 0, acc = 0, acc_ph = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(&tempty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * (uint32_t)g.N_pad;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&full[it_s], ph);
          tc_fence_after();
          const uint32_t sA = smem_u32(stage0 + (size_t)it_s * stage_bytes);
          const uint32_t sB = sA + 2 * a_tile;
          const int nks = min(kKC / 16, (g.K_pad - c * kKC) / 16);
          for (int ks = 0; ks < nks; ++ks) {
            const uint64_t a_hi = make_smem_desc(sA + ks * 256, 128, sbo);
            const uint64_t a_lo = make_smem_desc(sA + a_tile + ks * 256, 128, sbo);
            const uint64_t b_hi = make_smem_desc(sB + ks * 256, 128, sbo);
            const uint64_t b_lo = make_smem_desc(sB + b_tile + ks * 256, 128, sbo);
            umma_bf16(d, a_lo, b_hi, idesc, (c | ks) != 0);
            umma_bf16(d, a_hi, b_lo, idesc, 1);
            umma_bf16(d, a_hi, b_hi, idesc, 1);
          }
          umma_commit(&empty[it_s]);  // smem stage free once these MMAs have read it
          if (++it_s == (uint32_t)S) { it_s = 0; ph ^= 1; }
        }
        umma_commit(&tfull[acc]);  // accumulator complete
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else {
    // ================= epilogue: TMEM -> registers -> bias / relu / mask -> global =================
    uint32_t acc = 0, acc_ph = 0;
    const bool vec_st = (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      mbar_wait(&tfull[acc], acc_ph);
      tc_fence_after();
      const int64_t m = t * 128 + warp * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * (uint32_t)g.N_pad;
      for (int n0 = 0; n0 < g.N_pad; n0 += 16) {
        float v[16];
        tmem_ld16(trow + n0, v);
        tmem_ld_wait();
        if (m < g.M && n0 < g.N_store) {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int n = n0 + q;
            float y = v[q];
            if (n < g.N_store) {
              if (g.bias) y += g.bias[n];
              if (g.relu) y = fmaxf(y, 0.f);
              if (g.mask && !(g.mask[m * g.ldmask + n] > 0.f)) y = 0.f;
            }
            v[q] = y;
          }
          float* dst = g.C + m * g.ldc + n0;
          if (vec_st && n0 + 16 <= g.N_store) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              reinterpret_cast<float4*>(dst)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          } else {
            for (int q = 0; q < 16; ++q)
              if (n0 + q < g.N_store) dst[q] = v[q];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, g.tmem_cols);
  }
}

static uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

static int launch_rowgemm(cudaStream_t st, RowGemmArgs g) {
  if (g.M == 0) return 0;
  TF_CHECK_ARG(g.N_pad % 16 == 0 && g.N_pad >= 16 && g.N_pad <= 256, "tc rowgemm: N_pad=%d unsupported", g.N_pad);
  TF_CHECK_ARG(g.K_pad % 16 == 0 && g.K_pad >= 16, "tc rowgemm: K_pad=%d unsupported", g.K_pad);
  const size_t stage = 2 * (size_t)tile_bytes(128, kKC) + 2 * (size_t)tile_bytes(g.N_pad, kKC);
  int stages = (int)std::min<size_t>(6, (227 * 1024 - 1024) / stage);
  TF_CHECK_ARG(stages >= 2, "tc rowgemm: tile too large for shared memory");
  g.stages = stages;
  g.tmem_cols = pow2_cols(2 * g.N_pad);
  const size_t smem = 1024 + stages * stage;
  TF_CHECK_CUDA(cudaFuncSetAttribute(k_tc_rowgemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = (g.M + 127) / 128;
  const unsigned grid = (unsigned)std::min<int64_t>(ntiles, kSMs);
  k_tc_rowgemm<<<grid, kRowThreads, smem, st>>>(g);
  TF_CHECK_LAUNCH();
  return 0;
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
  int stages;
  uint32_t tmem_cols;
};

constexpr int kRedThreads = 128 + 32 + 256;  // warps 0-3 epilogue, 4 MMA, 5-12 producers

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
      mbar_init(&full[s], 256);
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
    // ===== producers: transposing loads, 8 consecutive rows -> one 16-byte k-chunk =====
    const int pt = tid - 160;
    const int a_items = 128 * (kRC / 8), b_items = g.N_pad * (kRC / 8);
    uint32_t it_s = 0, ph = 0;
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait(&empty[it_s], ph ^ 1);
      unsigned char* sA = stage0 + (size_t)it_s * stage_bytes;
      unsigned char* sB = sA + 2 * a_tile;
      const int64_t r0 = r_begin + (int64_t)c * kRC;
      for (int item = pt; item < a_items + b_items; item += 256) {
        const bool isA = item < a_items;
        const int e = isA ? item : item - a_items;
        const int ncols = isA ? 128 : g.N_pad;
        const int col = e % ncols, j = e / ncols;  // j = row group of 8
        const float* src = isA ? g.G : g.X;
        const int64_t ld = isA ? g.ldg : g.ldx;
        const int valid = isA ? g.Mg : g.Nx;
        float x[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int64_t r = r0 + j * 8 + q;
          x[q] = (col < valid && r < r_end) ? src[r * ld + col] : 0.f;
        }
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t off = tile_offset(col, j * 8, kRC);
        unsigned char* base = isA ? sA : sB;
        const uint32_t tb = isA ? a_tile : b_tile;
        *reinterpret_cast<uint4*>(base + off) = hi;
        *reinterpret_cast<uint4*>(base + tb + off) = lo;
      }
      fence_proxy_async();
      mbar_arrive(&full[it_s]);
      if (++it_s == (uint32_t)S) { it_s = 0; ph ^= 1; }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      const uint32_t sbo = kRC * 16;
      uint32_t it_s = 0, ph = 0;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&full[it_s], ph);
        tc_fence_after();
        const uint32_t sA = smem_u32(stage0 + (size_t)it_s * stage_bytes);
        const uint32_t sB = sA + 2 * a_tile;
        for (int ks = 0; ks < kRC / 16; ++ks) {
          const uint64_t a_hi = make_smem_desc(sA + ks * 256, 128, sbo);
          const uint64_t a_lo = make_smem_desc(sA + a_tile + ks * 256, 128, sbo);
          for (int n0 = 0; n0 < g.N_pad; n0 += 256) {
            const int nn = min(256, g.N_pad - n0);
            const uint32_t idesc = make_idesc_bf16(128, nn);
            // rows n0.. of the B tile start (n0/8) core-matrix rows further
            const uint32_t boff = (uint32_t)(n0 >> 3) * sbo + ks * 256;
            const uint64_t b_hi = make_smem_desc(sB + boff, 128, sbo);
            const uint64_t b_lo = make_smem_desc(sB + b_tile + boff, 128, sbo);
            const uint32_t d = tmem_base + n0;
            umma_bf16(d, a_lo, b_hi, idesc, (c | ks) != 0);
            umma_bf16(d, a_hi, b_lo, idesc, 1);
            umma_bf16(d, a_hi, b_hi, idesc, 1);
          }
        }
        umma_commit(&empty[it_s]);
        if (++it_s == (uint32_t)S) { it_s = 0; ph ^= 1; }
      }
      umma_commit(tfull);
    }
  } else {
    // ===== epilogue: RED-add the 128 x N tile into out[n*ldo + m] =====
    if (nchunks > 0) {
      mbar_wait(tfull, 0);
      tc_fence_after();
      const int m = warp * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int n0 = 0; n0 < g.N_pad; n0 += 16) {
        float v[16];
        tmem_ld16(trow + n0, v);
        tmem_ld_wait();
        if (m < g.Mg) {
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (n0 + q < g.Nx) atomicAdd(g.out + (int64_t)(n0 + q) * g.ldo + m, v[q]);
        }
      }
    }
  }
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
  const size_t stage = 2 * (size_t)tile_bytes(128, kRC) + 2 * (size_t)tile_bytes(g.N_pad, kRC);
  int stages = (int)std::min<size_t>(6, (227 * 1024 - 1024) / stage);
  TF_CHECK_ARG(stages >= 2, "tc redgemm: tile too large for shared memory");
  g.stages = stages;
  g.tmem_cols = pow2_cols(g.N_pad);
  int64_t ctas = std::min<int64_t>(kSMs, std::max<int64_t>(1, g.rows / 256));
  g.rows_per_cta = round_up64(ceil_div64(g.rows, ctas), kRC);
  ctas = ceil_div64(g.rows, g.rows_per_cta);
  const size_t smem = 1024 + stages * stage;
  TF_CHECK_CUDA(cudaFuncSetAttribute(k_tc_redgemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_tc_redgemm<<<(unsigned)ctas, kRedThreads, smem, st>>>(g);
  TF_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// test entry points (tests/test_gpu_tc.py): plain GEMMs through the tensor-core kernels
// ---------------------------------------------------------------------------------------------
int tc_rowgemm_test(cudaStream_t st, const float* A, int64_t M, int K, const float* W, int N, const float* bias, int relu,
                    const float* mask, float* C, void* scratch, size_t scratch_bytes) {
  const int K_pad = round_up(K, 16), N_pad = round_up(N, 16);
  TF_CHECK_ARG(packed_weight_bytes(K_pad, N_pad) <= scratch_bytes, "scratch too small");
  PackArgs p{W, N, 1, K, N, K_pad, N_pad, (unsigned char*)scratch};
  const int items = N_pad * ((K_pad + kKC - 1) / kKC) * (kKC / 8);
  k_pack_weights<<<(items + 255) / 256, 256, 0, st>>>(p);
  TF_CHECK_LAUNCH();
  RowGemmArgs g{};
  g.A = A; g.lda = K; g.M = M; g.K_valid = K; g.K_pad = K_pad; g.Bp = (const unsigned char*)scratch; g.N_pad = N_pad;
  g.C = C; g.ldc = N; g.N_store = N; g.bias = bias; g.mask = mask; g.ldmask = N; g.relu = relu;
  return launch_rowgemm(st, g);
}

int tc_redgemm_test(cudaStream_t st, const float* G, int Mg, const float* X, int Nx, int64_t rows, float* out) {
  RedGemmArgs g{};
  g.G = G; g.ldg = Mg; g.Mg = Mg; g.X = X; g.ldx = Nx; g.Nx = Nx; g.N_pad = round_up(Nx, 16); g.rows = rows;
  g.out = out; g.ldo = Mg;
  return launch_redgemm(st, g);
}

}  // namespace tf
