// FeatureMlp on CUDA cores in true fp32 (networks.py:38-121).  This is the exact-arithmetic
// path (what the JAX CPU oracle computes) and the on-device reference for the tcgen05 path.
#include <algorithm>

#include <stdlib.h>

#include "mlp.cuh"

namespace tf {

int64_t mlp_ws_floats(const MlpShape& s, int64_t M) {
  int64_t Mp = round_up64(M, 128);
  const int64_t wg = mlp_fused_supported(s) ? std::min<int64_t>(Mp / 128, kSMs) * (320 + s.Ca) * 128 : 0;
  return Mp * ((int64_t)2 * round_up(s.squash, 16) + 2 * round_up(s.enc, 16) + 4 * s.units + 8 + 16) + 64 + wg +
         (int64_t)round_up64((int64_t)mlp_tc_wpack_bytes(s), 256) / 4;
}
MlpWs mlp_ws_carve(const MlpShape& s, int64_t M, float* base) {
  int64_t Mp = round_up64(M, 128);
  MlpWs w;
  w.ldf = round_up(s.squash, 16);
  w.ldx = round_up(s.enc, 16);
  float* p = base;
  w.wpack = reinterpret_cast<unsigned char*>(p);
  w.wpack_bytes = mlp_tc_wpack_bytes(s);
  p += round_up64((int64_t)w.wpack_bytes, 256) / 4;
  w.f = p;   p += Mp * w.ldf;
  w.df = p;  p += Mp * w.ldf;
  w.x = p;   p += Mp * w.ldx;
  w.dx = p;  p += Mp * w.ldx;
  w.h1 = p;  p += Mp * s.units;
  w.h2 = p;  p += Mp * s.units;
  w.dp2 = p; p += Mp * s.units;
  w.dp1 = p; p += Mp * s.units;
  w.bits1 = reinterpret_cast<uint32_t*>(p); p += Mp * 8;
  w.aux = p; p += Mp * 16 + 64;
  w.wg_partial = p;
  return w;
}

// ---------------------------------------------------------------------------------------------
// Generic strided SGEMM: C(m,n) = epilogue( sum_k A(m,k) * B(k,n) )
// ---------------------------------------------------------------------------------------------
struct GemmArgs {
  const float* A; int64_t sam, sak;
  const float* B; int64_t sbk, sbn;
  float* C; int64_t ldc;
  const float* bias;                  // per column or null
  const float* mask; int64_t ldmask;  // keep C(m,n) only where mask(m,n) > 0, or null
  int64_t M, N, K;
  int64_t k_chunk;  // split-K chunk (gridDim.z chunks); atomic accumulate when gridDim.z > 1
  int relu;
};

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) k_sgemm(GemmArgs g) {
  constexpr int BK = 16;
  constexpr int NTX = BN / TN;  // threads along n
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int tx = t % NTX, ty = t / NTX;
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.k_chunk;
  const int64_t kend = min(g.K, kbeg + g.k_chunk);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  constexpr int A_PER = BM * BK / 256, B_PER = (BN * BK + 255) / 256;
  const bool a_kc = (g.sak == 1), b_nc = (g.sbn == 1);

  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    float ra[A_PER], rb[B_PER];
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int e = t + 256 * i, mm, kk;
      if (a_kc) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      int64_t m = m0 + mm, k = k0 + kk;
      ra[i] = (m < g.M && k < kend) ? g.A[m * g.sam + k * g.sak] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      int e = t + 256 * i, nn, kk;
      if (b_nc) { nn = e % BN; kk = e / BN; } else { kk = e % BK; nn = e / BK; }
      int64_t n = n0 + nn, k = k0 + kk;
      rb[i] = (e < BN * BK && n < g.N && k < kend) ? g.B[k * g.sbk + n * g.sbn] : 0.f;
    }
    __syncthreads();  // previous tile fully consumed
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int e = t + 256 * i, mm, kk;
      if (a_kc) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      As[kk][mm] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      int e = t + 256 * i, nn, kk;
      if (b_nc) { nn = e % BN; kk = e / BN; } else { kk = e % BK; nn = e / BK; }
      if (e < BN * BK) Bs[kk][nn] = rb[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int h = 0; h < TM / 4; ++h) {
        float4 v = *reinterpret_cast<const float4*>(&As[kk][h * (BM / 2) + ty * 4]);
        a[4 * h + 0] = v.x; a[4 * h + 1] = v.y; a[4 * h + 2] = v.z; a[4 * h + 3] = v.w;
      }
#pragma unroll
      for (int h = 0; h < TN / 4; ++h) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[kk][h * (BN / 2) + tx * 4]);
        b[4 * h + 0] = v.x; b[4 * h + 1] = v.y; b[4 * h + 2] = v.z; b[4 * h + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }

  const bool atomic = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t m = m0 + (i / 4) * (BM / 2) + ty * 4 + (i % 4);
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int64_t n = n0 + (j / 4) * (BN / 2) + tx * 4 + (j % 4);
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += g.bias[n];
      if (g.relu) v = fmaxf(v, 0.f);
      if (g.mask && !(g.mask[m * g.ldmask + n] > 0.f)) v = 0.f;
      if (atomic)
        atomicAdd(&g.C[m * g.ldc + n], v);
      else
        g.C[m * g.ldc + n] = v;
    }
  }
}

// TM == 4 / TN == 4 use a single half: rows ty*4.. (BM/2 offset unused). Guard the layout.
template <int BM, int BN, int TM, int TN>
static int launch_sgemm(cudaStream_t st, GemmArgs g, int splits) {
  static_assert((TM == 8 || BM / TM * 4 == BM) && (TN == 8 || BN / TN * 4 == BN), "tile layout");
  if (g.M == 0 || g.N == 0) return 0;
  splits = max(1, splits);
  g.k_chunk = round_up64(ceil_div64(g.K, splits), 16);
  int z = (int)ceil_div64(g.K, g.k_chunk);
  dim3 grid((unsigned)ceil_div64(g.M, BM), (unsigned)ceil_div64(g.N, BN), (unsigned)z);
  k_sgemm<BM, BN, TM, TN><<<grid, 256, 0, st>>>(g);
  TF_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Fourier encoding (networks.py:13-35, :68-76)
// ---------------------------------------------------------------------------------------------
// FAST (tensor-core path): even frequency levels by sincosf (one shared range reduction; cos(a) stands in for the
// reference's sin(a + pi/2), equal up to the rounding of a + pi/2), odd levels by the double-angle identities
// (~3 ulp) - half the transcendental work.  !FAST (exact fp32 path): the reference's expressions verbatim.
template <bool FAST>
__global__ void __launch_bounds__(256) k_encode_fwd(const float* __restrict__ f, const float* __restrict__ viewdirs,
                                                    float* __restrict__ x, int64_t M, int rows_per_ray, MlpShape s, int ldf,
                                                    int ldx) {
  const int D = s.squash + 3;
  // 32 lanes per row (D <= 32 source dimensions; a division-free mapping: the 64-bit div/mod by D was ~40 of the
  // kernel's ~90 instructions per item), or the flat mapping for wider inputs
  int64_t m;
  int d;
  if (D <= 32) {
    m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    d = threadIdx.x & 31;
    if (m >= M || d >= D) return;
  } else {
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= M * D) return;
    m = item / D;
    d = (int)(item % D);
  }
  float* row = x + m * ldx;
  float val;
  int F, off;
  if (d == 0)
    for (int q = s.enc; q < ldx; ++q) row[q] = 0.f;  // padding columns stay finite
  if (d < s.squash) {
    val = f[m * ldf + d];
    F = s.Ff;
    off = D + d * 2 * s.Ff;
  } else {
    val = viewdirs[(m / rows_per_ray) * 3 + (d - s.squash)];
    F = s.Fv;
    off = D + s.squash * 2 * s.Ff + (d - s.squash) * 2 * s.Fv;
  }
  row[d] = val;
  const float half_pi = 1.57079632679489661923f;
  float scale = 1.0f;
  if (FAST) {
    for (int j = 0; j < F; j += 2) {
      float sn, cs;
      sincosf(val * scale, &sn, &cs);
      // levels j and j+1 are adjacent in the row: 8-byte stores halve the number of store requests (the kernel is
      // bound by the rate of scattered 4-byte stores, not by the arithmetic)
      if (j + 1 < F && ((reinterpret_cast<uintptr_t>(row + off + j) | reinterpret_cast<uintptr_t>(row + off + F + j)) & 7) == 0) {
        *reinterpret_cast<float2*>(row + off + j) = make_float2(sn, 2.0f * sn * cs);
        *reinterpret_cast<float2*>(row + off + F + j) = make_float2(cs, fmaf(-2.0f * sn, sn, 1.0f));
      } else {
        row[off + j] = sn;
        row[off + F + j] = cs;
        if (j + 1 < F) {
          row[off + j + 1] = 2.0f * sn * cs;
          row[off + F + j + 1] = fmaf(-2.0f * sn, sn, 1.0f);
        }
      }
      scale *= 4.0f;
    }
    return;
  }
  for (int j = 0; j < F; ++j) {
    float in = val * scale;  // exact: power-of-two scaling
    row[off + j] = sinf(in);
    row[off + F + j] = sinf(__fadd_rn(in, half_pi));
    scale *= 2.0f;
  }
}

template <bool FAST>
__global__ void __launch_bounds__(256) k_encode_bwd(const float* __restrict__ f, const float* __restrict__ dx,
                                                    float* __restrict__ df, int64_t M, MlpShape s, int ldf, int ldx) {
  int64_t m;
  int d;
  if (s.squash <= 32) {  // 32 lanes per row, no division (see k_encode_fwd)
    m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    d = threadIdx.x & 31;
    if (m >= M || d >= s.squash) return;
  } else {
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= M * s.squash) return;
    m = item / s.squash;
    d = (int)(item % s.squash);
  }
  const float* row = dx + m * ldx;
  float val = f[m * ldf + d];
  float g = row[d];
  int off = s.squash + 3 + d * 2 * s.Ff;
  const float half_pi = 1.57079632679489661923f;
  float scale = 1.0f;
  if (FAST) {  // d/din sin(in) = cos(in), d/din sin(in + pi/2) = -sin(in); odd levels by double angle
    for (int j = 0; j < s.Ff; j += 2) {
      float sn, cs;
      sincosf(val * scale, &sn, &cs);
      if (j + 1 < s.Ff && ((reinterpret_cast<uintptr_t>(row + off + j) | reinterpret_cast<uintptr_t>(row + off + s.Ff + j)) & 7) == 0) {
        const float2 ds = __ldg(reinterpret_cast<const float2*>(row + off + j));
        const float2 dc = __ldg(reinterpret_cast<const float2*>(row + off + s.Ff + j));
        g += scale * (ds.x * cs - dc.x * sn);
        g += 2.0f * scale * (ds.y * fmaf(-2.0f * sn, sn, 1.0f) - dc.y * (2.0f * sn * cs));
      } else {
        g += scale * (row[off + j] * cs - row[off + s.Ff + j] * sn);
        if (j + 1 < s.Ff)
          g += 2.0f * scale * (row[off + j + 1] * fmaf(-2.0f * sn, sn, 1.0f) - row[off + s.Ff + j + 1] * (2.0f * sn * cs));
      }
      scale *= 4.0f;
    }
    df[m * ldf + d] = g;
    return;
  }
  for (int j = 0; j < s.Ff; ++j) {
    float in = val * scale;
    g += scale * (row[off + j] * cosf(in) + row[off + s.Ff + j] * cosf(__fadd_rn(in, half_pi)));
    scale *= 2.0f;
  }
  df[m * ldf + d] = g;
}

// Row-staged forward variant for wide rows (tensor-core path, squash + 3 <= 32): one warp per row; lane d evaluates
// source dimension d into a shared-memory copy of the row, then the warp stores the row with coalesced 16-byte accesses.
// The element-wise kernel above issues one 4- or 8-byte request per value at a 2F-float stride and is bound by the
// request rate (dozer, 400-float rows: 285 us for 325 MB with 4-byte stores, 150 us with 8-byte stores, 75 us staged).
// Measured: wins for ldx = 400 (dozer), loses 4 us for ldx = 160 (lego); the reverse kernel gains nothing from staging.
__global__ void __launch_bounds__(256) k_encode_fwd_rows(const float* __restrict__ f, const float* __restrict__ viewdirs,
                                                         float* __restrict__ x, int64_t M, int rows_per_ray, MlpShape s,
                                                         int ldf, int ldx) {
  extern __shared__ __align__(16) float srow[];  // [8][ldx]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t m = (int64_t)blockIdx.x * 8 + wib;
  if (m >= M) return;
  float* row = srow + wib * ldx;
  const int D = s.squash + 3;
  for (int q = s.enc + lane; q < ldx; q += 32) row[q] = 0.f;  // padding columns stay finite
  if (lane < D) {
    float val;
    int F, off;
    if (lane < s.squash) {
      val = f[m * ldf + lane];
      F = s.Ff;
      off = D + lane * 2 * s.Ff;
    } else {
      val = viewdirs[(m / rows_per_ray) * 3 + (lane - s.squash)];
      F = s.Fv;
      off = D + s.squash * 2 * s.Ff + (lane - s.squash) * 2 * s.Fv;
    }
    row[lane] = val;
    float scale = 1.0f;
    for (int j = 0; j < F; j += 2) {  // even levels by sincosf, odd levels by the double-angle identities
      float sn, cs;
      sincosf(val * scale, &sn, &cs);
      row[off + j] = sn;
      row[off + F + j] = cs;
      if (j + 1 < F) {
        row[off + j + 1] = 2.0f * sn * cs;
        row[off + F + j + 1] = fmaf(-2.0f * sn, sn, 1.0f);
      }
      scale *= 4.0f;
    }
  }
  __syncwarp();
  float4* dst = reinterpret_cast<float4*>(x + m * ldx);
  const float4* src = reinterpret_cast<const float4*>(row);
  for (int q = lane; q < (ldx >> 2); q += 32) dst[q] = src[q];
}

// ---------------------------------------------------------------------------------------------
// Output layer: FiLM (networks.py:103-111) + Dense 3 + sigmoid (:114-120). One warp per row,
// lane owns hidden units 4*lane..4*lane+3 (units == 128).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_out_fwd(const float* __restrict__ h2, const float* __restrict__ w3,
                                                 const float* __restrict__ b3, const float* __restrict__ embed,
                                                 const uint32_t* __restrict__ cams, float* __restrict__ rgb, int64_t M,
                                                 int rows_per_ray) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float w[4][3];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < 3; ++c) w[j][c] = w3[(4 * lane + j) * 3 + c];
  for (int64_t m = warp; m < M; m += nwarps) {
    float4 h = *reinterpret_cast<const float4*>(h2 + m * 128 + 4 * lane);
    float hv[4] = {h.x, h.y, h.z, h.w};
    if (embed != nullptr && lane >= 16) {
      const float* em = embed + (int64_t)cams[m / rows_per_ray] * 128;
      float4 sc = *reinterpret_cast<const float4*>(em + 4 * lane - 64);
      float4 sh = *reinterpret_cast<const float4*>(em + 4 * lane);
      hv[0] = __fadd_rn(__fmul_rn(sc.x, hv[0]), sh.x);
      hv[1] = __fadd_rn(__fmul_rn(sc.y, hv[1]), sh.y);
      hv[2] = __fadd_rn(__fmul_rn(sc.z, hv[2]), sh.z);
      hv[3] = __fadd_rn(__fmul_rn(sc.w, hv[3]), sh.w);
    }
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float a = hv[0] * w[0][c];
      a = fmaf(hv[1], w[1][c], a);
      a = fmaf(hv[2], w[2][c], a);
      a = fmaf(hv[3], w[3][c], a);
      o[c] = warp_sum(a);
    }
    if (lane < 3) {
      float y = (lane == 0 ? o[0] : (lane == 1 ? o[1] : o[2])) + b3[lane];
      rgb[3 * m + lane] = 1.0f / (1.0f + expf(-y));
    }
  }
}

// Reverse of the output layer. Each warp walks a contiguous chunk of rows (four rows in flight),
// keeps the dW3 / db3 / dEmbed partial sums in registers; dW3/db3 are reduced across the block in
// shared memory so only one set of atomics per block reaches L2 (the 387 addresses are a hot
// spot otherwise). Embedding rows are flushed whenever the camera index changes: consecutive
// rows of one ray share the camera.
__global__ void __launch_bounds__(256) k_out_bwd(const float* __restrict__ h2, const float* __restrict__ w3,
                                                 const float* __restrict__ embed, const uint32_t* __restrict__ cams,
                                                 const float* __restrict__ rgb, const float* __restrict__ d_rgb,
                                                 float* __restrict__ dp2, float* __restrict__ dw3, float* __restrict__ db3,
                                                 float* __restrict__ dembed, int64_t M, int rows_per_ray, int64_t chunk) {
  __shared__ float red[8][388];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int64_t warp = (int64_t)blockIdx.x * 8 + wib;
  int64_t mbeg = warp * chunk, mend = min(M, mbeg + chunk);
  float w[4][3], gw[4][3];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      w[j][c] = w3[(4 * lane + j) * 3 + c];
      gw[j][c] = 0.f;
    }
  float gb[3] = {0.f, 0.f, 0.f};
  float gsc[4] = {0.f, 0.f, 0.f, 0.f}, gsh[4] = {0.f, 0.f, 0.f, 0.f};
  const bool film = embed != nullptr;
  int64_t cur_cam = -1;
  auto flush_embed = [&]() {
    if (film && cur_cam >= 0 && lane >= 16) {
      float* de = dembed + cur_cam * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(de + 4 * lane - 64 + j, gsc[j]);
        atomicAdd(de + 4 * lane + j, gsh[j]);
        gsc[j] = 0.f;
        gsh[j] = 0.f;
      }
    }
  };
  auto process = [&](int64_t m, const float4& h, const float* y3, const float* dy3) {
    float hv[4] = {h.x, h.y, h.z, h.w};   // pre-FiLM (post-relu)
    float hf[4] = {h.x, h.y, h.z, h.w};   // post-FiLM
    float sc[4] = {1.f, 1.f, 1.f, 1.f};
    if (film) {
      int64_t cam = cams[m / rows_per_ray];
      if (cam != cur_cam) {
        flush_embed();
        cur_cam = cam;
      }
      if (lane >= 16) {
        const float* em = embed + cam * 128;
        float4 s4 = *reinterpret_cast<const float4*>(em + 4 * lane - 64);
        float4 t4 = *reinterpret_cast<const float4*>(em + 4 * lane);
        sc[0] = s4.x; sc[1] = s4.y; sc[2] = s4.z; sc[3] = s4.w;
        hf[0] = __fadd_rn(__fmul_rn(s4.x, hv[0]), t4.x);
        hf[1] = __fadd_rn(__fmul_rn(s4.y, hv[1]), t4.y);
        hf[2] = __fadd_rn(__fmul_rn(s4.z, hv[2]), t4.z);
        hf[3] = __fadd_rn(__fmul_rn(s4.w, hv[3]), t4.w);
      }
    }
    float dy[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      dy[c] = dy3[c] * y3[c] * (1.0f - y3[c]);
      gb[c] += dy[c];
    }
    float dh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dh[j] = dy[0] * w[j][0] + dy[1] * w[j][1] + dy[2] * w[j][2];
#pragma unroll
      for (int c = 0; c < 3; ++c) gw[j][c] = fmaf(hf[j], dy[c], gw[j][c]);
    }
    if (film && lane >= 16) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        gsc[j] = fmaf(dh[j], hv[j], gsc[j]);
        gsh[j] += dh[j];
        dh[j] *= sc[j];
      }
    }
    float4 o;
    o.x = hv[0] > 0.f ? dh[0] : 0.f;
    o.y = hv[1] > 0.f ? dh[1] : 0.f;
    o.z = hv[2] > 0.f ? dh[2] : 0.f;
    o.w = hv[3] > 0.f ? dh[3] : 0.f;
    *reinterpret_cast<float4*>(dp2 + m * 128 + 4 * lane) = o;
  };
  // kOutBwdRows independent rows in flight per warp (all their loads are issued before the first row is processed;
  // rows are still processed in order, so the partial sums - and the results - do not depend on the depth).  The
  // kernel is one wave of ~14 warps per SM: with 2 rows (1 KB per warp) it had ~14 KB per SM in flight against the
  // ~35 KB that saturate HBM at its latency.
  constexpr int kOutBwdRows = 4;
  int64_t m = mbeg;
  for (; m + kOutBwdRows <= mend; m += kOutBwdRows) {
    float4 h[kOutBwdRows];
    float y[kOutBwdRows][3], dy_in[kOutBwdRows][3];
#pragma unroll
    for (int r = 0; r < kOutBwdRows; ++r) {
      h[r] = *reinterpret_cast<const float4*>(h2 + (m + r) * 128 + 4 * lane);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        y[r][c] = rgb[3 * (m + r) + c];
        dy_in[r][c] = d_rgb[3 * (m + r) + c];
      }
    }
#pragma unroll
    for (int r = 0; r < kOutBwdRows; ++r) process(m + r, h[r], y[r], dy_in[r]);
  }
  for (; m < mend; ++m) {
    float4 ha = *reinterpret_cast<const float4*>(h2 + m * 128 + 4 * lane);
    float ya[3], da[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      ya[c] = rgb[3 * m + c];
      da[c] = d_rgb[3 * m + c];
    }
    process(m, ha, ya, da);
  }
  flush_embed();
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < 3; ++c) red[wib][(4 * lane + j) * 3 + c] = gw[j][c];
  if (lane == 0) {
    red[wib][384] = gb[0];
    red[wib][385] = gb[1];
    red[wib][386] = gb[2];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 387; i += 256) {
    float a = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) a += red[q][i];
    if (i < 384)
      atomicAdd(&dw3[i], a);
    else
      atomicAdd(&db3[i - 384], a);
  }
}

// Column sums (bias gradients): out[n] += sum_m G[m][n], n < 128.
__global__ void __launch_bounds__(256) k_colsum128(const float* __restrict__ G, float* __restrict__ out, int64_t M,
                                                   int64_t chunk) {
  __shared__ float part[128];
  const int n = threadIdx.x & 127, half = threadIdx.x >> 7;
  int64_t mbeg = (int64_t)blockIdx.x * chunk, mend = min(M, mbeg + chunk);
  float a = 0.f;
  for (int64_t m = mbeg + half; m < mend; m += 2) a += G[m * 128 + n];
  if (half == 1) part[n] = a;
  __syncthreads();
  if (half == 0) atomicAdd(&out[n], a + part[n]);
}

// ---------------------------------------------------------------------------------------------
int mlp_encode_fwd(cudaStream_t st, const MlpShape& s, const MlpWs& ws, const float* viewdirs, int64_t M, int rows_per_ray,
                   bool fast) {
  const unsigned grid = (unsigned)(s.squash + 3 <= 32 ? ceil_div64(M, 8) : ceil_div64(M * (s.squash + 3), 256));
  const bool rows = fast && s.squash + 3 <= 32 && ws.ldx > 256 && ws.ldx % 4 == 0 && (size_t)8 * ws.ldx * 4 <= 48 * 1024 &&
                    (reinterpret_cast<uintptr_t>(ws.x) & 15) == 0 && !getenv("TENSORF_ENC_ELEMENTWISE");
  if (rows)
    k_encode_fwd_rows<<<(unsigned)ceil_div64(M, 8), 256, (size_t)8 * ws.ldx * 4, st>>>(ws.f, viewdirs, ws.x, M, rows_per_ray, s,
                                                                                     ws.ldf, ws.ldx);
  else if (fast)
    k_encode_fwd<true><<<grid, 256, 0, st>>>(ws.f, viewdirs, ws.x, M, rows_per_ray, s, ws.ldf, ws.ldx);
  else
    k_encode_fwd<false><<<grid, 256, 0, st>>>(ws.f, viewdirs, ws.x, M, rows_per_ray, s, ws.ldf, ws.ldx);
  TF_CHECK_LAUNCH();
  return 0;
}
int mlp_encode_bwd(cudaStream_t st, const MlpShape& s, const MlpWs& ws, int64_t M, bool fast) {
  const unsigned grid = (unsigned)(s.squash <= 32 ? ceil_div64(M, 8) : ceil_div64(M * s.squash, 256));
  if (fast)
    k_encode_bwd<true><<<grid, 256, 0, st>>>(ws.f, ws.dx, ws.df, M, s, ws.ldf, ws.ldx);
  else
    k_encode_bwd<false><<<grid, 256, 0, st>>>(ws.f, ws.dx, ws.df, M, s, ws.ldf, ws.ldx);
  TF_CHECK_LAUNCH();
  return 0;
}
int mlp_out_fwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const MlpWs& ws, const uint32_t* cams, int64_t M,
                int rows_per_ray, float* rgb) {
  int64_t warps = std::min<int64_t>(M, (int64_t)kSMs * 64);
  k_out_fwd<<<(unsigned)ceil_div64(warps * 32, 256), 256, 0, st>>>(ws.h2, p.w3, p.b3, s.ncam ? p.embed : nullptr, cams, rgb, M,
                                                                    rows_per_ray);
  TF_CHECK_LAUNCH();
  return 0;
}
int mlp_out_bwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const MlpWs& ws, const uint32_t* cams, int64_t M,
                int rows_per_ray, const float* rgb, const float* d_rgb, const MlpGrads& gr) {
  // chunk = whole rays (one embedding flush per ray), ~64 rows per warp, 8 warps per block
  int64_t target = 64;
  if (const char* e = getenv("TENSORF_OUTBWD_CHUNK")) target = std::max(1, atoi(e));
  int64_t chunk = ceil_div64(std::max<int64_t>(target, rows_per_ray), rows_per_ray) * rows_per_ray;
  int64_t warps = ceil_div64(M, chunk);
  k_out_bwd<<<(unsigned)ceil_div64(warps, 8), 256, 0, st>>>(ws.h2, p.w3, s.ncam ? p.embed : nullptr, cams, rgb, d_rgb,
                                                                    ws.dp2, gr.w3, gr.b3, gr.embed, M, rows_per_ray, chunk);
  TF_CHECK_LAUNCH();
  return 0;
}
int mlp_colsum128(cudaStream_t st, const float* G, float* out, int64_t M) {
  const int64_t cs_chunk = std::max<int64_t>(256, ceil_div64(M, 4 * kSMs));
  k_colsum128<<<(unsigned)ceil_div64(M, cs_chunk), 256, 0, st>>>(G, out, M, cs_chunk);
  TF_CHECK_LAUNCH();
  return 0;
}
// All MLP gradient leaves are zeroed by ONE launch (8 separate memsets cost ~2 us each in launch gaps).
struct ZeroJobs {
  float* p[8];
  int64_t n[8];
};
__global__ void __launch_bounds__(256) k_zero_multi(ZeroJobs z) {
  float* p = z.p[blockIdx.y];
  const int64_t n = z.n[blockIdx.y];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0.f;
}
int mlp_zero_grads(cudaStream_t st, const MlpShape& s, const MlpGrads& gr) {
  const int U = s.units;
  ZeroJobs z{};
  int k = 0;
  auto add = [&](float* p, int64_t n) {
    if (p && n > 0) {
      z.p[k] = p;
      z.n[k] = n;
      ++k;
    }
  };
  add(gr.w0, (int64_t)s.Ca * s.squash);
  add(gr.w1, (int64_t)s.enc * U);
  add(gr.b1, U);
  add(gr.w2, (int64_t)U * U);
  add(gr.b2, U);
  add(gr.w3, (int64_t)U * 3);
  add(gr.b3, 3);
  if (s.ncam) add(gr.embed, (int64_t)s.ncam * U);
  if (k == 0) return 0;
  k_zero_multi<<<dim3(64, k), 256, 0, st>>>(z);
  TF_CHECK_LAUNCH();
  return 0;
}

int mlp_simt_fwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const float* feat, const float* viewdirs,
                 const uint32_t* cams, int64_t M, int rows_per_ray, const MlpWs& ws, float* rgb) {
  if (M == 0) return 0;
  GemmArgs g{};
  // Dense_0: f = feat @ W0 (no bias)
  g = GemmArgs{feat, s.Ca, 1, p.w0, s.squash, 1, ws.f, ws.ldf, nullptr, nullptr, 0, M, s.squash, s.Ca, 0, 0};
  TF_RETURN_IF_ERROR((launch_sgemm<128, 32, 4, 4>(st, g, 1)));
  TF_RETURN_IF_ERROR(mlp_encode_fwd(st, s, ws, viewdirs, M, rows_per_ray));
  // Dense_1 + relu
  g = GemmArgs{ws.x, ws.ldx, 1, p.w1, s.units, 1, ws.h1, s.units, p.b1, nullptr, 0, M, s.units, s.enc, 0, 1};
  TF_RETURN_IF_ERROR((launch_sgemm<128, 128, 8, 8>(st, g, 1)));
  // Dense_2 + relu
  g = GemmArgs{ws.h1, s.units, 1, p.w2, s.units, 1, ws.h2, s.units, p.b2, nullptr, 0, M, s.units, s.units, 0, 1};
  TF_RETURN_IF_ERROR((launch_sgemm<128, 128, 8, 8>(st, g, 1)));
  return mlp_out_fwd(st, s, p, ws, cams, M, rows_per_ray, rgb);
}

int mlp_simt_bwd(cudaStream_t st, const MlpShape& s, const MlpParams& p, const float* feat, const float* viewdirs,
                 const uint32_t* cams, int64_t M, int rows_per_ray, const MlpWs& ws, const float* rgb, const float* d_rgb,
                 float* d_feat, const MlpGrads& gr) {
  const int U = s.units;
  if (!gr.prezeroed) TF_RETURN_IF_ERROR(mlp_zero_grads(st, s, gr));
  if (M == 0) return 0;
  (void)viewdirs;
  TF_RETURN_IF_ERROR(mlp_out_bwd(st, s, p, ws, cams, M, rows_per_ray, rgb, d_rgb, gr));
  const int splits = (int)std::min<int64_t>(2 * kSMs, std::max<int64_t>(1, M / 256));
  GemmArgs g{};
  // dW2 = h1^T dp2 ; db2
  g = GemmArgs{ws.h1, 1, U, ws.dp2, U, 1, gr.w2, U, nullptr, nullptr, 0, U, U, M, 0, 0};
  TF_RETURN_IF_ERROR((launch_sgemm<128, 128, 8, 8>(st, g, splits)));
  TF_RETURN_IF_ERROR(mlp_colsum128(st, ws.dp2, gr.b2, M));
  // dp1 = (dp2 @ W2^T) * (h1 > 0)
  g = GemmArgs{ws.dp2, U, 1, p.w2, 1, U, ws.dp1, U, nullptr, ws.h1, U, M, U, U, 0, 0};
  TF_RETURN_IF_ERROR((launch_sgemm<128, 128, 8, 8>(st, g, 1)));
  // dW1 = x^T dp1 ; db1
  g = GemmArgs{ws.x, 1, ws.ldx, ws.dp1, U, 1, gr.w1, U, nullptr, nullptr, 0, s.enc, U, M, 0, 0};
  TF_RETURN_IF_ERROR((launch_sgemm<128, 128, 8, 8>(st, g, splits)));
  TF_RETURN_IF_ERROR(mlp_colsum128(st, ws.dp1, gr.b1, M));
  // dx = dp1 @ W1^T
  g = GemmArgs{ws.dp1, U, 1, p.w1, 1, U, ws.dx, ws.ldx, nullptr, nullptr, 0, M, s.enc, U, 0, 0};
  TF_RETURN_IF_ERROR((launch_sgemm<128, 128, 8, 8>(st, g, 1)));
  // df through the Fourier features
  TF_RETURN_IF_ERROR(mlp_encode_bwd(st, s, ws, M));
  // dW0 = feat^T df
  g = GemmArgs{feat, 1, s.Ca, ws.df, ws.ldf, 1, gr.w0, s.squash, nullptr, nullptr, 0, s.Ca, s.squash, M, 0, 0};
  TF_RETURN_IF_ERROR((launch_sgemm<128, 32, 4, 4>(st, g, splits)));
  // d_feat = df @ W0^T
  g = GemmArgs{ws.df, ws.ldf, 1, p.w0, 1, s.squash, d_feat, s.Ca, nullptr, nullptr, 0, M, s.Ca, s.squash, 0, 0};
  TF_RETURN_IF_ERROR((launch_sgemm<128, 128, 8, 8>(st, g, 1)));
  return 0;
}

}  // namespace tf
