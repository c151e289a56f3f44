// TensorVM kernels: layout pack/unpack and the standalone `TensorVM.interpolate`
// forward / reverse (tensor_vm.py:42-89, :140-167, :226-250).
#include <stdlib.h>

#include <algorithm>

#include "vm.cuh"

namespace tf {

// ---- pack: channel-first (C, T) -> texel-major (T, Cp) ---------------------------------------
// One CTA moves 32 texels x all channels through shared memory: reads are coalesced along the
// texel axis (the reference's contiguous axis), writes are one contiguous 32*Cp-float run.
// Up to 4 arrays (lines and planes of the density and appearance factors) go out in ONE launch: the
// per-array kernels were launch-bound at 128^3 (4 launches of 5-12 us for 12.7 MB).
struct PackJob {
  const float* src;
  float* dst;
  int C, Cp;
  int64_t T;               // texels per pair
  int64_t tiles_per_pair;  // ceil(T / 32)
  int64_t block_begin;     // first CTA of this job
};
struct PackJobsVm {
  PackJob j[4];
  int n;
};

// TW = texels per tile.  The channel-first side is read / written in runs of TW*4 bytes per channel, each in a different
// DRAM page (channel stride = T*4 bytes): TW = 128 moves four lines per page visit, which matters once the factors no
// longer fit in L2 (300^3: 139 MB per direction).
template <bool UNPACK, int TW>
__global__ void __launch_bounds__(256) k_pack(const __grid_constant__ PackJobsVm jobs) {
  extern __shared__ float tile[];  // [Cp][TW + 1]
  int ji = 0;
#pragma unroll 1
  while (ji + 1 < jobs.n && (int64_t)blockIdx.x >= jobs.j[ji + 1].block_begin) ++ji;
  const PackJob& J = jobs.j[ji];
  const int C = J.C, Cp = J.Cp;
  const int64_t T = J.T;
  const int64_t tile_id = (int64_t)blockIdx.x - J.block_begin;
  const int P = (int)(tile_id / J.tiles_per_pair);
  const int64_t t0 = (tile_id % J.tiles_per_pair) * TW;
  constexpr int RPB = 256 / TW;  // channel rows handled per pass
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int n = (int)min((int64_t)TW, T - t0) * Cp;
  // (texel, channel) of packed element i = threadIdx.x + 256*k, advanced without a division per element
  const int dq = 256 / Cp, dr = 256 % Cp;
  int tq = threadIdx.x / Cp, c = threadIdx.x % Cp;
  if (!UNPACK) {
    const float* s = J.src + (int64_t)P * C * T;
    for (int ch = ty; ch < Cp; ch += RPB) {
      int64_t t = t0 + tx;
      tile[ch * (TW + 1) + tx] = (ch < C && t < T) ? s[(int64_t)ch * T + t] : 0.0f;
    }
    __syncthreads();
    float* d = J.dst + ((int64_t)P * T + t0) * Cp;
    for (int i = threadIdx.x; i < n; i += 256) {
      d[i] = tile[c * (TW + 1) + tq];
      tq += dq;
      c += dr;
      if (c >= Cp) {
        c -= Cp;
        ++tq;
      }
    }
  } else {
    const float* s = J.src + ((int64_t)P * T + t0) * Cp;
    for (int i = threadIdx.x; i < n; i += 256) {
      tile[c * (TW + 1) + tq] = s[i];
      tq += dq;
      c += dr;
      if (c >= Cp) {
        c -= Cp;
        ++tq;
      }
    }
    __syncthreads();
    float* d = J.dst + (int64_t)P * C * T;
    for (int ch = ty; ch < C; ch += RPB) {
      int64_t t = t0 + tx;
      if (t < T) d[(int64_t)ch * T + t] = tile[ch * (TW + 1) + tx];
    }
  }
}

// factors[i] = {vector, matrix, packed, C}; G shared.  UNPACK: packed -> vector/matrix.
template <bool UNPACK, int TW>
static int launch_pack_tw(cudaStream_t st, int nf, const float* const* vec, const float* const* mat, const float* const* packed,
                          const int* C, int G) {
  PackJobsVm jobs{};
  int64_t blocks = 0;
  int maxCp = 0;
  for (int f = 0; f < nf; ++f) {
    const int Cp = packed_cp(C[f]);
    maxCp = std::max(maxCp, Cp);
    for (int part = 0; part < 2; ++part) {
      PackJob& J = jobs.j[jobs.n++];
      const float* chan_first = part == 0 ? vec[f] : mat[f];
      const float* pk = packed[f] + (part == 0 ? 0 : packed_line_floats(C[f], G));
      J.src = UNPACK ? pk : chan_first;
      J.dst = const_cast<float*>(UNPACK ? chan_first : pk);
      J.C = C[f];
      J.Cp = Cp;
      J.T = part == 0 ? G : (int64_t)G * G;
      J.tiles_per_pair = ceil_div64(J.T, TW);
      J.block_begin = blocks;
      blocks += 3 * J.tiles_per_pair;
    }
  }
  const size_t smem = (size_t)maxCp * (TW + 1) * sizeof(float);
  TF_CHECK_ARG(smem <= 48 * 1024, "channel dim too large for pack kernel (Cp=%d)", maxCp);
  TF_CHECK_ARG(blocks < ((int64_t)1 << 31), "pack: grid too large");
  k_pack<UNPACK, TW><<<(unsigned)blocks, 256, smem, st>>>(jobs);
  TF_CHECK_LAUNCH();
  return 0;
}
template <bool UNPACK>
static int launch_pack_multi(cudaStream_t st, int nf, const float* const* vec, const float* const* mat, const float* const* packed,
                             const int* C, int G) {
  int maxCp = 0;
  for (int f = 0; f < nf; ++f) maxCp = std::max(maxCp, packed_cp(C[f]));
  // measured on B200 (pack + unpack): 128^3: TW 32: 31.8 us, 64: 26.6, 128: 26.7; 300^3: 110.9, 74.9, 71.6
  int tw = (size_t)maxCp * 129 * sizeof(float) <= 48 * 1024 ? 128 : ((size_t)maxCp * 65 * sizeof(float) <= 48 * 1024 ? 64 : 32);
  if (const char* e = getenv("TENSORF_PACK_TW")) tw = atoi(e);
  if (tw == 128) return launch_pack_tw<UNPACK, 128>(st, nf, vec, mat, packed, C, G);
  if (tw == 64) return launch_pack_tw<UNPACK, 64>(st, nf, vec, mat, packed, C, G);
  return launch_pack_tw<UNPACK, 32>(st, nf, vec, mat, packed, C, G);
}

int vm_pack(cudaStream_t st, const float* vector, const float* matrix, float* packed, int C, int G) {
  const float* pk = packed;
  return launch_pack_multi<false>(st, 1, &vector, &matrix, &pk, &C, G);
}
int vm_unpack(cudaStream_t st, const float* packed, float* vector, float* matrix, int C, int G) {
  const float* v = vector;
  const float* m = matrix;
  return launch_pack_multi<true>(st, 1, &v, &m, &packed, &C, G);
}
// density + appearance factors in one launch
int vm_pack2(cudaStream_t st, const float* v0, const float* m0, float* p0, int C0, const float* v1, const float* m1, float* p1, int C1,
             int G) {
  const float* v[2] = {v0, v1};
  const float* m[2] = {m0, m1};
  const float* p[2] = {p0, p1};
  const int C[2] = {C0, C1};
  return launch_pack_multi<false>(st, 2, v, m, p, C, G);
}
int vm_unpack2(cudaStream_t st, const float* p0, float* v0, float* m0, int C0, const float* p1, float* v1, float* m1, int C1, int G) {
  const float* v[2] = {v0, v1};
  const float* m[2] = {m0, m1};
  const float* p[2] = {p0, p1};
  const int C[2] = {C0, C1};
  return launch_pack_multi<true>(st, 2, v, m, p, C, G);
}

// ---- standalone interpolate -----------------------------------------------------------------
// Work item = (sample b, float4 channel group v).  Items are flattened so a warp covers whole
// texel fragments with contiguous lanes (a texel is nvec consecutive float4s): every load
// instruction moves full 32-byte sectors.  feature_major=1 writes rows (B, 3C); feature_major=0
// writes the reference's (3C, B).
__global__ void __launch_bounds__(256) k_vm_interp_fwd(const float* __restrict__ packed, const float* __restrict__ ijk,
                                                       float* __restrict__ out, int C, int Cp, int G, int64_t B,
                                                       int feature_major) {
  const int nvec = Cp >> 2;
  int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= B * nvec) return;
  int64_t b = item / nvec;
  int v = (int)(item % nvec);
  float gm1 = (float)(G - 1);
  float x[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) x[a] = grid_coord(ijk[(int64_t)a * B + b], gm1);
  VmTaps taps;
  make_vm_taps(taps, x, G);
#pragma unroll
  for (int P = 0; P < 3; ++P) {
    PairAddr pa = pair_addr(taps, P, G, Cp, v);
    float4 lin, bil;
    pair_values(packed, pa, lin, bil);
    float4 f = f4_mul(lin, bil);
    float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = 4 * v + j;
      if (c < C) {
        if (feature_major)
          out[b * (3 * C) + P * C + c] = fv[j];
        else
          out[((int64_t)(P * C + c)) * B + b] = fv[j];
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_vm_interp_bwd(const float* __restrict__ packed, const float* __restrict__ ijk,
                                                       const float* __restrict__ d_out, float* __restrict__ d_packed,
                                                       int C, int Cp, int G, int64_t B, int feature_major) {
  const int nvec = Cp >> 2;
  int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= B * nvec) return;
  int64_t b = item / nvec;
  int v = (int)(item % nvec);
  float gm1 = (float)(G - 1);
  float x[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) x[a] = grid_coord(ijk[(int64_t)a * B + b], gm1);
  VmTaps taps;
  make_vm_taps(taps, x, G);
#pragma unroll
  for (int P = 0; P < 3; ++P) {
    float gv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = 4 * v + j;
      gv[j] = 0.f;
      if (c < C) gv[j] = feature_major ? d_out[b * (3 * C) + P * C + c] : d_out[((int64_t)(P * C + c)) * B + b];
    }
    float4 g = make_float4(gv[0], gv[1], gv[2], gv[3]);
    PairAddr pa = pair_addr(taps, P, G, Cp, v);
    float4 lin, bil;
    pair_values(packed, pa, lin, bil);
    float4 gl = f4_mul(g, bil);  // d/d lin
    float4 gb = f4_mul(g, lin);  // d/d bil
    red_add_v4(d_packed + pa.l0, f4_scale(gl, pa.wl0));
    red_add_v4(d_packed + pa.l1, f4_scale(gl, pa.wl1));
    red_add_v4(d_packed + pa.m00, f4_scale(gb, pa.w00));
    red_add_v4(d_packed + pa.m01, f4_scale(gb, pa.w01));
    red_add_v4(d_packed + pa.m10, f4_scale(gb, pa.w10));
    red_add_v4(d_packed + pa.m11, f4_scale(gb, pa.w11));
  }
}

int vm_interp_fwd(cudaStream_t st, const float* packed, const float* ijk, float* out, int C, int G, int64_t B,
                  int feature_major) {
  if (B == 0) return 0;
  int Cp = packed_cp(C);
  int64_t items = B * (Cp / 4);
  k_vm_interp_fwd<<<(unsigned)ceil_div64(items, 256), 256, 0, st>>>(packed, ijk, out, C, Cp, G, B, feature_major);
  TF_CHECK_LAUNCH();
  return 0;
}
int vm_interp_bwd(cudaStream_t st, const float* packed, const float* ijk, const float* d_out, float* d_packed, int C,
                  int G, int64_t B, int feature_major) {
  if (B == 0) return 0;
  int Cp = packed_cp(C);
  int64_t items = B * (Cp / 4);
  k_vm_interp_bwd<<<(unsigned)ceil_div64(items, 256), 256, 0, st>>>(packed, ijk, d_out, d_packed, C, Cp, G, B,
                                                                     feature_major);
  TF_CHECK_LAUNCH();
  return 0;
}

}  // namespace tf
