// Device functions shared by every kernel on the path: ray sampling, coordinate
// normalisation and the VM-decomposition taps.  Every kernel (forward gather, backward
// scatter, appearance lookup) calls the SAME functions so the coordinates are bit-identical
// between the forward and the reverse pass.
#pragma once
#include "common.cuh"

namespace tf {

// One axis of jax.scipy.ndimage.map_coordinates(order=1, mode="nearest")
// (tensor_vm.py:226-250): lower=floor(x); weights from the UNCLIPPED floor; indices clipped.
struct Tap {
  int i0, i1;
  float w0, w1;
};

__device__ __forceinline__ Tap make_tap(float x, int G) {
  Tap t;
  float l = floorf(x);
  t.w1 = __fsub_rn(x, l);        // upper_weight = x - lower
  t.w0 = __fsub_rn(1.0f, t.w1);  // lower_weight = 1 - upper_weight
  int i = __float2int_rz(l);     // saturating; the clip below absorbs out-of-range values
  t.i0 = min(max(i, 0), G - 1);
  t.i1 = min(max(i, -1), G - 2) + 1;  // clip(i+1, 0, G-1) without overflow
  return t;
}

// [-1,1] -> [0, G-1]: tensor_vm.py:150-153.
__device__ __forceinline__ float grid_coord(float q, float gm1) {
  return __fmul_rn(__fmul_rn(__fadd_rn(q, 1.0f), 0.5f), gm1);
}

struct RayParams {
  float o[3], d[3];
  float t0, step;  // bounded scene only
};

struct SceneParams {
  float a0[3], inv_ext_unused[3], ext[3];  // aabb[0], (aabb[1]-aabb[0])
  int N, G, contracted;
  const float* jitter;   // (N,) bounded / (R,N) contracted
  const float* base_ts;  // (N,) contracted
  const float* deltas;   // (N,) contracted
};

__device__ __forceinline__ void load_scene(SceneParams& sc, const float* aabb, int N, int G, int contracted,
                                           const float* jitter, const float* base_ts, const float* deltas) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    sc.a0[a] = aabb[a];
    sc.ext[a] = __fsub_rn(aabb[3 + a], aabb[a]);
  }
  sc.N = N;
  sc.G = G;
  sc.contracted = contracted;
  sc.jitter = jitter;
  sc.base_ts = base_ts;
  sc.deltas = deltas;
}

// render.py:399-434 ray_segment_from_bounding_box + :369 step size.
__device__ __forceinline__ void load_ray(RayParams& ray, const SceneParams& sc, const float* origins,
                                         const float* directions, const float* aabb, int r) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    ray.o[a] = origins[3 * r + a];
    ray.d[a] = directions[3 * r + a];
  }
  ray.t0 = 0.f;
  ray.step = 0.f;
  if (!sc.contracted) {
    float tmin = -INFINITY, tmax = INFINITY;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float den = __fadd_rn(ray.d[a], 1e-8f);
      float tl = __fdiv_rn(__fsub_rn(aabb[a], ray.o[a]), den);
      float th = __fdiv_rn(__fsub_rn(aabb[3 + a], ray.o[a]), den);
      tmin = fmaxf(tmin, fminf(tl, th));
      tmax = fminf(tmax, fmaxf(tl, th));
    }
    float t_min = fmaxf(0.0f, tmin);
    float t_max = tmax;
    float t_max_c = fmaxf(t_max, __fadd_rn(t_min, 1e-3f));
    bool valid = t_min < t_max;
    float t0 = valid ? t_min : 0.0f;
    float t1 = valid ? t_max_c : 1e-3f;
    ray.t0 = t0;
    ray.step = __fdiv_rn(__fsub_rn(t1, t0), (float)sc.N);
  }
}

// Distance along the ray and step size of sample s: render.py:372-381 (bounded) or
// render.py:151-161 (contracted).
__device__ __forceinline__ float sample_t(const RayParams& ray, const SceneParams& sc, int r, int s, float& delta) {
  if (sc.contracted) {
    delta = sc.deltas[s];
    float u = sc.jitter[(int64_t)r * sc.N + s];
    return __fadd_rn(sc.base_ts[s], __fmul_rn(delta, u));
  }
  delta = ray.step;
  float ts = __fadd_rn((float)s, sc.jitter[s]);
  ts = __fmul_rn(ts, ray.step);
  return __fadd_rn(ray.t0, ts);
}

// World point -> contraction (render.py:164-171) -> [-1,1] (render.py:193-195) -> grid
// coordinates (tensor_vm.py:150-153). x[] = continuous grid coordinate per world axis.
__device__ __forceinline__ void sample_grid_coords(const RayParams& ray, const SceneParams& sc, float t, float x[3]) {
  float p[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) p[a] = __fadd_rn(ray.o[a], __fmul_rn(t, ray.d[a]));
  if (sc.contracted) {
    float n = fmaxf(fabsf(p[0]), fmaxf(fabsf(p[1]), fabsf(p[2])));
    if (!(n <= 1.0f)) {
      float f = __fsub_rn(2.0f, __fdiv_rn(1.0f, n));
#pragma unroll
      for (int a = 0; a < 3; ++a) p[a] = __fdiv_rn(__fmul_rn(f, p[a]), n);
    }
  }
  float gm1 = (float)(sc.G - 1);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float q = __fmul_rn(__fsub_rn(__fdiv_rn(__fsub_rn(p[a], sc.a0[a]), sc.ext[a]), 0.5f), 2.0f);
    x[a] = grid_coord(q, gm1);
  }
}

// Axis roles of the three vector-matrix pairs (tensor_vm.py:50-52 with :155-160):
//   pair 0: line x, plane (y,z);  pair 1: line z, plane (x,y);  pair 2: line y, plane (z,x).
// The plane's first coordinate is the row (stride G), the second is contiguous.
struct VmTaps {
  Tap ax[3];  // per world axis
};
__device__ __forceinline__ void make_vm_taps(VmTaps& t, const float x[3], int G) {
#pragma unroll
  for (int a = 0; a < 3; ++a) t.ax[a] = make_tap(x[a], G);
}
__device__ __forceinline__ constexpr int pair_line_axis(int P) { return P == 0 ? 0 : (P == 1 ? 2 : 1); }
__device__ __forceinline__ constexpr int pair_row_axis(int P) { return P == 0 ? 1 : (P == 1 ? 0 : 2); }
__device__ __forceinline__ constexpr int pair_col_axis(int P) { return P == 0 ? 2 : (P == 1 ? 1 : 0); }

__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_fma(float4 a, float s, float4 c) {
  return make_float4(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z), fmaf(a.w, s, c.w));
}

// Pointers to the six texel fragments (float4 group v) one pair touches.
struct PairAddr {
  // float offsets into the packed buffer; 32-bit: entry points reject factors with >= 2^31 packed floats
  // (64-bit offsets cost two to three instructions per address in kernels that are issue-bound)
  int l0, l1, m00, m01, m10, m11;
  float wl0, wl1, w00, w01, w10, w11;
};
__device__ __forceinline__ PairAddr pair_addr(const VmTaps& t, int P, int G, int Cp, int v) {
  const Tap& L = t.ax[pair_line_axis(P)];
  const Tap& A = t.ax[pair_row_axis(P)];
  const Tap& B = t.ax[pair_col_axis(P)];
  PairAddr pa;
  const int lbase = P * G * Cp + 4 * v;
  pa.l0 = lbase + L.i0 * Cp;
  pa.l1 = lbase + L.i1 * Cp;
  const int mbase = 3 * G * Cp + P * G * G * Cp + 4 * v;
  pa.m00 = mbase + (A.i0 * G + B.i0) * Cp;
  pa.m01 = mbase + (A.i0 * G + B.i1) * Cp;
  pa.m10 = mbase + (A.i1 * G + B.i0) * Cp;
  pa.m11 = mbase + (A.i1 * G + B.i1) * Cp;
  pa.wl0 = L.w0;
  pa.wl1 = L.w1;
  // map_coordinates multiplies the weights first, then the gathered value.
  pa.w00 = __fmul_rn(A.w0, B.w0);
  pa.w01 = __fmul_rn(A.w0, B.w1);
  pa.w10 = __fmul_rn(A.w1, B.w0);
  pa.w11 = __fmul_rn(A.w1, B.w1);
  return pa;
}

// lin = w0*v[i0] + w1*v[i1];  bil = sum over (lo,lo),(lo,hi),(hi,lo),(hi,hi) (tensor_vm.py:226-250)
__device__ __forceinline__ void pair_values(const float* __restrict__ packed, const PairAddr& pa, float4& lin, float4& bil) {
  float4 l0 = ld4(packed + pa.l0), l1 = ld4(packed + pa.l1);
  float4 m00 = ld4(packed + pa.m00), m01 = ld4(packed + pa.m01);
  float4 m10 = ld4(packed + pa.m10), m11 = ld4(packed + pa.m11);
  lin = f4_fma(l1, pa.wl1, f4_scale(l0, pa.wl0));
  bil = f4_scale(m00, pa.w00);
  bil = f4_fma(m01, pa.w01, bil);
  bil = f4_fma(m10, pa.w10, bil);
  bil = f4_fma(m11, pa.w11, bil);
}

// Order-preserving float -> uint key: a > b (as floats, -0 < +0) <=> key(a) > key(b).
__device__ __forceinline__ uint32_t ordered_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// jax.nn.softplus = logaddexp(z, 0) (render.py:207)
__device__ __forceinline__ float softplus_f(float z) { return fmaxf(z, 0.0f) + log1pf(expf(-fabsf(z))); }

}  // namespace tf
