// Optimiser step + grid resampling launchers (optim_kernels.cu).
#pragma once
#include <cmath>

#include "common.cuh"

namespace tf {

#ifdef __CUDACC__
// One element, in optax's operation order with every rounding kept (no FMA contraction):
//   mu' = (1-b1)*g + b1*mu ; nu' = (1-b2)*g^2 + b2*nu              (optax update_moment)
//   u   = (mu'/bc1) / (sqrt(nu'/bc2 + eps_root) + eps)               (bias_correction, scale_by_adam)
//   p'  = p + lr_decay * (neg_lr * u)                                (scale, training.py:186, apply_updates)
// `A` carries the scalars b1, b2, one_minus_b1, one_minus_b2, eps, eps_root, bc1, bc2, lr_decay
// (shared by k_adam and the peer-memory kernel, so both produce the same bits).
template <class A>
__device__ __forceinline__ void adam_one(float& p, float g, float& mu, float& nu, const A& a, float neg_lr) {
  mu = __fadd_rn(__fmul_rn(a.one_minus_b1, g), __fmul_rn(a.b1, mu));
  nu = __fadd_rn(__fmul_rn(a.one_minus_b2, __fmul_rn(g, g)), __fmul_rn(a.b2, nu));
  const float mh = __fdiv_rn(mu, a.bc1), nh = __fdiv_rn(nu, a.bc2);
  const float u = __fdiv_rn(mh, __fadd_rn(__fsqrt_rn(__fadd_rn(nh, a.eps_root)), a.eps));
  p = __fadd_rn(p, __fmul_rn(a.lr_decay, __fmul_rn(neg_lr, u)));
}
#endif

int64_t adam_scratch_bytes(const int64_t* sizes, int n_leaves);
int adam_step(cudaStream_t st, const tensorf_adam_desc* d, const int64_t* sizes, float* const* params, const float* const* grads,
              float* const* mu, float* const* nu, const float* neg_lrs, float* grad_norm, void* scratch,
              int64_t scratch_bytes);

// Peer-memory reduce + Adam + broadcast (peer_kernels.cu, SURVEY 8e fused follow-up).
void peer_shard(int64_t total, int rank, int world, int64_t* begin, int64_t* end);
int64_t peer_adam_scratch_bytes(int64_t shard_floats);
int adam_step_peer(cudaStream_t st, const tensorf_peer_adam_desc* d, const int64_t* leaf_offsets, const float* neg_lrs,
                   const float* const* grad_peers, float* const* param_peers, const float* grad_mc, float* param_mc,
                   float* mu_shard, float* nu_shard, float* const* norm_slot_peers, void* scratch, int64_t scratch_bytes);
int peer_set_max_ctas(int max_ctas);
int peer_allreduce(cudaStream_t st, int rank, int world, int64_t total, float* const* peers, float* mc,
                   uint32_t* const* signal_peers, uint32_t* local_flags, uint32_t epoch);
int peer_grad_norm(cudaStream_t st, const float* norm_slots, int world, float* grad_norm);

int64_t vm_resize_scratch_bytes(int C, int G_in, int G_out);
int vm_resize(cudaStream_t st, const float* vector_in, const float* matrix_in, int C, int G_in, int G_out, float* vector_out,
              float* matrix_out, void* scratch, int64_t scratch_bytes);

}  // namespace tf
