// SURVEY §8e fused follow-up: the exchange step of sharded training and the optimiser step that follows
// it, as ONE kernel over peer memory instead of ncclAllReduce(grads) + k_adam on every rank.
//
//   rank r owns elements [begin, end) of the flat parameter vector (1/P of it):
//     g   = sum over ranks of grad_p[e]      P2P loads in rank order (identical on every run), or ONE
//                                            multimem.ld_reduce: the NVSwitch adds the P copies in flight
//     Adam step with the owner's shard of mu / nu (moment traffic and memory / P)
//     p'  -> every rank's parameter buffer   P2P stores, or ONE multimem.st broadcast by the switch
//     sum g^2 over the shard -> slot [r] of every rank (optax.global_norm comes from the P slots)
//
// Bytes over NVLink per rank: (P-1)/P * 4T in (gradients) + (P-1)/P * 4T out (parameters), the same as a ring
// allreduce, but Adam's 28 B/parameter local pass shrinks to 28/P and no rank ever holds the reduced gradient.
// Why the scatter kernels do NOT RED straight into the owner's peer buffer (the survey's first suggestion):
// k_scatter_walk issues one 16-byte RED per texel VISIT (~0.4 GB of REDs per step at lego 128^3); the local L2
// combines them into a 12.7 MB dense gradient, so reducing after the L2 moves 30x fewer bytes over the links.
// Ordering across ranks is the caller's (one barrier before, one after; see include/tensorf_b200.h).
#include <algorithm>

#include "optim.cuh"

namespace tf {

constexpr int kPeerThreads = 256;
constexpr int kPeerV4PerThread = 2;                               // float4 per thread and tile
constexpr int kPeerTile = kPeerThreads * kPeerV4PerThread * 4;    // elements per CTA tile
constexpr int kPeerMaxLeaves = TENSORF_PEER_MAX_LEAVES + 1;       // + the padding pseudo-leaf

struct PeerAdamArgs {
  const float* g[TENSORF_PEER_MAX_WORLD];
  float* p[TENSORF_PEER_MAX_WORLD];
  float* slots[TENSORF_PEER_MAX_WORLD];
  const float* g_mc;
  float* p_mc;
  float* mu;   // shard-local: element e lives at e - begin
  float* nu;
  int64_t off[kPeerMaxLeaves + 1];
  float neg_lr[kPeerMaxLeaves];
  int n_leaves;  // including the padding pseudo-leaf (neg_lr 0, excluded from the norm)
  int n_real;
  int rank, world;
  int64_t begin, end;
  int n_tiles;
  float b1, b2, one_minus_b1, one_minus_b2, eps, eps_root, bc1, bc2, lr_decay;
  float* partial;        // [gridDim.x]
  unsigned int* ticket;  // zeroed before launch
};

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}

// W > 0: world size known at compile time (all peer loads of a float4 in flight together); W == 0: runtime loop.
// MC: gradients through multimem.ld_reduce and parameters through multimem.st.
template <int W, bool MC>
__global__ void __launch_bounds__(kPeerThreads) k_adam_peer(const __grid_constant__ PeerAdamArgs a) {
  __shared__ float s_red[kPeerThreads / 32];
  __shared__ bool s_last;
  const int world = W > 0 ? W : a.world;
  const int64_t real_end = a.off[a.n_real];
  float ss = 0.f;
  int li = 0;
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int64_t base = a.begin + (int64_t)tile * kPeerTile;
    float4 g4[kPeerV4PerThread], p4[kPeerV4PerThread], m4[kPeerV4PerThread], v4[kPeerV4PerThread];
    bool on[kPeerV4PerThread];
#pragma unroll
    for (int i = 0; i < kPeerV4PerThread; ++i) {  // every load of the tile in flight before any arithmetic
      const int64_t e = base + ((int64_t)i * kPeerThreads + threadIdx.x) * 4;
      on[i] = e < a.end;
      if (!on[i]) continue;
      if (MC) {
        g4[i] = multimem_ld_reduce_add(a.g_mc + e);
      } else if (W > 0) {
        float4 t[W > 0 ? W : 1];
#pragma unroll
        for (int r = 0; r < W; ++r) t[r] = __ldcg(reinterpret_cast<const float4*>(a.g[r] + e));
        g4[i] = t[0];
#pragma unroll
        for (int r = 1; r < W; ++r) g4[i] = add4(g4[i], t[r]);  // rank order, every rounding kept
      } else {
        g4[i] = __ldcg(reinterpret_cast<const float4*>(a.g[0] + e));
        for (int r = 1; r < world; ++r) g4[i] = add4(g4[i], __ldcg(reinterpret_cast<const float4*>(a.g[r] + e)));
      }
      p4[i] = *reinterpret_cast<const float4*>(a.p[a.rank] + e);
      m4[i] = *reinterpret_cast<const float4*>(a.mu + (e - a.begin));
      v4[i] = *reinterpret_cast<const float4*>(a.nu + (e - a.begin));
    }
#pragma unroll
    for (int i = 0; i < kPeerV4PerThread; ++i) {
      if (!on[i]) continue;
      const int64_t e = base + ((int64_t)i * kPeerThreads + threadIdx.x) * 4;
      while (li + 1 < a.n_leaves && e >= a.off[li + 1]) ++li;  // e only grows along a thread's walk
      float nl[4];
      if (e + 3 < a.off[li + 1]) {
        nl[0] = nl[1] = nl[2] = nl[3] = a.neg_lr[li];
      } else {  // the float4 straddles a leaf boundary (leaves need not be multiples of 4 long)
        int lj = li;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          while (lj + 1 < a.n_leaves && e + j >= a.off[lj + 1]) ++lj;
          nl[j] = a.neg_lr[lj];
        }
      }
      float* gp = reinterpret_cast<float*>(&g4[i]);
      float* pp = reinterpret_cast<float*>(&p4[i]);
      float* mp = reinterpret_cast<float*>(&m4[i]);
      float* vp = reinterpret_cast<float*>(&v4[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (e + j < real_end) {
          ss += gp[j] * gp[j];
          adam_one(pp[j], gp[j], mp[j], vp[j], a, nl[j]);
        }
      }
      *reinterpret_cast<float4*>(a.mu + (e - a.begin)) = m4[i];
      *reinterpret_cast<float4*>(a.nu + (e - a.begin)) = v4[i];
      if (MC) {
        multimem_st(a.p_mc + e, p4[i]);
      } else if (W > 0) {
#pragma unroll
        for (int r = 0; r < W; ++r) *reinterpret_cast<float4*>(a.p[r] + e) = p4[i];
      } else {
        for (int r = 0; r < world; ++r) *reinterpret_cast<float4*>(a.p[r] + e) = p4[i];
      }
    }
  }
  // ---- sum of squares of the reduced gradient over this rank's shard: per-CTA partials, the last CTA to
  // arrive adds them in index order (fp64) and stores the shard's sum into slot [rank] of every rank ----
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kPeerThreads / 32; ++w) t += s_red[w];
    a.partial[blockIdx.x] = t;
    __threadfence();
    s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double acc = 0.0;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += kPeerThreads) acc += (double)__ldcg(a.partial + i);
  __shared__ double s_acc[kPeerThreads];
  s_acc[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kPeerThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_acc[threadIdx.x] += s_acc[threadIdx.x + o];
    __syncthreads();
  }
  if ((int)threadIdx.x < world) a.slots[threadIdx.x][a.rank] = (float)s_acc[0];
}

// All-reduce only (no optimiser step): rank r sums its range of every rank's buffer and stores the sum back into
// every rank's buffer, in place (ranges are disjoint, so no rank reads what another is writing).  Used where the
// caller keeps its own optimiser (bench.py's timed step, jax-side optax) but wants the exchange on the launch
// stream over peer memory instead of a library collective on a second stream.
struct PeerReduceArgs {
  float* x[TENSORF_PEER_MAX_WORLD];
  float* x_mc;
  int64_t begin, end;
  int n_tiles;
  int world;
  // in-kernel cross-rank ordering (sig[0] == nullptr: the caller brackets the launch with its own barriers)
  uint32_t* sig[TENSORF_PEER_MAX_WORLD];  // rank p's signal pad: [0,16) "gradients ready", [16,32) "stores landed"
  uint32_t* local;                        // this rank only: [0] gate, [1] arrival counter
  uint32_t epoch;                         // call number, starts at 1, the same on every rank
  int rank;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool epoch_reached(uint32_t seen, uint32_t epoch) { return (int32_t)(seen - epoch) >= 0; }

// Entry: block 0 tells every rank "my buffer is complete" (everything enqueued before this kernel has finished) and
// waits for the same from every rank, then opens the gate for the other blocks of this grid.  All blocks are
// co-resident (grid <= 4 per SM, see peer_allreduce) and block 0 is dispatched first: the spinning blocks cannot
// starve it, and block 0 only depends on the block 0 of every other rank's kernel.
__device__ __forceinline__ void peer_entry_barrier(const PeerReduceArgs& a, int world, uint32_t epoch) {
  if (blockIdx.x == 0) {
    if ((int)threadIdx.x < world) {
      st_release_sys(a.sig[threadIdx.x] + a.rank, epoch);
      while (!epoch_reached(ld_acquire_sys(a.sig[a.rank] + threadIdx.x), epoch)) {
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(a.local, epoch);
  } else {
    if (threadIdx.x == 0) {
      while (!epoch_reached(ld_acquire_gpu(a.local), epoch)) {
      }
    }
    __syncthreads();
  }
}
// Exit: the last block of this grid to finish its stores tells every rank "my stores have landed" and waits for the
// same from every rank, so the kernel (and with it the stream) only completes once every buffer holds every sum.
__device__ __forceinline__ void peer_exit_barrier(const PeerReduceArgs& a, int world, uint32_t epoch) {
  __shared__ bool s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    s_last = atomicAdd(a.local + 1, 1u) == gridDim.x - 1;
    if (s_last) a.local[1] = 0;  // every block has arrived: reset for the next call
    __threadfence_system();
  }
  __syncthreads();
  if (!s_last) return;
  if ((int)threadIdx.x < world) {
    st_release_sys(a.sig[threadIdx.x] + 16 + a.rank, epoch);
    while (!epoch_reached(ld_acquire_sys(a.sig[a.rank] + 16 + threadIdx.x), epoch)) {
    }
  }
  // device-side call counter (epoch argument 0): every block of this grid read it at entry and has arrived here, so the
  // last block may advance it for the next launch on the stream
  if (a.epoch == 0) {
    __syncthreads();
    if (threadIdx.x == 0) a.local[2] = epoch;
  }
}
constexpr int kPeerRedV4 = 4;  // float4 per thread and tile
constexpr int kPeerRedTile = kPeerThreads * kPeerRedV4 * 4;

template <int W, bool MC>
__global__ void __launch_bounds__(kPeerThreads) k_peer_allreduce(const __grid_constant__ PeerReduceArgs a) {
  const int world = W > 0 ? W : a.world;
  const bool sync = a.sig[0] != nullptr;
  // epoch 0: the call number lives in device memory (local[2] = calls completed so far), which makes the launch
  // CUDA-graph capturable - every replay sees the next epoch without a new launch argument
  const uint32_t epoch = !sync ? 0u : (a.epoch ? a.epoch : ld_acquire_gpu(a.local + 2) + 1u);
  if (sync) peer_entry_barrier(a, world, epoch);
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int64_t base = a.begin + (int64_t)tile * kPeerRedTile;
    float4 acc[kPeerRedV4];
#pragma unroll
    for (int i = 0; i < kPeerRedV4; ++i) {
      const int64_t e = base + ((int64_t)i * kPeerThreads + threadIdx.x) * 4;
      if (e >= a.end) continue;
      if (MC) {
        acc[i] = multimem_ld_reduce_add(a.x_mc + e);
      } else if (W > 0) {
        float4 t[W > 0 ? W : 1];
#pragma unroll
        for (int r = 0; r < W; ++r) t[r] = __ldcg(reinterpret_cast<const float4*>(a.x[r] + e));
        acc[i] = t[0];
#pragma unroll
        for (int r = 1; r < W; ++r) acc[i] = add4(acc[i], t[r]);
      } else {
        acc[i] = __ldcg(reinterpret_cast<const float4*>(a.x[0] + e));
        for (int r = 1; r < world; ++r) acc[i] = add4(acc[i], __ldcg(reinterpret_cast<const float4*>(a.x[r] + e)));
      }
    }
#pragma unroll
    for (int i = 0; i < kPeerRedV4; ++i) {
      const int64_t e = base + ((int64_t)i * kPeerThreads + threadIdx.x) * 4;
      if (e >= a.end) continue;
      if (MC) {
        multimem_st(a.x_mc + e, acc[i]);
      } else if (W > 0) {
#pragma unroll
        for (int r = 0; r < W; ++r) *reinterpret_cast<float4*>(a.x[r] + e) = acc[i];
      } else {
        for (int r = 0; r < world; ++r) *reinterpret_cast<float4*>(a.x[r] + e) = acc[i];
      }
    }
  }
  if (sync) peer_exit_barrier(a, world, epoch);
}

__global__ void k_peer_grad_norm(const float* __restrict__ slots, int world, float* __restrict__ out) {
  double acc = 0.0;
  for (int r = 0; r < world; ++r) acc += (double)__ldcg(slots + r);
  out[0] = (float)sqrt(acc);
}

void peer_shard(int64_t total, int rank, int world, int64_t* begin, int64_t* end) {
  const int64_t units = (total + 3) / 4;
  const int64_t base = units / world, rem = units % world;
  const int64_t b = rank * base + (rank < rem ? rank : rem);
  const int64_t e = b + base + (rank < rem ? 1 : 0);
  *begin = b * 4 < total ? b * 4 : total;
  *end = e * 4 < total ? e * 4 : total;
}

static int peer_grid(int64_t shard_floats) {
  const int64_t tiles = ceil_div64(shard_floats, kPeerTile);
  const int64_t cap = (int64_t)kSMs * 8;  // persistent: at most 8 CTAs of 256 threads per SM
  return (int)(tiles < 1 ? 1 : (tiles < cap ? tiles : cap));
}

int64_t peer_adam_scratch_bytes(int64_t shard_floats) {
  if (shard_floats < 0) return -1;
  return 16 + (int64_t)peer_grid(shard_floats) * (int64_t)sizeof(float);
}

int adam_step_peer(cudaStream_t st, const tensorf_peer_adam_desc* d, const int64_t* leaf_offsets, const float* neg_lrs,
                   const float* const* grad_peers, float* const* param_peers, const float* grad_mc, float* param_mc,
                   float* mu_shard, float* nu_shard, float* const* norm_slot_peers, void* scratch, int64_t scratch_bytes) {
  TF_CHECK_ARG(d && leaf_offsets && neg_lrs && grad_peers && param_peers && norm_slot_peers, "adam_peer: null argument");
  const int nl = d->adam.n_leaves;
  TF_CHECK_ARG(nl >= 1 && nl <= TENSORF_PEER_MAX_LEAVES, "adam_peer: n_leaves=%d outside [1,%d]", nl, TENSORF_PEER_MAX_LEAVES);
  TF_CHECK_ARG(d->world >= 1 && d->world <= TENSORF_PEER_MAX_WORLD, "adam_peer: world=%d outside [1,%d]", d->world,
               TENSORF_PEER_MAX_WORLD);
  TF_CHECK_ARG(d->rank >= 0 && d->rank < d->world, "adam_peer: rank %d outside world of %d", d->rank, d->world);
  TF_CHECK_ARG(d->total >= 0 && d->total % 4 == 0, "adam_peer: total=%lld must be a non-negative multiple of 4",
               (long long)d->total);
  TF_CHECK_ARG(d->shard_begin >= 0 && d->shard_begin <= d->shard_end && d->shard_end <= d->total && d->shard_begin % 4 == 0 &&
                   d->shard_end % 4 == 0,
               "adam_peer: shard [%lld,%lld) is not a 4-aligned range inside [0,%lld]", (long long)d->shard_begin,
               (long long)d->shard_end, (long long)d->total);
  TF_CHECK_ARG(d->adam.bias_correction1 > 0.f && d->adam.bias_correction2 > 0.f,
               "adam_peer: bias corrections must be > 0 (step count >= 1)");
  TF_CHECK_ARG(leaf_offsets[0] == 0, "adam_peer: leaf_offsets[0] must be 0");
  for (int i = 0; i < nl; ++i)
    TF_CHECK_ARG(leaf_offsets[i + 1] >= leaf_offsets[i], "adam_peer: leaf_offsets must be non-decreasing (leaf %d)", i);
  TF_CHECK_ARG(leaf_offsets[nl] <= d->total && d->total - leaf_offsets[nl] < 4,
               "adam_peer: leaves cover %lld floats but total=%lld (padding must be < 4 floats)", (long long)leaf_offsets[nl],
               (long long)d->total);
  TF_CHECK_ARG((grad_mc == nullptr) == (param_mc == nullptr), "adam_peer: grad_mc and param_mc must both be set or both be NULL");
  PeerAdamArgs a{};
  for (int r = 0; r < d->world; ++r) {
    TF_CHECK_ARG(grad_peers[r] && param_peers[r] && norm_slot_peers[r], "adam_peer: rank %d has a null peer buffer", r);
    TF_CHECK_ARG(((reinterpret_cast<uintptr_t>(grad_peers[r]) | reinterpret_cast<uintptr_t>(param_peers[r])) & 15) == 0,
                 "adam_peer: rank %d flat buffers must be 16-byte aligned", r);
    a.g[r] = grad_peers[r];
    a.p[r] = param_peers[r];
    a.slots[r] = norm_slot_peers[r];
  }
  TF_CHECK_ARG(((reinterpret_cast<uintptr_t>(grad_mc) | reinterpret_cast<uintptr_t>(param_mc)) & 15) == 0,
               "adam_peer: multicast addresses must be 16-byte aligned");
  const int64_t shard = d->shard_end - d->shard_begin;
  TF_CHECK_ARG(shard == 0 || (mu_shard && nu_shard), "adam_peer: null moment shard");
  TF_CHECK_ARG(((reinterpret_cast<uintptr_t>(mu_shard) | reinterpret_cast<uintptr_t>(nu_shard)) & 15) == 0,
               "adam_peer: moment shards must be 16-byte aligned");
  const int grid = peer_grid(shard);
  TF_CHECK_ARG(scratch && scratch_bytes >= 16 + (int64_t)grid * (int64_t)sizeof(float),
               "adam_peer: scratch too small (%lld bytes; see tensorf_peer_adam_scratch_bytes)", (long long)scratch_bytes);
  TF_CHECK_ARG(ceil_div64(shard, kPeerTile) < ((int64_t)1 << 31), "adam_peer: shard too large");
  a.g_mc = grad_mc;
  a.p_mc = param_mc;
  a.mu = mu_shard;
  a.nu = nu_shard;
  for (int i = 0; i <= nl; ++i) a.off[i] = leaf_offsets[i];
  for (int i = 0; i < nl; ++i) a.neg_lr[i] = neg_lrs[i];
  a.n_real = nl;
  a.n_leaves = nl;
  if (leaf_offsets[nl] < d->total) {  // padding pseudo-leaf: untouched by Adam (see real_end), outside the norm
    a.off[nl + 1] = d->total;
    a.neg_lr[nl] = 0.f;
    a.n_leaves = nl + 1;
  }
  a.rank = d->rank;
  a.world = d->world;
  a.begin = d->shard_begin;
  a.end = d->shard_end;
  a.n_tiles = (int)ceil_div64(shard, kPeerTile);
  a.b1 = d->adam.b1; a.b2 = d->adam.b2;
  a.one_minus_b1 = 1.0f - d->adam.b1; a.one_minus_b2 = 1.0f - d->adam.b2;
  a.eps = d->adam.eps; a.eps_root = d->adam.eps_root;
  a.bc1 = d->adam.bias_correction1; a.bc2 = d->adam.bias_correction2;
  a.lr_decay = d->adam.lr_decay;
  a.ticket = reinterpret_cast<unsigned int*>(scratch);
  a.partial = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scratch) + 16);
  StageTimer t(st, "adam_peer");
  TF_CHECK_CUDA(cudaMemsetAsync(a.ticket, 0, 16, st));
  count_launch();
  const bool mc = grad_mc != nullptr;
#define TF_PEER_LAUNCH(W_)                                                         \
  do {                                                                             \
    if (mc) k_adam_peer<W_, true><<<grid, kPeerThreads, 0, st>>>(a);               \
    else k_adam_peer<W_, false><<<grid, kPeerThreads, 0, st>>>(a);                 \
  } while (0)
  switch (d->world) {
    case 1: TF_PEER_LAUNCH(1); break;
    case 2: TF_PEER_LAUNCH(2); break;
    case 4: TF_PEER_LAUNCH(4); break;
    case 8: TF_PEER_LAUNCH(8); break;
    default: TF_PEER_LAUNCH(0); break;
  }
#undef TF_PEER_LAUNCH
  TF_CHECK_LAUNCH();
  return 0;
}

// Upper bound on the CTAs of the next exchange kernels launched from this thread (0 = default, 4 per SM): an exchange
// that runs on a side stream BESIDE a compute kernel should trickle through a few SMs instead of taking issue slots on all.
static thread_local int g_peer_max_ctas = 0;
int peer_set_max_ctas(int max_ctas) {
  TF_CHECK_ARG(max_ctas >= 0, "peer_set_max_ctas: negative");
  g_peer_max_ctas = max_ctas;
  return 0;
}

int peer_allreduce(cudaStream_t st, int rank, int world, int64_t total, float* const* peers, float* mc,
                   uint32_t* const* signal_peers, uint32_t* local_flags, uint32_t epoch) {
  TF_CHECK_ARG(peers, "peer_allreduce: null argument");
  TF_CHECK_ARG((signal_peers == nullptr) == (local_flags == nullptr),
               "peer_allreduce: signal_peers and local_flags must both be set (in-kernel ordering) or both be NULL");
  TF_CHECK_ARG(world >= 1 && world <= TENSORF_PEER_MAX_WORLD, "peer_allreduce: world=%d outside [1,%d]", world,
               TENSORF_PEER_MAX_WORLD);
  TF_CHECK_ARG(rank >= 0 && rank < world, "peer_allreduce: rank %d outside world of %d", rank, world);
  TF_CHECK_ARG(total >= 0 && total % 4 == 0, "peer_allreduce: total=%lld must be a non-negative multiple of 4", (long long)total);
  PeerReduceArgs a{};
  for (int r = 0; r < world; ++r) {
    TF_CHECK_ARG(peers[r] && (reinterpret_cast<uintptr_t>(peers[r]) & 15) == 0,
                 "peer_allreduce: rank %d buffer must be non-NULL and 16-byte aligned", r);
    a.x[r] = peers[r];
  }
  TF_CHECK_ARG((reinterpret_cast<uintptr_t>(mc) & 15) == 0, "peer_allreduce: multicast address must be 16-byte aligned");
  a.x_mc = mc;
  a.world = world;
  a.rank = rank;
  if (signal_peers) {
    for (int r = 0; r < world; ++r) {
      TF_CHECK_ARG(signal_peers[r], "peer_allreduce: rank %d has a null signal pad", r);
      a.sig[r] = signal_peers[r];
    }
    a.local = local_flags;
    a.epoch = epoch;
  }
  peer_shard(total, rank, world, &a.begin, &a.end);
  const int64_t shard = a.end - a.begin;
  if (shard == 0 && !signal_peers) return 0;  // with in-kernel ordering every rank must still take part
  const int64_t tiles = ceil_div64(shard, kPeerRedTile);
  TF_CHECK_ARG(tiles < ((int64_t)1 << 31), "peer_allreduce: shard too large");
  a.n_tiles = (int)tiles;
  // 4 CTAs of 256 threads x <= 48 registers per SM: every block of the grid is resident at once, so the blocks
  // spinning on the in-kernel gate never keep block 0 (dispatched first in any case) off an SM
  int64_t cap = (int64_t)kSMs * 4;
  if (g_peer_max_ctas > 0) cap = std::min<int64_t>(cap, g_peer_max_ctas);
  const int grid = (int)(tiles < 1 ? 1 : (tiles < cap ? tiles : cap));
  StageTimer t(st, "peer_allreduce");
  const bool use_mc = mc != nullptr;
#define TF_PEER_LAUNCH(W_)                                                        \
  do {                                                                            \
    if (use_mc) k_peer_allreduce<W_, true><<<grid, kPeerThreads, 0, st>>>(a);     \
    else k_peer_allreduce<W_, false><<<grid, kPeerThreads, 0, st>>>(a);           \
  } while (0)
  switch (world) {
    case 1: TF_PEER_LAUNCH(1); break;
    case 2: TF_PEER_LAUNCH(2); break;
    case 4: TF_PEER_LAUNCH(4); break;
    case 8: TF_PEER_LAUNCH(8); break;
    default: TF_PEER_LAUNCH(0); break;
  }
#undef TF_PEER_LAUNCH
  TF_CHECK_LAUNCH();
  return 0;
}

int peer_grad_norm(cudaStream_t st, const float* norm_slots, int world, float* grad_norm) {
  TF_CHECK_ARG(norm_slots && grad_norm, "peer_grad_norm: null argument");
  TF_CHECK_ARG(world >= 1 && world <= TENSORF_PEER_MAX_WORLD, "peer_grad_norm: world=%d outside [1,%d]", world,
               TENSORF_PEER_MAX_WORLD);
  k_peer_grad_norm<<<1, 1, 0, st>>>(norm_slots, world, grad_norm);
  TF_CHECK_LAUNCH();
  return 0;
}

}  // namespace tf
