// K-AR: data-parallel gradient all-reduce FUSED with the AdamW update, over NVLink symmetric memory.
//
// Baseline being replaced: DDP's bucketed NCCL all-reduce of the bf16 gradients
// followed by a separate torch AdamW step (/root/reference/GRPO/grpo_trainer.py:690-692).
//
// Every rank keeps its flat gradient buffer and flat parameter buffer in symmetric memory (mapped into
// all peers).  One kernel per optimizer step, rank r owns elements [lo, lo+n):
//     g      = sum_p grad_p[i]                    P2P ld.global from the peers' buffers over NVLink, or one
//                                                 multimem.ld_reduce (in-switch NVLS reduction, fp32 accumulate)
//     m,v,p  = AdamW(g * scale)                   moments sharded ZeRO-1 style: only the owner keeps them
//     param_q[i] = p  for every rank q            P2P st.global into the peers' buffers, or one multimem.st
// i.e. reduce-scatter + optimizer + all-gather with the reduced gradient never leaving registers.
// Cross-rank ordering (grads complete before the kernel, params visible after it) is provided by the
// symmetric-memory signal-pad barriers launched on the same stream around the kernel.
#include "common.cuh"
#include "kernels.h"

namespace nrl {

constexpr int kMaxWorld = 16;

struct PeerPtrs {
  const __nv_bfloat16* grad[kMaxWorld];
  __nv_bfloat16* param[kMaxWorld];
};

NRL_DEVICE void unpack8(const uint4& a, float (&o)[8]) {
  uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 f = unpack_bf16x2(w[j]);
    o[2 * j] = f.x;
    o[2 * j + 1] = f.y;
  }
}
NRL_DEVICE uint4 pack8(const float (&o)[8]) {
  return make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
}
NRL_DEVICE uint4 ld_peer_v4(const void* p) {        // peer memory: bypass L1, relaxed system scope
  uint4 r;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
NRL_DEVICE void st_peer_v4(void* p, const uint4& v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// NVLS: one load returns the sum over all ranks' copies (bf16x2 lanes, fp32 accumulation in the switch)
NRL_DEVICE uint4 multimem_ld_reduce_bf16x8(const void* mc_ptr) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(mc_ptr) : "memory");
  return r;
}
NRL_DEVICE void multimem_st_bf16x8(void* mc_ptr, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};"
               ::"l"(mc_ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename MomentT>
NRL_DEVICE void load_moment8(const MomentT* p, float (&o)[8]);
template <>
NRL_DEVICE void load_moment8<float>(const float* p, float (&o)[8]) {
  float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
template <>
NRL_DEVICE void load_moment8<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[8]) {
  unpack8(*reinterpret_cast<const uint4*>(p), o);
}
template <typename MomentT>
NRL_DEVICE void store_moment8(MomentT* p, const float (&o)[8]);
template <>
NRL_DEVICE void store_moment8<float>(float* p, const float (&o)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(o[0], o[1], o[2], o[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(o[4], o[5], o[6], o[7]);
}
template <>
NRL_DEVICE void store_moment8<__nv_bfloat16>(__nv_bfloat16* p, const float (&o)[8]) {
  *reinterpret_cast<uint4*>(p) = pack8(o);
}

NRL_DEVICE void adamw8(float (&p)[8], const float (&g)[8], float (&m)[8], float (&v)[8], const AdamHyper& h) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float gj = g[j] * h.grad_scale;
    m[j] = h.beta1 * m[j] + (1.f - h.beta1) * gj;
    v[j] = h.beta2 * v[j] + (1.f - h.beta2) * gj * gj;
    p[j] = p[j] * (1.f - h.lr * h.wd) - h.step_size * (m[j] / (sqrtf(v[j] * h.inv_bc2) + h.eps));
  }
}

// P2P variant.  lo / n in elements (multiples of 8); moments are indexed from 0 (the owner's shard).
template <typename MomentT, int WORLD>
__global__ void __launch_bounds__(256) allreduce_adam_p2p_kernel(PeerPtrs ptrs, MomentT* __restrict__ m,
                                                                 MomentT* __restrict__ v, float* __restrict__ master,
                                                                 long lo, long n, int rank, AdamHyper h) {
  const long nvec = n / 8;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long e = lo + i * 8;
    uint4 raw[WORLD];
#pragma unroll
    for (int p = 0; p < WORLD; ++p) {
      const int q = (rank + p) % WORLD;                 // stagger peers so links are loaded evenly
      raw[p] = ld_peer_v4(ptrs.grad[q] + e);
    }
    float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int p = 0; p < WORLD; ++p) {
      float t[8];
      unpack8(raw[p], t);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += t[j];
    }
    float p8[8], m8[8], v8[8];
    // fp32 master copy of the owned shard (ZeRO-1): the bf16 parameter is only its rounded broadcast, so updates
    // far below one bf16 ulp (lr 6e-6 on |w| ~ 0.02) accumulate instead of rounding away
    if (master != nullptr) load_moment8<float>(master + i * 8, p8);
    else unpack8(*reinterpret_cast<const uint4*>(ptrs.param[rank] + e), p8);
    load_moment8<MomentT>(m + i * 8, m8);
    load_moment8<MomentT>(v + i * 8, v8);
    adamw8(p8, g, m8, v8, h);
    store_moment8<MomentT>(m + i * 8, m8);
    store_moment8<MomentT>(v + i * 8, v8);
    if (master != nullptr) store_moment8<float>(master + i * 8, p8);
    const uint4 out = pack8(p8);
#pragma unroll
    for (int p = 0; p < WORLD; ++p) {
      const int q = (rank + p) % WORLD;
      st_peer_v4(ptrs.param[q] + e, out);
    }
  }
}

// NVLS multicast variant: grad_mc / param_mc are the multicast addresses of the symmetric buffers.
template <typename MomentT>
__global__ void __launch_bounds__(256) allreduce_adam_mc_kernel(const __nv_bfloat16* grad_mc, __nv_bfloat16* param_mc,
                                                                const __nv_bfloat16* __restrict__ param_local,
                                                                MomentT* __restrict__ m, MomentT* __restrict__ v,
                                                                float* __restrict__ master, long lo, long n, AdamHyper h) {
  const long nvec = n / 8;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long e = lo + i * 8;
    float g[8], p8[8], m8[8], v8[8];
    unpack8(multimem_ld_reduce_bf16x8(grad_mc + e), g);
    if (master != nullptr) load_moment8<float>(master + i * 8, p8);
    else unpack8(*reinterpret_cast<const uint4*>(param_local + e), p8);
    load_moment8<MomentT>(m + i * 8, m8);
    load_moment8<MomentT>(v + i * 8, v8);
    adamw8(p8, g, m8, v8, h);
    store_moment8<MomentT>(m + i * 8, m8);
    store_moment8<MomentT>(v + i * 8, v8);
    if (master != nullptr) store_moment8<float>(master + i * 8, p8);
    multimem_st_bf16x8(param_mc + e, pack8(p8));
  }
}

// Plain fused all-reduce (sum, in place, every rank ends with the total) -- used for the A/B against NCCL
// and by ops that need the reduced tensor itself.  Two-shot: each rank reduces its slice and writes it to all.
template <int WORLD>
__global__ void __launch_bounds__(256) allreduce_p2p_kernel(PeerPtrs ptrs, long lo, long n, int rank, float scale) {
  const long nvec = n / 8;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long e = lo + i * 8;
    uint4 raw[WORLD];
#pragma unroll
    for (int p = 0; p < WORLD; ++p) raw[p] = ld_peer_v4(ptrs.grad[(rank + p) % WORLD] + e);
    float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int p = 0; p < WORLD; ++p) {
      float t[8];
      unpack8(raw[p], t);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += t[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= scale;
    const uint4 out = pack8(g);
#pragma unroll
    for (int p = 0; p < WORLD; ++p) st_peer_v4(ptrs.param[(rank + p) % WORLD] + e, out);
  }
}

template <typename MomentT>
static cudaError_t launch_p2p(const PeerPtrs& ptrs, void* m, void* v, float* master, long lo, long n, int world, int rank,
                              const AdamHyper& h, int blocks, cudaStream_t s) {
  auto M = static_cast<MomentT*>(m);
  auto V = static_cast<MomentT*>(v);
  switch (world) {
    case 1: allreduce_adam_p2p_kernel<MomentT, 1><<<blocks, 256, 0, s>>>(ptrs, M, V, master, lo, n, rank, h); break;
    case 2: allreduce_adam_p2p_kernel<MomentT, 2><<<blocks, 256, 0, s>>>(ptrs, M, V, master, lo, n, rank, h); break;
    case 4: allreduce_adam_p2p_kernel<MomentT, 4><<<blocks, 256, 0, s>>>(ptrs, M, V, master, lo, n, rank, h); break;
    case 8: allreduce_adam_p2p_kernel<MomentT, 8><<<blocks, 256, 0, s>>>(ptrs, M, V, master, lo, n, rank, h); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace nrl

using namespace nrl;

extern "C" cudaError_t nrl_allreduce_adam(const void* const* grad_ptrs, void* const* param_ptrs, const void* grad_mc,
                                          void* param_mc, void* m, void* v, float* master, long lo, long n, int world,
                                          int rank, int moments_bf16, int use_multicast, AdamHyper h, int max_blocks,
                                          cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  if (n % 8 != 0 || lo % 8 != 0 || world > kMaxWorld) return cudaErrorInvalidValue;
  long blocks = (n / 8 + 255) / 256;
  if (blocks > max_blocks) blocks = max_blocks;
  if (use_multicast) {
    auto G = static_cast<const __nv_bfloat16*>(grad_mc);
    auto P = static_cast<__nv_bfloat16*>(param_mc);
    auto PL = static_cast<const __nv_bfloat16*>(param_ptrs[rank]);
    if (moments_bf16)
      allreduce_adam_mc_kernel<__nv_bfloat16><<<static_cast<int>(blocks), 256, 0, s>>>(
          G, P, PL, static_cast<__nv_bfloat16*>(m), static_cast<__nv_bfloat16*>(v), master, lo, n, h);
    else
      allreduce_adam_mc_kernel<float><<<static_cast<int>(blocks), 256, 0, s>>>(G, P, PL, static_cast<float*>(m),
                                                                               static_cast<float*>(v), master, lo, n, h);
    return cudaGetLastError();
  }
  PeerPtrs ptrs;
  for (int i = 0; i < world; ++i) {
    ptrs.grad[i] = static_cast<const __nv_bfloat16*>(grad_ptrs[i]);
    ptrs.param[i] = static_cast<__nv_bfloat16*>(param_ptrs[i]);
  }
  return moments_bf16 ? launch_p2p<__nv_bfloat16>(ptrs, m, v, master, lo, n, world, rank, h, static_cast<int>(blocks), s)
                      : launch_p2p<float>(ptrs, m, v, master, lo, n, world, rank, h, static_cast<int>(blocks), s);
}

extern "C" cudaError_t nrl_allreduce_sum(void* const* buf_ptrs, long lo, long n, int world, int rank, float scale,
                                         int max_blocks, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  if (n % 8 != 0 || lo % 8 != 0 || world > kMaxWorld) return cudaErrorInvalidValue;
  long blocks = (n / 8 + 255) / 256;
  if (blocks > max_blocks) blocks = max_blocks;
  PeerPtrs ptrs;
  for (int i = 0; i < world; ++i) {
    ptrs.grad[i] = static_cast<const __nv_bfloat16*>(buf_ptrs[i]);
    ptrs.param[i] = static_cast<__nv_bfloat16*>(buf_ptrs[i]);
  }
  const int b = static_cast<int>(blocks);
  switch (world) {
    case 1: allreduce_p2p_kernel<1><<<b, 256, 0, s>>>(ptrs, lo, n, rank, scale); break;
    case 2: allreduce_p2p_kernel<2><<<b, 256, 0, s>>>(ptrs, lo, n, rank, scale); break;
    case 4: allreduce_p2p_kernel<4><<<b, 256, 0, s>>>(ptrs, lo, n, rank, scale); break;
    case 8: allreduce_p2p_kernel<8><<<b, 256, 0, s>>>(ptrs, lo, n, rank, scale); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
