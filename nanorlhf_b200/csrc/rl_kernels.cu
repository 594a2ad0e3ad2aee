// RL-specific fused kernels:
//   K-GAE  gae_scan          reverse linear-recurrence scan (GAE / discounted suffix sums), warp per sequence
//   K-LOSS policy_loss       PPO-clip (+ GRPO k3-KL) forward + gradient + all statistics in ONE pass
//          value_loss        clipped value loss forward + gradient
//   K18    adamw_flat        AdamW over flat bf16 parameter / gradient buffers with fp32 or bf16 moments
// Reference: the python time-step loop (/root/reference/GRPO/grpo_trainer.py:611-617, PPO/ppo_trainer.py:
// 688-697), ~15 elementwise launches per micro-batch (:662-688) and torch AdamW (:692).
#include "common.cuh"
#include "kernels.h"

namespace nrl {

// ---- K-GAE ------------------------------------------------------------------------------------
// A_t = delta_t + c * A_{t+1},  delta_t = r_t + gamma * V_{t+1} - V_t  (V == nullptr -> delta = r, c = gamma*lam)
// One warp per row; 32 time-steps per iteration with a weighted reverse Hillis-Steele scan.
__global__ void gae_scan_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                float* __restrict__ adv, float* __restrict__ returns, int B, int T, float gamma,
                                float lam) {
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (row >= B) return;
  const int lane = threadIdx.x & 31;
  const float c = gamma * lam;
  // c^(2^k) for the scan, c^(32-lane) for the carry
  float cpow[5];
  cpow[0] = c;
#pragma unroll
  for (int k = 1; k < 5; ++k) cpow[k] = cpow[k - 1] * cpow[k - 1];
  const float* r = rewards + static_cast<size_t>(row) * T;
  const float* v = values ? values + static_cast<size_t>(row) * T : nullptr;
  float carry = 0.f;
  const int nchunks = (T + 31) / 32;
  for (int ch = nchunks - 1; ch >= 0; --ch) {
    const int t = ch * 32 + lane;
    float delta = 0.f, vt = 0.f;
    if (t < T) {
      delta = r[t];
      if (v) {
        vt = v[t];
        float vn = (t + 1 < T) ? v[t + 1] : 0.f;
        delta += gamma * vn - vt;
      }
    }
    float x = delta;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      float y = __shfl_down_sync(0xffffffffu, x, 1 << k);
      if (lane + (1 << k) < 32) x = fmaf(cpow[k], y, x);
    }
    // x now = sum_{j>=0, lane+j<32} c^j delta_{t+j}; add c^(32-lane) * carry
    float cw = powf(c, static_cast<float>(32 - lane));
    if (c == 1.f) cw = 1.f;
    x = fmaf(cw, carry, x);
    if (t < T) {
      adv[static_cast<size_t>(row) * T + t] = x;
      if (returns) returns[static_cast<size_t>(row) * T + t] = x + vt;
    }
    carry = __shfl_sync(0xffffffffu, x, 0);
  }
}

// ---- K-LOSS -----------------------------------------------------------------------------------
// acc[0]=sum(loss*m) [1]=sum(m) [2]=sum(clipped*m) [3]=sum(diff^2*m) [4]=sum(diff^2) [5]=sum(ratio)
// [6]=sum(ratio*m) [7]=sum(k) [8]=sum(k*m) [9]=n
// grad_unnorm[i] = d(per-token loss)/d(new_logp) * m   (caller divides by sum(m))
__global__ void policy_loss_kernel(const float* __restrict__ new_lp, const float* __restrict__ old_lp,
                                   const float* __restrict__ adv, const uint8_t* __restrict__ mask,
                                   const float* __restrict__ ref_lp, float cliprange, float kl_coef, long n,
                                   float* __restrict__ grad_unnorm, float* __restrict__ acc) {
  float s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float nl = new_lp[i], ol = old_lp[i], a = adv[i];
    const float m = mask[i] ? 1.f : 0.f;
    const float diff = nl - ol;
    const float ratio = __expf(diff);
    const float l1 = -a * ratio;
    const float rc = fminf(fmaxf(ratio, 1.f - cliprange), 1.f + cliprange);
    const float l2 = -a * rc;
    float per = fmaxf(l1, l2);
    // d max(l1,l2)/d nl: l1 branch -> -a*ratio ; l2 branch -> 0 when clipped, -a*ratio when not clipped
    const bool clipped_active = (l2 > l1);
    float g = clipped_active ? ((rc == ratio) ? -a * ratio : 0.f) : -a * ratio;
    float k = 0.f;
    if (ref_lp != nullptr) {
      k = nl - ref_lp[i];
      const float e = __expf(-k);
      per += kl_coef * (e + k - 1.f);
      g += kl_coef * (1.f - e);
    }
    grad_unnorm[i] = g * m;
    s[0] += per * m; s[1] += m; s[2] += clipped_active ? m : 0.f;
    s[3] += diff * diff * m; s[4] += diff * diff; s[5] += ratio; s[6] += ratio * m; s[7] += k; s[8] += k * m;
  }
#pragma unroll
  for (int j = 0; j < 9; ++j) s[j] = warp_sum(s[j]);
  __shared__ float red[9][8];
  const int w = threadIdx.x >> 5;
  if (lane_id() == 0)
#pragma unroll
    for (int j = 0; j < 9; ++j) red[j][w] = s[j];
  __syncthreads();
  if (threadIdx.x < 9) {
    float t = 0.f;
    for (int q = 0; q < (blockDim.x >> 5); ++q) t += red[threadIdx.x][q];
    atomicAdd(&acc[threadIdx.x], t);
  }
}

// acc[0]=sum(max(l1,l2)*m) [1]=sum(m) [2]=sum((l2>l1)*m); grad_unnorm = d(0.5*max)/d vpred * m
__global__ void value_loss_kernel(const float* __restrict__ vpred, const float* __restrict__ vold,
                                  const float* __restrict__ ret, const uint8_t* __restrict__ mask, float clip, long n,
                                  float* __restrict__ grad_unnorm, float* __restrict__ acc) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float v = vpred[i], vo = vold[i], R = ret[i];
    const float m = mask[i] ? 1.f : 0.f;
    const float vc = fmaxf(fminf(v, vo + clip), vo - clip);
    const float l1 = (v - R) * (v - R), l2 = (vc - R) * (vc - R);
    const bool use2 = l2 > l1;
    float g = use2 ? ((vc == v) ? (v - R) : 0.f) : (v - R);
    grad_unnorm[i] = g * m;     // d(0.5 * l)/dv
    s0 += fmaxf(l1, l2) * m; s1 += m; s2 += use2 ? m : 0.f;
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane_id() == 0) {
    atomicAdd(&acc[0], s0);
    atomicAdd(&acc[1], s1);
    atomicAdd(&acc[2], s2);
  }
}

// ---- AdamW over flat buffers --------------------------------------------------------------------
template <typename MomentT>
struct MomentIO;
template <>
struct MomentIO<float> {
  static NRL_DEVICE void load8(const float* p, float (&o)[8]) {
    float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
  static NRL_DEVICE void store8(float* p, const float (&o)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(o[4], o[5], o[6], o[7]);
  }
};
template <>
struct MomentIO<__nv_bfloat16> {
  static NRL_DEVICE void load8(const __nv_bfloat16* p, float (&o)[8]) {
    uint4 a = *reinterpret_cast<const uint4*>(p);
    uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = unpack_bf16x2(w[j]);
      o[2 * j] = f.x; o[2 * j + 1] = f.y;
    }
  }
  static NRL_DEVICE void store8(__nv_bfloat16* p, const float (&o)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                              pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
  }
};

NRL_DEVICE void adamw_update8(float (&p)[8], const float (&g)[8], float (&m)[8], float (&v)[8], const AdamHyper& h) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float gj = g[j] * h.grad_scale;
    m[j] = h.beta1 * m[j] + (1.f - h.beta1) * gj;
    v[j] = h.beta2 * v[j] + (1.f - h.beta2) * gj * gj;
    const float denom = sqrtf(v[j] * h.inv_bc2) + h.eps;
    p[j] = p[j] * (1.f - h.lr * h.wd) - h.step_size * (m[j] / denom);
  }
}

template <typename MomentT>
__global__ void adamw_flat_kernel(__nv_bfloat16* __restrict__ param, const __nv_bfloat16* __restrict__ grad,
                                  MomentT* __restrict__ m, MomentT* __restrict__ v, float* __restrict__ master, long n,
                                  AdamHyper h) {
  const long nvec = n / 8;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    float p8[8], g8[8], m8[8], v8[8];
    if (master != nullptr) MomentIO<float>::load8(master + i * 8, p8);      // fp32 master weights (see comm.cu)
    else MomentIO<__nv_bfloat16>::load8(param + i * 8, p8);
    MomentIO<__nv_bfloat16>::load8(grad + i * 8, g8);
    MomentIO<MomentT>::load8(m + i * 8, m8);
    MomentIO<MomentT>::load8(v + i * 8, v8);
    adamw_update8(p8, g8, m8, v8, h);
    MomentIO<__nv_bfloat16>::store8(param + i * 8, p8);
    if (master != nullptr) MomentIO<float>::store8(master + i * 8, p8);
    MomentIO<MomentT>::store8(m + i * 8, m8);
    MomentIO<MomentT>::store8(v + i * 8, v8);
  }
}

}  // namespace nrl

using namespace nrl;

extern "C" cudaError_t nrl_gae_scan(const float* rewards, const float* values, float* adv, float* returns, int B,
                                    int T, float gamma, float lam, cudaStream_t s) {
  if (B == 0 || T == 0) return cudaSuccess;
  const int warps = 4;
  gae_scan_kernel<<<(B + warps - 1) / warps, warps * 32, 0, s>>>(rewards, values, adv, returns, B, T, gamma, lam);
  return cudaGetLastError();
}

extern "C" cudaError_t nrl_policy_loss(const float* new_lp, const float* old_lp, const float* adv,
                                       const uint8_t* mask, const float* ref_lp, float cliprange, float kl_coef,
                                       long n, float* grad_unnorm, float* acc, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  long blocks = (n + 255) / 256;
  if (blocks > 296) blocks = 296;
  policy_loss_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(new_lp, old_lp, adv, mask, ref_lp, cliprange, kl_coef,
                                                              n, grad_unnorm, acc);
  return cudaGetLastError();
}

extern "C" cudaError_t nrl_value_loss(const float* vpred, const float* vold, const float* ret, const uint8_t* mask,
                                      float clip, long n, float* grad_unnorm, float* acc, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  long blocks = (n + 255) / 256;
  if (blocks > 296) blocks = 296;
  value_loss_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(vpred, vold, ret, mask, clip, n, grad_unnorm, acc);
  return cudaGetLastError();
}

extern "C" cudaError_t nrl_adamw_flat(void* param, const void* grad, void* m, void* v, float* master, long n,
                                      int moments_bf16, AdamHyper h, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  if (n % 8 != 0) return cudaErrorInvalidValue;
  long blocks = (n / 8 + 255) / 256;
  if (blocks > 148L * 8) blocks = 148L * 8;
  auto P = static_cast<__nv_bfloat16*>(param);
  auto G = static_cast<const __nv_bfloat16*>(grad);
  if (moments_bf16)
    adamw_flat_kernel<__nv_bfloat16><<<static_cast<int>(blocks), 256, 0, s>>>(
        P, G, static_cast<__nv_bfloat16*>(m), static_cast<__nv_bfloat16*>(v), master, n, h);
  else
    adamw_flat_kernel<float><<<static_cast<int>(blocks), 256, 0, s>>>(P, G, static_cast<float*>(m),
                                                                       static_cast<float*>(v), master, n, h);
  return cudaGetLastError();
}
