// K5 sampler: temperature + nucleus (top-p) + seeded categorical draw, or arg-max at temperature 0.
//
// One CTA per sequence, three streaming passes over the row of logits (never sorted, never copied):
//   1. online max / sum-exp                                   -> softmax normaliser
//   2. probability-mass histogram over 2048 log2-spaced bins  -> top-p threshold bin (mass from the top)
//   3. draw u ~ Philox(seed, row, step) in [0, kept mass) and walk the kept tokens in index order
// The kept set is {p >= lower edge of the threshold bin}: the smallest probability kept is within one
// bin width (2^(1/32) - 1 = 2.2 %) of the exact top-p cut-off.
// Reference: vLLM SamplingParams(temperature, top_p=0.95, seed=...) in vllm_generate
// (/root/reference/GRPO/grpo_trainer.py:127) and the T=0 greedy pass of ReMax (remax_trainer.py:167).
#include <curand_kernel.h>

#include "common.cuh"
#include "kernels.h"

namespace nrl {

constexpr int kSampThreads = 1024;
constexpr int kBins = 2048;
constexpr float kBinsPerOctave = 32.f;     // bin = floor(-log2(p) * 32), clamped: covers p down to 2^-64

template <typename T>
struct RowVec;
template <>
struct RowVec<float> {
  static constexpr int N = 4;
  static NRL_DEVICE void load(const float* z, int vec, float (&o)[4]) {
    float4 a = reinterpret_cast<const float4*>(z)[vec];
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  }
};
template <>
struct RowVec<__nv_bfloat16> {
  static constexpr int N = 8;
  static NRL_DEVICE void load(const __nv_bfloat16* z, int vec, float (&o)[8]) {
    uint4 a = reinterpret_cast<const uint4*>(z)[vec];
    uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = unpack_bf16x2(w[j]);
      o[2 * j] = f.x; o[2 * j + 1] = f.y;
    }
  }
};

NRL_DEVICE float block_reduce_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if (lane_id() == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) t = warp_sum(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}
NRL_DEVICE float block_reduce_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if (lane_id() == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -INFINITY;
  if (threadIdx.x < 32) t = warp_max(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}

template <typename T>
__global__ void __launch_bounds__(kSampThreads) sample_top_p_kernel(const T* __restrict__ logits, long row_stride,
                                                                    int V, float inv_temp, float top_p,
                                                                    unsigned long long seed, unsigned long long step,
                                                                    const int* __restrict__ row_ids,
                                                                    const int* __restrict__ row_steps,
                                                                    int* __restrict__ out_tokens) {
  __shared__ float red[32];
  __shared__ float hist[kBins];
  __shared__ float chunk_sum[kSampThreads];
  __shared__ int s_cut_bin;
  __shared__ float s_kept_mass;
  __shared__ int s_token;
  const int row = blockIdx.x;
  const T* z = logits + static_cast<long>(row) * row_stride;
  const int tid = threadIdx.x;
  // Thread t owns vectors t, t+1024, ... (coalesced 16-byte loads).  The categorical walk of pass 3 uses
  // the fixed order "thread-major, then vector order" -- any fixed order is a valid sampling order.
  constexpr int VN = RowVec<T>::N;
  const int nvec = (V + VN - 1) / VN;

  // ---- pass 1: max and sum-exp ----
  float mx = -INFINITY;
  for (int v = tid; v < nvec; v += kSampThreads) {
    float x[VN];
    RowVec<T>::load(z, v, x);
#pragma unroll
    for (int j = 0; j < VN; ++j)
      if (v * VN + j < V) mx = fmaxf(mx, x[j] * inv_temp);
  }
  mx = block_reduce_max(mx, red);
  float se = 0.f;
  for (int v = tid; v < nvec; v += kSampThreads) {
    float x[VN];
    RowVec<T>::load(z, v, x);
#pragma unroll
    for (int j = 0; j < VN; ++j)
      if (v * VN + j < V) se += __expf(x[j] * inv_temp - mx);
  }
  se = block_reduce_sum(se, red);
  const float inv_se = 1.f / se;
  const float log2_inv_se = __log2f(inv_se);

  // ---- pass 2: mass histogram ----
  for (int i = tid; i < kBins; i += kSampThreads) hist[i] = 0.f;
  __syncthreads();
  {
    // run-length aggregation: consecutive tokens of a thread that land in the same bin are summed in
    // registers, so a flat distribution (random-init model) costs one shared atomic per thread, not 150
    int cur_bin = -1;
    float cur_mass = 0.f;
    for (int v = tid; v < nvec; v += kSampThreads) {
      float x[VN];
      RowVec<T>::load(z, v, x);
#pragma unroll
      for (int j = 0; j < VN; ++j)
        if (v * VN + j < V) {
          float zl = x[j] * inv_temp - mx;                        // ln p = zl - ln(se)
          float p = __expf(zl) * inv_se;
          float nlog2 = -(zl * 1.4426950408889634f + log2_inv_se);
          int b = min(kBins - 1, max(0, static_cast<int>(nlog2 * kBinsPerOctave)));
          if (b != cur_bin) {
            if (cur_bin >= 0) atomicAdd(&hist[cur_bin], cur_mass);
            cur_bin = b;
            cur_mass = 0.f;
          }
          cur_mass += p;
        }
    }
    if (cur_bin >= 0) atomicAdd(&hist[cur_bin], cur_mass);
  }
  __syncthreads();
  if (tid < 32) {
    // walk bins from the most probable; find the first bin where cumulative mass reaches top_p
    float cum = 0.f;
    int cut = kBins - 1;
    bool found = false;
    for (int base = 0; base < kBins && !found; base += 32) {
      float incl = hist[base + tid];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float y = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += y;
      }
      unsigned ball = __ballot_sync(0xffffffffu, cum + incl >= top_p);
      if (ball) {
        cut = base + __ffs(ball) - 1;
        found = true;
      } else {
        cum += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    if (tid == 0) s_cut_bin = cut;
  }
  __syncthreads();
  const int cut_bin = s_cut_bin;

  // ---- pass 3: draw and walk ----
  float mine = 0.f;
  for (int v = tid; v < nvec; v += kSampThreads) {
    float x[VN];
    RowVec<T>::load(z, v, x);
#pragma unroll
    for (int j = 0; j < VN; ++j)
      if (v * VN + j < V) {
        float zl = x[j] * inv_temp - mx;
        float nlog2 = -(zl * 1.4426950408889634f + log2_inv_se);
        int b = min(kBins - 1, max(0, static_cast<int>(nlog2 * kBinsPerOctave)));
        if (b <= cut_bin) mine += __expf(zl) * inv_se;
      }
  }
  chunk_sum[tid] = mine;
  __syncthreads();
  if (tid < 32) {
    // warp 0: prefix over the 1024 per-thread masses (32 each), pick the owner thread
    float local[32];
    float tsum = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { local[i] = chunk_sum[tid * 32 + i]; tsum += local[i]; }
    float incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float y = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += y;
    }
    const float total = __shfl_sync(0xffffffffu, incl, 31);
    float u = 0.f;
    if (tid == 0) {
      curandStatePhilox4_32_10_t st;
      curand_init(seed, static_cast<unsigned long long>(row_ids ? row_ids[row] : row),
                  step + (row_steps ? static_cast<unsigned long long>(row_steps[row]) : 0ull), &st);
      u = curand_uniform(&st) * total;              // (0, total]
    }
    u = __shfl_sync(0xffffffffu, u, 0);
    const float excl = incl - tsum;
    unsigned ball = __ballot_sync(0xffffffffu, (incl >= u) && (tsum > 0.f));
    int wl = ball ? (__ffs(ball) - 1) : 31;
    if (tid == wl) {
      float acc = excl;
      int owner = tid * 32 + 31;
      float resid = 0.f;
      bool hit = false;
      for (int i = 0; i < 32; ++i) {
        if (!hit && local[i] > 0.f && acc + local[i] >= u) { owner = tid * 32 + i; resid = u - acc; hit = true; }
        if (!hit) acc += local[i];
      }
      if (!hit) {   // numerical slack: fall back to the last thread with mass
        for (int i = 31; i >= 0; --i) if (local[i] > 0.f) { owner = tid * 32 + i; resid = local[i]; break; }
      }
      s_token = owner;
      s_kept_mass = resid;
    }
  }
  __syncthreads();
  if (tid == s_token) {
    const float resid = s_kept_mass;
    float acc = 0.f;
    int tok = -1, last_kept = -1;
    for (int v = tid; v < nvec && tok < 0; v += kSampThreads) {
      float x[VN];
      RowVec<T>::load(z, v, x);
#pragma unroll
      for (int j = 0; j < VN; ++j)
        if (tok < 0 && v * VN + j < V) {
          float zl = x[j] * inv_temp - mx;
          float nlog2 = -(zl * 1.4426950408889634f + log2_inv_se);
          int b = min(kBins - 1, max(0, static_cast<int>(nlog2 * kBinsPerOctave)));
          if (b <= cut_bin) {
            last_kept = v * VN + j;
            acc += __expf(zl) * inv_se;
            if (acc >= resid) tok = v * VN + j;
          }
        }
    }
    if (tok < 0) tok = last_kept >= 0 ? last_kept : 0;
    out_tokens[row] = tok;
  }
}

template <typename T>
__global__ void __launch_bounds__(kSampThreads) argmax_kernel(const T* __restrict__ logits, long row_stride, int V,
                                                              int* __restrict__ out_tokens) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  const int row = blockIdx.x;
  const T* z = logits + static_cast<long>(row) * row_stride;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  constexpr int VN = RowVec<T>::N;
  const int nvec = (V + VN - 1) / VN;
  for (int v = threadIdx.x; v < nvec; v += kSampThreads) {
    float x[VN];
    RowVec<T>::load(z, v, x);
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      int i = v * VN + j;
      if (i < V && (x[j] > best || (x[j] == best && i < bi))) { best = x[j]; bi = i; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane_id() == 0) { s_val[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = s_val[threadIdx.x];
    bi = s_idx[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (threadIdx.x == 0) out_tokens[row] = bi;
  }
}

}  // namespace nrl

using namespace nrl;

extern "C" cudaError_t nrl_sample(const void* logits, int is_bf16, long row_stride, int rows, int V, float temperature,
                                  float top_p, unsigned long long seed, unsigned long long step, const int* row_ids,
                                  const int* row_steps, int* out_tokens, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  if (temperature == 0.f) {
    if (is_bf16)
      argmax_kernel<__nv_bfloat16><<<rows, kSampThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(logits), row_stride, V, out_tokens);
    else
      argmax_kernel<float><<<rows, kSampThreads, 0, s>>>(static_cast<const float*>(logits), row_stride, V, out_tokens);
  } else {
    float inv_t = 1.f / temperature;
    if (top_p >= 1.f) top_p = 2.f;     // keep everything
    if (is_bf16)
      sample_top_p_kernel<__nv_bfloat16><<<rows, kSampThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(logits), row_stride, V,
                                                                       inv_t, top_p, seed, step, row_ids, row_steps, out_tokens);
    else
      sample_top_p_kernel<float><<<rows, kSampThreads, 0, s>>>(static_cast<const float*>(logits), row_stride, V, inv_t,
                                                               top_p, seed, step, row_ids, row_steps, out_tokens);
  }
  return cudaGetLastError();
}
