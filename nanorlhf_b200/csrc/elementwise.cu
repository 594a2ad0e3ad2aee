// Bandwidth-bound fused elementwise / row-reduce kernels (K4): RMSNorm (+residual), RoPE, SwiGLU.
// All 128-bit vectorised, one pass over the data, fp32 math, bf16 I/O.
// Reference equivalents: HF eager RMSNorm / rotate_half RoPE / SiLU*mul in the training path and
// vLLM's fused_add_rms_norm / rotary_embedding / silu_and_mul in the rollout path
// (SURVEY.md section 2.5 K4).
#include "common.cuh"
#include "kernels.h"

namespace nrl {

// ---- RMSNorm ---------------------------------------------------------------------------------
// y = x * rsqrt(mean(x^2) + eps) * w ; optionally x <- x + residual first (writes the sum back to
// `residual_out`).  One CTA (128 threads) per row, row cached in registers (d <= 8192).
template <bool kAddResidual>
__global__ void __launch_bounds__(128) rmsnorm_kernel(const __nv_bfloat16* __restrict__ x,
                                                      const __nv_bfloat16* __restrict__ residual,
                                                      const __nv_bfloat16* __restrict__ w,
                                                      __nv_bfloat16* __restrict__ y,
                                                      __nv_bfloat16* __restrict__ residual_out,
                                                      float* __restrict__ rstd_out, int d, float eps) {
  const int row = blockIdx.x;
  const int nvec = d / 8;
  const uint4* xv = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * d);
  const uint4* rv = kAddResidual ? reinterpret_cast<const uint4*>(residual + static_cast<size_t>(row) * d) : nullptr;
  constexpr int kMaxVec = 8;   // 8 vec * 128 threads * 8 elems = 8192
  float vals[kMaxVec][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    int idx = threadIdx.x + i * 128;
    if (idx < nvec) {
      uint4 a = xv[idx];
      uint32_t aw[4] = {a.x, a.y, a.z, a.w};
      uint32_t bw[4] = {0, 0, 0, 0};
      if (kAddResidual) {
        uint4 b = rv[idx];
        bw[0] = b.x; bw[1] = b.y; bw[2] = b.z; bw[3] = b.w;
      }
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = unpack_bf16x2(aw[j]);
        if (kAddResidual) {
          float2 g = unpack_bf16x2(bw[j]);
          f.x += g.x; f.y += g.y;
          ow[j] = pack_bf16x2(f.x, f.y);
          f = unpack_bf16x2(ow[j]);        // the normalised value is computed from the rounded sum
        }
        vals[i][2 * j] = f.x;
        vals[i][2 * j + 1] = f.y;
        ss += f.x * f.x + f.y * f.y;
      }
      if (kAddResidual)
        reinterpret_cast<uint4*>(residual_out + static_cast<size_t>(row) * d)[idx] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
  __shared__ float red[4];
  ss = warp_sum(ss);
  if (lane_id() == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  ss = red[0] + red[1] + red[2] + red[3];
  const float rstd = rsqrtf(ss / d + eps);
  if (rstd_out != nullptr && threadIdx.x == 0) rstd_out[row] = rstd;
  const uint4* wv = reinterpret_cast<const uint4*>(w);
  uint4* yv = reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * d);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    int idx = threadIdx.x + i * 128;
    if (idx < nvec) {
      uint4 ww = wv[idx];
      uint32_t wr[4] = {ww.x, ww.y, ww.z, ww.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 g = unpack_bf16x2(wr[j]);
        o[j] = pack_bf16x2(vals[i][2 * j] * rstd * g.x, vals[i][2 * j + 1] * rstd * g.y);
      }
      yv[idx] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// dx = rstd * (g*w - xhat * mean(g*w*xhat)) ; xhat = x * rstd
__global__ void __launch_bounds__(128) rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                          const __nv_bfloat16* __restrict__ w,
                                                          const __nv_bfloat16* __restrict__ gy,
                                                          const float* __restrict__ rstd_in,
                                                          __nv_bfloat16* __restrict__ gx, int d) {
  const int row = blockIdx.x;
  const int nvec = d / 8;
  const float rstd = rstd_in[row];
  const uint4* xv = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * d);
  const uint4* gv = reinterpret_cast<const uint4*>(gy + static_cast<size_t>(row) * d);
  const uint4* wv = reinterpret_cast<const uint4*>(w);
  constexpr int kMaxVec = 8;
  float xh[kMaxVec][8], gw[kMaxVec][8];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    int idx = threadIdx.x + i * 128;
    if (idx < nvec) {
      uint4 a = xv[idx], g = gv[idx], ww = wv[idx];
      uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gr[4] = {g.x, g.y, g.z, g.w}, wr[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 fx = unpack_bf16x2(aw[j]), fg = unpack_bf16x2(gr[j]), fw = unpack_bf16x2(wr[j]);
        xh[i][2 * j] = fx.x * rstd;
        xh[i][2 * j + 1] = fx.y * rstd;
        gw[i][2 * j] = fg.x * fw.x;
        gw[i][2 * j + 1] = fg.y * fw.y;
        dot += gw[i][2 * j] * xh[i][2 * j] + gw[i][2 * j + 1] * xh[i][2 * j + 1];
      }
    }
  }
  __shared__ float red[4];
  dot = warp_sum(dot);
  if (lane_id() == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  const float mean_dot = (red[0] + red[1] + red[2] + red[3]) / d;
  uint4* ov = reinterpret_cast<uint4*>(gx + static_cast<size_t>(row) * d);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    int idx = threadIdx.x + i * 128;
    if (idx < nvec) {
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[j] = pack_bf16x2(rstd * (gw[i][2 * j] - xh[i][2 * j] * mean_dot),
                           rstd * (gw[i][2 * j + 1] - xh[i][2 * j + 1] * mean_dot));
      ov[idx] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ---- RoPE (rotate-half convention) ------------------------------------------------------------
// x: [T, H, D] with an arbitrary row stride (so q/k can be views of a fused qkv buffer);
// cos/sin: [T, D/2] fp32.  sin_sign = -1 gives the backward (inverse rotation).
__global__ void rope_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                            const float* __restrict__ cos_t, const float* __restrict__ sin_t, int T, int H, int D,
                            long x_stride_t, long y_stride_t, float sin_sign) {
  const int half = D / 2;
  const int vec_per_head = half / 8;                 // 8 pairs per thread-iteration
  const long total = static_cast<long>(T) * H * vec_per_head;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    int v = i % vec_per_head;
    long th = i / vec_per_head;
    int h = th % H;
    long t = th / H;
    const __nv_bfloat16* xp = x + t * x_stride_t + static_cast<long>(h) * D + v * 8;
    __nv_bfloat16* yp = y + t * y_stride_t + static_cast<long>(h) * D + v * 8;
    uint4 lo = *reinterpret_cast<const uint4*>(xp);
    uint4 hi = *reinterpret_cast<const uint4*>(xp + half);
    const float4* cp = reinterpret_cast<const float4*>(cos_t + t * half + v * 8);
    const float4* sp = reinterpret_cast<const float4*>(sin_t + t * half + v * 8);
    float4 c0 = cp[0], c1 = cp[1], s0 = sp[0], s1 = sp[1];
    float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    uint32_t lw[4] = {lo.x, lo.y, lo.z, lo.w}, hw[4] = {hi.x, hi.y, hi.z, hi.w}, ol[4], oh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 a = unpack_bf16x2(lw[j]), b = unpack_bf16x2(hw[j]);
      float sa = s[2 * j] * sin_sign, sb = s[2 * j + 1] * sin_sign;
      ol[j] = pack_bf16x2(a.x * c[2 * j] - b.x * sa, a.y * c[2 * j + 1] - b.y * sb);
      oh[j] = pack_bf16x2(b.x * c[2 * j] + a.x * sa, b.y * c[2 * j + 1] + a.y * sb);
    }
    *reinterpret_cast<uint4*>(yp) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
    *reinterpret_cast<uint4*>(yp + half) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
  }
}

// ---- SwiGLU -----------------------------------------------------------------------------------
// gate_up: [T, 2F] = [gate | up]; out[T, F] = silu(gate) * up
__global__ void swiglu_kernel(const __nv_bfloat16* __restrict__ gate, const __nv_bfloat16* __restrict__ up, long in_stride,
                              __nv_bfloat16* __restrict__ out, long T, int F) {
  const int vecs = F / 8;
  const long total = T * vecs;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    long t = i / vecs;
    int v = i % vecs;
    uint4 g = *reinterpret_cast<const uint4*>(gate + t * in_stride + v * 8);
    uint4 u = *reinterpret_cast<const uint4*>(up + t * in_stride + v * 8);
    uint32_t gw[4] = {g.x, g.y, g.z, g.w}, uw[4] = {u.x, u.y, u.z, u.w}, o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 a = unpack_bf16x2(gw[j]), b = unpack_bf16x2(uw[j]);
      float r0 = a.x / (1.f + __expf(-a.x)) * b.x;
      float r1 = a.y / (1.f + __expf(-a.y)) * b.y;
      o[j] = pack_bf16x2(r0, r1);
    }
    *reinterpret_cast<uint4*>(out + t * F + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// d gate = g * up * (sig + gate*sig*(1-sig)) ; d up = g * silu(gate)
__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ gate, const __nv_bfloat16* __restrict__ up, long in_stride,
                                  const __nv_bfloat16* __restrict__ gout, __nv_bfloat16* __restrict__ dgate,
                                  __nv_bfloat16* __restrict__ dup, long out_stride, long T, int F) {
  const int vecs = F / 8;
  const long total = T * vecs;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    long t = i / vecs;
    int v = i % vecs;
    uint4 g = *reinterpret_cast<const uint4*>(gate + t * in_stride + v * 8);
    uint4 u = *reinterpret_cast<const uint4*>(up + t * in_stride + v * 8);
    uint4 go = *reinterpret_cast<const uint4*>(gout + t * F + v * 8);
    uint32_t gw[4] = {g.x, g.y, g.z, g.w}, uw[4] = {u.x, u.y, u.z, u.w}, ow[4] = {go.x, go.y, go.z, go.w};
    uint32_t dg[4], du[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 a = unpack_bf16x2(gw[j]), b = unpack_bf16x2(uw[j]), o = unpack_bf16x2(ow[j]);
      float s0 = 1.f / (1.f + __expf(-a.x)), s1 = 1.f / (1.f + __expf(-a.y));
      dg[j] = pack_bf16x2(o.x * b.x * (s0 + a.x * s0 * (1.f - s0)), o.y * b.y * (s1 + a.y * s1 * (1.f - s1)));
      du[j] = pack_bf16x2(o.x * a.x * s0, o.y * a.y * s1);
    }
    *reinterpret_cast<uint4*>(dgate + t * out_stride + v * 8) = make_uint4(dg[0], dg[1], dg[2], dg[3]);
    *reinterpret_cast<uint4*>(dup + t * out_stride + v * 8) = make_uint4(du[0], du[1], du[2], du[3]);
  }
}

static inline int grid_for(long total, int threads) {
  long b = (total + threads - 1) / threads;
  long cap = 148L * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace nrl

using namespace nrl;

extern "C" cudaError_t nrl_rmsnorm(const void* x, const void* residual, const void* w, void* y, void* residual_out,
                                   float* rstd, int rows, int d, float eps, cudaStream_t s) {
  if (d % 8 != 0 || d > 8192) return cudaErrorInvalidValue;
  if (rows == 0) return cudaSuccess;
  auto X = static_cast<const __nv_bfloat16*>(x);
  auto R = static_cast<const __nv_bfloat16*>(residual);
  auto W = static_cast<const __nv_bfloat16*>(w);
  auto Y = static_cast<__nv_bfloat16*>(y);
  auto RO = static_cast<__nv_bfloat16*>(residual_out);
  if (residual != nullptr)
    rmsnorm_kernel<true><<<rows, 128, 0, s>>>(X, R, W, Y, RO, rstd, d, eps);
  else
    rmsnorm_kernel<false><<<rows, 128, 0, s>>>(X, nullptr, W, Y, nullptr, rstd, d, eps);
  return cudaGetLastError();
}

// y = LayerNorm(x + residual) * w + b  (inference path of the reward model: the add and the norm in one pass).
// One warp per row, d <= 2048, row cached in registers, two-pass (mean, then centred variance) in fp32.
__global__ void __launch_bounds__(256) add_layernorm_kernel(const __nv_bfloat16* __restrict__ x,
                                                            const __nv_bfloat16* __restrict__ residual,
                                                            const __nv_bfloat16* __restrict__ w,
                                                            const __nv_bfloat16* __restrict__ b,
                                                            __nv_bfloat16* __restrict__ y, int rows, int d, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31, nvec = d / 8;
  const uint4* xv = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * d);
  const uint4* rv = residual ? reinterpret_cast<const uint4*>(residual + static_cast<size_t>(row) * d) : nullptr;
  constexpr int kMaxVec = 8;                      // 8 vec * 32 lanes * 8 elems = 2048
  float vals[kMaxVec][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      const uint4 a = xv[idx];
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
      uint32_t bw[4] = {0, 0, 0, 0};
      if (rv != nullptr) {
        const uint4 r4 = rv[idx];
        bw[0] = r4.x; bw[1] = r4.y; bw[2] = r4.z; bw[3] = r4.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = unpack_bf16x2(aw[j]);
        if (rv != nullptr) {
          const float2 g = unpack_bf16x2(bw[j]);
          f = unpack_bf16x2(pack_bf16x2(f.x + g.x, f.y + g.y));      // the sum is a bf16 tensor in the eager model
        }
        vals[i][2 * j] = f.x;
        vals[i][2 * j + 1] = f.y;
        sum += f.x + f.y;
      }
    }
  }
  const float mean = warp_sum(sum) / d;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (lane + i * 32 < nvec)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float c = vals[i][j] - mean;
        var += c * c;
      }
  const float rstd = rsqrtf(warp_sum(var) / d + eps);
  const uint4* wv = reinterpret_cast<const uint4*>(w);
  const uint4* bv = reinterpret_cast<const uint4*>(b);
  uint4* yv = reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * d);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      const uint4 w4 = wv[idx], b4 = bv[idx];
      const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 wf = unpack_bf16x2(ww[j]), bf = unpack_bf16x2(bb[j]);
        ow[j] = pack_bf16x2((vals[i][2 * j] - mean) * rstd * wf.x + bf.x, (vals[i][2 * j + 1] - mean) * rstd * wf.y + bf.y);
      }
      yv[idx] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
}

extern "C" cudaError_t nrl_add_layernorm(const void* x, const void* residual, const void* w, const void* b, void* y, int rows,
                                         int d, float eps, cudaStream_t s) {
  if (d % 8 != 0 || d > 2048) return cudaErrorInvalidValue;
  if (rows == 0) return cudaSuccess;
  add_layernorm_kernel<<<(rows + 7) / 8, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x),
                                                     static_cast<const __nv_bfloat16*>(residual),
                                                     static_cast<const __nv_bfloat16*>(w), static_cast<const __nv_bfloat16*>(b),
                                                     static_cast<__nv_bfloat16*>(y), rows, d, eps);
  return cudaGetLastError();
}

extern "C" cudaError_t nrl_rmsnorm_bwd(const void* x, const void* w, const void* gy, const float* rstd, void* gx,
                                       int rows, int d, cudaStream_t s) {
  if (d % 8 != 0 || d > 8192) return cudaErrorInvalidValue;
  if (rows == 0) return cudaSuccess;
  rmsnorm_bwd_kernel<<<rows, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w),
                                          static_cast<const __nv_bfloat16*>(gy), rstd,
                                          static_cast<__nv_bfloat16*>(gx), d);
  return cudaGetLastError();
}

extern "C" cudaError_t nrl_rope(const void* x, void* y, const float* cos_t, const float* sin_t, int T, int H, int D,
                                long x_stride_t, long y_stride_t, float sin_sign, cudaStream_t s) {
  if (D % 16 != 0) return cudaErrorInvalidValue;
  long total = static_cast<long>(T) * H * (D / 16);
  if (total == 0) return cudaSuccess;
  rope_kernel<<<grid_for(total, 256), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x),
                                                   static_cast<__nv_bfloat16*>(y), cos_t, sin_t, T, H, D, x_stride_t,
                                                   y_stride_t, sin_sign);
  return cudaGetLastError();
}

extern "C" cudaError_t nrl_swiglu(const void* gate, const void* up, long in_stride, void* out, long T, int F,
                                  cudaStream_t s) {
  if (F % 8 != 0 || in_stride % 8 != 0) return cudaErrorInvalidValue;
  if (T == 0) return cudaSuccess;
  swiglu_kernel<<<grid_for(T * (F / 8), 256), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(gate),
                                                          static_cast<const __nv_bfloat16*>(up), in_stride,
                                                          static_cast<__nv_bfloat16*>(out), T, F);
  return cudaGetLastError();
}

extern "C" cudaError_t nrl_swiglu_bwd(const void* gate, const void* up, long in_stride, const void* gout, void* dgate,
                                      void* dup, long out_stride, long T, int F, cudaStream_t s) {
  if (F % 8 != 0 || in_stride % 8 != 0 || out_stride % 8 != 0) return cudaErrorInvalidValue;
  if (T == 0) return cudaSuccess;
  swiglu_bwd_kernel<<<grid_for(T * (F / 8), 256), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(gate), static_cast<const __nv_bfloat16*>(up), in_stride,
      static_cast<const __nv_bfloat16*>(gout), static_cast<__nv_bfloat16*>(dgate), static_cast<__nv_bfloat16*>(dup),
      out_stride, T, F);
  return cudaGetLastError();
}
