// fp8 (e4m3) quantisation kernels for the fp8 rollout path (rollout_dtype="fp8"):
//   quant_rows_e4m3   x[M,K] bf16 -> q[M,K] e4m3 + scale[M] (amax / 448 per row): activations per token,
//                     weights per output channel (the sampler arena is quantised once per weight refresh)
// The matching GEMM is the tcgen05 kind::f8f6f4 instantiation in gemm_sm100.cu (scales applied in its epilogue).
#include "common.cuh"
#include "kernels.h"

namespace nrl {

__global__ void __launch_bounds__(256) quant_rows_e4m3_kernel(const __nv_bfloat16* __restrict__ x, long x_stride,
                                                              uint8_t* __restrict__ q, long q_stride,
                                                              float* __restrict__ scale, int K) {
  const long row = blockIdx.x;
  const __nv_bfloat16* xr = x + row * x_stride;
  const int nvec = K / 8;
  float amax = 0.f;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    uint4 a = reinterpret_cast<const uint4*>(xr)[v];
    uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = unpack_bf16x2(w[j]);
      amax = fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y)));
    }
  }
  __shared__ float red[8];
  amax = warp_max(amax);
  if (lane_id() == 0) red[threadIdx.x >> 5] = amax;
  __syncthreads();
  amax = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) amax = fmaxf(amax, (i < (blockDim.x >> 5)) ? red[i] : 0.f);
  const float sc = fmaxf(amax, 1e-12f) / 448.f;
  const float inv = 1.f / sc;
  if (threadIdx.x == 0) scale[row] = sc;
  uint8_t* qr = q + row * q_stride;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    uint4 a = reinterpret_cast<const uint4*>(xr)[v];
    uint32_t w[4] = {a.x, a.y, a.z, a.w};
    uint16_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = unpack_bf16x2(w[j]);
      o[j] = __nv_cvt_float2_to_fp8x2(make_float2(f.x * inv, f.y * inv), __NV_SATFINITE, __NV_E4M3);
    }
    uint2 packed = make_uint2(static_cast<uint32_t>(o[0]) | (static_cast<uint32_t>(o[1]) << 16),
                              static_cast<uint32_t>(o[2]) | (static_cast<uint32_t>(o[3]) << 16));
    reinterpret_cast<uint2*>(qr)[v] = packed;
  }
}

}  // namespace nrl

extern "C" cudaError_t nrl_quant_rows_e4m3(const void* x, long x_stride, void* q, long q_stride, float* scale, int M, int K,
                                           cudaStream_t s) {
  if (M == 0) return cudaSuccess;
  if (K % 8 != 0 || x_stride % 8 != 0 || q_stride % 8 != 0) return cudaErrorInvalidValue;
  nrl::quant_rows_e4m3_kernel<<<M, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), x_stride, static_cast<uint8_t*>(q),
                                                q_stride, scale, K);
  return cudaGetLastError();
}
