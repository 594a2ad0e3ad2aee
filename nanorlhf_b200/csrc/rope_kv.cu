// Decode-step fusion of RoPE(q), RoPE(k) and the KV page write ("RoPE fused on K at write").
//
// The decode step used to launch three kernels per layer for this (rope on q, rope on k, page writer); at 1024 sequences
// each moves a few MB, i.e. they are launch-latency sized.  Here ONE warp handles one (sequence, head) of the fused qkv
// projection output [S, (Hq + 2 Hkv) * 128]:
//   q head   : rotate in place;
//   k head j : rotate in place (prefix-attention consumers see the rotated k), then store k_j and v_j of this token into
//              their page slot -- bf16 pages [blk][Hkv][16][128], or e4m3 pages with one scale per (token, head):
//              K [blk][Hkv][16][128], V TRANSPOSED [blk][Hkv][128][16] (the layout attention_decode_fp8.cu reads).
// cos / sin come as [S, 64] fp32 tables (computed once per step from the positions).
#include "common.cuh"
#include "kernels.h"

namespace nrl {

constexpr int kRD = 128, kRPage = 16;

template <bool KV8>
__global__ void __launch_bounds__(256) rope_kv_write_kernel(__nv_bfloat16* __restrict__ qkv, long stride_s,
                                                            const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                                            void* __restrict__ k_cache, void* __restrict__ v_cache,
                                                            float* __restrict__ k_scale, float* __restrict__ v_scale,
                                                            const int* __restrict__ slot_mapping, int S, int Hq, int Hkv) {
  const long widx = blockIdx.x * static_cast<long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int heads = Hq + Hkv;
  if (widx >= static_cast<long>(S) * heads) return;
  const int lane = threadIdx.x & 31;
  const int s = widx / heads, h = widx % heads;
  __nv_bfloat16* x = qkv + s * stride_s + static_cast<long>(h) * kRD;           // q heads then k heads are contiguous
  const float2 c = *reinterpret_cast<const float2*>(cos_t + static_cast<long>(s) * 64 + 2 * lane);
  const float2 sn = *reinterpret_cast<const float2*>(sin_t + static_cast<long>(s) * 64 + 2 * lane);
  const float2 lo = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + 2 * lane));
  const float2 hi = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + 64 + 2 * lane));
  const float r0 = lo.x * c.x - hi.x * sn.x, r1 = lo.y * c.y - hi.y * sn.y;     // dims 2l, 2l+1
  const float r2 = hi.x * c.x + lo.x * sn.x, r3 = hi.y * c.y + lo.y * sn.y;     // dims 64+2l, 65+2l
  const uint32_t olo = pack_bf16x2(r0, r1), ohi = pack_bf16x2(r2, r3);
  *reinterpret_cast<uint32_t*>(x + 2 * lane) = olo;
  *reinterpret_cast<uint32_t*>(x + 64 + 2 * lane) = ohi;
  if (h < Hq) return;
  // ---- k head j = h - Hq: page write of k_j (rotated) and v_j ----
  const int j = h - Hq;
  const int slot = slot_mapping[s];
  if (slot < 0) return;
  const long page = static_cast<long>(slot / kRPage) * Hkv + j;
  const int tok = slot % kRPage;
  const __nv_bfloat16* vsrc = qkv + s * stride_s + static_cast<long>(Hq + Hkv + j) * kRD + lane * 4;
  const uint2 vraw = *reinterpret_cast<const uint2*>(vsrc);
  if (!KV8) {
    __nv_bfloat16* kd = static_cast<__nv_bfloat16*>(k_cache) + (page * kRPage + tok) * kRD;
    __nv_bfloat16* vd = static_cast<__nv_bfloat16*>(v_cache) + (page * kRPage + tok) * kRD;
    *reinterpret_cast<uint32_t*>(kd + 2 * lane) = olo;
    *reinterpret_cast<uint32_t*>(kd + 64 + 2 * lane) = ohi;
    *reinterpret_cast<uint2*>(vd + lane * 4) = vraw;
  } else {
    // quantise from the bf16-rounded values (what the unfused writer sees)
    const float2 k01 = unpack_bf16x2(olo), k23 = unpack_bf16x2(ohi);
    float amax = warp_max(fmaxf(fmaxf(fabsf(k01.x), fabsf(k01.y)), fmaxf(fabsf(k23.x), fabsf(k23.y))));
    float sc = fmaxf(amax, 1e-12f) / 448.f, inv = 1.f / sc;
    uint8_t* kq = static_cast<uint8_t*>(k_cache) + (page * kRPage + tok) * kRD;
    *reinterpret_cast<uint16_t*>(kq + 2 * lane) = __nv_cvt_float2_to_fp8x2(make_float2(k01.x * inv, k01.y * inv), __NV_SATFINITE, __NV_E4M3);
    *reinterpret_cast<uint16_t*>(kq + 64 + 2 * lane) = __nv_cvt_float2_to_fp8x2(make_float2(k23.x * inv, k23.y * inv), __NV_SATFINITE, __NV_E4M3);
    if (lane == 0) k_scale[page * kRPage + tok] = sc;
    const float2 a = unpack_bf16x2(vraw.x), b = unpack_bf16x2(vraw.y);
    amax = warp_max(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(b.x), fabsf(b.y))));
    sc = fmaxf(amax, 1e-12f) / 448.f;
    inv = 1.f / sc;
    const uint16_t q01 = __nv_cvt_float2_to_fp8x2(make_float2(a.x * inv, a.y * inv), __NV_SATFINITE, __NV_E4M3);
    const uint16_t q23 = __nv_cvt_float2_to_fp8x2(make_float2(b.x * inv, b.y * inv), __NV_SATFINITE, __NV_E4M3);
    uint8_t* vq = static_cast<uint8_t*>(v_cache) + page * (kRPage * kRD) + (lane * 4) * kRPage + tok;      // V^T[d][tok]
    vq[0] = static_cast<uint8_t>(q01 & 0xFF);
    vq[kRPage] = static_cast<uint8_t>(q01 >> 8);
    vq[2 * kRPage] = static_cast<uint8_t>(q23 & 0xFF);
    vq[3 * kRPage] = static_cast<uint8_t>(q23 >> 8);
    if (lane == 0) v_scale[page * kRPage + tok] = sc;
  }
}

}  // namespace nrl

extern "C" cudaError_t nrl_rope_kv_write(void* qkv, long stride_s, const float* cos_t, const float* sin_t, void* k_cache, void* v_cache,
                                         float* k_scale, float* v_scale, const int* slot_mapping, int S, int Hq, int Hkv, int head_dim,
                                         int page, int kv8, cudaStream_t s) {
  using namespace nrl;
  if (head_dim != kRD || page != kRPage) return cudaErrorInvalidValue;
  if (S == 0) return cudaSuccess;
  const long warps = static_cast<long>(S) * (Hq + Hkv);
  const int blocks = static_cast<int>((warps + 7) / 8);
  if (kv8)
    rope_kv_write_kernel<true><<<blocks, 256, 0, s>>>(static_cast<__nv_bfloat16*>(qkv), stride_s, cos_t, sin_t, k_cache, v_cache, k_scale,
                                                     v_scale, slot_mapping, S, Hq, Hkv);
  else
    rope_kv_write_kernel<false><<<blocks, 256, 0, s>>>(static_cast<__nv_bfloat16*>(qkv), stride_s, cos_t, sin_t, k_cache, v_cache, nullptr,
                                                      nullptr, slot_mapping, S, Hq, Hkv);
  return cudaGetLastError();
}
