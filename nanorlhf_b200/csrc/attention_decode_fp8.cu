// fp8 (e4m3) paged KV cache: page writer + decode attention that reads HALF the bytes of the bf16 path.
//
// Decode attention is the dominant cost of the rollout and it is HBM-bound (attention_decode.cu runs at the copy
// roofline), so the only way to make it faster is to move fewer bytes.  Round 1 had an fp8 variant that dequantised
// whole pages through shared memory and was issue-bound (slower than bf16).  This version dequantises IN REGISTERS,
// on the operand fragments of the tensor-core instruction itself:
//
//   * K page  [16 tokens][128 d] e4m3 (2 KB);  V page stored TRANSPOSED [128 d][16 tokens] e4m3 (2 KB);
//     one fp32 scale per (token, kv head) for K and for V (64 + 64 B)   ->  4.1 KB per page instead of 8 KB.
//   * both MMAs are mma.sync m16n8k16 f16 (e4m3 values are exact in f16).  The B fragment of that instruction wants, per
//     thread, the pairs k = {2t, 2t+1} and {2t+8, 2t+9} of one column.  The contraction index is a dummy, so it is RENUMBERED:
//     virtual k {2t, 2t+1, 2t+8, 2t+9} := physical {4t, 4t+1, 4t+2, 4t+3}.  One 32-bit shared-memory load then holds exactly
//     the four e4m3 values a thread needs, and two `cvt.rn.f16x2.e4m3x2` turn it into the two B registers -- no unpacking,
//     no ldmatrix, no shuffles.  Q (the A operand of S = Q K^T) is loaded with the same renumbering of d; for O = P V the
//     renumbered index is the token, which fixes which physical key each score column must hold: lane group g of score tile
//     j loads key 4 (g >> 1) + 2 j + (g & 1), so the accumulator registers of S already sit where the A fragment of P needs them.
//   * per-token scales: S[:, tok] *= k_scale[tok] after the first MMA, P[:, tok] *= v_scale[tok] * 2^10 before the second
//     (the 2^10 keeps small probabilities out of the f16 subnormal range; it is divided out with the softmax denominator).
//   * per-warp 4-stage cp.async page pipeline, all G = Hq / Hkv query heads per KV read, split-KV merge -- as in the
//     bf16 kernel; 3 CTAs per SM (68 KB each) keep ~200 KB of loads in flight per SM.
//
// Reference path replaced: vLLM PagedAttention / FlashInfer decode inside llm.generate
// (/root/reference/GRPO/grpo_trainer.py:142; SURVEY.md section 2.5 K2, "fp8 KV option").
#include "common.cuh"
#include "kernels.h"

namespace nrl {

constexpr int kHd = 128;
constexpr int kPg = 16;
constexpr int kWarps8 = 4;
constexpr int kStages8 = 4;
constexpr int kRawPage = kPg * kHd;                 // 2048 bytes of e4m3 (K or V of one page, one kv head)
constexpr int kStage8 = 2 * kRawPage + 2 * 64;      // K, V^T, k_scale[16], v_scale[16]
constexpr float kPScale = 1024.f;

NRL_DEVICE void mma_f16_16816(float (&d)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  // rows 8..15 of the A tile are padding (G <= 8 query heads): a1 = a3 = 0
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}
NRL_DEVICE uint32_t e4m3x2_to_f16x2(uint32_t v16) {
  uint32_t r;
  asm volatile("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(r) : "h"(static_cast<uint16_t>(v16)));
  return r;
}
NRL_DEVICE uint32_t pack_f16x2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
NRL_DEVICE uint32_t lds32(uint32_t addr) {
  uint32_t r;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(addr));
  return r;
}
// K rows are 128 B = 8 chunks of 16 B; chunk c of token row r lives at chunk (c ^ sigma(r)), sigma(r) = 2 (r >> 2) + (r & 1):
// the 8 rows a warp reads together (4 (g >> 1) + 2 j + (g & 1), g = 0..7) have sigma = g, i.e. 8 distinct chunks -> 32 banks.
NRL_DEVICE int k_sigma(int r) { return ((r >> 2) << 1) | (r & 1); }

// ---- page writer: one warp per (token, kv head) ----------------------------------------------------------------
// kq [num_blocks, Hkv, 16, 128]   vq [num_blocks, Hkv, 128, 16] (transposed)   ks, vs [num_blocks, Hkv, 16]
__global__ void kv_cache_write_fp8_kernel(const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                                          long k_stride_t, long v_stride_t, uint8_t* __restrict__ kq,
                                          uint8_t* __restrict__ vq, float* __restrict__ ks, float* __restrict__ vs,
                                          const int* __restrict__ slot_mapping, const int* __restrict__ src_index,
                                          int pairs, int Hkv) {
  const long widx = blockIdx.x * static_cast<long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (widx >= static_cast<long>(pairs) * Hkv) return;
  const int lane = threadIdx.x & 31;
  const int pair = widx / Hkv, h = widx % Hkv;
  const int slot = slot_mapping[pair];
  if (slot < 0) return;
  const long t = src_index ? src_index[pair] : pair;
  const long page = static_cast<long>(slot / kPg) * Hkv + h;
  const int tok = slot % kPg;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const __nv_bfloat16* src = (which == 0 ? k + t * k_stride_t : v + t * v_stride_t) + h * kHd + lane * 4;
    const uint2 raw = *reinterpret_cast<const uint2*>(src);
    const float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y);
    float amax = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(b.x), fabsf(b.y)));
    amax = warp_max(amax);
    const float sc = fmaxf(amax, 1e-12f) / 448.f, inv = 1.f / sc;
    const uint16_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a.x * inv, a.y * inv), __NV_SATFINITE, __NV_E4M3);
    const uint16_t hi = __nv_cvt_float2_to_fp8x2(make_float2(b.x * inv, b.y * inv), __NV_SATFINITE, __NV_E4M3);
    if (which == 0) {
      *reinterpret_cast<uint32_t*>(kq + (page * kPg + tok) * kHd + lane * 4) = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
      if (lane == 0) ks[page * kPg + tok] = sc;
    } else {
      uint8_t* dst = vq + page * kRawPage + (lane * 4) * kPg + tok;        // V^T[d][tok]
      dst[0] = static_cast<uint8_t>(lo & 0xFF);
      dst[kPg] = static_cast<uint8_t>(lo >> 8);
      dst[2 * kPg] = static_cast<uint8_t>(hi & 0xFF);
      dst[3 * kPg] = static_cast<uint8_t>(hi >> 8);
      if (lane == 0) vs[page * kPg + tok] = sc;
    }
  }
}

struct Decode8Params {
  const __nv_bfloat16* q;
  const uint8_t *kq, *vq;
  const float *ks, *vs;
  const int* block_tables;
  const int* context_lens;
  __nv_bfloat16* out;
  float* part_o;
  float* part_ml;
  long q_stride_s;
  int max_blocks, Hq, Hkv, G, splits;
  float scale_log2;
};

__global__ void __launch_bounds__(kWarps8 * 32, 3) paged_decode_fp8_kernel(Decode8Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int seq = blockIdx.x, kvh = blockIdx.y, split = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int ctx = p.context_lens[seq];
  const int n_pages = (ctx + kPg - 1) / kPg;
  const int pages_per_split = (n_pages + p.splits - 1) / p.splits;
  const int page_lo = split * pages_per_split;
  const int page_hi = min(n_pages, page_lo + pages_per_split);
  const int* bt = p.block_tables + static_cast<long>(seq) * p.max_blocks;

  // ---- Q fragments straight from global: row g (query head g of this kv head), d = 16 ks + 4 t .. + 3, as f16 ----
  uint32_t qa0[8], qa2[8];
  {
    const bool real = g < p.G;
    const __nv_bfloat16* qrow = p.q + seq * p.q_stride_s + (kvh * p.G + (real ? g : 0)) * kHd;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint2 raw = *reinterpret_cast<const uint2*>(qrow + ks * 16 + t4 * 4);
      const float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y);
      qa0[ks] = real ? pack_f16x2(a.x, a.y) : 0u;
      qa2[ks] = real ? pack_f16x2(b.x, b.y) : 0u;
    }
  }

  uint8_t* my_smem = smem + warp * kStages8 * kStage8;
  const int my_first = page_lo + warp;
  const int my_count = (my_first < page_hi) ? (page_hi - my_first + kWarps8 - 1) / kWarps8 : 0;

  auto issue = [&](int it) {
    if (it < my_count) {
      const int page = my_first + it * kWarps8;
      const long pg = static_cast<long>(bt[page]) * p.Hkv + kvh;
      const uint8_t* ksrc = p.kq + pg * kRawPage;
      const uint8_t* vsrc = p.vq + pg * kRawPage;
      uint8_t* kd = my_smem + (it % kStages8) * kStage8;
      uint8_t* vd = kd + kRawPage;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int chunk = lane + j * 32;          // 128 chunks of 16 B per 2 KB page
        const int r = chunk >> 3, c = chunk & 7;
        cp_async_16(kd + r * 128 + ((c ^ k_sigma(r)) << 4), ksrc + chunk * 16);
        cp_async_16(vd + chunk * 16, vsrc + chunk * 16);
      }
      if (lane < 8) {
        const float* ssrc = (lane < 4 ? p.ks : p.vs) + pg * kPg + (lane & 3) * 4;
        cp_async_16(vd + kRawPage + lane * 16, ssrc);
      }
    }
    cp_async_commit();
  };

  float o[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;     // row g

#pragma unroll
  for (int s = 0; s < kStages8 - 1; ++s) issue(s);
  for (int it = 0; it < my_count; ++it) {
    issue(it + kStages8 - 1);
    cp_async_wait<kStages8 - 1>();
    __syncwarp();
    const uint32_t kb = smem_u32(my_smem + (it % kStages8) * kStage8);
    const uint32_t vb = kb + kRawPage;
    const uint32_t sb = vb + kRawPage;
    const int tok0 = (my_first + it * kWarps8) * kPg;

    // ---- S = Q K^T : score tile j holds physical keys 4 t + 2 j + {0, 1} in its columns {2t, 2t+1} ----
    float s[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int krow = ((g >> 1) << 2) + 2 * j + (g & 1);            // sigma(krow) == g
      const uint32_t rbase = kb + krow * 128 + t4 * 4;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t w = lds32(rbase + ((ks ^ g) << 4));
        mma_f16_16816(s[j], qa0[ks], qa2[ks], e4m3x2_to_f16x2(w & 0xFFFFu), e4m3x2_to_f16x2(w >> 16));
      }
    }
    // ---- scale, mask, online softmax on row g ----
    float4 ksc, vsc;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(ksc.x), "=f"(ksc.y), "=f"(ksc.z), "=f"(ksc.w) : "r"(sb + t4 * 16));
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(vsc.x), "=f"(vsc.y), "=f"(vsc.z), "=f"(vsc.w) : "r"(sb + 64 + t4 * 16));
    const float kscale[2][2] = {{ksc.x, ksc.y}, {ksc.z, ksc.w}};
    const float vscale[2][2] = {{vsc.x, vsc.y}, {vsc.z, vsc.w}};
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int tok = tok0 + t4 * 4 + 2 * j + e;
        s[j][e] = (tok >= ctx) ? -INFINITY : s[j][e] * kscale[j][e] * p.scale_log2;
        tmax = fmaxf(tmax, s[j][e]);
      }
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
    const float m_new = fmaxf(m_run, tmax);
    const float corr = (m_new == -INFINITY) ? 1.f : exp2f(m_run - m_new);
    float psum = 0.f;
    float pr[2][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        pr[j][e] = (m_new == -INFINITY) ? 0.f : exp2f(s[j][e] - m_new);
        psum += pr[j][e];
      }
    psum += __shfl_xor_sync(0xffffffffu, psum, 1);
    psum += __shfl_xor_sync(0xffffffffu, psum, 2);
    l_run = l_run * corr + psum;
    m_run = m_new;
#pragma unroll
    for (int i = 0; i < 16; ++i) { o[i][0] *= corr; o[i][1] *= corr; }
    // ---- O += (P * v_scale) V : A fragment = this thread's own score registers ----
    const uint32_t pa0 = pack_f16x2(pr[0][0] * vscale[0][0] * kPScale, pr[0][1] * vscale[0][1] * kPScale);   // tokens 4t, 4t+1
    const uint32_t pa2 = pack_f16x2(pr[1][0] * vscale[1][0] * kPScale, pr[1][1] * vscale[1][1] * kPScale);   // tokens 4t+2, 4t+3
    const uint32_t vrow = vb + g * kPg + t4 * 4;                          // V^T[d = 8 nd + g][tokens 4t .. 4t+3]
#pragma unroll
    for (int nd = 0; nd < 16; ++nd) {
      const uint32_t w = lds32(vrow + nd * 8 * kPg);
      mma_f16_16816(o[nd], pa0, pa2, e4m3x2_to_f16x2(w & 0xFFFFu), e4m3x2_to_f16x2(w >> 16));
    }
    __syncwarp();
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- merge the 4 warps through shared memory (re-using the pipeline buffers: 4 x 8 x 128 fp32 = 16 KB) ----
  float* mo = reinterpret_cast<float*>(smem);                 // [warp][8][128]
  float* mml = mo + kWarps8 * 8 * kHd;                        // [warp][8][2]
#pragma unroll
  for (int nd = 0; nd < 16; ++nd) {
    mo[(warp * 8 + g) * kHd + nd * 8 + t4 * 2] = o[nd][0];
    mo[(warp * 8 + g) * kHd + nd * 8 + t4 * 2 + 1] = o[nd][1];
  }
  if (t4 == 0) {
    mml[(warp * 8 + g) * 2] = m_run;
    mml[(warp * 8 + g) * 2 + 1] = l_run;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.G * kHd; i += blockDim.x) {
    const int r = i / kHd, d = i % kHd;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kWarps8; ++w) M = fmaxf(M, mml[(w * 8 + r) * 2]);
    float L = 0.f, acc = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps8; ++w) {
      const float mw = mml[(w * 8 + r) * 2];
      const float c = (mw == -INFINITY) ? 0.f : exp2f(mw - M);
      L += mml[(w * 8 + r) * 2 + 1] * c;
      acc += mo[(w * 8 + r) * kHd + d] * c;
    }
    acc *= (1.f / kPScale);
    if (p.splits == 1) {
      p.out[(static_cast<long>(seq) * p.Hq + kvh * p.G + r) * kHd + d] = __float2bfloat16(L > 0.f ? acc / L : 0.f);
    } else {
      const long base = ((static_cast<long>(seq) * p.Hkv + kvh) * p.splits + split) * 8 + r;
      p.part_o[base * kHd + d] = acc;
      if (d == 0) {
        p.part_ml[base * 2] = M;
        p.part_ml[base * 2 + 1] = L;
      }
    }
  }
}

__global__ void decode8_merge_splits_kernel(Decode8Params p) {
  const int seq = blockIdx.x, kvh = blockIdx.y;
  for (int i = threadIdx.x; i < p.G * kHd; i += blockDim.x) {
    const int r = i / kHd, d = i % kHd;
    const long base = ((static_cast<long>(seq) * p.Hkv + kvh) * p.splits) * 8 + r;
    float M = -INFINITY;
    for (int s = 0; s < p.splits; ++s) M = fmaxf(M, p.part_ml[(base + s * 8) * 2]);
    float L = 0.f, acc = 0.f;
    for (int s = 0; s < p.splits; ++s) {
      const float ms = p.part_ml[(base + s * 8) * 2];
      const float c = (ms == -INFINITY) ? 0.f : exp2f(ms - M);
      L += p.part_ml[(base + s * 8) * 2 + 1] * c;
      acc += p.part_o[(base + s * 8) * kHd + d] * c;
    }
    p.out[(static_cast<long>(seq) * p.Hq + kvh * p.G + r) * kHd + d] = __float2bfloat16(L > 0.f ? acc / L : 0.f);
  }
}

}  // namespace nrl

using namespace nrl;

extern "C" cudaError_t nrl_kv_cache_write_fp8(const void* k, const void* v, long k_stride_t, long v_stride_t, void* kq, void* vq,
                                              float* ks, float* vs, const int* slot_mapping, const int* src_index, int pairs,
                                              int Hkv, int head_dim, int page, cudaStream_t s) {
  if (head_dim != kHd || page != kPg) return cudaErrorInvalidValue;
  if (pairs == 0) return cudaSuccess;
  const long warps = static_cast<long>(pairs) * Hkv;
  kv_cache_write_fp8_kernel<<<static_cast<int>((warps + 7) / 8), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(k), static_cast<const __nv_bfloat16*>(v), k_stride_t, v_stride_t,
      static_cast<uint8_t*>(kq), static_cast<uint8_t*>(vq), ks, vs, slot_mapping, src_index, pairs, Hkv);
  return cudaGetLastError();
}

extern "C" cudaError_t nrl_paged_decode_fp8(const void* q, long q_stride_s, const void* kq, const void* vq, const float* ks,
                                            const float* vs, const int* block_tables, const int* context_lens, void* out,
                                            float* part_o, float* part_ml, int S, int Hq, int Hkv, int head_dim, int page,
                                            int max_blocks, int splits, float scale, cudaStream_t s) {
  if (head_dim != kHd || page != kPg || Hq % Hkv != 0 || Hq / Hkv > 8) return cudaErrorInvalidValue;
  if (S == 0) return cudaSuccess;
  Decode8Params p;
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.kq = static_cast<const uint8_t*>(kq);
  p.vq = static_cast<const uint8_t*>(vq);
  p.ks = ks; p.vs = vs;
  p.block_tables = block_tables; p.context_lens = context_lens;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.part_o = part_o; p.part_ml = part_ml;
  p.q_stride_s = q_stride_s;
  p.max_blocks = max_blocks; p.Hq = Hq; p.Hkv = Hkv; p.G = Hq / Hkv; p.splits = splits;
  p.scale_log2 = scale * 1.4426950408889634f;
  const int smem = kWarps8 * kStages8 * kStage8;          // 67,584 B (>= the 16.3 KB the warp merge needs)
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(paged_decode_fp8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  paged_decode_fp8_kernel<<<dim3(S, Hkv, splits), kWarps8 * 32, smem, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (splits > 1) {
    decode8_merge_splits_kernel<<<dim3(S, Hkv), 128, 0, s>>>(p);
    e = cudaGetLastError();
  }
  return e;
}
