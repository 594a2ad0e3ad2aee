// fp8 (e4m3) KV cache for the fp8 rollout path: page writer with per-(token, kv-head) scales and the paged decode
// attention kernel that consumes it.  Same structure as attention_decode.cu (one CTA per (sequence, kv head, split),
// per-warp 3-stage cp.async page pipeline, all G query heads per KV read) but a page is 2 KB per K / V instead of
// 4 KB, i.e. the HBM-bound decode step reads half the bytes.  Pages are converted e4m3 -> fp16 in shared memory
// right before the tensor-core math (mma.sync f16); the per-token scales are applied to the score column (K) and
// the probability column (V) in fp32, so no per-element rescaling is needed.
#include "common.cuh"
#include "kernels.h"

namespace nrl {

constexpr int kHd = 128;
constexpr int kPg = 16;
constexpr int kWarps8 = 4;
constexpr int kStages8 = 3;
constexpr int kRawPage = kPg * kHd;                 // 2048 bytes of e4m3 (K or V of one page, one kv head)
constexpr int kRawStage = 2 * kRawPage + 2 * 64;    // K, V, k_scale[16], v_scale[16]
constexpr int kCvtTile = kPg * kHd * 2;             // 4096 bytes of fp16

NRL_DEVICE void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
NRL_DEVICE void ldsm4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
NRL_DEVICE void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
NRL_DEVICE uint32_t swz8(int r, int c) { return static_cast<uint32_t>(r * 256 + ((c ^ (r & 7)) << 4)); }
NRL_DEVICE uint32_t e4m3x2_to_f16x2(uint16_t v) {
  uint32_t r;
  asm volatile("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(r) : "h"(v));
  return r;
}
NRL_DEVICE uint32_t pack_f16x2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- page writer: one warp per (pair, kv head) ----------------------------------------------------------------
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
  const long row = (static_cast<long>(slot / kPg) * Hkv + h) * kPg + (slot % kPg);
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const __nv_bfloat16* src = (which == 0 ? k + t * k_stride_t : v + t * v_stride_t) + h * kHd + lane * 4;
    uint2 raw = *reinterpret_cast<const uint2*>(src);
    float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y);
    float amax = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(b.x), fabsf(b.y)));
    amax = warp_max(amax);
    const float sc = fmaxf(amax, 1e-12f) / 448.f, inv = 1.f / sc;
    uint16_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a.x * inv, a.y * inv), __NV_SATFINITE, __NV_E4M3);
    uint16_t hi = __nv_cvt_float2_to_fp8x2(make_float2(b.x * inv, b.y * inv), __NV_SATFINITE, __NV_E4M3);
    uint8_t* dst = (which == 0 ? kq : vq) + row * kHd + lane * 4;
    *reinterpret_cast<uint32_t*>(dst) = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
    if (lane == 0) (which == 0 ? ks : vs)[row] = sc;
  }
}

struct Decode8Params {
  const __nv_bfloat16* q;
  const uint8_t *kq, *vq;          // [num_blocks, Hkv, 16, 128] e4m3
  const float *ks, *vs;            // [num_blocks, Hkv, 16]
  const int* block_tables;
  const int* context_lens;
  __nv_bfloat16* out;
  float* part_o;
  float* part_ml;
  long q_stride_s;
  int max_blocks, Hq, Hkv, G, splits;
  float scale_log2;
};

__global__ void __launch_bounds__(kWarps8 * 32, 2) paged_decode_fp8_kernel(Decode8Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // [warp][ stage x (Kraw | Vraw | ks | vs) | Kf16 tile | Vf16 tile ] , then Q tile [16][128] fp16
  constexpr int kPerWarp = kStages8 * kRawStage + 2 * kCvtTile;
  uint8_t* q_tile = smem + kWarps8 * kPerWarp;
  const int seq = blockIdx.x, kvh = blockIdx.y, split = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int ctx = p.context_lens[seq];
  const int n_pages = (ctx + kPg - 1) / kPg;
  const int pages_per_split = (n_pages + p.splits - 1) / p.splits;
  const int page_lo = split * pages_per_split;
  const int page_hi = min(n_pages, page_lo + pages_per_split);
  const int* bt = p.block_tables + static_cast<long>(seq) * p.max_blocks;

  // Q tile in fp16, zero padded to 16 rows
  for (int i = threadIdx.x; i < 16 * 16; i += blockDim.x) {
    const int r = i >> 4, c = i & 15;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < p.G) {
      uint4 raw = *reinterpret_cast<const uint4*>(p.q + seq * p.q_stride_s + (kvh * p.G + r) * kHd + c * 8);
      uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = unpack_bf16x2(w[j]);
        w[j] = pack_f16x2(f.x, f.y);
      }
      val = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4*>(q_tile + swz8(r, c)) = val;
  }
  __syncthreads();
  uint32_t qf[8][4];
  {
    const uint32_t qbase = smem_u32(q_tile);
    const int mrow = (lane & 7) + ((lane >> 3) & 1) * 8, mcol = lane >> 4;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) ldsm4(qf[ks], qbase + swz8(mrow, ks * 2 + mcol));
  }

  uint8_t* my = smem + warp * kPerWarp;
  uint8_t* kt = my + kStages8 * kRawStage;         // converted K tile (fp16, swizzled)
  uint8_t* vt = kt + kCvtTile;
  const int my_first = page_lo + warp;
  const int my_count = (my_first < page_hi) ? (page_hi - my_first + kWarps8 - 1) / kWarps8 : 0;

  auto issue = [&](int it) {
    if (it < my_count) {
      const int page = my_first + it * kWarps8;
      const long blk = bt[page];
      const long pg = blk * p.Hkv + kvh;
      uint8_t* st = my + (it % kStages8) * kRawStage;
#pragma unroll
      for (int j = 0; j < 4; ++j) {                  // 128 chunks of 16 B per raw page
        const int chunk = lane + j * 32;
        cp_async_16(st + chunk * 16, p.kq + pg * kRawPage + chunk * 16);
        cp_async_16(st + kRawPage + chunk * 16, p.vq + pg * kRawPage + chunk * 16);
      }
      if (lane < 4) cp_async_16(st + 2 * kRawPage + lane * 16, p.ks + pg * kPg + lane * 4);
      else if (lane < 8) cp_async_16(st + 2 * kRawPage + 64 + (lane - 4) * 16, p.vs + pg * kPg + (lane - 4) * 4);
    }
    cp_async_commit();
  };

  float o[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  issue(0);
  issue(1);
  for (int it = 0; it < my_count; ++it) {
    issue(it + 2);
    cp_async_wait<2>();
    __syncwarp();
    const uint8_t* st = my + (it % kStages8) * kRawStage;
    // ---- e4m3 -> fp16 into the swizzled tiles ----
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int chunk = lane + j * 32;               // 16 fp8 values: row = chunk / 8, 16-dim group = chunk % 8
      const int r = chunk >> 3, cg = chunk & 7;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        uint4 raw = *reinterpret_cast<const uint4*>(st + which * kRawPage + chunk * 16);
        uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
        uint32_t h[8];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          h[2 * q4] = e4m3x2_to_f16x2(static_cast<uint16_t>(w[q4] & 0xffff));
          h[2 * q4 + 1] = e4m3x2_to_f16x2(static_cast<uint16_t>(w[q4] >> 16));
        }
        uint8_t* dstt = which == 0 ? kt : vt;
        *reinterpret_cast<uint4*>(dstt + swz8(r, cg * 2)) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(dstt + swz8(r, cg * 2 + 1)) = make_uint4(h[4], h[5], h[6], h[7]);
      }
    }
    __syncwarp();
    const float* kscale = reinterpret_cast<const float*>(st + 2 * kRawPage);
    const float* vscale = kscale + 16;
    const uint32_t kb = smem_u32(kt), vb = smem_u32(vt);
    const int tok0 = (my_first + it * kWarps8) * kPg;

    float s[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t kf[4];
      const int mrow = (lane & 7) + (lane >> 4) * 8, mcol = (lane >> 3) & 1;
      ldsm4(kf, kb + swz8(mrow, ks * 2 + mcol));
      mma_f16_16816(s[0], qf[ks], kf[0], kf[1]);
      mma_f16_16816(s[1], qf[ks], kf[2], kf[3]);
    }
    float tmax = -INFINITY;
    float vsc[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int tl = nt * 8 + t4 * 2 + j;
        const int tok = tok0 + tl;
        vsc[nt][j] = vscale[tl];
        s[nt][j] = (tok >= ctx) ? -INFINITY : s[nt][j] * (kscale[tl] * p.scale_log2);
        tmax = fmaxf(tmax, s[nt][j]);
      }
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
    const float m_new = fmaxf(m_run, tmax);
    const float corr = (m_new == -INFINITY) ? 1.f : exp2f(m_run - m_new);
    float psum = 0.f, pr[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float pv = (m_new == -INFINITY) ? 0.f : exp2f(s[nt][j] - m_new);
        psum += pv;
        pr[nt][j] = pv * vsc[nt][j];               // fold the V dequant scale into the probability column
      }
    psum += __shfl_xor_sync(0xffffffffu, psum, 1);
    psum += __shfl_xor_sync(0xffffffffu, psum, 2);
    l_run = l_run * corr + psum;
    m_run = m_new;
#pragma unroll
    for (int i = 0; i < 16; ++i) { o[i][0] *= corr; o[i][1] *= corr; }
    uint32_t pa[4];
    pa[0] = pack_f16x2(pr[0][0], pr[0][1]);
    pa[1] = 0u;
    pa[2] = pack_f16x2(pr[1][0], pr[1][1]);
    pa[3] = 0u;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      uint32_t vf[4];
      const int mrow = (lane & 7) + ((lane >> 3) & 1) * 8, mcol = lane >> 4;
      ldsm4_t(vf, vb + swz8(mrow, nd * 2 + mcol));
      mma_f16_16816(o[nd * 2], pa, vf[0], vf[1]);
      mma_f16_16816(o[nd * 2 + 1], pa, vf[2], vf[3]);
    }
    __syncwarp();
  }
  cp_async_wait<0>();
  __syncthreads();

  float* mo = reinterpret_cast<float*>(smem);
  float* mml = mo + kWarps8 * 8 * kHd;
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
  const int smem = kWarps8 * (kStages8 * kRawStage + 2 * kCvtTile) + 16 * kHd * 2;
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
